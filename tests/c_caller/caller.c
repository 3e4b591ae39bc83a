/* A plain C caller of the drop-in library, written against the headers in
 * include/ under the reference's own header names (the way the reference's
 * test/test.c:1-72 is written against inc/): caller-allocated outputs with that
 * file's sizes, scene_load + compute_paths.  Prints a few order-independent
 * numbers that tests/test_library_abi.py and tests/test_gpu_parity.py compare
 * with the oracle.  usage: caller scene.hrt num_paths num_bounces f_GHz */
#include "compute_paths.h" /* compute_paths, ChannelInfo, RaysInfo */
#include "scene.h"         /* Scene, scene_load */
#include "vec3.h"          /* Vec3 */
#include "ray.h"           /* Ray */

#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void *zalloc(size_t n) { void *p = calloc(n ? n : 1, 1); if (!p) { perror("calloc"); exit(8); } return p; }

int main(int argc, char **argv)
{
  if (argc < 5) { fprintf(stderr, "usage: %s scene.hrt num_paths num_bounces f_GHz\n", argv[0]); return 1; }
  const size_t R = 2, T = 1, P = (size_t)atol(argv[2]), B = (size_t)atol(argv[3]);
  const float f = (float)atof(argv[4]);
  Vec3 rx_pos[2] = {{0.f, 0.f, .5f}, {0.4f, -0.3f, 1.25f}}, tx_pos[1] = {{0.f, 0.f, .5f}};
  Vec3 rx_vel[2] = {{0.f, 0.f, 0.f}, {1.f, 0.f, 0.f}}, tx_vel[1] = {{0.f, 2.f, 0.f}};
  const size_t nl = R * T, ns = R * T * B * P;
  ChannelInfo los = { 1, zalloc(nl * sizeof(Vec3)), zalloc(nl * sizeof(Vec3)), zalloc(nl * 4), zalloc(nl * 4),
                      zalloc(nl * 4), zalloc(nl * 4), zalloc(nl * 4), zalloc(nl * 4) };
  RaysInfo rlos = { 1, 1, zalloc(nl * sizeof(Ray)), zalloc(nl / 8 + 1) };
  ChannelInfo sc = { (uint32_t)(B * P), zalloc(ns * sizeof(Vec3)), zalloc(P * sizeof(Vec3)), zalloc(ns * 4), zalloc(ns * 4),
                     zalloc(ns * 4), zalloc(ns * 4), zalloc(ns * 4), zalloc(ns * 4) };
  RaysInfo rsc = { (uint32_t)(B + 1), (uint32_t)P, zalloc(T * (B + 1) * P * sizeof(Ray)), zalloc(T * (B + 1) * (P / 8 + 1)) };
  Scene scene = scene_load(argv[1]);
  /* argv[5]: number of calls on the same scene (exercises the uploaded-scene cache and the
   * replacement of Mesh.ns); the outputs of the last call are reported */
  const int calls = argc > 5 ? atoi(argv[5]) : 1;
  for (int k = 0; k < calls; ++k)
    compute_paths(&scene, rx_pos, tx_pos, rx_vel, tx_vel, f, R, T, P, B, &los, &rlos, &sc, &rsc);
  /* FNV-1a over every scatter output array the call determines completely */
  uint64_t h = 0xCBF29CE484222325ull;
  {
    const void *arr[] = { sc.directions_rx, sc.a_te_re, sc.a_te_im, sc.a_tm_re, sc.a_tm_im, sc.tau, sc.freq_shift, rsc.rays, rsc.rays_active };
    const size_t len[] = { ns * sizeof(Vec3), ns * 4, ns * 4, ns * 4, ns * 4, ns * 4, ns * 4, T * (B + 1) * P * sizeof(Ray), T * (B + 1) * (P / 8 + 1) };
    for (int a = 0; a < 9; ++a) {
      const unsigned char *b = (const unsigned char *)arr[a];
      for (size_t i = 0; i < len[a]; ++i) { h ^= b[i]; h *= 0x100000001B3ull; }
    }
  }
  uint64_t n_paths = 0, tau_bits = 0;
  for (size_t i = 0; i < ns; ++i)
    if (sc.tau[i] != 0.f) { uint32_t w; memcpy(&w, &sc.tau[i], 4); ++n_paths; tau_bits += w; }
  uint32_t los_bits[2];
  memcpy(los_bits, los.tau, 8);
  printf("{\"paths\": %" PRIu64 ", \"tau_bits\": %" PRIu64 ", \"los_tau_bits\": [%u, %u], \"normal0_z\": %.9g, \"hash\": \"%016" PRIx64 "\"}\n",
         n_paths, tau_bits, los_bits[0], los_bits[1], scene.meshes[0].ns ? scene.meshes[0].ns[0].z : 0.0, h);
  free_scene(&scene);
  return 0;
}
