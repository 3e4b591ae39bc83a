/* ref_harness.c -- headless driver around a compute_paths() library.
 * TEST INFRASTRUCTURE ONLY (bench.py cpu_baseline / --impl reference).
 *
 * Plays the role of the reference's test/test.c:10-87 without the GLUT viewer:
 * allocates the caller-owned outputs with the sizes of test/test.c:29-60,
 * loads a scene, times compute_paths() alone with CLOCK_MONOTONIC and prints
 * one JSON object: seconds, ray-bounces (primary closest-hit queries, derived
 * from the RaysInfo activity masks) and valid/occluded path counts.
 *
 * usage: ref_harness scene.hrt f_GHz P B R T  rx(3R floats) tx(3T floats) [reps]
 * Linked against oracle/_ref/libhrt_ref.so (the unmodified reference).
 */
#include "hermespy_rt.h"
#include <stdio.h>
#include <time.h>

static double now(void)
{ struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

int main(int argc, char **argv)
{
  if (argc < 7) { fprintf(stderr, "usage: %s scene f_GHz P B R T rx... tx... [reps]\n", argv[0]); return 2; }
  const char *path = argv[1];
  float f = strtof(argv[2], NULL);
  size_t P = strtoull(argv[3], NULL, 10), B = strtoull(argv[4], NULL, 10);
  size_t R = strtoull(argv[5], NULL, 10), T = strtoull(argv[6], NULL, 10);
  if ((size_t)argc < 7 + 3 * R + 3 * T) { fprintf(stderr, "missing positions\n"); return 2; }
  Vec3 *rx = calloc(R, sizeof(Vec3)), *tx = calloc(T, sizeof(Vec3));
  Vec3 *rv = calloc(R, sizeof(Vec3)), *tv = calloc(T, sizeof(Vec3));
  int a = 7;
  for (size_t i = 0; i < R; ++i) { rx[i].x = strtof(argv[a++], 0); rx[i].y = strtof(argv[a++], 0); rx[i].z = strtof(argv[a++], 0); }
  for (size_t i = 0; i < T; ++i) { tx[i].x = strtof(argv[a++], 0); tx[i].y = strtof(argv[a++], 0); tx[i].z = strtof(argv[a++], 0); }
  int reps = a < argc ? atoi(argv[a]) : 1;

  size_t nl = R * T, ns = R * T * B * P;
  ChannelInfo los = { 1, calloc(nl, 12), calloc(nl, 12), calloc(nl, 4), calloc(nl, 4),
                      calloc(nl, 4), calloc(nl, 4), calloc(nl, 4), calloc(nl, 4) };
  ChannelInfo sc = { (uint32_t)(B * P), calloc(ns, 12), calloc(ns, 12), calloc(ns, 4), calloc(ns, 4),
                     calloc(ns, 4), calloc(ns, 4), calloc(ns, 4), calloc(ns, 4) };
  RaysInfo rl = { 1, 1, calloc(nl, sizeof(Ray)), calloc(nl / 8 + 1, 1) };
  size_t rows = T * (B + 1) + 1;   /* +1 row: the reference overruns by design */
  RaysInfo rs = { (uint32_t)(B + 1), (uint32_t)P, calloc(rows * P, sizeof(Ray)), calloc(rows * (P / 8 + 1), 1) };

  Scene scene = scene_load(path);
  double best = 1e30;
  for (int r = 0; r < reps; ++r) {
    for (uint32_t m = 0; m < scene.num_meshes; ++m) { free(scene.meshes[m].ns); scene.meshes[m].ns = NULL; }
    double t0 = now();
    compute_paths(&scene, rx, tx, rv, tv, f, R, T, P, B, &los, &rl, &sc, &rs);
    double dt = now() - t0;
    if (dt < best) best = dt;
  }

  /* ray-bounces: rays alive at the start of each bounce.  Row 0 of the mask is
   * all ones; row (t*B+b+1) holds the state after bounce b -- for T == 1 only
   * (SURVEY appendix A-9), so count per-TX only when T == 1. */
  unsigned long long rb = 0, valid = 0, occl = 0;
  if (T == 1) {
    rb = P;
    for (size_t b = 0; b + 1 < B; ++b) {
      const uint8_t *row = rs.rays_active + (b + 1) * (P / 8 + 1);
      for (size_t p = 0; p < P; ++p) rb += (row[p >> 3] >> (p & 7)) & 1;
    }
  }
  (void)valid; (void)occl;
  printf("{\"seconds\": %.6f, \"ray_bounces\": %llu, \"P\": %zu, \"B\": %zu, \"R\": %zu, \"T\": %zu}\n",
         best, rb, P, B, R, T);
  return 0;
}
