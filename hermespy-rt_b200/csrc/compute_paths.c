/* compute_paths.c -- the drop-in entry point (host C).
 *
 * Same symbol, arguments, units, output layout and ownership rules as the
 * reference's compute_paths() (inc/compute_paths.h:59-74,
 * src/compute_paths.c:419-757); the work is done by the CUDA layer behind
 * include/hrt_cuda.h.  What stays on the host, as in the reference:
 *   - the material table at the carrier frequency (precompute_materials,
 *     :171-206) -- 17 rows through the host libm, then uploaded;
 *   - Mesh.ns: the reference mallocs per-mesh normals into the caller's scene
 *     (:208-224); callers may read them (the viewer does), so they are filled
 *     from the GPU-computed normals (cached with the uploaded scene).  A
 *     previous array is freed before the new one is attached -- the reference
 *     leaks it on repeated calls; freeing the last one is free_scene()'s job.
 * Errors follow the reference convention: message on stderr and exit(8)
 * (inc/common.h:20-25).  There is no CPU fallback.
 *
 * Environment:
 *   HRT_DEVICES=0,1,...   CUDA devices of the implicit context: the job is sharded
 *                         over them inside the library (hrt_multi_*); "all" = every
 *                         device of the box.  Default: HRT_DEVICE or device 0
 *   HRT_DEVICE=<n>        single CUDA device (default 0)
 *   HRT_NO_RAYSINFO=1     do not fill raysInfo_scat (saves the largest copy)
 *   HRT_NO_SCENE_CACHE=1  re-upload the scene and rebuild the BVH on every call
 *   HRT_STRICT_UNTOUCHED=1  leave every word of the scatter gains, delays and directions that the
 *                         reference leaves untouched (slots of rays that left the scene, directions of
 *                         occluded slots) untouched as well, instead of writing 0 -- costs a second
 *                         host copy of those arrays ("Unwritten outputs")
 */
#include "../../include/hrt_cuda.h"

#include <stdio.h>

static hrt_multi *g_ctx = NULL;
static char g_ctx_devices[256] = "";
/* the scene the devices currently hold: content hash and its normals (reference
 * :208-224), so that repeated calls on the same scene neither re-upload it nor
 * rebuild the BVH (the reference's binding re-reads the file on every call,
 * compute_paths_pybind11.cpp:119) */
static uint64_t g_scene_hash = 0;
static Vec3 *g_scene_normals = NULL;
static size_t g_scene_tris = 0;

static void die(const char *what, const char *detail)
{
  fprintf(stderr, "hermespy_rt: %s: %s\n", what, detail ? detail : "");
  exit(8);
}

static void drop_ctx(void)
{
  if (g_ctx) { hrt_multi_destroy(g_ctx); g_ctx = NULL; }
  free(g_scene_normals); g_scene_normals = NULL; g_scene_hash = 0; g_scene_tris = 0;
}

static hrt_multi *implicit_ctx(void)
{
  char want[256];
  const char *ds = getenv("HRT_DEVICES"), *d1 = getenv("HRT_DEVICE");
  snprintf(want, sizeof want, "%s", ds && *ds ? ds : (d1 && *d1 ? d1 : "0"));
  if (g_ctx && !strcmp(want, g_ctx_devices)) return g_ctx;
  drop_ctx();
  int devs[64], n = 0;
  if (!strcmp(want, "all")) n = 0;
  else {
    const char *q = want;
    while (*q && n < 64) {
      char *end;
      const long v = strtol(q, &end, 10);
      if (end == q) die("bad HRT_DEVICES", want);
      devs[n++] = (int)v;
      q = end;
      while (*q == ',' || *q == ' ') ++q;
    }
    if (n == 0) die("bad HRT_DEVICES", want);
  }
  if (hrt_multi_create(n ? devs : NULL, n, &g_ctx) != HRT_OK) die("cannot create CUDA context", hrt_multi_last_error(NULL));
  snprintf(g_ctx_devices, sizeof g_ctx_devices, "%s", want);
  static int registered = 0;
  if (!registered) { atexit(drop_ctx); registered = 1; }
  return g_ctx;
}

/* change detection, not cryptography: four independent xor-rotate-multiply lanes over 32-byte pieces (a byte-wise
 * FNV chain runs at < 1 GB/s -- 30 ms per call on a million-triangle scene), FNV-1a for the tail */
static uint64_t fnv(const void *p, size_t n, uint64_t h)
{
  const unsigned char *b = (const unsigned char *)p;
  if (n >= 64) {
    const uint64_t K = 0x9E3779B97F4A7C15ull;
    uint64_t a0 = h ^ 0x243F6A8885A308D3ull, a1 = h + 0x13198A2E03707344ull, a2 = ~h, a3 = h * 0x100000001B3ull + 1u;
    while (n >= 32) {
      uint64_t w[4];
      memcpy(w, b, 32);
      a0 ^= w[0]; a0 = (a0 << 29 | a0 >> 35) * K;
      a1 ^= w[1]; a1 = (a1 << 31 | a1 >> 33) * K;
      a2 ^= w[2]; a2 = (a2 << 27 | a2 >> 37) * K;
      a3 ^= w[3]; a3 = (a3 << 33 | a3 >> 31) * K;
      b += 32; n -= 32;
    }
    h = (a0 ^ (a1 << 17 | a1 >> 47)) * K;
    h = (h ^ (a2 << 31 | a2 >> 33)) * K;
    h = (h ^ (a3 << 45 | a3 >> 19)) * K;
    h ^= h >> 29;
  }
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001B3ull; }
  return h;
}

/* content hash of everything the GPU side reads from the scene */
static uint64_t scene_hash(const Scene *scene)
{
  uint64_t h = fnv(&scene->num_meshes, 4, 0xCBF29CE484222325ull);
  for (uint32_t m = 0; m < scene->num_meshes; ++m) {
    const Mesh *me = &scene->meshes[m];
    h = fnv(&me->num_vertices, 4, h); h = fnv(&me->num_triangles, 4, h);
    h = fnv(me->vs, (size_t)me->num_vertices * sizeof(Vec3), h);
    h = fnv(me->is, (size_t)me->num_triangles * 12, h);
    h = fnv(&me->material_index, 4, h); h = fnv(&me->velocity, sizeof(Vec3), h);
  }
  return h ? h : 1;
}

/* scene -> GPUs (flatten, normals, BVH; skipped when they already hold this
 * scene) and the material table at f; fills Mesh.ns like the reference's
 * precompute_normals -- a previous array is freed first (the reference leaks it) */
static hrt_multi *prepare(Scene *scene, float carrier_frequency_GHz)
{
  hrt_multi *ctx = implicit_ctx();
  size_t total = 0;
  for (uint32_t m = 0; m < scene->num_meshes; ++m) total += scene->meshes[m].num_triangles;
  const uint64_t h = scene_hash(scene);
  if (h != g_scene_hash || total != g_scene_tris || !g_scene_normals || getenv("HRT_NO_SCENE_CACHE")) {
    free(g_scene_normals);
    g_scene_normals = (Vec3 *)malloc((total ? total : 1) * sizeof(Vec3));
    if (!g_scene_normals) die("compute_paths", "out of memory");
    g_scene_hash = 0;
    if (hrt_multi_scene_upload(ctx, scene, g_scene_normals) != HRT_OK) die("scene upload failed", hrt_multi_last_error(ctx));
    g_scene_hash = h; g_scene_tris = total;
  }
  size_t off = 0;
  for (uint32_t m = 0; m < scene->num_meshes; ++m) {
    Mesh *me = &scene->meshes[m];
    free(me->ns);
    me->ns = (Vec3 *)malloc((me->num_triangles ? me->num_triangles : 1) * sizeof(Vec3));
    if (!me->ns) die("compute_paths", "out of memory");
    memcpy(me->ns, g_scene_normals + off, (size_t)me->num_triangles * sizeof(Vec3));
    off += me->num_triangles;
  }

  /* material constants at this frequency: only the materials the scene uses
   * are (re)computed, the others stay zero (reference :176-180) */
  HrtMaterialDerived table[NUM_G_MATERIALS];
  memset(table, 0, sizeof table);
  for (uint32_t m = 0; m < scene->num_meshes; ++m) {
    uint32_t mi = scene->meshes[m].material_index;
    hrt_materials_derive(mi, carrier_frequency_GHz, &table[mi]);
  }
  if (hrt_multi_materials_set(ctx, table) != HRT_OK) die("material upload failed", hrt_multi_last_error(ctx));
  return ctx;
}

static void run_dense(
    Scene *scene, Vec3 *rx_pos, Vec3 *tx_pos, Vec3 *rx_vel, Vec3 *tx_vel,
    float carrier_frequency_GHz,
    size_t num_rx, size_t num_tx, size_t num_rays, size_t num_bounces,
    ChannelInfo *chanInfo_los, RaysInfo *raysInfo_los,
    ChannelInfo *chanInfo_scat, RaysInfo *raysInfo_scat, float *a_te_c64, float *a_tm_c64)
{
  if (!scene || !chanInfo_scat) die("compute_paths", "NULL scene or scatter output");
  hrt_multi *ctx = prepare(scene, carrier_frequency_GHz);

  HrtRunParams p;
  memset(&p, 0, sizeof p);
  p.num_rx = num_rx; p.num_tx = num_tx; p.num_paths = num_rays; p.num_bounces = num_bounces;
  p.carrier_frequency_GHz = carrier_frequency_GHz;
  p.rx_pos = rx_pos; p.tx_pos = tx_pos; p.rx_vel = rx_vel; p.tx_vel = tx_vel;
  p.flags = HRT_FLAG_DENSE;
  if (raysInfo_scat && !getenv("HRT_NO_RAYSINFO")) p.flags |= HRT_FLAG_RAYSINFO;
  p.los = chanInfo_los; p.rays_los = raysInfo_los;
  p.scat = chanInfo_scat; p.rays_scat = raysInfo_scat;
  if (a_te_c64) { p.flags |= HRT_FLAG_DENSE_C64; p.scat_a_te_c64 = a_te_c64; p.scat_a_tm_c64 = a_tm_c64; }
  if (!getenv("HRT_STRICT_UNTOUCHED")) {
    if (hrt_multi_run(ctx, &p) != HRT_OK) die("compute_paths failed", hrt_multi_last_error(ctx));
    return;
  }
  /* Strict mode: the run fills temporaries, then only the words the reference writes are merged into
   * the caller's arrays -- gains and delay where the ray was alive (:685-709), the arrival direction
   * where the receiver was not occluded (:707).  freq_shift is written in full by the reference as well
   * (its memcpy chain, :494-508) and RaysInfo rows are copied whole (:732-743): those go straight out. */
  const size_t n = num_rx * num_tx * num_bounces * num_rays;
  float *tmp = (float *)malloc((n ? n : 1) * 8 * sizeof(float));     /* 4 gains (or 2 x complex), tau, 3 direction words */
  uint8_t *state = (uint8_t *)calloc(n ? n : 1, 1);
  if (!tmp || !state) die("compute_paths", "out of memory (strict mode)");
  ChannelInfo t = *chanInfo_scat;
  t.a_te_re = tmp; t.a_te_im = tmp + n; t.a_tm_re = tmp + 2 * n; t.a_tm_im = tmp + 3 * n;
  t.tau = tmp + 4 * n; t.directions_rx = (Vec3 *)(tmp + 5 * n);
  p.scat = &t;
  if (a_te_c64) { p.scat_a_te_c64 = tmp; p.scat_a_tm_c64 = tmp + 2 * n; }
  p.flags |= HRT_FLAG_TRACE;
  p.trace_slot_state = state;
  if (hrt_multi_run(ctx, &p) != HRT_OK) die("compute_paths failed", hrt_multi_last_error(ctx));
  for (size_t i = 0; i < n; ++i) {
    if (!state[i]) continue;                                   /* ray no longer alive: nothing written */
    if (a_te_c64) {
      a_te_c64[2 * i] = tmp[2 * i]; a_te_c64[2 * i + 1] = tmp[2 * i + 1];
      a_tm_c64[2 * i] = tmp[2 * n + 2 * i]; a_tm_c64[2 * i + 1] = tmp[2 * n + 2 * i + 1];
    } else {
      chanInfo_scat->a_te_re[i] = t.a_te_re[i]; chanInfo_scat->a_te_im[i] = t.a_te_im[i];
      chanInfo_scat->a_tm_re[i] = t.a_tm_re[i]; chanInfo_scat->a_tm_im[i] = t.a_tm_im[i];
    }
    chanInfo_scat->tau[i] = t.tau[i];
    if (state[i] == 1) chanInfo_scat->directions_rx[i] = t.directions_rx[i];   /* 2: occluded, direction untouched */
  }
  free(tmp); free(state);
}

void compute_paths(
    Scene *scene, Vec3 *rx_pos, Vec3 *tx_pos, Vec3 *rx_vel, Vec3 *tx_vel,
    float carrier_frequency_GHz,
    size_t num_rx, size_t num_tx, size_t num_rays, size_t num_bounces,
    ChannelInfo *chanInfo_los, RaysInfo *raysInfo_los,
    ChannelInfo *chanInfo_scat, RaysInfo *raysInfo_scat)
{
  run_dense(scene, rx_pos, tx_pos, rx_vel, tx_vel, carrier_frequency_GHz, num_rx, num_tx, num_rays, num_bounces,
            chanInfo_los, raysInfo_los, chanInfo_scat, raysInfo_scat, NULL, NULL);
}

/* compute_paths() with the scatter gains as interleaved complex64 (include/hrt_cuda.h) */
void compute_paths_c64(
    Scene *scene, Vec3 *rx_pos, Vec3 *tx_pos, Vec3 *rx_vel, Vec3 *tx_vel,
    float carrier_frequency_GHz,
    size_t num_rx, size_t num_tx, size_t num_rays, size_t num_bounces,
    ChannelInfo *chanInfo_los, RaysInfo *raysInfo_los,
    ChannelInfo *chanInfo_scat, RaysInfo *raysInfo_scat,
    float *a_te_c64, float *a_tm_c64)
{
  if (!a_te_c64 || !a_tm_c64) die("compute_paths_c64", "NULL complex output");
  run_dense(scene, rx_pos, tx_pos, rx_vel, tx_vel, carrier_frequency_GHz, num_rx, num_tx, num_rays, num_bounces,
            chanInfo_los, raysInfo_los, chanInfo_scat, raysInfo_scat, a_te_c64, a_tm_c64);
}

/* Streaming consumer of the same path set: instead of
 * one record per path, the channel impulse response per (rx, tx) -- what a
 * caller of compute_paths() forms next from ChannelInfo -- accumulated on the
 * GPU.  cir[((rx * num_tx + tx) * num_bins + bin) * 4 + {te_re, te_im, tm_re,
 * tm_im}], bin = floor((tau - tau0_s) / dt_s); the array is overwritten.
 * Returns the number of paths whose delay fell outside the window. */
size_t compute_cir(
    Scene *scene, Vec3 *rx_pos, Vec3 *tx_pos, Vec3 *rx_vel, Vec3 *tx_vel,
    float carrier_frequency_GHz,
    size_t num_rx, size_t num_tx, size_t num_rays, size_t num_bounces,
    float tau0_s, float dt_s, size_t num_bins, float *cir)
{
  if (!scene || !cir || !num_bins) die("compute_cir", "NULL scene or output");
  hrt_multi *ctx = prepare(scene, carrier_frequency_GHz);
  const size_t nl = num_rx * num_tx;
  memset(cir, 0, nl * num_bins * 4 * sizeof(float));

  /* LoS through the same run: small host arrays in the reference's layout */
  ChannelInfo los;
  memset(&los, 0, sizeof los);
  float *buf = (float *)calloc(nl * 12, sizeof(float));
  if (!buf) die("compute_cir", "out of memory");
  los.num_rays = 1;
  los.directions_rx = (Vec3 *)buf; los.directions_tx = (Vec3 *)(buf + 3 * nl);
  los.a_te_re = buf + 6 * nl; los.a_te_im = buf + 7 * nl; los.a_tm_re = buf + 8 * nl; los.a_tm_im = buf + 9 * nl;
  los.tau = buf + 10 * nl; los.freq_shift = buf + 11 * nl;

  HrtRunParams p;
  memset(&p, 0, sizeof p);
  p.num_rx = num_rx; p.num_tx = num_tx; p.num_paths = num_rays; p.num_bounces = num_bounces;
  p.carrier_frequency_GHz = carrier_frequency_GHz;
  p.rx_pos = rx_pos; p.tx_pos = tx_pos; p.rx_vel = rx_vel; p.tx_vel = tx_vel;
  p.flags = HRT_FLAG_CIR;
  p.los = &los;
  p.cir = cir; p.cir_tau0_s = tau0_s; p.cir_dt_s = dt_s; p.cir_bins = (uint32_t)num_bins;
  if (hrt_multi_run(ctx, &p) != HRT_OK) die("compute_cir failed", hrt_multi_last_error(ctx));
  free(buf);
  HrtRunStats st;
  hrt_multi_get_stats(ctx, &st);
  return (size_t)st.cir_dropped;
}

/* The valid scatter paths of the same path set as compact records (HrtPathRecord,
 * include/hrt_cuda.h) instead of dense arrays: what the reference writes into
 * the slots of living rays and unoccluded receivers, nothing for the others.
 * Stores at most `capacity` records into `paths` (arbitrary order) and returns
 * the number of valid paths found (> capacity: a subset was stored). */
size_t compute_path_list(
    Scene *scene, Vec3 *rx_pos, Vec3 *tx_pos, Vec3 *rx_vel, Vec3 *tx_vel,
    float carrier_frequency_GHz,
    size_t num_rx, size_t num_tx, size_t num_rays, size_t num_bounces,
    HrtPathRecord *paths, size_t capacity)
{
  if (!scene || !paths || !capacity) die("compute_path_list", "NULL scene or output");
  hrt_multi *ctx = prepare(scene, carrier_frequency_GHz);
  HrtRunParams p;
  memset(&p, 0, sizeof p);
  p.num_rx = num_rx; p.num_tx = num_tx; p.num_paths = num_rays; p.num_bounces = num_bounces;
  p.carrier_frequency_GHz = carrier_frequency_GHz;
  p.rx_pos = rx_pos; p.tx_pos = tx_pos; p.rx_vel = rx_vel; p.tx_vel = tx_vel;
  p.flags = HRT_FLAG_PATHLIST;
  uint64_t found = 0;
  p.paths = paths; p.paths_capacity = capacity; p.paths_count = &found;
  if (hrt_multi_run(ctx, &p) != HRT_OK) die("compute_path_list failed", hrt_multi_last_error(ctx));
  return (size_t)found;
}
