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
 *     from the GPU-computed normals.  Like the reference this allocates on
 *     every call and leaves freeing to free_scene().
 * Errors follow the reference convention: message on stderr and exit(8)
 * (inc/common.h:20-25).  There is no CPU fallback.
 *
 * Environment:
 *   HRT_DEVICE=<n>        CUDA device of the implicit context (default 0)
 *   HRT_NO_RAYSINFO=1     do not fill raysInfo_scat (saves the largest copy)
 */
#include "../../include/hrt_cuda.h"

#include <stdio.h>

static hrt_ctx *g_ctx = NULL;
static int g_ctx_device = -1;

static void die(const char *what, const char *detail)
{
  fprintf(stderr, "hermespy_rt: %s: %s\n", what, detail ? detail : "");
  exit(8);
}

static void drop_ctx(void) { if (g_ctx) { hrt_ctx_destroy(g_ctx); g_ctx = NULL; } }

static hrt_ctx *implicit_ctx(void)
{
  int dev = 0;
  const char *s = getenv("HRT_DEVICE");
  if (s) dev = atoi(s);
  if (g_ctx && g_ctx_device == dev) return g_ctx;
  drop_ctx();
  if (hrt_ctx_create(dev, &g_ctx) != HRT_OK) die("cannot create CUDA context", hrt_last_error(NULL));
  g_ctx_device = dev;
  static int registered = 0;
  if (!registered) { atexit(drop_ctx); registered = 1; }
  return g_ctx;
}

/* scene -> GPU (flatten, normals, BVH) and the material table at f; fills
 * Mesh.ns like the reference's precompute_normals */
static hrt_ctx *prepare(Scene *scene, float carrier_frequency_GHz)
{
  hrt_ctx *ctx = implicit_ctx();
  size_t total = 0;
  for (uint32_t m = 0; m < scene->num_meshes; ++m) total += scene->meshes[m].num_triangles;
  Vec3 *normals = (Vec3 *)malloc((total ? total : 1) * sizeof(Vec3));
  if (!normals) die("compute_paths", "out of memory");
  if (hrt_scene_upload(ctx, scene, normals) != HRT_OK) die("scene upload failed", hrt_last_error(ctx));
  size_t off = 0;
  for (uint32_t m = 0; m < scene->num_meshes; ++m) {
    Mesh *me = &scene->meshes[m];
    me->ns = (Vec3 *)malloc((me->num_triangles ? me->num_triangles : 1) * sizeof(Vec3));
    if (!me->ns) die("compute_paths", "out of memory");
    memcpy(me->ns, normals + off, (size_t)me->num_triangles * sizeof(Vec3));
    off += me->num_triangles;
  }
  free(normals);

  /* material constants at this frequency: only the materials the scene uses
   * are (re)computed, the others stay zero (reference :176-180) */
  HrtMaterialDerived table[NUM_G_MATERIALS];
  memset(table, 0, sizeof table);
  for (uint32_t m = 0; m < scene->num_meshes; ++m) {
    uint32_t mi = scene->meshes[m].material_index;
    hrt_materials_derive(mi, carrier_frequency_GHz, &table[mi]);
  }
  if (hrt_materials_set(ctx, table) != HRT_OK) die("material upload failed", hrt_last_error(ctx));
  return ctx;
}

void compute_paths(
    Scene *scene, Vec3 *rx_pos, Vec3 *tx_pos, Vec3 *rx_vel, Vec3 *tx_vel,
    float carrier_frequency_GHz,
    size_t num_rx, size_t num_tx, size_t num_rays, size_t num_bounces,
    ChannelInfo *chanInfo_los, RaysInfo *raysInfo_los,
    ChannelInfo *chanInfo_scat, RaysInfo *raysInfo_scat)
{
  if (!scene || !chanInfo_scat) die("compute_paths", "NULL scene or scatter output");
  hrt_ctx *ctx = prepare(scene, carrier_frequency_GHz);

  HrtRunParams p;
  memset(&p, 0, sizeof p);
  p.num_rx = num_rx; p.num_tx = num_tx; p.num_paths = num_rays; p.num_bounces = num_bounces;
  p.carrier_frequency_GHz = carrier_frequency_GHz;
  p.rx_pos = rx_pos; p.tx_pos = tx_pos; p.rx_vel = rx_vel; p.tx_vel = tx_vel;
  p.shard_world = 1;
  p.flags = HRT_FLAG_DENSE;
  if (raysInfo_scat && !getenv("HRT_NO_RAYSINFO")) p.flags |= HRT_FLAG_RAYSINFO;
  p.los = chanInfo_los; p.rays_los = raysInfo_los;
  p.scat = chanInfo_scat; p.rays_scat = raysInfo_scat;
  if (hrt_run(ctx, &p) != HRT_OK) die("compute_paths failed", hrt_last_error(ctx));
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
  hrt_ctx *ctx = prepare(scene, carrier_frequency_GHz);
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
  p.shard_world = 1;
  p.flags = HRT_FLAG_CIR;
  p.los = &los;
  p.cir = cir; p.cir_tau0_s = tau0_s; p.cir_dt_s = dt_s; p.cir_bins = (uint32_t)num_bins;
  if (hrt_run(ctx, &p) != HRT_OK) die("compute_cir failed", hrt_last_error(ctx));
  free(buf);
  HrtRunStats st;
  hrt_get_stats(ctx, &st);
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
  hrt_ctx *ctx = prepare(scene, carrier_frequency_GHz);
  HrtRunParams p;
  memset(&p, 0, sizeof p);
  p.num_rx = num_rx; p.num_tx = num_tx; p.num_paths = num_rays; p.num_bounces = num_bounces;
  p.carrier_frequency_GHz = carrier_frequency_GHz;
  p.rx_pos = rx_pos; p.tx_pos = tx_pos; p.rx_vel = rx_vel; p.tx_vel = tx_vel;
  p.shard_world = 1;
  p.flags = HRT_FLAG_PATHLIST;
  uint64_t found = 0;
  p.paths = paths; p.paths_capacity = capacity; p.paths_count = &found;
  if (hrt_run(ctx, &p) != HRT_OK) die("compute_path_list failed", hrt_last_error(ctx));
  return (size_t)found;
}
