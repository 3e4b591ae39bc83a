/* hrt_cuda.cu -- CUDA (sm_100a) implementation behind include/hrt_cuda.h.
 *
 * Kernels (DESIGN.md has the data layout and the roofline of each):
 *   (kernels: hrt_build_kernels.cuh, hrt_run_kernels.cuh)
 *   k_tri_setup / k_morton / k_gather / k_karras / k_refit / k_emit
 *        scene -> SoA triangle records + LBVH, replaces the reference's scene
 *        walk (src/compute_paths.c:208-224, :253-258)
 *   k_raygen     Fibonacci launch directions            (reference :443-451)
 *   k_init       per-ray state + output initialisation  (reference :460-508)
 *   k_los        line-of-sight pairs                    (reference :514-577)
 *   k_bounce     one wavefront step per bounce depth: closest hit, Fresnel,
 *                gain/delay update, reflection, ballot/prefix-sum compaction
 *                of the survivors                       (reference :599-664)
 *   k_scatter    per-(hit, rx) shadow query + scattering coefficients +
 *                outputs / reductions                   (reference :670-723)
 *
 * The arithmetic lives in hrt_core.cuh; this file is launch logic, memory
 * management and the C ABI.  Nothing here computes paths on the CPU.
 */
#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/hrt_cuda.h"
#include "hrt_bvh.cuh"
#include "hrt_rxmap.cuh"
#include "hrt_ext.cuh"

/* host C helpers (host_math.c): glibc double trig, as the reference uses */
extern "C" void hrt_host_launch_dir(uint64_t path, uint64_t num_paths, float out[3]);

static_assert(sizeof(HrtMaterialDerived) == sizeof(HrtMaterial), "material ABI");
static_assert(sizeof(Ray) == 24 && sizeof(Vec3) == 12, "reference ABI");
static_assert(sizeof(HrtPairSummary) == 48 && sizeof(HrtBounceSummary) == 32, "summary ABI");
static_assert(sizeof(HrtPathRecord) == 48, "path record ABI");
static_assert(sizeof(HrtRefractRecord) == 48, "refraction record ABI");

/* raw scattering parameters of g_materials for the opt-in extensions (hrt_ext.cuh) */
__constant__ HrtExtTable c_ext;

#ifndef HRT_BLOCK
#define HRT_BLOCK 512   /* 2 blocks of 512 threads per SM: 64 registers, ~85 KB shared memory each */
#endif
#ifndef HRT_MAP_BLOCK
#define HRT_MAP_BLOCK 576   /* k_scatter with receiver maps: 2 blocks of 576 threads, 56 registers, 36 warps per SM.  Sweep (k_scatter ms on C4):
                               512 / 576 / 640 / 672 threads x 2, 1024 x 1 -> 237.6 / 231.4 / 232.2 / 237.6 / 235.4.  The tree-walking kernels
                               stay at 512 (576: C5 663 -> 681 ms, C4 through the BVH 436 -> 444 ms: spills) */
#endif
#ifndef HRT_GLOBAL_BLOCK
#define HRT_GLOBAL_BLOCK 448   /* k_scatter walking a scene in global memory (C5): 2 blocks of 448 threads, 72 registers: 671 -> 654 ms
                                  (512 x 2 at 64 registers spills more; 384 x 3 at 56: 696 ms) */
#endif
#ifndef HRT_MAP_MIN_BLOCKS
#define HRT_MAP_MIN_BLOCKS 2   /* (same 36 warps in smaller blocks, 384 x 3 / 288 x 4 / 192 x 6: 232.0 / 232.9 / 232.7 ms against 231.5) */
#endif
#ifndef HRT_MIN_BLOCKS
#define HRT_MIN_BLOCKS 2   /* => 64 registers (1024 threads/SM): best of the sweep in profiles/r1_sweeps.md; __launch_bounds__ min blocks/SM of the two traversal kernels */
#endif
#define HRT_SMEM_SCENE_LIMIT (100 * 1024)  /* stage 8 octant node copies + triangles in shared memory below this */
#define HRT_AMB_CAP 65536

#define HRT_SORT_STREAMS 4
static char g_create_error[256] = "";

struct SceneDev {
  const float4 *nodes;
  const float4 *tris;
  const uint32_t *tri_gid;
  const uint32_t *mesh_of;     /* by triangle id */
  const uint32_t *mesh_mat;    /* by mesh */
  const float *mesh_vel;       /* 3 per mesh */
  uint32_t num_tris, num_nodes;
  uint32_t octants;            /* copies of the binary nodes (1: they only feed the collapse and the refit) */
  int root_ref;
  const float4 *wnodes;        /* 4-wide nodes the kernels traverse (hrt_bvh.cuh), wide_octants copies */
  uint32_t num_wide, wide_octants;   /* copies: 8 (one per direction octant) or 1 */
  uint32_t wstride;                  /* float4s per node slot: 7 (scenes staged in shared memory) or 8 (128-byte aligned) */
  int wroot;
  float key_lo[3], key_scale[3];   /* vertex bounds -> 10-bit grid of the hit-order keys */
  float key_log;                   /* > 0: logarithmic x/y grid around the TX, cells per octave */
};

/* receiver maps (hrt_rxmap.cuh).  Cell word: (offset of the cell's list inside the
 * receiver's item range << 8) | length. */
struct RxMapDev {
  const uint32_t *cells;       /* [R][6 G G] */
  const uint32_t *items;       /* [R][items_per_rx]: slot | q_hi << 16 | q_lo << 24 (depth bounds, hrt_rxmap.cuh) */
  const float *inv_step;       /* [R] depth quantisation of each receiver's items */
  uint32_t G, items_per_rx;
};

struct RunDev {
  uint32_t R, T, B;
  uint32_t n;            /* paths per TX in this chunk      */
  uint32_t n_alloc;      /* row pitch of all per-path arrays */
  uint64_t P;            /* num_paths of the whole job       */
  uint64_t l0;           /* first shard-local path index of the chunk */
  uint32_t rank, world;
  uint64_t blk;
  HrtRunConst k;
  const float *rx_pos, *tx_pos, *rx_vel, *tx_vel;
  float *dirs;           /* [n_alloc][3] launch directions of the chunk */
  /* Ray records, 64 B = 4 x float4 each, [T][n_alloc], double buffered by depth:
   *   (o.xyz, d.x) (d.yz, te_r, te_i) (tm_r, tm_i, tau, theta) (leaf slot, path, -, -)
   * k_bounce(depth) reads rec[depth & 1] through the queue and appends the
   * survivors -- reflected ray, updated gains/delay, incidence angle, hit slot --
   * to rec[(depth + 1) & 1] in queue order: one coalesced 64-byte write per
   * survivor, one 64-byte gather per reader. */
  float4 *rec[2];
  Ray *rays;             /* [B+1][T][n_alloc] RaysInfo rows (HRT_FLAG_RAYSINFO only) */
  uint8_t *dead_at;      /* [T][n_alloc] bounce at which the ray left the scene, 255 alive (RAYSINFO only) */
  uint32_t *queue[2];    /* [T][n_alloc] record indices, in work order */
  uint32_t *queue_alt;   /* [T][n_alloc] sort output, swapped with queue[k] on the host */
  uint32_t *qkey, *qkey_alt; /* [T][n_alloc] Morton code of the hit point, order of the next queue */
  uint32_t *qcount;      /* [B+1][T] queue sizes, then [B][T] k_bounce and [B][T] k_scatter work cursors */
  float *out_f[6];       /* [R][T][B][n_alloc] te_re te_im tm_re tm_im tau freq */
  uint32_t gain_stride;  /* 1; 2 with HRT_FLAG_DENSE_C64: te_im = te_re + 1 (interleaved complex), likewise tm */
  float *out_dir;        /* [R][T][B][n_alloc][3] */
  uint32_t *tr_hit;      /* [T][B][n_alloc] */
  float *tr_t;
  uint8_t *tr_state;     /* [R][T][B][n_alloc] */
  HrtPairSummary *pair;  /* [R][T][B] */
  HrtBounceSummary *bounce; /* [T][B] */
  uint32_t *amb_list, *amb_count;
  uint32_t *dkey, *dkey2, *perm, *perm2;  /* direction sort of the chunk's paths */
  unsigned long long *counters;  /* [16] instrumented build; [15] = CIR paths outside the window */
  float4 *plist;         /* [plist_cap][3] HrtPathRecord, HRT_FLAG_PATHLIST */
  unsigned long long plist_cap, *plist_count;
  float *cir;            /* [R][T][cir_bins][4] HRT_FLAG_CIR */
  float cir_tau0, cir_inv_dt;
  uint32_t cir_bins;
  uint32_t flags;
  RxMapDev map;          /* k_scatter<..., MAP>: receiver maps of this run's receivers */
  float4 *refr;          /* [refr_cap][3] HrtRefractRecord, HRT_FLAG_EXT_REFRACT (count: counters[13]) */
  unsigned long long refr_cap;
};

/* global path index of shard-local index L (blocks dealt round-robin) */
__host__ __device__ static inline uint64_t hrt_gpath(uint64_t L, uint32_t rank, uint32_t world, uint64_t blk)
{
  if (world <= 1) return L;
  const uint64_t q = L / blk, r = L - q * blk;
  return (q * world + rank) * blk + r;
}

/* ------------------------------------------------------------------ context */

/* Host side of the dense device -> host path.  cudaMemcpy into pageable memory
 * runs at 21 GB/s on this box (2.3 GB/s when the destination pages have never
 * been touched: one thread takes every page fault); PCIe delivers 57 GB/s into
 * pinned memory and eight host threads copy at 77 GB/s (33 GB/s into fresh
 * pages) -- scripts/d2h_bandwidth.py.  So large outputs go through two pinned
 * staging buffers and a small pool of copy threads, double buffered. */
struct HostPool {
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv, cv_done;
  std::function<void(int, int)> fn;
  int gen = 0, pending = 0, n = 0;
  bool stop = false;
  void start(int k)
  {
    n = k;
    for (int w = 0; w < k; ++w)
      th.emplace_back([this, w] {
        int seen = 0;
        for (;;) {
          std::function<void(int, int)> f;
          {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return stop || gen != seen; });
            if (stop) return;
            seen = gen; f = fn;
          }
          f(w, n);
          { std::lock_guard<std::mutex> lk(mu); if (--pending == 0) cv_done.notify_all(); }
        }
      });
  }
  void run(const std::function<void(int, int)> &f)
  {
    std::unique_lock<std::mutex> lk(mu);
    fn = f; pending = n; ++gen;
    cv.notify_all();
    cv_done.wait(lk, [&] { return pending == 0; });
  }
  ~HostPool()
  {
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    cv.notify_all();
    for (auto &t : th) t.join();
  }
};

struct CopyTile { const char *dev; size_t dpitch; char *host; size_t hpitch, width, rows; };
#define HRT_STAGE_BYTES ((size_t)128 << 20)
#define HRT_STAGE_MIN_TOTAL ((size_t)16 << 20)   /* smaller outputs: plain cudaMemcpy2DAsync */

struct hrt_ctx {
  int device;
  cudaStream_t stream;
  cudaEvent_t ev[8];
  cudaEvent_t *evpool; size_t evpool_n;   /* per-launch timing events, grown on demand */
  char err[512];
  int leaf_max;
  float pad_ulps;

  /* scene */
  bool have_scene;
  uint32_t num_tris, num_meshes, num_nodes, octants;
  int root_ref;
  float pad, scene_max_abs;
  float scene_lo[3], scene_hi[3];
  float4 *d_tris; uint32_t *d_tri_gid; uint32_t *d_mesh_of; uint32_t *d_mesh_mat; float *d_mesh_vel;
  float4 *d_nodes;
  /* 4-wide nodes collapsed from d_nodes (build_wide) and the collapse scratch */
  float4 *d_wnodes; size_t cap_wnodes;
  int *d_wparent; uint8_t *d_weven; uint32_t *d_widx, *d_wtotal; size_t cap_wscratch;
  uint32_t num_wide, wide_octants, wstride; int wroot;
  /* builder arrays kept for re-padding */
  int *d_kl, *d_kr, *d_kfirst, *d_klast, *d_newidx;
  float *d_box;            /* [(2n-1)][6] lo.xyz hi.xyz; inner nodes then leaves */
  bool sah;                /* tree built by the binned-SAH builder (raw arrays below) */
  uint32_t cap_raw, cap_nodes;   /* capacities (triangles) of d_raw_* and d_nodes: rebuilds reuse them */
  int2 *d_raw_ref;         /* [num_nodes] child refs */
  float *d_raw_box;        /* [num_nodes][12] both children's boxes, unpadded */
  float build_ms; int build_levels;
  /* geometry kept for hrt_scene_advance: vertices, rebased indices, vertex ->
   * mesh, triangle records and boxes in (mesh, face) order, first node of
   * every build level (children always live on the next level) */
  float *d_verts; uint32_t *d_idx3; uint32_t *d_vmesh; float4 *d_recs; float *d_tboxes; unsigned *d_bounds;
  size_t num_verts;
  int level_first[256];
  float max_speed;

  /* receiver maps (hrt_rxmap.cuh) of the last run's receivers, reused while the
   * receivers, the scene and the padding stay the same */
  uint32_t *d_map_cells; uint32_t *d_map_items; uint32_t *d_map_cursor;   /* cursor[R], then status[1], then inv_step[R] (float) */
  size_t cap_map_cells, cap_map_items, cap_map_cursor;
  uint32_t map_G, map_items_per_rx, map_R; bool map_valid;
  uint64_t map_key, map_failed_key, scene_version; uint64_t map_key0;   /* receivers + scene version, without G */
  float map_build_ms;

  bool have_mats;
  HrtMaterialTable mats;

  /* run buffers */
  size_t cap_n, cap_T, cap_R, cap_B; uint32_t cap_flags;
  RunDev rd;
  float *d_pos;            /* rx_pos, tx_pos, rx_vel, tx_vel */
  size_t cap_pos;
  void *sort_tmp; size_t sort_tmp_bytes;
  /* side streams for the per-transmitter hit sorts of one depth (small sorts overlap), each with its
   * own CUB scratch of sort_side_bytes */
  cudaStream_t sort_stream[HRT_SORT_STREAMS]; void *sort_side_tmp[HRT_SORT_STREAMS]; size_t sort_side_bytes;
  cudaEvent_t sort_fork, sort_join[HRT_SORT_STREAMS];
  /* k_scatter of depth b runs on scat_stream[b & 1]: its ragged end overlaps the next depth's k_bounce, sort and
   * the start of the next k_scatter.  scat_ready: hits of a depth are in order; scat_done[k]: last k_scatter on stream k */
  cudaStream_t scat_stream[2]; cudaEvent_t scat_ready, scat_done[2];
  void *d_los;             /* HrtLosOut[R*T] */
  size_t cap_los;
  float *d_cir; size_t cap_cir;
  float4 *d_refr; size_t cap_refr;
  float4 *d_plist; size_t cap_plist;
  HostPool *pool; char *stage[2]; cudaEvent_t stage_ev[2];
  float *h_patch; float *d_patch;     /* pinned / device staging of host-recomputed launch directions */
  uint8_t tail_dead[8];               /* RaysInfo: bounce at which paths 0..7 of TX 1 died in the last run (appendix A-9) */
  HrtRunStats stats;
};

static int fail(hrt_ctx *c, int code, const char *fmt, ...)
{
  va_list ap; va_start(ap, fmt);
  if (c) vsnprintf(c->err, sizeof c->err, fmt, ap);
  else   vsnprintf(g_create_error, sizeof g_create_error, fmt, ap);
  va_end(ap);
  return code;
}

#define CK(call)                                                                   \
  do { cudaError_t e_ = (call);                                                    \
       if (e_ != cudaSuccess)                                                      \
         return fail(ctx, HRT_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,   \
                     cudaGetErrorString(e_)); } while (0)

template <class T> static cudaError_t dev_alloc(T **p, size_t count)
{ *p = nullptr; return cudaMalloc((void **)p, (count ? count : 1) * sizeof(T)); }
template <class T> static void dev_free(T *&p) { if (p) cudaFree((void *)p); p = nullptr; }

#include "hrt_build_kernels.cuh"
#include "hrt_run_kernels.cuh"

/* ------------------------------------------------------------------- API */

extern "C" int hrt_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" const char *hrt_last_error(const hrt_ctx *ctx) { return ctx ? ctx->err : g_create_error; }

extern "C" int hrt_ctx_create(int device, hrt_ctx **out)
{
  hrt_ctx *ctx = nullptr;
  if (!out) return HRT_E_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(nullptr, HRT_E_NO_DEVICE, "no CUDA device available (%s); this library has no CPU path",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0 || device >= n) return fail(nullptr, HRT_E_ARG, "device %d out of range (0..%d)", device, n - 1);
  hrt_ctx *c = (hrt_ctx *)calloc(1, sizeof(hrt_ctx));
  if (!c) return fail(nullptr, HRT_E_NOMEM, "out of host memory");
  c->device = device;
  c->leaf_max = 2;   /* measured best on B200 (profiles/r1_v3_sweeps.txt) */
  c->pad_ulps = 64.f;
  if (const char *s = getenv("HRT_LEAF_MAX")) { int v = atoi(s); if (v >= 1 && v <= HRT_LEAF_MAX_CAP) c->leaf_max = v; }
  if (const char *s = getenv("HRT_BVH_PAD_ULPS")) { float v = strtof(s, nullptr); if (v >= 1.f) c->pad_ulps = v; }
  if ((e = cudaSetDevice(device)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    fail(nullptr, HRT_E_CUDA, "cannot initialise device %d: %s", device, cudaGetErrorString(e));
    free(c); return HRT_E_CUDA;
  }
  for (int i = 0; i < 8; ++i) cudaEventCreate(&c->ev[i]);
  {
    /* the builder's scratch comes from the device's default memory pool (build_sah): keep up to
     * 1 GB of it cached between rebuilds instead of returning it at every synchronisation */
    cudaMemPool_t mp;
    if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) {
      uint64_t keep = 1ull << 30;
      cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
  }
  *out = c;
  (void)ctx;
  return HRT_OK;
}

static void free_scene_dev(hrt_ctx *c)
{
  dev_free(c->d_tris); dev_free(c->d_tri_gid); dev_free(c->d_mesh_of); dev_free(c->d_mesh_mat);
  dev_free(c->d_mesh_vel); dev_free(c->d_nodes); dev_free(c->d_kl); dev_free(c->d_kr);
  dev_free(c->d_kfirst); dev_free(c->d_klast); dev_free(c->d_newidx); dev_free(c->d_box);
  dev_free(c->d_raw_ref); dev_free(c->d_raw_box); c->cap_raw = c->cap_nodes = 0;
  dev_free(c->d_verts); dev_free(c->d_idx3); dev_free(c->d_vmesh); dev_free(c->d_recs); dev_free(c->d_tboxes); dev_free(c->d_bounds);
  dev_free(c->d_wnodes); dev_free(c->d_wparent); dev_free(c->d_weven); dev_free(c->d_widx); dev_free(c->d_wtotal);
  c->cap_wnodes = c->cap_wscratch = 0; c->num_wide = 0;
  c->have_scene = false;
}

static void free_run_dev(hrt_ctx *c)
{
  RunDev &r = c->rd;
  dev_free(r.dirs); dev_free(r.rays); dev_free(r.rec[0]); dev_free(r.rec[1]);
  dev_free(r.dead_at); dev_free(r.queue[0]); dev_free(r.queue[1]);
  dev_free(r.queue_alt); dev_free(r.qkey); dev_free(r.qkey_alt);
  dev_free(r.qcount);
  if (r.gain_stride == 2) { r.out_f[1] = nullptr; r.out_f[3] = nullptr; }     /* aliases of out_f[0] / out_f[2] */
  for (int k = 0; k < 6; ++k) dev_free(r.out_f[k]);
  r.gain_stride = 1;
  dev_free(r.out_dir); dev_free(r.tr_hit); dev_free(r.tr_t); dev_free(r.tr_state);
  dev_free(r.pair); dev_free(r.bounce); dev_free(r.amb_list); dev_free(r.amb_count);
  dev_free(r.dkey); dev_free(r.dkey2); dev_free(r.perm); dev_free(r.perm2);
  dev_free(r.counters);
  c->cap_n = 0;
}

extern "C" void hrt_ctx_destroy(hrt_ctx *c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  free_scene_dev(c); free_run_dev(c);
  dev_free(c->d_pos);
  dev_free(c->d_map_cells); dev_free(c->d_map_items); dev_free(c->d_map_cursor);
  dev_free(c->d_patch);
  if (c->h_patch) { cudaFreeHost(c->h_patch); c->h_patch = nullptr; }
  if (c->d_los) { cudaFree(c->d_los); c->d_los = nullptr; }
  if (c->d_cir) { cudaFree(c->d_cir); c->d_cir = nullptr; }
  if (c->d_plist) { cudaFree(c->d_plist); c->d_plist = nullptr; }
  if (c->d_refr) { cudaFree(c->d_refr); c->d_refr = nullptr; }
  delete c->pool; c->pool = nullptr;
  for (int k = 0; k < 2; ++k) if (c->stage[k]) { cudaFreeHost(c->stage[k]); cudaEventDestroy(c->stage_ev[k]); c->stage[k] = nullptr; }
  if (c->sort_tmp) { cudaFree(c->sort_tmp); c->sort_tmp = nullptr; }
  for (int k = 0; k < HRT_SORT_STREAMS; ++k) {
    if (c->sort_side_tmp[k]) { cudaFree(c->sort_side_tmp[k]); c->sort_side_tmp[k] = nullptr; }
    if (c->sort_stream[k]) { cudaStreamDestroy(c->sort_stream[k]); c->sort_stream[k] = nullptr; cudaEventDestroy(c->sort_join[k]); }
  }
  if (c->sort_fork) { cudaEventDestroy(c->sort_fork); c->sort_fork = nullptr; }
  if (c->scat_ready) {
    cudaEventDestroy(c->scat_ready); c->scat_ready = nullptr;
    for (int k = 0; k < 2; ++k) { cudaStreamDestroy(c->scat_stream[k]); cudaEventDestroy(c->scat_done[k]); c->scat_stream[k] = nullptr; }
  }
  for (int i = 0; i < 8; ++i) cudaEventDestroy(c->ev[i]);
  for (size_t i = 0; i < c->evpool_n; ++i) cudaEventDestroy(c->evpool[i]);
  free(c->evpool);
  cudaStreamDestroy(c->stream);
  free(c);
}

static inline unsigned nblk(size_t n, unsigned bs = 256) { return (unsigned)((n + bs - 1) / bs); }

/* binary nodes in d_nodes -> the 4-wide nodes the kernels traverse (hrt_bvh.cuh):
 * parents, depth parity, numbering of the even-depth nodes, emission of the
 * octant copies.  Works for either builder; redone whenever the boxes change
 * (re-padding, refit).  Everything on `st`. */
static int build_wide(hrt_ctx *ctx, cudaStream_t st)
{
  const uint32_t n = ctx->num_nodes;
  ctx->wroot = ctx->root_ref; ctx->num_wide = 0;
  if (n == 0) return HRT_OK;
  if (ctx->cap_wscratch < n) {
    dev_free(ctx->d_wparent); dev_free(ctx->d_weven); dev_free(ctx->d_widx); dev_free(ctx->d_wtotal);
    ctx->cap_wscratch = 0;
    CK(dev_alloc(&ctx->d_wparent, n)); CK(dev_alloc(&ctx->d_weven, n)); CK(dev_alloc(&ctx->d_widx, n)); CK(dev_alloc(&ctx->d_wtotal, 1));
    ctx->cap_wscratch = n;
  }
  k_wide_parent<<<nblk(n), 256, 0, st>>>(ctx->d_nodes, (int)n, ctx->d_wparent);
  k_wide_depth<<<nblk(n), 256, 0, st>>>((int)n, ctx->d_wparent, ctx->d_weven);
  k_wide_scan<<<1, 1024, 0, st>>>((int)n, ctx->d_weven, ctx->d_widx, ctx->d_wtotal);
  CK(cudaGetLastError());
  uint32_t total = 0;
  CK(cudaMemcpyAsync(&total, ctx->d_wtotal, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  ctx->num_wide = total;
  ctx->wide_octants = octant_copies(total);
  /* scenes that are read from global memory get 128-byte node slots: a visit touches 4 sectors instead of
   * up to 5 of a node that straddles them (the shared-memory layout stays packed: bank spread, capacity) */
  ctx->wstride = (ctx->wide_octants == 8 && scene_smem_bytes(total, ctx->num_tris, 8) <= HRT_SMEM_SCENE_LIMIT) || getenv("HRT_WIDE_PACKED")
                   ? HRT_WIDE_F4 : 8u;
  const size_t need = (size_t)total * ctx->wstride * ctx->wide_octants;
  if (ctx->cap_wnodes < need) {
    dev_free(ctx->d_wnodes); ctx->cap_wnodes = 0;
    CK(dev_alloc(&ctx->d_wnodes, need)); ctx->cap_wnodes = need;
  }
  k_wide_emit<<<nblk(n), 256, 0, st>>>(ctx->d_nodes, (int)n, ctx->d_weven, ctx->d_widx, ctx->d_wnodes,
                                       (size_t)total * ctx->wstride, ctx->wide_octants, ctx->wstride);
  CK(cudaGetLastError());
  ctx->wroot = 0;
  return HRT_OK;
}

/* (re)emit the traversal nodes with the given padding, on stream `st` */
static int emit_nodes(hrt_ctx *ctx, float pad, cudaStream_t st)
{
  const int n = (int)ctx->num_tris;
  ctx->pad = pad;
  ctx->scene_version++;          /* geometry, tree or padding changed: receiver maps are stale */
  if (ctx->num_nodes == 0) { ctx->wroot = ctx->root_ref; ctx->num_wide = 0; return HRT_OK; }
  if (ctx->sah) {
    k_emit_raw<<<nblk(ctx->num_nodes), 256, 0, st>>>((int)ctx->num_nodes, ctx->d_raw_ref, ctx->d_raw_box, pad,
                                                    ctx->d_nodes, ctx->octants, ctx->num_nodes);
  } else {
    k_emit<<<nblk(n - 1), 256, 0, st>>>(n, ctx->d_kl, ctx->d_kr, ctx->d_kfirst, ctx->d_klast,
                                        ctx->d_newidx + n /* used flags live behind newidx */,
                                        ctx->d_newidx, ctx->d_box, ctx->leaf_max, pad, ctx->d_nodes, ctx->octants, ctx->num_nodes);
  }
  CK(cudaGetLastError());
  return build_wide(ctx, st);
}

/* binned-SAH build over the triangle boxes; fills d_tris / d_tri_gid in leaf
 * order and the raw node arrays.  See the kernels above. */
static int build_sah(hrt_ctx *ctx, uint32_t n, const float *d_boxes, const float4 *d_recs)
{
  cudaStream_t st = ctx->stream;
  const size_t max_work = (size_t)n / (size_t)(ctx->leaf_max + 1) + 2;
  uint32_t *d_idx[2] = { nullptr, nullptr }; int *d_wof[2] = { nullptr, nullptr };
  SahWork *d_work[2] = { nullptr, nullptr }; SahBin *d_bins = nullptr; int *d_cnt = nullptr;
  int rc = HRT_OK, level = 0, count = 1, cur = 0;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
#define CKB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(ctx, HRT_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); goto out; } } while (0)
  CKB(cudaEventCreate(&e0)); CKB(cudaEventCreate(&e1));
  CKB(cudaEventRecord(e0, st));
  /* scratch from the stream's memory pool: a rebuild next to GBs of live run buffers must not pay
   * cudaMalloc / cudaFree (measured: 0.7 s instead of 50 ms on the 958k-triangle scene) */
#define SCRATCH(ptr, count) CKB(cudaMallocAsync((void **)&(ptr), (size_t)(count) * sizeof(*(ptr)), st))
  for (int k = 0; k < 2; ++k) { SCRATCH(d_idx[k], n); SCRATCH(d_wof[k], n); SCRATCH(d_work[k], max_work); }
  SCRATCH(d_bins, max_work * 3 * SAH_BINS); SCRATCH(d_cnt, 2);
#undef SCRATCH
  if (ctx->cap_raw < n) {
    dev_free(ctx->d_raw_ref); dev_free(ctx->d_raw_box); ctx->cap_raw = 0;
    CKB(dev_alloc(&ctx->d_raw_ref, n)); CKB(dev_alloc(&ctx->d_raw_box, (size_t)n * 12));
    ctx->cap_raw = n;
  }
  {
    SahWork root; memset(&root, 0, sizeof root);
    root.start = 0; root.end = n; root.inner = 0;
    for (int k = 0; k < 3; ++k) { root.cb[k] = 0xFFFFFFFFu; root.cb[3 + k] = 0u; root.box[k] = ctx->scene_lo[k]; root.box[3 + k] = ctx->scene_hi[k]; }
    const int c0[2] = { 1, 0 };   /* inner nodes allocated, work items of the next level */
    CKB(cudaMemcpyAsync(d_work[0], &root, sizeof root, cudaMemcpyHostToDevice, st));
    CKB(cudaMemcpyAsync(d_cnt, c0, sizeof c0, cudaMemcpyHostToDevice, st));
    CKB(cudaStreamSynchronize(st));   /* root and c0 are stack temporaries */
  }
  k_sah_init<<<nblk(n), 256, 0, st>>>(n, d_boxes, d_idx[0], d_wof[0], d_work[0]);
  CKB(cudaGetLastError());
  ctx->level_first[0] = 0;
  while (count > 0) {
    if ((size_t)count > max_work) { rc = fail(ctx, HRT_E_STATE, "SAH builder: work list overflow"); goto out; }
    k_sah_prep<<<nblk(count), 256, 0, st>>>(count, d_work[cur], d_bins);
    k_sah_bin<<<nblk(n), 256, 0, st>>>(n, d_idx[cur], d_wof[cur], d_work[cur], d_boxes, d_bins);
    k_sah_split<<<nblk(count, 64), 64, 0, st>>>(count, level, ctx->leaf_max, d_work[cur], d_bins, d_work[cur ^ 1], d_cnt,
                                               ctx->d_raw_ref, ctx->d_raw_box);
    k_sah_part<<<nblk(n), 256, 0, st>>>(n, d_idx[cur], d_wof[cur], d_work[cur], d_work[cur ^ 1], d_boxes,
                                        d_idx[cur ^ 1], d_wof[cur ^ 1]);
    CKB(cudaGetLastError());
    int hc[2];
    CKB(cudaMemcpyAsync(hc, d_cnt, sizeof hc, cudaMemcpyDeviceToHost, st));
    CKB(cudaMemsetAsync(d_cnt + 1, 0, 4, st));
    CKB(cudaStreamSynchronize(st));
    ctx->num_nodes = (uint32_t)hc[0];
    count = hc[1]; cur ^= 1; ++level;
    ctx->level_first[level] = hc[0] - count;      /* the nodes just allocated are the next level */
    if (level > 200) { rc = fail(ctx, HRT_E_STATE, "SAH builder did not terminate"); goto out; }
  }
  k_gather_idx<<<nblk(n), 256, 0, st>>>(d_idx[cur], n, d_recs, ctx->d_tris, ctx->d_tri_gid);
  CKB(cudaGetLastError());
  CKB(cudaEventRecord(e1, st));
  CKB(cudaStreamSynchronize(st));
  cudaEventElapsedTime(&ctx->build_ms, e0, e1);
  ctx->build_levels = level;
out:
#undef CKB
  for (int k = 0; k < 2; ++k) { cudaFreeAsync(d_idx[k], st); cudaFreeAsync(d_wof[k], st); cudaFreeAsync(d_work[k], st); }
  cudaFreeAsync(d_bins, st); cudaFreeAsync(d_cnt, st);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  return rc;
}

extern "C" int hrt_scene_upload(hrt_ctx *ctx, const Scene *scene, Vec3 *normals_out)
{
  if (!ctx || !scene || !scene->meshes || scene->num_meshes == 0) return fail(ctx, HRT_E_ARG, "scene is empty");
  CK(cudaSetDevice(ctx->device));
  free_scene_dev(ctx);

  /* host: concatenate vertices / indices (indices rebased), validate */
  size_t nv = 0, nt = 0;
  for (uint32_t m = 0; m < scene->num_meshes; ++m) { nv += scene->meshes[m].num_vertices; nt += scene->meshes[m].num_triangles; }
  if (nt >= (1u << 28)) return fail(ctx, HRT_E_ARG, "too many triangles (%zu)", nt);
  const uint32_t M = scene->num_meshes;
  float *h_v = (float *)malloc((nv ? nv : 1) * 12);
  uint32_t *h_i = (uint32_t *)malloc((nt ? nt : 1) * 12);
  uint32_t *h_mesh_of = (uint32_t *)malloc((nt ? nt : 1) * 4);
  uint32_t *h_mat = (uint32_t *)malloc(M * 4);
  float *h_vel = (float *)malloc(M * 12);
  if (!h_v || !h_i || !h_mesh_of || !h_mat || !h_vel) { free(h_v); free(h_i); free(h_mesh_of); free(h_mat); free(h_vel); return fail(ctx, HRT_E_NOMEM, "out of host memory"); }
  size_t vo = 0, to = 0; float max_abs = 0.f; int bad = 0;
  float blo[3] = { 3e38f, 3e38f, 3e38f }, bhi[3] = { -3e38f, -3e38f, -3e38f };
  for (uint32_t m = 0; m < M && !bad; ++m) {
    const Mesh *me = &scene->meshes[m];
    if (me->material_index >= NUM_G_MATERIALS) { bad = 1; break; }
    memcpy(h_v + 3 * vo, me->vs, (size_t)me->num_vertices * 12);
    for (size_t k = 0; k < (size_t)3 * me->num_vertices; ++k) {
      const float v = h_v[3 * vo + k], a = fabsf(v);
      if (a > max_abs) max_abs = a;
      if (!(a == a)) bad = 2;
      if (v < blo[k % 3]) blo[k % 3] = v;
      if (v > bhi[k % 3]) bhi[k % 3] = v;
    }
    for (size_t k = 0; k < (size_t)3 * me->num_triangles; ++k) {
      const uint32_t ix = me->is[k];
      if (ix >= me->num_vertices) { bad = 3; break; }
      h_i[3 * to + k] = ix + (uint32_t)vo;
    }
    for (uint32_t f = 0; f < me->num_triangles; ++f) h_mesh_of[to + f] = m;
    h_mat[m] = me->material_index;
    h_vel[3 * m] = me->velocity.x; h_vel[3 * m + 1] = me->velocity.y; h_vel[3 * m + 2] = me->velocity.z;
    vo += me->num_vertices; to += me->num_triangles;
  }
  if (bad) {
    free(h_v); free(h_i); free(h_mesh_of); free(h_mat); free(h_vel);
    return fail(ctx, HRT_E_ARG, bad == 1 ? "material index out of range" : bad == 2 ? "NaN vertex" : "vertex index out of range");
  }

  const uint32_t n = (uint32_t)nt;
  ctx->num_tris = n; ctx->num_meshes = M; ctx->scene_max_abs = max_abs;
  for (int k = 0; k < 3; ++k) { ctx->scene_lo[k] = nv ? blo[k] : 0.f; ctx->scene_hi[k] = nv ? bhi[k] : 1.f; }
  float *&d_v = ctx->d_verts; uint32_t *&d_i = ctx->d_idx3; float4 *&d_recs = ctx->d_recs; float *&d_boxes = ctx->d_tboxes;
  unsigned *&d_bounds = ctx->d_bounds; uint64_t *d_keys = nullptr, *d_keys2 = nullptr; int *d_parent = nullptr;
  uint32_t *h_vmesh = nullptr;
  unsigned *d_arrive = nullptr; void *d_tmp = nullptr; size_t tmp_bytes = 0;
  int rc = HRT_OK;
#define CKG(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(ctx, HRT_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); goto done; } } while (0)
  {
    cudaStream_t st = ctx->stream;
    CKG(dev_alloc(&d_v, nv * 3)); CKG(dev_alloc(&d_i, (size_t)n * 3));
    CKG(dev_alloc(&ctx->d_vmesh, nv));
    ctx->num_verts = nv;
    h_vmesh = (uint32_t *)malloc((nv ? nv : 1) * 4);
    if (!h_vmesh) { rc = fail(ctx, HRT_E_NOMEM, "out of host memory"); goto done; }
    {
      size_t o = 0; float ms = 0.f;
      for (uint32_t m = 0; m < M; ++m) {
        for (uint32_t k = 0; k < scene->meshes[m].num_vertices; ++k) h_vmesh[o++] = m;
        const Vec3 v = scene->meshes[m].velocity;
        ms = fmaxf(ms, fmaxf(fabsf(v.x), fmaxf(fabsf(v.y), fabsf(v.z))));
      }
      ctx->max_speed = ms;
    }
    CKG(cudaMemcpyAsync(ctx->d_vmesh, h_vmesh, nv * 4, cudaMemcpyHostToDevice, st));
    CKG(dev_alloc(&d_recs, (size_t)n * 3)); CKG(dev_alloc(&d_boxes, (size_t)n * 6));
    CKG(dev_alloc(&d_bounds, 6));
    CKG(dev_alloc(&ctx->d_tris, (size_t)n * 3)); CKG(dev_alloc(&ctx->d_tri_gid, n));
    CKG(dev_alloc(&ctx->d_mesh_of, n)); CKG(dev_alloc(&ctx->d_mesh_mat, M)); CKG(dev_alloc(&ctx->d_mesh_vel, (size_t)M * 3));
    const size_t nn = n > 1 ? n - 1 : 1;
    CKG(cudaMemcpyAsync(d_v, h_v, nv * 12, cudaMemcpyHostToDevice, st));
    CKG(cudaMemcpyAsync(d_i, h_i, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    CKG(cudaMemcpyAsync(ctx->d_mesh_of, h_mesh_of, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CKG(cudaMemcpyAsync(ctx->d_mesh_mat, h_mat, (size_t)M * 4, cudaMemcpyHostToDevice, st));
    CKG(cudaMemcpyAsync(ctx->d_mesh_vel, h_vel, (size_t)M * 12, cudaMemcpyHostToDevice, st));
    const unsigned init_bounds[6] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u };
    CKG(cudaMemcpyAsync(d_bounds, init_bounds, sizeof init_bounds, cudaMemcpyHostToDevice, st));
    ctx->num_nodes = 0; ctx->root_ref = 0; ctx->octants = 1; ctx->wide_octants = 8; ctx->num_wide = 0; ctx->wroot = 0;
    if (n > 0) {
      k_tri_setup<<<nblk(n), 256, 0, st>>>(d_v, d_i, n, d_recs, d_boxes, d_bounds);
      ctx->sah = !getenv("HRT_BVH_LBVH");
      if (ctx->sah && (int)n > ctx->leaf_max) {
        CKG(cudaGetLastError());
        const int brc = build_sah(ctx, n, d_boxes, d_recs);
        if (brc) { rc = brc; goto done; }
        ctx->octants = 1;
        CKG(dev_alloc(&ctx->d_nodes, (size_t)n * 4)); ctx->cap_nodes = n;      /* <= n - 1 inner nodes: a rebuild reuses it */
        goto built;
      }
      /* Morton/Karras builder (HRT_BVH_LBVH=1, and scenes of <= leaf_max triangles) */
      ctx->sah = false;
      CKG(dev_alloc(&d_keys, n)); CKG(dev_alloc(&d_keys2, n));
      CKG(dev_alloc(&ctx->d_kl, nn)); CKG(dev_alloc(&ctx->d_kr, nn)); CKG(dev_alloc(&ctx->d_kfirst, nn));
      CKG(dev_alloc(&ctx->d_klast, nn)); CKG(dev_alloc(&ctx->d_newidx, 2 * (size_t)n + 2));
      CKG(dev_alloc(&ctx->d_box, (2 * (size_t)n) * 6)); CKG(dev_alloc(&d_parent, 2 * (size_t)n));
      CKG(dev_alloc(&d_arrive, nn));
      k_morton<<<nblk(n), 256, 0, st>>>(d_boxes, n, d_bounds, d_keys);
      CKG(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_keys, d_keys2, (int)n, 0, 64, st));
      CKG(cudaMalloc(&d_tmp, tmp_bytes ? tmp_bytes : 1));
      CKG(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_keys, d_keys2, (int)n, 0, 64, st));
      k_gather<<<nblk(n), 256, 0, st>>>(d_keys2, n, d_recs, d_boxes, ctx->d_tris, ctx->d_tri_gid, ctx->d_box);
      CKG(cudaGetLastError());
      if ((int)n <= ctx->leaf_max) {
        ctx->root_ref = hrt_leaf_ref(0u, n);
      } else {
        int *d_used = ctx->d_newidx + n;
        CKG(cudaMemsetAsync(d_arrive, 0, nn * sizeof(unsigned), st));
        k_karras<<<nblk(n - 1), 256, 0, st>>>(d_keys2, (int)n, ctx->d_kl, ctx->d_kr, ctx->d_kfirst, ctx->d_klast, d_parent);
        k_refit<<<nblk(n), 256, 0, st>>>((int)n, ctx->d_kl, ctx->d_kr, d_parent, ctx->d_box, d_arrive);
        k_mark<<<nblk(n - 1), 256, 0, st>>>((int)n, ctx->d_kfirst, ctx->d_klast, ctx->leaf_max, d_used);
        CKG(cudaGetLastError());
        void *d_tmp2 = nullptr; size_t tb2 = 0;
        CKG(cub::DeviceScan::ExclusiveSum(nullptr, tb2, d_used, ctx->d_newidx, (int)n - 1, st));
        CKG(cudaMalloc(&d_tmp2, tb2 ? tb2 : 1));
        cudaError_t es = cub::DeviceScan::ExclusiveSum(d_tmp2, tb2, d_used, ctx->d_newidx, (int)n - 1, st);
        int last_idx = 0, last_used = 0;
        if (es == cudaSuccess) es = cudaMemcpyAsync(&last_idx, ctx->d_newidx + (n - 2), 4, cudaMemcpyDeviceToHost, st);
        if (es == cudaSuccess) es = cudaMemcpyAsync(&last_used, d_used + (n - 2), 4, cudaMemcpyDeviceToHost, st);
        if (es == cudaSuccess) es = cudaStreamSynchronize(st);
        cudaFree(d_tmp2);
        CKG(es);
        ctx->num_nodes = (uint32_t)(last_idx + last_used);
        ctx->root_ref = 0;
        ctx->octants = 1;
        CKG(dev_alloc(&ctx->d_nodes, (size_t)ctx->num_nodes * 4 * ctx->octants));
      }
    }
built:
    ctx->pad = hrt_box_pad(max_abs, ctx->pad_ulps);
    {
      int erc = emit_nodes(ctx, ctx->pad, st);
      if (erc) { rc = erc; goto done; }
    }
    if (normals_out && n) {
      float4 *h_recs = (float4 *)malloc((size_t)n * 48);
      if (!h_recs) { rc = fail(ctx, HRT_E_NOMEM, "out of host memory"); goto done; }
      cudaError_t e2 = cudaMemcpyAsync(h_recs, d_recs, (size_t)n * 48, cudaMemcpyDeviceToHost, st);
      if (e2 == cudaSuccess) e2 = cudaStreamSynchronize(st);
      if (e2 == cudaSuccess)
        for (uint32_t g = 0; g < n; ++g) { normals_out[g].x = h_recs[3 * g + 2].y; normals_out[g].y = h_recs[3 * g + 2].z; normals_out[g].z = h_recs[3 * g + 2].w; }
      free(h_recs);
      CKG(e2);
    }
    CKG(cudaStreamSynchronize(st));
    ctx->have_scene = true;
  }
done:
#undef CKG
  free(h_vmesh);
  cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_parent); cudaFree(d_arrive); cudaFree(d_tmp);
  free(h_v); free(h_i); free(h_mesh_of); free(h_mat); free(h_vel);
  if (rc) free_scene_dev(ctx);
  return rc;
}

/* Advance every mesh by velocity * dt_s on the GPU and bring the BVH up to
 * date: in place (same topology, boxes recomputed level by level, bottom-up)
 * or, with rebuild != 0, by building a fresh SAH tree from the moved triangles. */
extern "C" int hrt_scene_advance(hrt_ctx *ctx, float dt_s, int rebuild)
{
  if (!ctx) return HRT_E_ARG;
  if (!ctx->have_scene) return fail(ctx, HRT_E_STATE, "hrt_scene_advance before hrt_scene_upload");
  if (!(dt_s == dt_s) || fabsf(dt_s) > 1e30f) return fail(ctx, HRT_E_ARG, "bad time step");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const uint32_t n = ctx->num_tris;
  if (n == 0 || ctx->max_speed == 0.f) return HRT_OK;
  if (!ctx->sah && !rebuild && ctx->num_nodes) return fail(ctx, HRT_E_STATE, "in-place refit needs the SAH tree (unset HRT_BVH_LBVH) or rebuild = 1");
  k_move_verts<<<nblk(ctx->num_verts), 256, 0, st>>>(ctx->d_verts, ctx->d_vmesh, ctx->d_mesh_vel, ctx->num_verts, dt_s);
  const unsigned init_bounds[6] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u };
  CK(cudaMemcpyAsync(ctx->d_bounds, init_bounds, sizeof init_bounds, cudaMemcpyHostToDevice, st));
  k_tri_setup<<<nblk(n), 256, 0, st>>>(ctx->d_verts, ctx->d_idx3, n, ctx->d_recs, ctx->d_tboxes, ctx->d_bounds);
  CK(cudaGetLastError());
  unsigned hb[6];
  CK(cudaMemcpyAsync(hb, ctx->d_bounds, sizeof hb, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  float max_abs = 0.f;
  for (int k = 0; k < 3; ++k) {
    ctx->scene_lo[k] = dec_f(hb[k]); ctx->scene_hi[k] = dec_f(hb[3 + k]);
    max_abs = fmaxf(max_abs, fmaxf(fabsf(ctx->scene_lo[k]), fabsf(ctx->scene_hi[k])));
  }
  ctx->scene_max_abs = max_abs;
  const float pad = fmaxf(ctx->pad, hrt_box_pad(max_abs, ctx->pad_ulps));
  if (rebuild && (int)n > ctx->leaf_max) {
    ctx->sah = true;
    const int rc = build_sah(ctx, n, ctx->d_tboxes, ctx->d_recs);
    if (rc) { ctx->have_scene = false; return rc; }
    ctx->octants = 1;
    if (ctx->cap_nodes < n) {
      dev_free(ctx->d_nodes); ctx->cap_nodes = 0;
      CK(dev_alloc(&ctx->d_nodes, (size_t)n * 4)); ctx->cap_nodes = n;
    }
    ctx->root_ref = 0;
  } else {
    /* same leaf order: refresh the triangle records, then the boxes */
    k_gather_idx<<<nblk(n), 256, 0, st>>>(ctx->d_tri_gid, n, ctx->d_recs, ctx->d_tris, ctx->d_tri_gid);
    for (int L = ctx->build_levels - 1; L >= 0 && ctx->num_nodes; --L) {
      const int f = ctx->level_first[L], l = ctx->level_first[L + 1];
      if (l > f) k_sah_refit<<<nblk(l - f), 256, 0, st>>>(f, l, ctx->d_raw_ref, ctx->d_raw_box, ctx->d_tri_gid, ctx->d_tboxes);
    }
    CK(cudaGetLastError());
  }
  const int erc = emit_nodes(ctx, pad, st);
  if (erc) return erc;
  CK(cudaStreamSynchronize(st));
  return HRT_OK;
}

extern "C" int hrt_materials_set(hrt_ctx *ctx, const HrtMaterialDerived table[NUM_G_MATERIALS])
{
  if (!ctx || !table) return HRT_E_ARG;
  memcpy(&ctx->mats, table, sizeof ctx->mats);
  ctx->have_mats = true;
  return HRT_OK;
}

extern "C" int hrt_get_stats(const hrt_ctx *ctx, HrtRunStats *out)
{
  if (!ctx || !out) return HRT_E_ARG;
  *out = ctx->stats;
  return HRT_OK;
}

extern "C" int hrt_fp32_peak(hrt_ctx *ctx, float *tflops_unfused, float *tflops_fma)
{
  if (!ctx) return HRT_E_ARG;
  CK(cudaSetDevice(ctx->device));
  float *d = nullptr;
  CK(dev_alloc(&d, 1));
  const int sms = [&] { int v = 148; cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, ctx->device); return v; }();
  const int blocks = sms * 8, iters = 1 << 15;
  const double flops = (double)blocks * 256 * 8 * 2.0 * iters;
  float best[2] = {0.f, 0.f};
  for (int rep = 0; rep < 4; ++rep)
    for (int v = 0; v < 2; ++v) {
      CK(cudaEventRecord(ctx->ev[6], ctx->stream));
      if (v) k_fp32_peak<true><<<blocks, 256, 0, ctx->stream>>>(d, iters, 0.999f, 1e-3f);
      else   k_fp32_peak<false><<<blocks, 256, 0, ctx->stream>>>(d, iters, 0.999f, 1e-3f);
      CK(cudaGetLastError());
      CK(cudaEventRecord(ctx->ev[7], ctx->stream));
      CK(cudaEventSynchronize(ctx->ev[7]));
      float ms = 0.f; cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]);
      const float tf = (float)(flops / (ms * 1e-3) / 1e12);
      if (rep && tf > best[v]) best[v] = tf;
    }
  cudaFree(d);
  if (tflops_unfused) *tflops_unfused = best[0];
  if (tflops_fma) *tflops_fma = best[1];
  return HRT_OK;
}

static SceneDev scene_dev(const hrt_ctx *c)
{
  SceneDev s;
  s.nodes = c->d_nodes; s.tris = c->d_tris; s.tri_gid = c->d_tri_gid; s.mesh_of = c->d_mesh_of;
  s.mesh_mat = c->d_mesh_mat; s.mesh_vel = c->d_mesh_vel;
  s.num_tris = c->num_tris; s.num_nodes = c->num_nodes; s.root_ref = c->root_ref; s.octants = c->octants;
  for (int k = 0; k < 3; ++k) {
    s.key_lo[k] = c->scene_lo[k];
    s.key_scale[k] = 1024.f / fmaxf(c->scene_hi[k] - c->scene_lo[k], 1e-20f);
  }
  {
    const float ext = fmaxf(c->scene_hi[0] - c->scene_lo[0], c->scene_hi[1] - c->scene_lo[1]);
    /* 511 cells for log2(1 + 2 * ext / 5 cm) octaves: covers a TX anywhere inside the bounds */
    s.key_log = (ext > 256.f && !getenv("HRT_KEY_UNIFORM")) ? 511.f / log2f(1.f + 2.f * ext * 20.f) : 0.f;
  }
  s.wnodes = c->d_wnodes; s.num_wide = c->num_wide; s.wide_octants = c->wide_octants; s.wroot = c->wroot;
  s.wstride = c->wstride ? c->wstride : HRT_WIDE_F4;
  return s;
}

/* make sure boxes are padded for ray origins as far out as max_abs */
static int ensure_pad(hrt_ctx *ctx, float max_abs, cudaStream_t st)
{
  const float need = hrt_box_pad(fmaxf(max_abs, ctx->scene_max_abs), ctx->pad_ulps);
  /* the nodes are rewritten on the stream the traversal kernels will run on: a
   * caller-supplied stream is ordered behind any earlier work on the context's own */
  if (need > ctx->pad) {
    if (st != ctx->stream) CK(cudaStreamSynchronize(ctx->stream));
    return emit_nodes(ctx, need, st);
  }
  return HRT_OK;
}

/* kernel dispatch over the <SMEM, BRUTE> variants */
#define DISPATCH2(KERNEL, smem, brute, grid, block, shbytes, st, ...)                         \
  do {                                                                                        \
    if (smem) { if (brute) KERNEL<true, true><<<grid, block, shbytes, st>>>(__VA_ARGS__);     \
                else       KERNEL<true, false><<<grid, block, shbytes, st>>>(__VA_ARGS__); }  \
    else      { if (brute) KERNEL<false, true><<<grid, block, shbytes, st>>>(__VA_ARGS__);    \
                else       KERNEL<false, false><<<grid, block, shbytes, st>>>(__VA_ARGS__); } \
  } while (0)

/* opt-in to > 48 KB of dynamic shared memory, once per (kernel, device, size): the
 * attribute call is not free and hrt_run is called per step */
template <class K> static cudaError_t allow_smem(K kernel, size_t bytes)
{
  if (bytes <= 48 * 1024) return cudaSuccess;
  static std::mutex mu;
  static std::vector<std::pair<std::pair<const void *, int>, size_t>> done;
  int dev = 0; cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  for (auto &d : done) if (d.first.first == (const void *)kernel && d.first.second == dev && d.second >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) done.push_back({{(const void *)kernel, dev}, bytes});
  return e;
}

typedef void (*BounceFn)(RunDev, SceneDev, HrtMaterialTable, uint32_t);
typedef void (*ScatterFn)(RunDev, SceneDev, HrtMaterialTable, uint32_t, uint32_t);

/* [smem][brute][count]; the instrumented build exists for BVH traversal only */
static BounceFn bounce_fn(bool smem, bool brute, bool count)
{
  static const BounceFn tab[2][2][2] = {
    { { k_bounce<false, false, false>, k_bounce<false, false, true> },
      { k_bounce<false, true, false>,  k_bounce<false, true, false> } },
    { { k_bounce<true, false, false>,  k_bounce<true, false, true> },
      { k_bounce<true, true, false>,   k_bounce<true, true, false> } } };
  return tab[smem][brute][count];
}
/* [smem][brute][warp][count] */
static ScatterFn scatter_fn(bool smem, bool brute, bool warp, bool count)
{
  static const ScatterFn tab[2][2][2][2] = {
    { { { k_scatter<false, false, false, false>, k_scatter<false, false, false, true> },
        { k_scatter<false, false, true, false>,  k_scatter<false, false, true, true> } },
      { { k_scatter<false, true, false, false>,  k_scatter<false, true, false, false> },
        { k_scatter<false, true, true, false>,   k_scatter<false, true, true, false> } } },
    { { { k_scatter<true, false, false, false>,  k_scatter<true, false, false, true> },
        { k_scatter<true, false, true, false>,   k_scatter<true, false, true, true> } },
      { { k_scatter<true, true, false, false>,   k_scatter<true, true, false, false> },
        { k_scatter<true, true, true, false>,    k_scatter<true, true, true, false> } } } };
  return tab[smem][brute][warp][count];
}
/* shadow queries through receiver maps (shared-memory scenes, no brute force): [warp][count], lean [warp] */
static ScatterFn scatter_fn_map(bool warp, bool count)
{
  static const ScatterFn tab[2][2] = {
    { k_scatter<true, false, false, false, false, true>, k_scatter<true, false, false, true, false, true> },
    { k_scatter<true, false, true, false, false, true>,  k_scatter<true, false, true, true, false, true> } };
  return tab[warp][count];
}
static ScatterFn scatter_fn_map_lean(bool warp)
{
  static const ScatterFn tab[2] = { k_scatter<true, false, false, false, true, true>, k_scatter<true, false, true, false, true, true> };
  return tab[warp];
}

/* Receiver maps for this run's receivers (hrt_rxmap.cuh), built on `st` unless the
 * cached ones still apply.  *use = false (and no error) when maps do not apply or
 * a list overflowed: the run then walks the BVH. */
static uint64_t hash_bytes(const void *p, size_t n, uint64_t h)
{
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 0x100000001B3ull; }
  return h;
}

/* Cost model of the shadow queries of one run (C4-like scenes, measured on B200; ms): the BVH walk, or a map
 * of G cells per face edge -- its build (once per receiver set and scene; counted at a quarter, on the
 * assumption that a receiver set is used for a few calls) plus the queries through it. */
static double rxmap_query_ms(uint32_t G, double Q) { return Q * (G >= 256 ? 2.05e-8 : G >= 128 ? 2.15e-8 : 2.5e-8); }
static double rxmap_build_ms(uint32_t G, size_t R, uint32_t num_tris)
{ return (double)R * (0.0075 + 0.037 * ((double)G / 256.0) * ((double)G / 256.0)) * (0.5 + 0.5 * (double)num_tris / 234.0); }   /* 64 RX: 2.85 / 1.09 / 0.63 ms at G = 256 / 128 / 64 */
static double bvh_query_ms(double Q) { return Q * 3.8e-8; }

/* G = 0: choose by the cost model for Q expected shadow queries (may decide for the BVH: *use stays false) */
static int ensure_rxmap(hrt_ctx *ctx, const HrtRunParams *p, const float *d_rx, cudaStream_t st, bool *use, uint32_t G, double Q)
{
  *use = false;
  const size_t R = p->num_rx;
  if (ctx->num_tris == 0 || ctx->num_tris > 65535) return HRT_OK;
  uint64_t key0 = hash_bytes(p->rx_pos, R * sizeof(Vec3), 0xCBF29CE484222325ull);
  key0 = hash_bytes(&ctx->scene_version, 8, key0);
  const bool cached = ctx->map_valid && ctx->map_key0 == key0 && ctx->map_R == R;
  if (G == 0) {
    /* the finest of 256 / 128 / 64 whose cell words stay below 1 GB, or coarser when the run is too short for the
     * finer build to pay (measured on C4, 64 receivers: k_scatter 237 / 248 / 290 ms, builds 2.85 / 1.09 / 0.63 ms) */
    uint32_t Gmax = 256;
    while (Gmax > 64 && R * 6 * (size_t)Gmax * Gmax * 4 > ((size_t)1 << 30)) Gmax >>= 1;
    double best = bvh_query_ms(Q);
    for (uint32_t g = Gmax; g >= 64; g >>= 1) {
      const double c = rxmap_build_ms(g, R, ctx->num_tris) * 0.25 + rxmap_query_ms(g, Q);
      if (c < best) { best = c; G = g; }
    }
    /* a map of these receivers and this scene is there already: free of charge */
    if (cached && rxmap_query_ms(ctx->map_G, Q) <= best) { *use = true; return HRT_OK; }
    if (G == 0) return HRT_OK;
  }
  const size_t cells = R * 6 * (size_t)G * G;
  if (cells * 4 > ((size_t)2 << 30)) return HRT_OK;   /* (k_scatter indexes cells and items with 32 bits) */
  uint64_t key = hash_bytes(&G, 4, key0);
  if (cached && ctx->map_G == G) { *use = true; return HRT_OK; }
  if (ctx->map_failed_key == key) return HRT_OK;                  /* did not fit last time either: walk the BVH */
  ctx->map_valid = false;
  uint32_t per_rx = (uint32_t)(6 * (size_t)G * G * 4);            /* items per receiver: 4 per cell on average, grown on overflow */
  if (const char *e = getenv("HRT_RXMAP_ITEMS_PER_CELL")) { int v = atoi(e); if (v >= 1 && v <= 64) per_rx = (uint32_t)(6 * (size_t)G * G * v); }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  for (int attempt = 0; attempt < 3; ++attempt) {
    if (per_rx >= (1u << 24)) per_rx = (1u << 24) - 1;
    if (R * (size_t)per_rx >= ((size_t)1 << 32)) break;          /* beyond the kernels' 32-bit item index: walk the BVH */
    if (ctx->cap_map_cells < cells) { dev_free(ctx->d_map_cells); ctx->cap_map_cells = 0; CK(dev_alloc(&ctx->d_map_cells, cells)); ctx->cap_map_cells = cells; }
    if (ctx->cap_map_items < R * (size_t)per_rx) { dev_free(ctx->d_map_items); ctx->cap_map_items = 0; CK(dev_alloc(&ctx->d_map_items, R * (size_t)per_rx)); ctx->cap_map_items = R * (size_t)per_rx; }
    if (ctx->cap_map_cursor < 2 * R + 1) { dev_free(ctx->d_map_cursor); ctx->cap_map_cursor = 0; CK(dev_alloc(&ctx->d_map_cursor, 2 * R + 1)); ctx->cap_map_cursor = 2 * R + 1; }
    if (!e0) { CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); }
    CK(cudaMemsetAsync(ctx->d_map_cursor, 0, (R + 1) * 4, st));
    {
      /* depth quantisation per receiver: 255 steps up to the farthest corner of the scene's bounds */
      std::vector<float> is(R);
      const V3 lo = v3(ctx->scene_lo[0], ctx->scene_lo[1], ctx->scene_lo[2]), hi = v3(ctx->scene_hi[0], ctx->scene_hi[1], ctx->scene_hi[2]);
      for (size_t r = 0; r < R; ++r) is[r] = hrt_rxmap_inv_step(v3(p->rx_pos[r].x, p->rx_pos[r].y, p->rx_pos[r].z), lo, hi);
      CK(cudaMemcpyAsync(ctx->d_map_cursor + R + 1, is.data(), R * 4, cudaMemcpyHostToDevice, st));
      CK(cudaStreamSynchronize(st));               /* `is` is a temporary */
    }
    CK(cudaEventRecord(e0, st));
    const SceneDev sc = scene_dev(ctx);
    const uint32_t nb = G / HRT_RXMAP_BLOCK;
    k_rxmap_build<<<dim3(nb * nb, 6, (unsigned)R), 64, 0, st>>>(sc, d_rx, G, 4.f * ctx->pad, ctx->d_map_cells, ctx->d_map_items,
                                                               per_rx, (const float *)(ctx->d_map_cursor + R + 1),
                                                               ctx->d_map_cursor, ctx->d_map_cursor + R);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1, st));
    uint32_t status = 0;
    CK(cudaMemcpyAsync(&status, ctx->d_map_cursor + R, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&ctx->map_build_ms, e0, e1);
    ctx->stats.kernel_launches++;
    if (!status) {
      ctx->map_valid = true; ctx->map_key = key; ctx->map_key0 = key0; ctx->map_R = (uint32_t)R; ctx->map_G = G; ctx->map_items_per_rx = per_rx;
      *use = true;
      break;
    }
    per_rx *= 4;                                                   /* some list did not fit: more room, once or twice */
  }
  if (!*use) ctx->map_failed_key = key;
  if (e0) { cudaEventDestroy(e0); cudaEventDestroy(e1); }
  return HRT_OK;
}

/* summary-only runs without instrumentation: the lean instantiations */
static ScatterFn scatter_fn_lean(bool smem, bool warp)
{
  static const ScatterFn tab[2][2] = {
    { k_scatter<false, false, false, false, true>, k_scatter<false, false, true, false, true> },
    { k_scatter<true, false, false, false, true>,  k_scatter<true, false, true, false, true> } };
  return tab[smem][warp];
}

static int sm_count(int device)
{
  int v = 148; cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device); return v;
}

extern "C" int hrt_closest_hits(hrt_ctx *ctx, const Ray *rays, size_t n, uint32_t flags,
                                uint32_t *tri, float *t, float *theta)
{
  if (!ctx || !ctx->have_scene) return fail(ctx, HRT_E_STATE, "no scene uploaded");
  if (!n) return HRT_OK;
  CK(cudaSetDevice(ctx->device));
  float max_abs = 0.f;
  for (size_t i = 0; i < n; ++i) { max_abs = fmaxf(max_abs, fmaxf(fabsf(rays[i].o.x), fmaxf(fabsf(rays[i].o.y), fabsf(rays[i].o.z)))); }
  int rc = ensure_pad(ctx, max_abs, ctx->stream); if (rc) return rc;
  Ray *d_r = nullptr; uint32_t *d_tri = nullptr; float *d_t = nullptr, *d_th = nullptr;
  cudaStream_t st = ctx->stream;
  CK(dev_alloc(&d_r, n)); CK(dev_alloc(&d_tri, n)); CK(dev_alloc(&d_t, n)); CK(dev_alloc(&d_th, n));
  CK(cudaMemcpyAsync(d_r, rays, n * sizeof(Ray), cudaMemcpyHostToDevice, st));
  const size_t sb = scene_smem_bytes(ctx->num_wide, ctx->num_tris, ctx->wide_octants);
  const bool smem = ctx->wide_octants == 8 && sb <= HRT_SMEM_SCENE_LIMIT && !getenv("HRT_NO_SMEM"), brute = (flags & HRT_FLAG_BRUTE_FORCE) != 0;
  const SceneDev sc = scene_dev(ctx);
  const unsigned grid = (unsigned)min((size_t)sm_count(ctx->device) * 2, (n + HRT_BLOCK - 1) / HRT_BLOCK);
  if (smem) {
    CK(allow_smem(k_closest<true, true>, sb)); CK(allow_smem(k_closest<true, false>, sb));
  }
  DISPATCH2(k_closest, smem, brute, grid, HRT_BLOCK, smem ? sb : 0, st, sc, d_r, (uint32_t)n, d_tri, d_t, d_th);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(tri, d_tri, n * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(t, d_t, n * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(theta, d_th, n * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  cudaFree(d_r); cudaFree(d_tri); cudaFree(d_t); cudaFree(d_th);
  return HRT_OK;
}

/* ---- run buffers ---- */

static int ensure_run_buffers(hrt_ctx *ctx, size_t n, size_t R, size_t T, size_t B, uint32_t flags)
{
  const uint32_t shape_flags = flags & (HRT_FLAG_DENSE | HRT_FLAG_RAYSINFO | HRT_FLAG_TRACE | HRT_FLAG_DENSE_C64);   /* the flags that own buffers */
  if (ctx->cap_n >= n && ctx->cap_R == R && ctx->cap_T == T && ctx->cap_B == B && ctx->cap_flags == shape_flags)
    return HRT_OK;
  free_run_dev(ctx);
  RunDev &r = ctx->rd;
  const size_t TN = T * n;
  CK(dev_alloc(&r.dirs, n * 3)); CK(dev_alloc(&r.rec[0], TN * 4)); CK(dev_alloc(&r.rec[1], TN * 4));
  if (flags & HRT_FLAG_RAYSINFO) { CK(dev_alloc(&r.rays, (B + 1) * TN)); CK(dev_alloc(&r.dead_at, TN)); }
  CK(dev_alloc(&r.queue[0], TN)); CK(dev_alloc(&r.queue[1], TN));
  CK(dev_alloc(&r.queue_alt, TN)); CK(dev_alloc(&r.qkey, TN)); CK(dev_alloc(&r.qkey_alt, TN));
  CK(dev_alloc(&r.qcount, (3 * B + 1) * T));   /* queue sizes + work cursors of k_bounce / k_scatter */
  CK(dev_alloc(&r.amb_list, HRT_AMB_CAP)); CK(dev_alloc(&r.amb_count, 1));
  CK(dev_alloc(&r.dkey, n)); CK(dev_alloc(&r.dkey2, n)); CK(dev_alloc(&r.perm, n)); CK(dev_alloc(&r.perm2, n));
  CK(dev_alloc(&r.counters, 16));
  const size_t slots = R * T * B * n;
  if (flags & HRT_FLAG_DENSE) {
    if (flags & HRT_FLAG_DENSE_C64) {
      CK(dev_alloc(&r.out_f[0], slots * 2)); CK(dev_alloc(&r.out_f[2], slots * 2));
      r.out_f[1] = r.out_f[0] + 1; r.out_f[3] = r.out_f[2] + 1; r.gain_stride = 2;
      CK(dev_alloc(&r.out_f[4], slots)); CK(dev_alloc(&r.out_f[5], slots));
    } else {
      for (int k = 0; k < 6; ++k) CK(dev_alloc(&r.out_f[k], slots));
    }
    CK(dev_alloc(&r.out_dir, slots * 3));
  }
  if (flags & HRT_FLAG_TRACE) {
    CK(dev_alloc(&r.tr_hit, T * B * n)); CK(dev_alloc(&r.tr_t, T * B * n)); CK(dev_alloc(&r.tr_state, slots));
  }
  CK(dev_alloc(&r.pair, R * T * B)); CK(dev_alloc(&r.bounce, T * B));
  ctx->cap_n = n; ctx->cap_R = R; ctx->cap_T = T; ctx->cap_B = B; ctx->cap_flags = shape_flags;
  return HRT_OK;
}

/* number of paths of [0, P) owned by this shard */
static uint64_t shard_count(uint64_t P, uint32_t rank, uint32_t world, uint64_t blk)
{
  if (world <= 1) return P;
  const uint64_t nblocks = (P + blk - 1) / blk;
  uint64_t mine = nblocks / world + ((nblocks % world) > rank ? 1 : 0);
  if (mine == 0) return 0;
  uint64_t cnt = mine * blk;
  /* the globally last block may be partial */
  if ((nblocks - 1) % world == rank) cnt -= nblocks * blk - P;
  return cnt;
}

extern "C" uint64_t hrt_shard_count(uint64_t num_paths, uint32_t rank, uint32_t world, uint64_t block)
{ return shard_count(num_paths, rank, world, block ? block : 1); }
extern "C" uint64_t hrt_shard_path(uint64_t local_index, uint32_t rank, uint32_t world, uint64_t block)
{ return hrt_gpath(local_index, rank, world, block ? block : 1); }

/* device rows [nrows][n_alloc] (elem bytes) -> host rows of pitch P at the
 * columns of the chunk's paths: appended to `tiles` as rectangles of at most
 * HRT_STAGE_BYTES (see run_copy_tiles) */
static cudaError_t d2h_columns(const hrt_ctx *ctx, std::vector<CopyTile> &tiles, void *host, const void *dev, size_t elem,
                               size_t nrows, const RunDev &rd)
{
  if (!host) return cudaSuccess;
  (void)ctx;
  uint64_t L = 0;
  while (L < rd.n) {
    /* contiguous run inside one shard block */
    uint64_t run = rd.n - L;
    if (rd.world > 1) { const uint64_t in_blk = (rd.l0 + L) % rd.blk; run = (rd.blk - in_blk < run) ? rd.blk - in_blk : run; }
    const uint64_t g = hrt_gpath(rd.l0 + L, rd.rank, rd.world, rd.blk);
    const size_t dpitch = (size_t)rd.n_alloc * elem, hpitch = (size_t)rd.P * elem;
    /* split: column pieces when one row exceeds the staging buffer, else row bands */
    const size_t max_cols = HRT_STAGE_BYTES / elem;
    for (uint64_t c0 = 0; c0 < run; c0 += max_cols) {
      const size_t cols = (size_t)((run - c0 < max_cols) ? run - c0 : max_cols);
      const size_t width = cols * elem;
      const size_t band = HRT_STAGE_BYTES / width ? HRT_STAGE_BYTES / width : 1;
      for (size_t r0 = 0; r0 < nrows; r0 += band) {
        CopyTile t;
        t.rows = (nrows - r0 < band) ? nrows - r0 : band;
        t.dev = (const char *)dev + r0 * dpitch + (L + c0) * elem; t.dpitch = dpitch;
        t.host = (char *)host + r0 * hpitch + (g + c0) * elem; t.hpitch = hpitch;
        t.width = width;
        tiles.push_back(t);
      }
    }
    L += run;
  }
  return cudaSuccess;
}

/* Executes the queued copies on stream `st` (i.e. after the kernels queued
 * before) and returns when the host arrays are complete. */
static cudaError_t run_copy_tiles(hrt_ctx *ctx, cudaStream_t st, const std::vector<CopyTile> &tiles)
{
  size_t total = 0;
  for (const CopyTile &t : tiles) total += t.rows * t.width;
  cudaError_t e = cudaSuccess;
  if (total < HRT_STAGE_MIN_TOTAL || getenv("HRT_NO_STAGING")) {
    for (const CopyTile &t : tiles) {
      e = cudaMemcpy2DAsync(t.host, t.hpitch, t.dev, t.dpitch, t.width, t.rows, cudaMemcpyDeviceToHost, st);
      if (e != cudaSuccess) return e;
    }
    return cudaStreamSynchronize(st);
  }
  if (!ctx->pool) {
    unsigned hw = std::thread::hardware_concurrency();
    int k = (int)(hw ? hw / 2 : 4);
    if (const char *s = getenv("HRT_COPY_THREADS")) k = atoi(s);
    k = k < 1 ? 1 : (k > 16 ? 16 : k);
    ctx->pool = new HostPool();
    ctx->pool->start(k);
  }
  for (int k = 0; k < 2; ++k)
    if (!ctx->stage[k]) {
      e = cudaHostAlloc((void **)&ctx->stage[k], HRT_STAGE_BYTES, cudaHostAllocDefault);
      if (e != cudaSuccess) { ctx->stage[k] = nullptr; return e; }
      e = cudaEventCreateWithFlags(&ctx->stage_ev[k], cudaEventDisableTiming);
      if (e != cudaSuccess) return e;
    }
  const size_t nt = tiles.size();
  for (size_t k = 0; k <= nt; ++k) {
    if (k < nt) {
      const CopyTile &t = tiles[k];
      e = cudaMemcpy2DAsync(ctx->stage[k & 1], t.width, t.dev, t.dpitch, t.width, t.rows, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaEventRecord(ctx->stage_ev[k & 1], st);
      if (e != cudaSuccess) return e;
    }
    if (k >= 1) {
      const CopyTile t = tiles[k - 1];
      const char *src = ctx->stage[(k - 1) & 1];
      e = cudaEventSynchronize(ctx->stage_ev[(k - 1) & 1]);
      if (e != cudaSuccess) return e;
      const size_t bytes = t.rows * t.width;
      ctx->pool->run([=](int w, int n) {
        /* worker w copies bytes [a, b) of the packed tile, row segment by row segment */
        size_t a = bytes * (size_t)w / (size_t)n, b = bytes * (size_t)(w + 1) / (size_t)n;
        while (a < b) {
          const size_t r = a / t.width, off = a % t.width;
          const size_t len = (t.width - off < b - a) ? t.width - off : b - a;
          memcpy(t.host + r * t.hpitch + off, src + a, len);
          a += len;
        }
      });
    }
  }
  return cudaSuccess;
}

/* RaysInfo activity masks, bits outside TX 0's path range */
static void raysinfo_tail_bits(const HrtRunParams *p, const uint8_t tail_dead[8])
{
  const size_t T = p->num_tx, B = p->num_bounces;
  const uint64_t P = p->num_paths;
  /* bits outside TX 0's path range: row 0 is all ones (:470); in later rows
   * the tail bits of the last byte stay set for T == 1 */
  uint8_t *A = p->rays_scat->rays_active;
  const size_t rowb = P / 8 + 1;
  memset(A, 0xff, rowb);
  for (size_t t = 0; t < T; ++t)
    for (size_t b = 0; b < B; ++b) {
      uint8_t *row = A + (t * B + b + 1) * rowb;
      for (uint64_t bit = P; bit < rowb * 8; ++bit) {
        /* global bit index `bit` = TX 1, path bit-P (if it exists): its state
         * when the reference copies the row, i.e. after bounce b for t >= 1,
         * after bounce b-1 for t == 0 */
        bool on = true;
        const uint64_t j = bit - P;
        if (T > 1 && j < P && j < 8) {
          const int done = (int)b - (t == 0 ? 1 : 0);     /* last bounce TX 1 has finished */
          on = done < 0 || (int)tail_dead[j] > done;
        }
        if (on) row[bit >> 3] |= (uint8_t)(1u << (bit & 7)); else row[bit >> 3] &= (uint8_t)~(1u << (bit & 7));
      }
    }
}

#define HRT_FLAG_INTERNAL_NO_TAIL 0x80000000u   /* hrt_multi.cu: leave the RaysInfo tail bits to the caller */
extern "C" void hrt_internal_raysinfo_tail(const hrt_ctx *rank0, const HrtRunParams *p)
{ if (rank0 && p && p->rays_scat && p->rays_scat->rays_active) raysinfo_tail_bits(p, rank0->tail_dead); }

/* per-(rx, tx, bounce) and per-(tx, bounce) tables of this call -> added to the caller's (host or device) */
static int flush_summaries(hrt_ctx *ctx, const HrtRunParams *p, const RunDev &rd, uint32_t flags, cudaStream_t st)
{
  const size_t R = p->num_rx, T = p->num_tx, B = p->num_bounces;
  HrtRunStats &S = ctx->stats;
  const size_t np = R * T * B, nb = T * B;
  if (flags & HRT_FLAG_SUMMARY_DEV) {
    /* added on the device, on the run's stream; pair records: words 4, 5 of 6 are doubles */
    k_add_u64<<<nblk(np * 6), 256, 0, st>>>((unsigned long long *)p->pair_summary, (const unsigned long long *)rd.pair, np * 6, 6u, 4u);
    k_add_u64<<<nblk(nb * 4), 256, 0, st>>>((unsigned long long *)p->bounce_summary, (const unsigned long long *)rd.bounce, nb * 4, 0u, 0u);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    S.kernel_launches += 2;
  } else {
    HrtPairSummary *hp = (HrtPairSummary *)malloc(np * sizeof(HrtPairSummary));
    HrtBounceSummary *hb = (HrtBounceSummary *)malloc(nb * sizeof(HrtBounceSummary));
    if (!hp || !hb) { free(hp); free(hb); return fail(ctx, HRT_E_NOMEM, "out of host memory"); }
    cudaError_t e = cudaMemcpyAsync(hp, rd.pair, np * sizeof(HrtPairSummary), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hb, rd.bounce, nb * sizeof(HrtBounceSummary), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) {
      for (size_t i = 0; i < np; ++i) {
        p->pair_summary[i].n_valid += hp[i].n_valid; p->pair_summary[i].n_occluded += hp[i].n_occluded;
        p->pair_summary[i].hit_hash += hp[i].hit_hash; p->pair_summary[i].tau_bits += hp[i].tau_bits;
        p->pair_summary[i].power_te += hp[i].power_te; p->pair_summary[i].power_tm += hp[i].power_tm;
      }
      for (size_t i = 0; i < nb; ++i) {
        p->bounce_summary[i].n_traced += hb[i].n_traced; p->bounce_summary[i].n_hit += hb[i].n_hit;
        p->bounce_summary[i].hit_hash += hb[i].hit_hash; p->bounce_summary[i].t_bits += hb[i].t_bits;
      }
    }
    free(hp); free(hb);
    CK(e);
  }
  return HRT_OK;
}

/* compact path list -> the caller's buffer (unless it already lives in device memory), and the count */
static int flush_path_list(hrt_ctx *ctx, const HrtRunParams *p, const RunDev &rd, uint32_t flags, cudaStream_t st)
{
  unsigned long long found = 0;
  CK(cudaMemcpyAsync(&found, rd.counters + 14, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  const unsigned long long kept = found < p->paths_capacity ? found : p->paths_capacity;
  if (!(flags & HRT_FLAG_PATHLIST_DEV)) {
    CK(cudaMemcpyAsync(p->paths, ctx->d_plist, (size_t)kept * sizeof(HrtPathRecord), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  *p->paths_count = found;
  return HRT_OK;
}

/* impulse response of this call -> added to the caller's array, LoS paths included (rank 0) */
static int flush_cir(hrt_ctx *ctx, const HrtRunParams *p, const RunDev &rd, uint32_t rank, cudaStream_t st)
{
  const size_t R = p->num_rx, T = p->num_tx;
  HrtRunStats &S = ctx->stats;
  const size_t ncir = R * T * (size_t)p->cir_bins * 4;
  float *h = (float *)malloc(ncir * sizeof(float));
  if (!h) return fail(ctx, HRT_E_NOMEM, "out of host memory");
  unsigned long long dropped = 0;
  cudaError_t e = cudaMemcpyAsync(h, ctx->d_cir, ncir * sizeof(float), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(&dropped, rd.counters + 15, 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) {
    for (size_t i = 0; i < ncir; ++i) p->cir[i] += h[i];
    S.cir_dropped = dropped;
    /* the LoS paths (reference :556-577), written to p->los above */
    if (p->los && rank == 0)
      for (size_t k = 0; k < R * T; ++k) {
        if (p->los->a_te_re[k] == 0.f && p->los->tau[k] == 0.f) continue;   /* blocked */
        const float fb = (p->los->tau[k] - p->cir_tau0_s) * rd.cir_inv_dt;
        if (fb >= 0.f && fb < (float)p->cir_bins) {
          float *dst = p->cir + (k * p->cir_bins + (size_t)fb) * 4;
          dst[0] += p->los->a_te_re[k]; dst[1] += p->los->a_te_im[k];
          dst[2] += p->los->a_tm_re[k]; dst[3] += p->los->a_tm_im[k];
        } else S.cir_dropped++;
      }
  }
  free(h);
  CK(e);
  return HRT_OK;
}

/* Line of sight (reference :514-577): one query per (rx, tx) pair on the GPU,
 * results and RaysInfo bits into the caller's host arrays. */
static int run_los(hrt_ctx *ctx, const HrtRunParams *p, const RunDev &rd, const SceneDev &sc, bool smem, bool brute,
                   size_t scene_sb, cudaStream_t st)
{
  const size_t R = p->num_rx, T = p->num_tx;
  HrtRunStats &S = ctx->stats;
  const size_t npair = R * T;
  if (ctx->cap_los < npair) { if (ctx->d_los) cudaFree(ctx->d_los); CK(cudaMalloc(&ctx->d_los, npair * sizeof(HrtLosOut))); ctx->cap_los = npair; }
  DISPATCH2(k_los, smem, brute, nblk(npair, HRT_BLOCK), HRT_BLOCK, smem ? scene_sb : 0, st, rd, sc, (HrtLosOut *)ctx->d_los);
  CK(cudaGetLastError());
  S.kernel_launches++; S.los_queries = npair;
  HrtLosOut *h = (HrtLosOut *)malloc(npair * sizeof(HrtLosOut));
  if (!h) return fail(ctx, HRT_E_NOMEM, "out of host memory");
  cudaError_t e = cudaMemcpyAsync(h, ctx->d_los, npair * sizeof(HrtLosOut), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { free(h); CK(e); }
  ChannelInfo *L = p->los;
  for (size_t k = 0; k < npair; ++k) {
    const size_t r = k / T, t = k % T;
    L->a_te_im[k] = L->a_tm_im[k] = 0.f;                                     /* reference :515-516 */
    if (p->rays_los && p->rays_los->rays) {                                  /* reference :526-528 */
      Ray *lr = &p->rays_los->rays[k];
      lr->o = p->tx_pos[t];
      lr->d = vec3_sub(&p->rx_pos[r], &lr->o);
    }
    uint8_t *bits = (p->rays_los && p->rays_los->rays_active) ? p->rays_los->rays_active : nullptr;
    if (h[k].state == 0) {                                                   /* blocked, reference :548-554 */
      L->a_te_re[k] = L->a_tm_re[k] = L->tau[k] = 0.f;
      if (bits) bits[k / 8] &= (uint8_t)~(1u << (k % 8));
      continue;
    }
    L->directions_tx[k].x = h[k].dir_tx.x; L->directions_tx[k].y = h[k].dir_tx.y; L->directions_tx[k].z = h[k].dir_tx.z;
    L->directions_rx[k].x = h[k].dir_rx.x; L->directions_rx[k].y = h[k].dir_rx.y; L->directions_rx[k].z = h[k].dir_rx.z;
    L->a_te_re[k] = L->a_tm_re[k] = h[k].a;
    L->tau[k] = h[k].tau;
    L->freq_shift[k] = h[k].freq;
    if (bits) bits[k / 8] |= (uint8_t)(1u << (k % 8));
  }
  free(h);
  return HRT_OK;
}

extern "C" int hrt_run(hrt_ctx *ctx, const HrtRunParams *p)
{
  if (!ctx || !p) return HRT_E_ARG;
  const auto host_t0 = std::chrono::steady_clock::now();
  auto host_ms = [&] { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - host_t0).count(); };
  if (!ctx->have_scene) return fail(ctx, HRT_E_STATE, "hrt_run before hrt_scene_upload");
  if (!ctx->have_mats) return fail(ctx, HRT_E_STATE, "hrt_run before hrt_materials_set");
  const size_t R = p->num_rx, T = p->num_tx, B = p->num_bounces;
  const uint64_t P = p->num_paths;
  if (!R || !T || !B || !P || !(p->carrier_frequency_GHz > 0.f)) return fail(ctx, HRT_E_ARG, "num_rx, num_tx, num_paths, num_bounces and the carrier frequency must be > 0");
  if (B > 254 || T > 65535 || R > (1u << 20)) return fail(ctx, HRT_E_ARG, "num_bounces <= 254, num_tx <= 65535, num_rx <= 2^20");
  if (!p->rx_pos || !p->tx_pos || !p->rx_vel || !p->tx_vel) return fail(ctx, HRT_E_ARG, "NULL position/velocity array");
  uint32_t flags = p->flags;
  if ((flags & HRT_FLAG_RAYSINFO) && !(flags & HRT_FLAG_DENSE)) return fail(ctx, HRT_E_ARG, "RAYSINFO needs DENSE");
  if ((flags & HRT_FLAG_DENSE) && !p->scat) return fail(ctx, HRT_E_ARG, "DENSE needs scat outputs");
  if ((flags & HRT_FLAG_DENSE_C64) && (!(flags & HRT_FLAG_DENSE) || !p->scat_a_te_c64 || !p->scat_a_tm_c64))
    return fail(ctx, HRT_E_ARG, "DENSE_C64 needs DENSE and both complex arrays");
  if ((flags & HRT_FLAG_RAYSINFO) && !p->rays_scat) flags &= ~HRT_FLAG_RAYSINFO;
  if ((flags & HRT_FLAG_SUMMARY) && (!p->pair_summary || !p->bounce_summary)) return fail(ctx, HRT_E_ARG, "SUMMARY needs both summary arrays");
  if ((flags & HRT_FLAG_HOST_DIRS) && !p->dirs) return fail(ctx, HRT_E_ARG, "HOST_DIRS needs dirs");
  if ((flags & HRT_FLAG_PATHLIST) && (!p->paths || !p->paths_count || !p->paths_capacity || P >= (1ull << 32)))
    return fail(ctx, HRT_E_ARG, "PATHLIST needs paths, paths_count, paths_capacity > 0 and num_paths < 2^32");
  if ((flags & HRT_FLAG_CIR) && (!p->cir || !p->cir_bins || !(p->cir_dt_s > 0.f))) return fail(ctx, HRT_E_ARG, "CIR needs cir, cir_bins > 0 and cir_dt_s > 0");
  const uint32_t world = p->shard_world ? p->shard_world : 1, rank = p->shard_rank;
  uint64_t blk = p->shard_block ? p->shard_block : (1u << 20);
  if (world > 1 && (rank >= world || (blk % 32))) return fail(ctx, HRT_E_ARG, "bad shard (rank %u of %u, block %llu must be a multiple of 32)", rank, world, (unsigned long long)blk);
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = p->stream ? (cudaStream_t)p->stream : ctx->stream;

  /* positions -> device; padding must cover every ray origin */
  float max_abs = 0.f;
  for (size_t i = 0; i < R; ++i) max_abs = fmaxf(max_abs, fmaxf(fabsf(p->rx_pos[i].x), fmaxf(fabsf(p->rx_pos[i].y), fabsf(p->rx_pos[i].z))));
  for (size_t i = 0; i < T; ++i) max_abs = fmaxf(max_abs, fmaxf(fabsf(p->tx_pos[i].x), fmaxf(fabsf(p->tx_pos[i].y), fabsf(p->tx_pos[i].z))));
  if (!(max_abs < 1e30f)) return fail(ctx, HRT_E_ARG, "non-finite position");
  int rc = ensure_pad(ctx, max_abs, st); if (rc) return rc;
  const size_t npos = 6 * (R + T);
  if (ctx->cap_pos < npos) { dev_free(ctx->d_pos); CK(dev_alloc(&ctx->d_pos, npos)); ctx->cap_pos = npos; }
  float *d_rx = ctx->d_pos, *d_tx = d_rx + 3 * R, *d_rxv = d_tx + 3 * T, *d_txv = d_rxv + 3 * R;
  CK(cudaMemcpyAsync(d_rx, p->rx_pos, R * 12, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_tx, p->tx_pos, T * 12, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_rxv, p->rx_vel, R * 12, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_txv, p->tx_vel, T * 12, cudaMemcpyHostToDevice, st));

  /* chunking of this shard's paths */
  const uint64_t n_shard = shard_count(P, rank, world, blk);
  size_t chunk = 1u << 23;
  if (const char *s = getenv("HRT_CHUNK")) { long long v = atoll(s); if (v >= 32) chunk = (size_t)v; }
  const uint32_t shape_flags_now = flags & (HRT_FLAG_DENSE | HRT_FLAG_RAYSINFO | HRT_FLAG_TRACE | HRT_FLAG_DENSE_C64);
  const bool buffers_fit = ctx->cap_n && ctx->cap_R == R && ctx->cap_T == T && ctx->cap_B == B && ctx->cap_flags == shape_flags_now;
  if (buffers_fit && ctx->cap_n >= (n_shard < chunk ? n_shard : chunk)) {
    /* the buffers of the previous run already hold a full chunk of this shape: no memory query */
  } else {
    /* per-ray state (two 64-byte records, three queue words, two key words per TX,
     * sort scratch) must fit: at most half of the free device memory */
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b) {
      const size_t have = free_b + (ctx->cap_n ? ctx->cap_n * ctx->cap_T * 160 : 0);   /* our own buffers are reusable */
      const size_t lim = (have / 2) / (T * 176 + 40);
      if (chunk > lim) chunk = lim < 32 ? 32 : lim;
    }
  }
  if (flags & (HRT_FLAG_DENSE | HRT_FLAG_TRACE)) {
    /* bound the dense staging buffers to ~6 GB */
    const size_t per_path = R * T * B * 40 + T * (B + 1) * 24 + 64 * T;
    const size_t lim = (size_t)6e9 / per_path;
    if (chunk > lim) chunk = lim < 32 ? 32 : lim;
  }
  /* whole warps of paths; a chunk need not be a whole number of shard blocks (k_raygen,
   * d2h_columns and the RaysInfo bookkeeping split runs at block boundaries themselves) */
  chunk &= ~(size_t)31;
  if (chunk == 0) chunk = 32;
  if (chunk > n_shard) chunk = n_shard ? n_shard : 32;
  rc = ensure_run_buffers(ctx, chunk, R, T, B, flags); if (rc) return rc;

  RunDev rd = ctx->rd;
  rd.R = (uint32_t)R; rd.T = (uint32_t)T; rd.B = (uint32_t)B; rd.P = P;
  rd.n_alloc = (uint32_t)ctx->cap_n; rd.rank = rank; rd.world = world; rd.blk = blk;
  rd.rx_pos = d_rx; rd.tx_pos = d_tx; rd.rx_vel = d_rxv; rd.tx_vel = d_txv;
  rd.flags = flags;
  const float f_hz = (float)((double)p->carrier_frequency_GHz * 1e9);          /* reference :483 */
  rd.k.fsl_k = 4.f * HRT_PI * f_hz / HRT_C0;                                   /* reference :484 */
  rd.k.dop_k = f_hz / HRT_C0;                                                  /* reference :488 */

  HrtRunStats &S = ctx->stats;
  memset(&S, 0, sizeof S);
  S.num_tris = ctx->num_tris; S.num_nodes = ctx->num_nodes; S.box_pad = ctx->pad;
  S.bvh_sah = ctx->sah; S.bvh_levels = (uint32_t)ctx->build_levels; S.bvh_build_ms = ctx->build_ms;
  const SceneDev sc = scene_dev(ctx);
  const size_t scene_sb = scene_smem_bytes(ctx->num_wide, ctx->num_tris, ctx->wide_octants);
  const bool smem = ctx->wide_octants == 8 && scene_sb <= HRT_SMEM_SCENE_LIMIT && !getenv("HRT_NO_SMEM");
  const bool brute = (flags & HRT_FLAG_BRUTE_FORCE) != 0;
  S.scene_in_smem = smem;
  const int sms = sm_count(ctx->device);

  /* scatter mapping: a thread per hit (receivers in sequence, coherent lanes
   * thanks to the direction sort) whenever there are enough hits to fill the
   * machine; a warp per hit (lanes over receivers) for few rays x many RX */
  bool warp_mode = R >= 8 && (uint64_t)T * (P < chunk ? P : chunk) < (uint64_t)sms * 8192;   /* crossover measured with scripts/mode_sweep.py */
  if (const char *m = getenv("HRT_SCATTER_MODE")) warp_mode = (m[0] == 'w') && R >= 2;
  const bool count = (flags & HRT_FLAG_COUNT) != 0 && !brute;
  const BounceFn f_bounce = bounce_fn(smem, brute, count);
  const bool lean = !brute && !count && !(flags & (HRT_FLAG_DENSE | HRT_FLAG_TRACE | HRT_FLAG_CIR | HRT_FLAG_PATHLIST | HRT_FLAG_EXT_LOBES));
  if ((flags & HRT_FLAG_EXT_REFRACT) && (!p->refr_rays || !p->refr_capacity || !p->refr_count)) return fail(ctx, HRT_E_ARG, "EXT_REFRACT needs refr_rays, refr_capacity and refr_count");
  if (flags & (HRT_FLAG_EXT_LOBES | HRT_FLAG_EXT_REFRACT)) {
    HrtExtTable ext;
    for (int i = 0; i < NUM_G_MATERIALS; ++i) {
      ext.m[i].s1 = g_materials[i].s1; ext.m[i].s2 = g_materials[i].s2; ext.m[i].s3 = g_materials[i].s3;
      ext.m[i].a1 = g_materials[i].s1_alpha; ext.m[i].a3 = g_materials[i].s3_alpha;
    }
    CK(cudaMemcpyToSymbolAsync(c_ext, &ext, sizeof ext, 0, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));                 /* `ext` is a stack temporary */
  }
  if ((flags & HRT_FLAG_PATHLIST_DEV) && !(flags & HRT_FLAG_PATHLIST)) return fail(ctx, HRT_E_ARG, "PATHLIST_DEV needs PATHLIST");
  /* receiver maps instead of the tree walk for the shadow queries: scenes that live in
   * shared memory, enough receivers for the build (a few ms) to pay off */
  bool use_map = false;
  {
    size_t min_rx = 8;
    if (const char *e = getenv("HRT_RXMAP_MIN_RX")) min_rx = (size_t)atoll(e);
    const char *force = getenv("HRT_RXMAP");          /* 0: never, 1: whenever possible */
    uint32_t G = 0;                                   /* 0: by the cost model of ensure_rxmap */
    if (const char *e = getenv("HRT_RXMAP_G")) { int v = atoi(e); if (v >= 8 && v <= 1024) G = (uint32_t)v & ~7u; }
    if (force && force[0] == '1' && G == 0) {         /* forced: the finest that fits */
      G = 256;
      while (G > 64 && R * 6 * (size_t)G * G * 4 > ((size_t)1 << 30)) G >>= 1;
    }
    /* expected shadow queries: about half of the rays survive each bounce */
    const double Q = (double)T * (double)n_shard * (double)R * (B < 3 ? (double)B : 2.5);
    /* below ~1e8 expected queries the map does not pay: measured crossover of a steady state ~3e7, of a first call
     * (build + allocations) ~3e8 (scripts/probe_mapsize.py) */
    double min_q = 1e8;
    if (const char *e = getenv("HRT_RXMAP_MIN_QUERIES")) min_q = atof(e);
    bool want = force ? force[0] == '1' : (R >= min_rx && Q >= min_q);
    if (!force && G != 0 && bvh_query_ms(Q) <= rxmap_build_ms(G, R, ctx->num_tris) * 0.25 + rxmap_query_ms(G, Q)) want = false;
    if (want && smem && !brute) { rc = ensure_rxmap(ctx, p, d_rx, st, &use_map, G, Q); if (rc) return rc; }
  }
  rd.map.cells = ctx->d_map_cells; rd.map.items = ctx->d_map_items; rd.map.G = ctx->map_G; rd.map.items_per_rx = ctx->map_items_per_rx;
  rd.map.inv_step = (const float *)(ctx->d_map_cursor + p->num_rx + 1);
  const ScatterFn f_scatter = use_map ? (lean ? scatter_fn_map_lean(warp_mode) : scatter_fn_map(warp_mode, count))
                                      : lean ? scatter_fn_lean(smem, warp_mode) : scatter_fn(smem, brute, warp_mode, count);
  S.rx_map = use_map; S.rx_map_build_ms = use_map ? ctx->map_build_ms : 0.f; S.rx_map_cells = use_map ? ctx->map_G : 0;
  /* scatter kernel shared memory: scene (with receiver maps: its triangle records only) + receivers + reduction table */
  const size_t scat_scene_sb = use_map ? scene_smem_bytes(0, ctx->num_tris, 0) : (smem ? scene_sb : 0);
  const size_t rx_sb = ((3 * R + 3) / 4) * 16 + ((flags & HRT_FLAG_SUMMARY) ? R * sizeof(PairAcc) : 0);
  const bool smem_rx_ok = scat_scene_sb + rx_sb <= 110 * 1024;
  const size_t scat_sb = scat_scene_sb + (smem_rx_ok ? rx_sb : 0);
  if (smem) {
    CK(allow_smem(f_bounce, scene_sb));
    CK(allow_smem(k_los<true, true>, scene_sb)); CK(allow_smem(k_los<true, false>, scene_sb));
  }
  CK(allow_smem(f_scatter, scat_sb));
  CK(cudaMemsetAsync(rd.counters, 0, 16 * sizeof(unsigned long long), st));
  rd.cir = nullptr; rd.cir_bins = 0; rd.cir_tau0 = 0.f; rd.cir_inv_dt = 0.f;
  rd.plist = nullptr; rd.plist_cap = 0; rd.plist_count = rd.counters + 14;
  if (flags & HRT_FLAG_PATHLIST) {
    if (flags & HRT_FLAG_PATHLIST_DEV) {
      rd.plist = (float4 *)p->paths;
    } else {
      if (ctx->cap_plist < p->paths_capacity) {
        if (ctx->d_plist) cudaFree(ctx->d_plist);
        ctx->d_plist = nullptr; ctx->cap_plist = 0;
        CK(dev_alloc(&ctx->d_plist, (size_t)p->paths_capacity * 3)); ctx->cap_plist = p->paths_capacity;
      }
      rd.plist = ctx->d_plist;
    }
    rd.plist_cap = p->paths_capacity;
  }
  rd.refr = nullptr; rd.refr_cap = 0;
  if (flags & HRT_FLAG_EXT_REFRACT) {
    if (ctx->cap_refr < p->refr_capacity) {
      if (ctx->d_refr) cudaFree(ctx->d_refr);
      ctx->d_refr = nullptr; ctx->cap_refr = 0;
      CK(dev_alloc(&ctx->d_refr, (size_t)p->refr_capacity * 3)); ctx->cap_refr = p->refr_capacity;
    }
    rd.refr = ctx->d_refr; rd.refr_cap = p->refr_capacity;
  }
  if (flags & HRT_FLAG_CIR) {
    const size_t ncir = R * T * (size_t)p->cir_bins * 4;
    if (ctx->cap_cir < ncir) {
      if (ctx->d_cir) cudaFree(ctx->d_cir);
      ctx->d_cir = nullptr; ctx->cap_cir = 0;
      CK(dev_alloc(&ctx->d_cir, ncir)); ctx->cap_cir = ncir;
    }
    CK(cudaMemsetAsync(ctx->d_cir, 0, ncir * sizeof(float), st));
    rd.cir = ctx->d_cir; rd.cir_bins = p->cir_bins; rd.cir_tau0 = p->cir_tau0_s; rd.cir_inv_dt = 1.f / p->cir_dt_s;
  }
  /* hit order matters for the thread-per-hit mapping (lanes = neighbouring hits);
   * with a warp per hit the lanes share their origin anyway, and small runs are
   * better off without the per-depth count read-back the sort needs */
  /* all 30 Morton bits order the hits (sorting 24 saves a radix pass, 2 ms, and costs k_scatter as much in coherence) */
  int hit_sort_low_bit = 0;
  if (const char *e = getenv("HRT_HIT_SORT_LOW_BIT")) { int v = atoi(e); if (v >= 0 && v <= 22) hit_sort_low_bit = v; }
  /* tiny runs (< 2^22 path x receiver slots per depth, k_scatter < ~0.1 ms): ordering the work costs more than it
   * saves -- a dozen launches of ~0.2 ms in total (configs[0]: 0.79 -> 0.55 ms per call) */
  const bool tiny = (uint64_t)T * n_shard * (R ? R : 1) < (1ull << 22) && !getenv("HRT_SORT_ALWAYS");
  const bool sort_hits = !getenv("HRT_NO_HIT_SORT") && !getenv("HRT_NO_SORT") && !tiny && (!warp_mode || getenv("HRT_HIT_SORT_ALWAYS"));
  if (sort_hits) {
    /* the chunk's direction sort below needs less: same pair types, 32 key bits */
    size_t tb = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, rd.qkey, rd.qkey_alt, rd.queue[0], rd.queue_alt, (int)ctx->cap_n, 0, 32, st));
    if (tb > ctx->sort_tmp_bytes) {
      if (ctx->sort_tmp) cudaFree(ctx->sort_tmp);
      ctx->sort_tmp = nullptr; ctx->sort_tmp_bytes = 0;
      CK(cudaMalloc(&ctx->sort_tmp, tb)); ctx->sort_tmp_bytes = tb;
    }
    /* several transmitters: their sorts of one depth run side by side (HRT_SORT_SERIAL=1: one after the other) */
    if (T > 1 && !getenv("HRT_SORT_SERIAL")) {
      if (!ctx->sort_fork) {
        CK(cudaEventCreateWithFlags(&ctx->sort_fork, cudaEventDisableTiming));
        for (int k = 0; k < HRT_SORT_STREAMS; ++k) {
          CK(cudaStreamCreateWithFlags(&ctx->sort_stream[k], cudaStreamNonBlocking));
          CK(cudaEventCreateWithFlags(&ctx->sort_join[k], cudaEventDisableTiming));
        }
      }
      if (tb > ctx->sort_side_bytes) {
        for (int k = 0; k < HRT_SORT_STREAMS; ++k) {
          if (ctx->sort_side_tmp[k]) cudaFree(ctx->sort_side_tmp[k]);
          ctx->sort_side_tmp[k] = nullptr;
        }
        ctx->sort_side_bytes = 0;
        for (int k = 0; k < HRT_SORT_STREAMS; ++k) CK(cudaMalloc(&ctx->sort_side_tmp[k], tb));
        ctx->sort_side_bytes = tb;
      }
    }
  }
  const bool sort_side = sort_hits && T > 1 && ctx->sort_fork && ctx->sort_side_bytes && !getenv("HRT_SORT_SERIAL");
  /* depth pipeline: k_bounce(b + 1) needs k_bounce(b) only, so k_scatter(b) runs beside it on a stream of its own
   * (buffers alternate by depth parity: k_bounce(b) waits for k_scatter(b - 2)).  HRT_NO_OVERLAP=1: one stream */
  const bool overlap = B > 1 && !getenv("HRT_NO_OVERLAP");
  if (overlap && !ctx->scat_ready) {
    CK(cudaEventCreateWithFlags(&ctx->scat_ready, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) {
      CK(cudaStreamCreateWithFlags(&ctx->scat_stream[k], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&ctx->scat_done[k], cudaEventDisableTiming));
    }
  }

  const float host_setup = host_ms();
  CK(cudaEventRecord(ctx->ev[0], st));

  /* ---- line of sight (rank 0 of the shard group only) ---- */
  if (p->los && rank == 0) { rc = run_los(ctx, p, rd, sc, smem, brute, scene_sb, st); if (rc) return rc; }
  CK(cudaEventRecord(ctx->ev[1], st));

  if (flags & HRT_FLAG_SUMMARY) {
    CK(cudaMemsetAsync(rd.pair, 0, R * T * B * sizeof(HrtPairSummary), st));
    CK(cudaMemsetAsync(rd.bounce, 0, T * B * sizeof(HrtBounceSummary), st));
  }

  uint32_t *h_counts = (uint32_t *)malloc((B + 1) * T * 4);
  uint8_t *h_dead = nullptr;
  if (flags & HRT_FLAG_RAYSINFO) h_dead = (uint8_t *)malloc(T * ctx->cap_n);
  if (!h_counts || ((flags & HRT_FLAG_RAYSINFO) && !h_dead)) { free(h_counts); free(h_dead); return fail(ctx, HRT_E_NOMEM, "out of host memory"); }
  float ms_bounce = 0.f, ms_scatter = 0.f;
  /* per-launch timing: three events per (chunk, bounce), resolved after the
   * run -- no synchronisation inside the loop */
  const size_t EV_CAP = 5 * 512;
  size_t ev_used = 0;
  uint8_t tail_dead[8] = {255, 255, 255, 255, 255, 255, 255, 255};  /* TX 1, paths 0..7 */
  rc = HRT_OK;
#define CKR(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(ctx, HRT_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); goto run_done; } } while (0)

  for (uint64_t l0 = 0; l0 < n_shard; l0 += chunk) {
    rd.l0 = l0;
    rd.n = (uint32_t)((n_shard - l0 < chunk) ? n_shard - l0 : chunk);
    const unsigned g1 = (unsigned)min((size_t)sms * 16, ((size_t)rd.n + 255) / 256);

    /* launch directions */
    if (flags & HRT_FLAG_HOST_DIRS) {
      uint64_t L = 0;
      while (L < rd.n) {
        uint64_t run = rd.n - L;
        if (world > 1) { const uint64_t in_blk = (l0 + L) % blk; run = (blk - in_blk < run) ? blk - in_blk : run; }
        const uint64_t g = hrt_gpath(l0 + L, rank, world, blk);
        CKR(cudaMemcpyAsync(rd.dirs + 3 * L, p->dirs + g, run * 12, cudaMemcpyHostToDevice, st));
        L += run;
      }
    } else {
      CKR(cudaMemsetAsync(rd.amb_count, 0, 4, st));
      k_raygen<<<g1, 256, 0, st>>>(rd);
      CKR(cudaGetLastError());
      S.kernel_launches++;
      uint32_t namb = 0;
      CKR(cudaMemcpyAsync(&namb, rd.amb_count, 4, cudaMemcpyDeviceToHost, st));
      CKR(cudaStreamSynchronize(st));
      if (namb) {
        /* recompute the flagged directions with the host libm (see hrt_launch_dir): indices down,
         * values up in ONE transfer through a pinned buffer, written in place by k_patch_dirs */
        if (namb > HRT_AMB_CAP) { rc = fail(ctx, HRT_E_STATE, "ambiguous-direction list overflow (%u)", namb); goto run_done; }
        if (!ctx->h_patch) {
          CKR(cudaHostAlloc((void **)&ctx->h_patch, (size_t)HRT_AMB_CAP * 16, cudaHostAllocDefault));
          CKR(dev_alloc(&ctx->d_patch, (size_t)HRT_AMB_CAP * 4));
        }
        uint32_t *idx = (uint32_t *)ctx->h_patch;              /* [namb] indices, then reused as (index, x, y, z) records */
        CKR(cudaMemcpyAsync(idx, rd.amb_list, namb * 4, cudaMemcpyDeviceToHost, st));
        CKR(cudaStreamSynchronize(st));
        for (uint32_t k = namb; k-- > 0;) {                    /* back to front: record k overwrites indices >= 4k only */
          const uint32_t l = idx[k];
          float d[3];
          hrt_host_launch_dir(hrt_gpath(l0 + l, rank, world, blk), P, d);
          float *rec = ctx->h_patch + 4 * (size_t)k;
          memcpy(rec, &l, 4); rec[1] = d[0]; rec[2] = d[1]; rec[3] = d[2];
        }
        CKR(cudaMemcpyAsync(ctx->d_patch, ctx->h_patch, (size_t)namb * 16, cudaMemcpyHostToDevice, st));
        k_patch_dirs<<<nblk(namb), 256, 0, st>>>(rd, (const float4 *)ctx->d_patch, namb);
        CKR(cudaGetLastError());
        S.kernel_launches++;
        S.ambiguous_dirs += namb;
      }
    }

    /* order the chunk's paths by launch direction */
    if (flags & HRT_FLAG_HOST_DIRS) { k_dirkeys<<<g1, 256, 0, st>>>(rd); CKR(cudaGetLastError()); S.kernel_launches++; }
    if (tiny || getenv("HRT_NO_SORT")) {
      CKR(cudaMemcpyAsync(rd.perm2, rd.perm, (size_t)rd.n * 4, cudaMemcpyDeviceToDevice, st));
    } else {
      size_t tb = 0;
      CKR(cub::DeviceRadixSort::SortPairs(nullptr, tb, rd.dkey, rd.dkey2, rd.perm, rd.perm2, (int)rd.n, 0, 32, st));
      if (tb > ctx->sort_tmp_bytes) {
        if (ctx->sort_tmp) cudaFree(ctx->sort_tmp);
        ctx->sort_tmp = nullptr; ctx->sort_tmp_bytes = 0;
        CKR(cudaMalloc(&ctx->sort_tmp, tb)); ctx->sort_tmp_bytes = tb;
      }
      CKR(cub::DeviceRadixSort::SortPairs(ctx->sort_tmp, tb, rd.dkey, rd.dkey2, rd.perm, rd.perm2, (int)rd.n, 0, 32, st));
      S.kernel_launches += 4;
    }
    CKR(cudaMemsetAsync(rd.qcount, 0, (3 * B + 1) * T * 4, st));
    k_init<<<g1, 256, 0, st>>>(rd);
    CKR(cudaGetLastError());
    S.kernel_launches++;

    for (uint32_t b = 0; b < B; ++b) {
      if (flags & HRT_FLAG_RAYSINFO)
        CKR(cudaMemcpyAsync(rd.rays + (size_t)(b + 1) * T * rd.n_alloc, rd.rays + (size_t)b * T * rd.n_alloc,
                            T * (size_t)rd.n_alloc * sizeof(Ray), cudaMemcpyDeviceToDevice, st));
      /* persistent grids: enough blocks to fill the machine, grid-stride inside */
      const dim3 gb((unsigned)min((size_t)((sms * HRT_MIN_BLOCKS + T - 1) / T), ((size_t)rd.n + HRT_BLOCK - 1) / HRT_BLOCK), (unsigned)T);
      const bool timed = ev_used + 5 <= EV_CAP;
      if (timed) {
        if (ctx->evpool_n < ev_used + 5) {
          cudaEvent_t *np_ = (cudaEvent_t *)realloc(ctx->evpool, (ev_used + 5) * sizeof(cudaEvent_t));
          if (!np_) { rc = fail(ctx, HRT_E_NOMEM, "out of host memory"); goto run_done; }
          ctx->evpool = np_;
          while (ctx->evpool_n < ev_used + 5) CKR(cudaEventCreate(&ctx->evpool[ctx->evpool_n++]));
        }
      }
      /* this k_bounce overwrites the queue and records k_scatter(b - 2) reads */
      if (overlap && b >= 2) CKR(cudaStreamWaitEvent(st, ctx->scat_done[b & 1], 0));
      if (timed) CKR(cudaEventRecord(ctx->evpool[ev_used], st));
      f_bounce<<<gb, HRT_BLOCK, smem ? scene_sb : 0, st>>>(rd, sc, ctx->mats, b);
      CKR(cudaGetLastError());
      if (timed) CKR(cudaEventRecord(ctx->evpool[ev_used + 1], st));
      if (sort_hits) {
        /* order the hits of this depth by position (see hit_key) */
        CKR(cudaMemcpyAsync(h_counts, rd.qcount + (size_t)(b + 1) * T, T * 4, cudaMemcpyDeviceToHost, st));
        CKR(cudaStreamSynchronize(st));
        uint32_t *&qcur = rd.queue[(b + 1) & 1];
        if (sort_side) CKR(cudaEventRecord(ctx->sort_fork, st));
        unsigned side_used = 0;
        for (size_t t = 0; t < T; ++t) {
          const size_t off = t * (size_t)rd.n_alloc;
          if (h_counts[t] == 0) continue;
          const int k = (int)(t % HRT_SORT_STREAMS);
          cudaStream_t ss = sort_side ? ctx->sort_stream[k] : st;
          if (sort_side && !(side_used & (1u << k))) { CKR(cudaStreamWaitEvent(ss, ctx->sort_fork, 0)); side_used |= 1u << k; }
          size_t tb = sort_side ? ctx->sort_side_bytes : ctx->sort_tmp_bytes;
          CKR(cub::DeviceRadixSort::SortPairs(sort_side ? ctx->sort_side_tmp[k] : ctx->sort_tmp, tb, rd.qkey + off, rd.qkey_alt + off,
                                              qcur + off, rd.queue_alt + off, (int)h_counts[t], hit_sort_low_bit, 30, ss));
          S.kernel_launches += 4;
        }
        for (int k = 0; k < HRT_SORT_STREAMS; ++k)
          if (side_used & (1u << k)) {
            CKR(cudaEventRecord(ctx->sort_join[k], ctx->sort_stream[k]));
            CKR(cudaStreamWaitEvent(st, ctx->sort_join[k], 0));
          }
        uint32_t *tmpq = qcur; qcur = rd.queue_alt; rd.queue_alt = tmpq;
      }
      if (timed) CKR(cudaEventRecord(ctx->evpool[ev_used + 2], st));
      cudaStream_t ss = overlap ? ctx->scat_stream[b & 1] : st;
      if (overlap) { CKR(cudaEventRecord(ctx->scat_ready, st)); CKR(cudaStreamWaitEvent(ss, ctx->scat_ready, 0)); }
      if (timed) CKR(cudaEventRecord(ctx->evpool[ev_used + 3], ss));
      const size_t units = warp_mode ? (size_t)rd.n * 32 : rd.n;
      const size_t sblk = use_map ? HRT_MAP_BLOCK : smem ? HRT_BLOCK : HRT_GLOBAL_BLOCK, smin = use_map ? HRT_MAP_MIN_BLOCKS : HRT_MIN_BLOCKS;
      const dim3 gs((unsigned)min((size_t)((sms * smin + T - 1) / T), (units + sblk - 1) / sblk), (unsigned)T);
      f_scatter<<<gs, (unsigned)sblk, scat_sb, ss>>>(rd, sc, ctx->mats, b, smem_rx_ok);
      CKR(cudaGetLastError());
      S.kernel_launches += 2;
      if (timed) { CKR(cudaEventRecord(ctx->evpool[ev_used + 4], ss)); ev_used += 5; }
      if (overlap) CKR(cudaEventRecord(ctx->scat_done[b & 1], ss));
    }
    /* everything after the depth loop (and the next chunk) comes after every k_scatter of this chunk */
    if (overlap) {
      CKR(cudaStreamWaitEvent(st, ctx->scat_done[(B - 1) & 1], 0));
      CKR(cudaStreamWaitEvent(st, ctx->scat_done[B & 1], 0));
    }

    if (flags & HRT_FLAG_SUMMARY) {
      k_fold_counts<<<nblk(T * B, 128), 128, 0, st>>>(rd);
      CKR(cudaGetLastError());
      S.kernel_launches++;
    }
    CKR(cudaMemcpyAsync(h_counts, rd.qcount, (B + 1) * T * 4, cudaMemcpyDeviceToHost, st));

    /* dense outputs -> the caller's arrays (columns of this chunk) */
    std::vector<CopyTile> tiles;
    if (flags & HRT_FLAG_DENSE) {
      ChannelInfo *O = p->scat;
      float *dst[6] = { O->a_te_re, O->a_te_im, O->a_tm_re, O->a_tm_im, O->tau, O->freq_shift };
      if (flags & HRT_FLAG_DENSE_C64) {
        CKR(d2h_columns(ctx, tiles, p->scat_a_te_c64, rd.out_f[0], 8, R * T * B, rd));
        CKR(d2h_columns(ctx, tiles, p->scat_a_tm_c64, rd.out_f[2], 8, R * T * B, rd));
        dst[0] = dst[1] = dst[2] = dst[3] = nullptr;
      }
      for (int k = 0; k < 6; ++k) CKR(d2h_columns(ctx, tiles, dst[k], rd.out_f[k], 4, R * T * B, rd));
      CKR(d2h_columns(ctx, tiles, O->directions_rx, rd.out_dir, 12, R * T * B, rd));
    }
    if (flags & HRT_FLAG_TRACE) {
      CKR(d2h_columns(ctx, tiles, p->trace_hit_tri, rd.tr_hit, 4, T * B, rd));
      CKR(d2h_columns(ctx, tiles, p->trace_hit_t, rd.tr_t, 4, T * B, rd));
      CKR(d2h_columns(ctx, tiles, p->trace_slot_state, rd.tr_state, 1, R * T * B, rd));
    }
    if (flags & HRT_FLAG_RAYSINFO) {
      /* reference row order (:589, :732-743): row 0 = initial rays of TX 0,
       * row t*B+b+1 = rays of TX t after bounce b */
      RaysInfo *RI = p->rays_scat;
      if (RI->rays) {
        CKR(d2h_columns(ctx, tiles, RI->rays, rd.rays, sizeof(Ray), 1, rd));
        for (size_t t = 0; t < T; ++t)
          for (size_t b = 0; b < B; ++b)
            CKR(d2h_columns(ctx, tiles, RI->rays + (t * B + b + 1) * P, rd.rays + ((b + 1) * T + t) * (size_t)rd.n_alloc,
                            sizeof(Ray), 1, rd));
      }
      CKR(cudaMemcpyAsync(h_dead, rd.dead_at, T * (size_t)rd.n_alloc, cudaMemcpyDeviceToHost, st));
    }
    CKR(run_copy_tiles(ctx, st, tiles));
    CKR(cudaStreamSynchronize(st));
    if ((flags & HRT_FLAG_RAYSINFO) && T > 1 && l0 == 0 && rank == 0)
      for (uint32_t j = 0; j < 8 && j < rd.n; ++j) tail_dead[j] = h_dead[rd.n_alloc + j];

    for (size_t b = 0; b < B; ++b)
      for (size_t t = 0; t < T; ++t) { S.ray_bounces += h_counts[b * T + t]; S.primary_hits += h_counts[(b + 1) * T + t]; }

    if ((flags & HRT_FLAG_RAYSINFO) && p->rays_scat->rays_active) {
      /* activity masks.  The reference copies the first P/8+1 bytes of its
       * global bitmask (bit index tx*P+path) into every row: TX 0's bits, plus
       * for T > 1 the first bits of TX 1 in the last byte (appendix A-9). */
      uint8_t *A = p->rays_scat->rays_active;
      const size_t rowb = P / 8 + 1;
      /* big chunks: the copy threads share the loop, each a range of whole bytes (chunks and shard blocks start at
       * multiples of 32 paths, so no byte is written by two of them) */
      const bool par = rd.n >= (1u << 18) && ctx->pool;
      const int parts = par ? ctx->pool->n : 1;
      auto mask_range = [&](int w) {
      const uint64_t L_end = w + 1 == parts ? rd.n : ((uint64_t)rd.n * (uint64_t)(w + 1) / (uint64_t)parts) & ~(uint64_t)31;
      uint64_t L = ((uint64_t)rd.n * (uint64_t)w / (uint64_t)parts) & ~(uint64_t)31;
      while (L < L_end) {
        const uint64_t g = hrt_gpath(l0 + L, rank, world, blk);
        /* whole bytes: 8 consecutive paths of one shard block (blocks are multiples of 32) */
        const bool whole = (g & 7) == 0 && L + 8 <= rd.n && (world <= 1 || (l0 + L) % blk + 8 <= blk);
        if (whole) {
          for (size_t b = 0; b < B; ++b) {
            uint8_t byte = 0;
            for (int i = 0; i < 8; ++i) byte |= (uint8_t)((h_dead[L + i] > b) << i);
            for (size_t t = 0; t < T; ++t) A[(t * B + b + 1) * rowb + (g >> 3)] = byte;
          }
          L += 8;
          continue;
        }
        const uint8_t dead0 = h_dead[L];
        for (size_t t = 0; t < T; ++t)
          for (size_t b = 0; b < B; ++b) {
            uint8_t *row = A + (t * B + b + 1) * rowb;
            const uint8_t bit = (uint8_t)(1u << (g & 7));
            if (dead0 > b) row[g >> 3] |= bit; else row[g >> 3] &= (uint8_t)~bit;
          }
        ++L;
      }
      };
      if (par) ctx->pool->run([&](int w, int) { mask_range(w); });
      else mask_range(0);
    }
  }
  S.shadow_queries = S.primary_hits * R;

  memcpy(ctx->tail_dead, tail_dead, 8);
  /* (a multi-device run sets the tail bits once all devices are done: hrt_internal_raysinfo_tail) */
  if ((flags & HRT_FLAG_RAYSINFO) && p->rays_scat->rays_active && rank == 0 && !(flags & HRT_FLAG_INTERNAL_NO_TAIL)) raysinfo_tail_bits(p, tail_dead);

  if (flags & HRT_FLAG_SUMMARY) { rc = flush_summaries(ctx, p, rd, flags, st); if (rc) goto run_done; }
  if (flags & HRT_FLAG_PATHLIST) { rc = flush_path_list(ctx, p, rd, flags, st); if (rc) goto run_done; }
  if (flags & HRT_FLAG_CIR) { rc = flush_cir(ctx, p, rd, rank, st); if (rc) goto run_done; }
  if (flags & HRT_FLAG_EXT_REFRACT) {
    unsigned long long found = 0;
    CKR(cudaMemcpyAsync(&found, rd.counters + 13, 8, cudaMemcpyDeviceToHost, st));
    CKR(cudaStreamSynchronize(st));
    const unsigned long long kept = found < p->refr_capacity ? found : p->refr_capacity;
    CKR(cudaMemcpyAsync(p->refr_rays, ctx->d_refr, (size_t)kept * sizeof(HrtRefractRecord), cudaMemcpyDeviceToHost, st));
    CKR(cudaStreamSynchronize(st));
    *p->refr_count = found;
  }

  if (count) {
    unsigned long long hc[16];
    CKR(cudaMemcpyAsync(hc, rd.counters, sizeof hc, cudaMemcpyDeviceToHost, st));
    CKR(cudaStreamSynchronize(st));
    for (int k = 0; k < 5; ++k) { S.work_bounce[k] = hc[k]; S.work_scatter[k] = hc[5 + k]; }
  }
  CKR(cudaEventRecord(ctx->ev[2], st));
  CKR(cudaEventSynchronize(ctx->ev[2]));
  {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[2]);
    cudaEventElapsedTime(&b, ctx->ev[0], ctx->ev[1]);
    float ms_sort = 0.f;
    /* with the depth pipeline a k_bounce / sort interval may contain the ragged end of the previous k_scatter,
     * and a k_scatter interval the start-up of the next one: the three sums can exceed ms_total slightly */
    for (size_t e = 0; e + 5 <= ev_used; e += 5) {
      float x = 0.f, y = 0.f, z = 0.f;
      cudaEventElapsedTime(&x, ctx->evpool[e], ctx->evpool[e + 1]);
      cudaEventElapsedTime(&z, ctx->evpool[e + 1], ctx->evpool[e + 2]);
      cudaEventElapsedTime(&y, ctx->evpool[e + 3], ctx->evpool[e + 4]);
      ms_bounce += x; ms_scatter += y; ms_sort += z;
    }
    S.ms_sort = ms_sort;
    S.n_bounce_launches = S.n_scatter_launches = (uint32_t)(ev_used / 5);
    S.ms_total = a; S.ms_bounce = ms_bounce; S.ms_scatter = ms_scatter;
    S.ms_other = b;
    S.host_ms_setup = host_setup; S.host_ms_total = host_ms();
  }
run_done:
#undef CKR
  free(h_counts); free(h_dead);
  return rc;
}
