/* hrt_cuda.h -- thin C-ABI over the CUDA (sm_100a) implementation.
 *
 * Plain C: opaque context, pointers, sizes and int status codes only.  The
 * host-side compute_paths() (hermespy-rt_b200/csrc/compute_paths.c) is written
 * against this interface; so are bench.py, the tests (through ctypes) and any
 * FFI binding.
 *
 * Which reference code each entry point stands in for:
 *   hrt_scene_upload ...... the scene walk of moeller_trumbore + precompute_normals
 *                           (src/compute_paths.c:208-224, :253-258), now a
 *                           flattened SoA triangle buffer + GPU-built BVH
 *   hrt_materials_set ..... g_materials_precomputed (src/compute_paths.c:169-206)
 *   hrt_run ............... compute_paths() proper (src/compute_paths.c:419-757)
 *                           for one shard of the (tx, path) space
 *   hrt_closest_hits ...... moeller_trumbore() (src/compute_paths.c:237-287)
 *                           for a batch of rays (unit-test / validation entry)
 *   hrt_scene_advance ..... no reference counterpart: moves the meshes by their
 *                           Mesh.velocity (inc/scene.h:21-22) and refits the BVH
 * Output modes of hrt_run beyond the reference's dense arrays (HRT_FLAG_*):
 * per-(rx, tx, bounce) summaries, the delay-binned impulse response, and the
 * compact list of valid paths -- see HrtRunParams.
 *
 * Every function returns HRT_OK (0) or a negative HRT_E_* code;
 * hrt_last_error() gives the message.  There is no CPU fallback anywhere:
 * without a CUDA device hrt_ctx_create() fails with HRT_E_NO_DEVICE.
 */
#ifndef HRT_CUDA_H
#define HRT_CUDA_H

#include "hermespy_rt.h"

#ifdef __cplusplus
extern "C" {
#endif

#define HRT_OK            0
#define HRT_E_NO_DEVICE  -1   /* no usable CUDA device / driver            */
#define HRT_E_CUDA       -2   /* a CUDA call failed (see hrt_last_error)   */
#define HRT_E_ARG        -3   /* invalid argument / scene                  */
#define HRT_E_NOMEM      -4   /* host or device allocation failed          */
#define HRT_E_STATE      -5   /* call order (e.g. run before scene upload) */

typedef struct hrt_ctx hrt_ctx;

/* Derived per-material constants (reference MaterialPrecomputed,
 * src/compute_paths.c:125-132) in the layout the kernels consume.  Filled on
 * the host by hrt_materials_derive() with the host libm, like the reference. */
typedef struct {
  float eta_abs2, eta_abs_inv_sqrt;
  float sqrt_re, sqrt_im;
  float inv_re, inv_im;
  float r, s, s1_alpha;
  float pad_[3];
} HrtMaterialDerived;

/* hrt_run flags */
#define HRT_FLAG_DENSE        0x01u  /* fill dense ChannelInfo arrays (reference layout) */
#define HRT_FLAG_RAYSINFO     0x02u  /* also export RaysInfo rows (needs DENSE)          */
#define HRT_FLAG_SUMMARY      0x04u  /* accumulate HrtPairSummary / HrtBounceSummary     */
#define HRT_FLAG_TRACE        0x08u  /* export hit triangle ids / slot states (tests)    */
#define HRT_FLAG_BRUTE_FORCE  0x10u  /* skip the BVH: test every triangle (validation)   */
#define HRT_FLAG_HOST_DIRS    0x20u  /* launch directions supplied by caller (dirs)      */
#define HRT_FLAG_SUMMARY_DEV  0x40u  /* summary pointers are DEVICE memory               */
#define HRT_FLAG_COUNT        0x80u  /* instrumented kernels: count box/triangle tests   */
#define HRT_FLAG_CIR         0x100u  /* accumulate the delay-binned impulse response    */
#define HRT_FLAG_PATHLIST    0x200u  /* emit the valid scatter paths as a compact list  */
#define HRT_FLAG_DENSE_C64   0x800u  /* with DENSE: scatter gains as interleaved complex64 (scat_a_te_c64 /
                                       scat_a_tm_c64) instead of the four re / im arrays of `scat` */
/* Opt-in extensions the reference has as TODOs (no reference behaviour; definitions in
 * csrc/hrt_ext.cuh, checked against the oracle's double-precision statement): */
#define HRT_FLAG_EXT_LOBES  0x1000u  /* scatter gains from the three-lobe pattern (Material.s1/s2/s3, s1_alpha,
                                        s3_alpha; reference TODO src/compute_paths.c:414) */
#define HRT_FLAG_EXT_REFRACT 0x2000u /* list the refraction ray of every hit with ITU-R P.2040-3 (31c)/(31d)
                                        gains (reference TODO :587, :726-728): refr_rays / refr_capacity / refr_count */
#define HRT_FLAG_PATHLIST_DEV 0x400u /* ... into DEVICE memory (`paths`), e.g. a buffer
                                        that NCCL gathers next; paths_count stays host */

/* Order-independent per-(rx, tx, bounce) reduction of the scatter paths.
 * Integer fields are exact and comparable bit for bit with a CPU run. */
typedef struct {
  uint64_t n_valid;      /* paths written with a gain (reference :692-722)        */
  uint64_t n_occluded;   /* slots zeroed by the 1 m occlusion rule (:683-691)     */
  uint64_t hit_hash;     /* sum of mix64(path << 32 | primary triangle id), valid */
  uint64_t tau_bits;     /* sum of the fp32 bit patterns of tau, valid paths      */
  double   power_te;     /* sum |a_te|^2, valid paths                             */
  double   power_tm;     /* sum |a_tm|^2                                          */
} HrtPairSummary;

/* Per-(tx, bounce) counters of the primary rays. */
typedef struct {
  uint64_t n_traced;     /* rays alive at bounce start = primary queries (:615)  */
  uint64_t n_hit;        /* of which hit something                                */
  uint64_t hit_hash;     /* sum of mix64(path << 32 | triangle id) over hits      */
  uint64_t t_bits;       /* sum of the fp32 bit patterns of the hit distances     */
} HrtBounceSummary;

/* One valid scatter path (HRT_FLAG_PATHLIST): the words the reference writes
 * into slot ((rx * num_tx + tx) * num_bounces + bounce) * num_paths + path of
 * its dense ChannelInfo arrays (src/compute_paths.c:698-722), without the
 * slots of dead rays and occluded receivers.  48 bytes. */
typedef struct {
  uint32_t path;         /* ray index within its transmitter                      */
  uint32_t rx;
  uint16_t tx, bounce;
  float a_te_re, a_te_im, a_tm_re, a_tm_im;
  float tau, freq_shift;
  Vec3  direction_rx;
} HrtPathRecord;

/* One refraction ray (HRT_FLAG_EXT_REFRACT): spawned where path `path` of
 * transmitter `tx` hits a surface at bounce `bounce`; origin 1e-4 m inside the
 * surface along the refracted direction (Snell, n = Re sqrt(eta)); gains = the
 * ray's gains before this bounce times T_TE / T_TM of eqs. (31c)/(31d), divided
 * by the free-space factor of the segment like the reflected ray's (:627-634). */
typedef struct {
  uint32_t path; uint16_t tx, bounce;
  float    o[3], d[3];
  float    t_te_re, t_te_im, t_tm_re, t_tm_im;
} HrtRefractRecord;


typedef struct {
  /* problem (reference compute_paths arguments, inc/compute_paths.h:59-74) */
  size_t num_rx, num_tx, num_paths, num_bounces;
  float  carrier_frequency_GHz;
  const Vec3 *rx_pos, *tx_pos, *rx_vel, *tx_vel;      /* host memory */

  /* shard: paths are dealt in blocks of shard_block to shard_world ranks;
   * this call processes the blocks b with b % shard_world == shard_rank.
   * shard_world = 1 (or 0) means the whole path range. */
  uint32_t shard_rank, shard_world;
  size_t   shard_block;

  uint32_t flags;

  /* HRT_FLAG_DENSE: caller-owned host arrays in the reference's layout.
   * Only the columns of this shard's paths are written. */
  ChannelInfo *los;        /* may be NULL: LoS skipped                      */
  RaysInfo    *rays_los;   /* may be NULL                                   */
  ChannelInfo *scat;
  RaysInfo    *rays_scat;  /* used with HRT_FLAG_RAYSINFO                   */

  /* HRT_FLAG_SUMMARY: [num_rx][num_tx][num_bounces] and [num_tx][num_bounces];
   * ADDED to (caller zeroes them).  Host memory unless HRT_FLAG_SUMMARY_DEV. */
  HrtPairSummary   *pair_summary;
  HrtBounceSummary *bounce_summary;

  /* HRT_FLAG_TRACE (host, full-size arrays, shard columns written):
   * hit_tri [T][B][P] u32 (HRT id / 0xFFFFFFFF miss / 0xFFFFFFFE idle),
   * hit_t [T][B][P] f32, slot_state [R][T][B][P] u8 (0/1 path/2 occluded) */
  uint32_t *trace_hit_tri;
  float    *trace_hit_t;
  uint8_t  *trace_slot_state;

  /* HRT_FLAG_HOST_DIRS: [num_paths] launch directions (host) */
  const Vec3 *dirs;

  /* CUDA stream to run on (cudaStream_t), NULL = the context's own stream */
  void *stream;

  /* HRT_FLAG_CIR: channel impulse response per (rx, tx), the reduction a
   * consumer of ChannelInfo performs next (sum of a * delta(t - tau) over all
   * paths and bounces), formed on the GPU so that C4/C5-sized runs need no
   * per-path output: cir[((rx * num_tx + tx) * cir_bins + bin) * 4 + k],
   * k = a_te_re, a_te_im, a_tm_re, a_tm_im summed over the valid scatter paths
   * (and the LoS path when `los` is given) with
   * bin = floor((tau - cir_tau0_s) / cir_dt_s) in [0, cir_bins).  Host memory,
   * ADDED to (caller zeroes).  fp32 accumulation in arbitrary order. */
  float   *cir;
  float    cir_tau0_s, cir_dt_s;
  uint32_t cir_bins;

  /* HRT_FLAG_PATHLIST: the valid scatter paths of this call as records, in no
   * particular order, compacted on the GPU (warp ballot + prefix sum, one
   * atomic per warp).  `paths` is host memory for `paths_capacity` records;
   * *paths_count receives the number of valid paths found -- when it exceeds the
   * capacity only `paths_capacity` of them (an arbitrary subset) were stored.
   * Needs num_paths < 2^32. */
  HrtPathRecord *paths;
  uint64_t       paths_capacity;
  uint64_t      *paths_count;

  /* HRT_FLAG_DENSE_C64: [num_rx][num_tx][num_bounces][num_paths] complex64 each, as
   * (re, im) float pairs -- what numpy / the reference's Python ChannelInfo hold
   * (compute_paths_pybind11.cpp:22-42 repacks re / im arrays into these on the
   * host; here the device writes them directly).  scat->a_*_re / _im are then
   * not touched and may be NULL. */
  float         *scat_a_te_c64, *scat_a_tm_c64;

  /* HRT_FLAG_EXT_REFRACT: host buffer of refr_capacity records, filled in no
   * particular order; *refr_count = number spawned (may exceed the capacity) */
  HrtRefractRecord *refr_rays;
  uint64_t          refr_capacity;
  uint64_t         *refr_count;
} HrtRunParams;

/* Counters and timings of the last hrt_run on a context. */
typedef struct {
  uint64_t ray_bounces;        /* primary closest-hit queries                     */
  uint64_t primary_hits;
  uint64_t shadow_queries;     /* per-(hit, rx) closest-hit queries               */
  uint64_t los_queries;
  uint64_t ambiguous_dirs;     /* launch directions recomputed on the host        */
  uint64_t kernel_launches;    /* kernels of this library launched by the run     */
  float    ms_total;           /* CUDA-event time of the whole run on its stream  */
  float    ms_bounce;          /* sum over k_bounce launches.  By default k_scatter of
                                  depth b runs beside k_bounce / the sort of depth b+1
                                  (depth pipeline): the event interval of a k_bounce
                                  then includes its wait for SM slots; HRT_NO_OVERLAP=1
                                  puts every kernel on one stream (per-kernel times)   */
  float    ms_scatter;         /* sum over k_scatter launches (dominant kernel)   */
  float    ms_other;
  uint32_t n_bounce_launches;  /* launches behind ms_bounce / ms_scatter           */
  uint32_t n_scatter_launches;
  uint32_t num_tris, num_nodes, scene_in_smem;
  float    box_pad;
  /* HRT_FLAG_COUNT runs only: tests performed by k_bounce / k_scatter:
   * [0] ray-box tests, [1..4] Moeller-Trumbore tests reaching stage A (det),
   * B (u), C (v), D (t) -- SURVEY section 8d's unit of algorithmic work */
  uint64_t work_bounce[5], work_scatter[5];
  /* the scene's BVH: 1 = binned-SAH builder, 0 = Morton/Karras (HRT_BVH_LBVH=1) */
  uint32_t bvh_sah, bvh_levels;
  float    bvh_build_ms;       /* GPU time of the last build */
  float    ms_sort;            /* hit-queue ordering, part of ms_total */
  uint64_t cir_dropped;        /* HRT_FLAG_CIR: valid paths whose delay fell outside the window */
  /* shadow queries went through receiver maps (csrc/hrt_rxmap.cuh) instead of the BVH:
   * 1/0, cells per cube-map face edge, GPU time of the map build (0 when the cached maps applied) */
  uint32_t rx_map, rx_map_cells;
  float    rx_map_build_ms;
  /* host wall clock of the call: entry -> first kernel of the run queued (argument checks,
   * position upload, buffer and map management), and entry -> return */
  float    host_ms_setup, host_ms_total;
} HrtRunStats;

int  hrt_device_count(void);
int  hrt_ctx_create(int device, hrt_ctx **out);
void hrt_ctx_destroy(hrt_ctx *ctx);
const char *hrt_last_error(const hrt_ctx *ctx);   /* ctx may be NULL: creation errors */

/* Flatten + upload the scene and build the BVH on the GPU.  If normals_out is
 * not NULL it receives one Vec3 per triangle in (mesh, face) order -- the
 * values the reference stores in Mesh.ns (src/compute_paths.c:208-224). */
int hrt_scene_upload(hrt_ctx *ctx, const Scene *scene, Vec3 *normals_out);

/* Moving scenes (Mesh.velocity, reference inc/scene.h:21-22; SURVEY section 8
 * row f3): advance every mesh of the uploaded scene by velocity * dt_s on the
 * GPU (v <- v + vel * dt, each operation rounded to fp32) and bring the BVH up
 * to date -- in place (same topology, boxes recomputed bottom-up: rebuild = 0)
 * or with a fresh build from the moved triangles (rebuild = 1).  Results are
 * those of loading a scene file with the moved vertices. */
int hrt_scene_advance(hrt_ctx *ctx, float dt_s, int rebuild);

/* Host-side: derive the constants of material `index` at f GHz (reference
 * precompute_materials, src/compute_paths.c:171-206). */
void hrt_materials_derive(uint32_t index, float carrier_frequency_GHz, HrtMaterialDerived *out);
int  hrt_materials_set(hrt_ctx *ctx, const HrtMaterialDerived table[NUM_G_MATERIALS]);

int hrt_run(hrt_ctx *ctx, const HrtRunParams *p);

/* One-call form of the path-list mode on the implicit context of compute_paths()
 * (hermespy-rt_b200/csrc/compute_paths.c): compute_paths' arguments, then the
 * caller's record buffer.  Returns the number of valid scatter paths found. */
/* compute_paths() with the scatter gains delivered as interleaved complex64
 * arrays (a_te_c64 / a_tm_c64: [num_rx][num_tx][num_bounces * num_rays] pairs of
 * floats): the layout of the reference's Python ChannelInfo.a_te / a_tm, written
 * by the device instead of being repacked on the host
 * (compute_paths_pybind11.cpp:22-42).  chanInfo_scat's a_*_re / a_*_im members are
 * ignored; everything else is as in compute_paths().  raysInfo_* may be NULL. */
void compute_paths_c64(
    Scene *scene, Vec3 *rx_pos, Vec3 *tx_pos, Vec3 *rx_vel, Vec3 *tx_vel,
    float carrier_frequency_GHz,
    size_t num_rx, size_t num_tx, size_t num_rays, size_t num_bounces,
    ChannelInfo *chanInfo_los, RaysInfo *raysInfo_los,
    ChannelInfo *chanInfo_scat, RaysInfo *raysInfo_scat,
    float *a_te_c64, float *a_tm_c64);

size_t compute_path_list(
    Scene *scene, Vec3 *rx_pos, Vec3 *tx_pos, Vec3 *rx_vel, Vec3 *tx_vel,
    float carrier_frequency_GHz,
    size_t num_rx, size_t num_tx, size_t num_rays, size_t num_bounces,
    HrtPathRecord *paths, size_t capacity);

/* The shard partition as pure host arithmetic (no GPU): how many of the
 * num_paths paths rank `rank` of `world` owns, and the global path index of its
 * local index L.  Every path belongs to exactly one rank. */
uint64_t hrt_shard_count(uint64_t num_paths, uint32_t rank, uint32_t world, uint64_t block);
uint64_t hrt_shard_path(uint64_t local_index, uint32_t rank, uint32_t world, uint64_t block);
int hrt_get_stats(const hrt_ctx *ctx, HrtRunStats *out);

/* Measures this GPU's sustained fp32 rate with separately rounded FMUL/FADD
 * (the arithmetic of the exact intersection code) and with FFMA, in Tflop/s.
 * Denominator of the intersection roofline (no fp32 figure exists in
 * MEASURED_PEAKS.json). */
int hrt_fp32_peak(hrt_ctx *ctx, float *tflops_unfused, float *tflops_fma);

/* Batch closest hit (host arrays): tri = id in (mesh, face) order or
 * 0xFFFFFFFF, t = distance or -1, theta = folded incidence angle or 0. */
int hrt_closest_hits(hrt_ctx *ctx, const Ray *rays, size_t n, uint32_t flags,
                     uint32_t *tri, float *t, float *theta);

/* ------------------------------------------------------------------------
 * Several GPUs of one box behind one call.
 * The job is sharded by ray inside the library: device i of n traces the
 * 65,536-path blocks b = i (mod n) of every transmitter's path range on its
 * own replica of the scene and BVH, driven by its own host thread; results are
 * those of hrt_run on one device.  Replaces nothing in the reference (which is
 * single-threaded): it is how compute_paths() uses a whole 8 x B200 box
 * (HRT_DEVICES=0,1,...).
 */
typedef struct hrt_multi hrt_multi;

/* devices == NULL: devices 0 .. n-1; n <= 0: every device of the box */
int  hrt_multi_create(const int *devices, int n, hrt_multi **out);
void hrt_multi_destroy(hrt_multi *m);
const char *hrt_multi_last_error(const hrt_multi *m);     /* m may be NULL: creation errors */
int  hrt_multi_num_devices(const hrt_multi *m);
hrt_ctx *hrt_multi_ctx(hrt_multi *m, int i);              /* the i-th device's context (borrowed) */
int  hrt_multi_scene_upload(hrt_multi *m, const Scene *scene, Vec3 *normals_out);   /* replicated, built in parallel */
int  hrt_multi_scene_advance(hrt_multi *m, float dt_s, int rebuild);
int  hrt_multi_materials_set(hrt_multi *m, const HrtMaterialDerived table[NUM_G_MATERIALS]);
int  hrt_multi_get_stats(const hrt_multi *m, HrtRunStats *out);   /* counters summed, times = slowest device */

/* One job over all devices, HOST results (every hrt_run flag except the *_DEV
 * ones; shard_* must be unset).  Dense arrays: each device writes its own
 * disjoint columns -- no collective.  Summaries / impulse response: ADDED to,
 * as with hrt_run.  Path list: stored up to its capacity, *paths_count = found. */
int  hrt_multi_run(hrt_multi *m, const HrtRunParams *p);

/* One job over all devices, DEVICE-RESIDENT results on EVERY device, exchanged
 * with ncclAllGather over NVLink (ncclCommInitAll, one communicator per device
 * in this process; NCCL is bound at run time, HRT_E_STATE if it is absent):
 *   pair_dev[i] / bounce_dev[i]  device memory on device i ([R][T][B] / [T][B]):
 *                                overwritten with the whole job's tables
 *   paths_dev[i] (or NULL)       device memory on device i, n * paths_capacity_each
 *                                records: segment j = device j's valid paths,
 *                                counts[j] of them (host array [n])
 * Summary and path-list results only. */
int  hrt_multi_run_gathered(hrt_multi *m, const HrtRunParams *p,
                            HrtPairSummary *const *pair_dev, HrtBounceSummary *const *bounce_dev,
                            HrtPathRecord *const *paths_dev, uint64_t paths_capacity_each, uint64_t *counts);
int  hrt_multi_nccl_version(const hrt_multi *m);          /* 0 until the gathered entry has run */

#ifdef __cplusplus
}
#endif
#endif /* HRT_CUDA_H */
