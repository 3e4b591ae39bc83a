/* hrt_multi.cu -- several GPUs of one box behind ONE C call (SURVEY section 8
 * rows (b) "hrt_ctx_create(devices[])" and (e)).
 *
 * The path shards by ray: device i of n traces the 65,536-path blocks
 * b = i (mod n) of every transmitter's path range (hrt_run's shard_rank /
 * shard_world), on its own replica of the scene and BVH, driven by its own
 * host thread.  There is no exchange while tracing.
 *   - dense outputs: every device writes its own disjoint columns of the
 *     caller's host arrays -- no collective at all;
 *   - host summaries / impulse responses / path lists: per-device results are
 *     combined on the host after the threads join;
 *   - device-resident consumers (hrt_multi_run_gathered): the per-device
 *     [R][T][B] summary tables and the compact path lists are exchanged with
 *     ncclAllGather over NVLink (one communicator per device, created with
 *     ncclCommInitAll in this single process), so that EVERY GPU ends up with
 *     the whole job's tables and every path record.
 *
 * NCCL is bound at run time (dlopen "libnccl.so.2"): the library itself has no
 * link-time dependency on it, and inside a process that already carries an
 * NCCL (e.g. PyTorch's) that same copy is used.  Only the gathered entry needs
 * it; everything else works without.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "../../include/hrt_cuda.h"

/* internal hooks of hrt_cuda.cu */
extern "C" void hrt_internal_raysinfo_tail(const hrt_ctx *rank0, const HrtRunParams *p);
#define HRT_FLAG_INTERNAL_NO_TAIL 0x80000000u

struct NcclApi {
  void *lib;
  const char *(*GetErrorString)(ncclResult_t);
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)(void);
  ncclResult_t (*GroupEnd)(void);
  ncclResult_t (*GetVersion)(int *);
};

static bool nccl_bind(NcclApi &a, char *err, size_t errn)
{
  memset(&a, 0, sizeof a);
  const char *names[] = { "libnccl.so.2", "libnccl.so" };
  for (const char *n : names) { a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.lib) break; }
  if (!a.lib) { snprintf(err, errn, "NCCL not found (dlopen libnccl.so.2: %s)", dlerror()); return false; }
#define BIND(field, sym) do { *(void **)&a.field = dlsym(a.lib, sym); if (!a.field) { snprintf(err, errn, "NCCL symbol %s missing", sym); return false; } } while (0)
  BIND(GetErrorString, "ncclGetErrorString"); BIND(CommInitAll, "ncclCommInitAll"); BIND(CommDestroy, "ncclCommDestroy");
  BIND(AllGather, "ncclAllGather"); BIND(GroupStart, "ncclGroupStart"); BIND(GroupEnd, "ncclGroupEnd"); BIND(GetVersion, "ncclGetVersion");
#undef BIND
  return true;
}

struct hrt_multi {
  int n;
  std::vector<int> devices;
  std::vector<hrt_ctx *> ctx;
  char err[512];
  /* NCCL, created on first use */
  bool nccl_tried, nccl_ok;
  NcclApi nccl;
  std::vector<ncclComm_t> comm;
  std::vector<cudaStream_t> stream;
  int nccl_version;
  /* per-device scratch of the gathered entry */
  std::vector<void *> d_pair, d_bounce;
  std::vector<void *> d_pair_all, d_bounce_all;
  size_t cap_tables;
};

static char g_multi_error[512] = "";

static int mfail(hrt_multi *m, int code, const char *fmt, ...)
{
  va_list ap; va_start(ap, fmt);
  vsnprintf(m ? m->err : g_multi_error, 512, fmt, ap);
  va_end(ap);
  return code;
}

extern "C" const char *hrt_multi_last_error(const hrt_multi *m) { return m ? m->err : g_multi_error; }
extern "C" int hrt_multi_num_devices(const hrt_multi *m) { return m ? m->n : 0; }
extern "C" hrt_ctx *hrt_multi_ctx(hrt_multi *m, int i) { return (m && i >= 0 && i < m->n) ? m->ctx[i] : nullptr; }
extern "C" int hrt_multi_nccl_version(const hrt_multi *m) { return (m && m->nccl_ok) ? m->nccl_version : 0; }

extern "C" int hrt_multi_create(const int *devices, int n, hrt_multi **out)
{
  if (!out) return HRT_E_ARG;
  *out = nullptr;
  const int avail = hrt_device_count();
  if (avail <= 0) return mfail(nullptr, HRT_E_NO_DEVICE, "no CUDA device available; this library has no CPU path");
  if (n <= 0) n = avail;                                   /* all devices of the box */
  hrt_multi *m = new hrt_multi();
  m->n = n; m->nccl_tried = m->nccl_ok = false; m->cap_tables = 0; m->nccl_version = 0; m->err[0] = 0;
  for (int i = 0; i < n; ++i) {
    const int dev = devices ? devices[i] : i;
    for (int j = 0; j < i; ++j)
      if (m->devices[j] == dev) { mfail(nullptr, HRT_E_ARG, "device %d listed twice", dev); goto fail; }
    hrt_ctx *c = nullptr;
    if (hrt_ctx_create(dev, &c) != HRT_OK) { mfail(nullptr, HRT_E_CUDA, "device %d: %s", dev, hrt_last_error(nullptr)); goto fail; }
    m->devices.push_back(dev); m->ctx.push_back(c);
  }
  *out = m;
  return HRT_OK;
fail:
  for (hrt_ctx *c : m->ctx) hrt_ctx_destroy(c);
  delete m;
  return HRT_E_ARG;
}

static void free_tables(hrt_multi *m)
{
  for (size_t i = 0; i < m->d_pair.size(); ++i) {
    cudaSetDevice(m->devices[i]);
    cudaFree(m->d_pair[i]); cudaFree(m->d_bounce[i]);
    cudaFree(m->d_pair_all[i]); cudaFree(m->d_bounce_all[i]);
  }
  m->d_pair.clear(); m->d_bounce.clear();
  m->d_pair_all.clear(); m->d_bounce_all.clear();
  m->cap_tables = 0;
}

extern "C" void hrt_multi_destroy(hrt_multi *m)
{
  if (!m) return;
  free_tables(m);
  if (m->nccl_ok) {
    for (int i = 0; i < m->n; ++i) { cudaSetDevice(m->devices[i]); m->nccl.CommDestroy(m->comm[i]); cudaStreamDestroy(m->stream[i]); }
  }
  for (hrt_ctx *c : m->ctx) hrt_ctx_destroy(c);
  delete m;
}

/* runs f(i) for every device on its own host thread; returns the first failure */
template <class F> static int for_each_device(hrt_multi *m, F f)
{
  std::vector<int> rc(m->n, HRT_OK);
  if (m->n == 1) { rc[0] = f(0); }
  else {
    std::vector<std::thread> th;
    for (int i = 0; i < m->n; ++i) th.emplace_back([&, i] { rc[i] = f(i); });
    for (auto &t : th) t.join();
  }
  for (int i = 0; i < m->n; ++i)
    if (rc[i] != HRT_OK) return mfail(m, rc[i], "device %d: %s", m->devices[i], hrt_last_error(m->ctx[i]));
  return HRT_OK;
}

extern "C" int hrt_multi_scene_upload(hrt_multi *m, const Scene *scene, Vec3 *normals_out)
{
  if (!m) return HRT_E_ARG;
  return for_each_device(m, [&](int i) { return hrt_scene_upload(m->ctx[i], scene, i == 0 ? normals_out : nullptr); });
}

extern "C" int hrt_multi_scene_advance(hrt_multi *m, float dt_s, int rebuild)
{
  if (!m) return HRT_E_ARG;
  return for_each_device(m, [&](int i) { return hrt_scene_advance(m->ctx[i], dt_s, rebuild); });
}

extern "C" int hrt_multi_materials_set(hrt_multi *m, const HrtMaterialDerived table[NUM_G_MATERIALS])
{
  if (!m) return HRT_E_ARG;
  for (int i = 0; i < m->n; ++i) { const int rc = hrt_materials_set(m->ctx[i], table); if (rc) return mfail(m, rc, "device %d: %s", m->devices[i], hrt_last_error(m->ctx[i])); }
  return HRT_OK;
}

/* sums of the per-device counters; times are the maximum over the devices */
extern "C" int hrt_multi_get_stats(const hrt_multi *m, HrtRunStats *out)
{
  if (!m || !out) return HRT_E_ARG;
  memset(out, 0, sizeof *out);
  for (int i = 0; i < m->n; ++i) {
    HrtRunStats s; hrt_get_stats(m->ctx[i], &s);
    if (i == 0) *out = s;
    else {
      out->ray_bounces += s.ray_bounces; out->primary_hits += s.primary_hits; out->shadow_queries += s.shadow_queries;
      out->los_queries += s.los_queries; out->ambiguous_dirs += s.ambiguous_dirs; out->kernel_launches += s.kernel_launches;
      out->cir_dropped += s.cir_dropped;
      for (int k = 0; k < 5; ++k) { out->work_bounce[k] += s.work_bounce[k]; out->work_scatter[k] += s.work_scatter[k]; }
#define MX(f) if (s.f > out->f) out->f = s.f
      MX(ms_total); MX(ms_bounce); MX(ms_scatter); MX(ms_other); MX(ms_sort); MX(host_ms_setup); MX(host_ms_total); MX(rx_map_build_ms);
#undef MX
    }
  }
  return HRT_OK;
}

/* One job over all devices with HOST results (any flag combination of hrt_run
 * except the *_DEV ones).  Caller-visible semantics are those of hrt_run on one
 * device: dense arrays filled, summaries / impulse responses ADDED to, path
 * list stored up to its capacity. */
extern "C" int hrt_multi_run(hrt_multi *m, const HrtRunParams *p)
{
  if (!m || !p) return HRT_E_ARG;
  if (p->flags & (HRT_FLAG_SUMMARY_DEV | HRT_FLAG_PATHLIST_DEV)) return mfail(m, HRT_E_ARG, "device-resident results: use hrt_multi_run_gathered");
  if (p->shard_world > 1) return mfail(m, HRT_E_ARG, "hrt_multi_run shards the job itself (shard_world must be 0 or 1)");
  const int n = m->n;
  if (n == 1) {
    const int rc = hrt_run(m->ctx[0], p);
    return rc ? mfail(m, rc, "device %d: %s", m->devices[0], hrt_last_error(m->ctx[0])) : HRT_OK;
  }
  const size_t R = p->num_rx, T = p->num_tx, B = p->num_bounces;
  const uint64_t blk = p->shard_block ? p->shard_block : (1u << 16);
  const size_t np = R * T * B, nb = T * B, ncir = (p->flags & HRT_FLAG_CIR) ? R * T * (size_t)p->cir_bins * 4 : 0;
  /* private host results per device for everything that is accumulated */
  std::vector<std::vector<HrtPairSummary>> pair(n);
  std::vector<std::vector<HrtBounceSummary>> bounce(n);
  std::vector<std::vector<float>> cir(n);
  std::vector<uint64_t> found(n, 0);
  const uint64_t cap_each = (p->flags & HRT_FLAG_PATHLIST) ? p->paths_capacity / (uint64_t)n : 0;
  if ((p->flags & HRT_FLAG_PATHLIST) && cap_each == 0) return mfail(m, HRT_E_ARG, "paths_capacity smaller than the number of devices");
  std::vector<HrtRunParams> q(n, *p);
  for (int i = 0; i < n; ++i) {
    q[i].shard_rank = (uint32_t)i; q[i].shard_world = (uint32_t)n; q[i].shard_block = blk;
    q[i].stream = nullptr;
    q[i].flags |= HRT_FLAG_INTERNAL_NO_TAIL;
    if (i) { q[i].los = nullptr; q[i].rays_los = nullptr; }
    if (p->flags & HRT_FLAG_SUMMARY) {
      pair[i].assign(np, HrtPairSummary()); bounce[i].assign(nb, HrtBounceSummary());
      memset(pair[i].data(), 0, np * sizeof(HrtPairSummary)); memset(bounce[i].data(), 0, nb * sizeof(HrtBounceSummary));
      q[i].pair_summary = pair[i].data(); q[i].bounce_summary = bounce[i].data();
    }
    if (ncir) { cir[i].assign(ncir, 0.f); q[i].cir = i ? cir[i].data() : p->cir; if (i == 0) cir[0].clear(); }
    if (p->flags & HRT_FLAG_PATHLIST) { q[i].paths = p->paths + (size_t)i * cap_each; q[i].paths_capacity = cap_each; q[i].paths_count = &found[i]; }
  }
  const int rc = for_each_device(m, [&](int i) { return hrt_run(m->ctx[i], &q[i]); });
  if (rc) return rc;
  if ((p->flags & HRT_FLAG_RAYSINFO) && p->rays_scat && p->rays_scat->rays_active) hrt_internal_raysinfo_tail(m->ctx[0], p);
  if (p->flags & HRT_FLAG_SUMMARY) {
    for (int i = 0; i < n; ++i) {
      for (size_t k = 0; k < np; ++k) {
        p->pair_summary[k].n_valid += pair[i][k].n_valid; p->pair_summary[k].n_occluded += pair[i][k].n_occluded;
        p->pair_summary[k].hit_hash += pair[i][k].hit_hash; p->pair_summary[k].tau_bits += pair[i][k].tau_bits;
        p->pair_summary[k].power_te += pair[i][k].power_te; p->pair_summary[k].power_tm += pair[i][k].power_tm;
      }
      for (size_t k = 0; k < nb; ++k) {
        p->bounce_summary[k].n_traced += bounce[i][k].n_traced; p->bounce_summary[k].n_hit += bounce[i][k].n_hit;
        p->bounce_summary[k].hit_hash += bounce[i][k].hit_hash; p->bounce_summary[k].t_bits += bounce[i][k].t_bits;
      }
    }
  }
  for (int i = 1; i < n && ncir; ++i)
    for (size_t k = 0; k < ncir; ++k) p->cir[k] += cir[i][k];
  if (p->flags & HRT_FLAG_PATHLIST) {
    /* the devices' segments moved together at the front of the caller's buffer */
    uint64_t total = 0, kept = 0;
    for (int i = 0; i < n; ++i) {
      const uint64_t k = found[i] < cap_each ? found[i] : cap_each;
      if (i && k) memmove(p->paths + kept, p->paths + (size_t)i * cap_each, (size_t)k * sizeof(HrtPathRecord));
      kept += k; total += found[i];
    }
    *p->paths_count = total;
  }
  return HRT_OK;
}

/* ------------------------------------------------ device-resident results */

static int ensure_nccl(hrt_multi *m)
{
  if (m->nccl_ok) return HRT_OK;
  if (m->nccl_tried) return mfail(m, HRT_E_STATE, "NCCL is not available in this process");
  m->nccl_tried = true;
  char why[256];
  if (!nccl_bind(m->nccl, why, sizeof why)) return mfail(m, HRT_E_STATE, "%s", why);
  m->comm.assign(m->n, nullptr); m->stream.assign(m->n, nullptr);
  const ncclResult_t r = m->nccl.CommInitAll(m->comm.data(), m->n, m->devices.data());
  if (r != ncclSuccess) return mfail(m, HRT_E_CUDA, "ncclCommInitAll: %s", m->nccl.GetErrorString(r));
  for (int i = 0; i < m->n; ++i) {
    cudaSetDevice(m->devices[i]);
    if (cudaStreamCreateWithFlags(&m->stream[i], cudaStreamNonBlocking) != cudaSuccess) return mfail(m, HRT_E_CUDA, "stream creation failed on device %d", m->devices[i]);
  }
  m->nccl.GetVersion(&m->nccl_version);
  m->nccl_ok = true;
  return HRT_OK;
}

/* out[k] = sum over the n gathered tables; words with (index % period) >= first_double are doubles */
__global__ void k_reduce_tables(unsigned long long *out, const unsigned long long *all, size_t words, int n, uint32_t period, uint32_t first_double)
{
  const size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (k >= words) return;
  if (period && (uint32_t)(k % period) >= first_double) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += ((const double *)all)[(size_t)i * words + k];
    ((double *)out)[k] = s;
  } else {
    unsigned long long s = 0;
    for (int i = 0; i < n; ++i) s += all[(size_t)i * words + k];
    out[k] = s;
  }
}

/* One job over all devices with DEVICE-RESIDENT results on EVERY device:
 *   pair_dev[i], bounce_dev[i]   device memory on device i, [R][T][B] / [T][B]:
 *                                overwritten with the whole job's tables
 *                                (all-gather of the per-device tables + local sum)
 *   paths_dev[i] (optional)      device memory on device i, n * paths_capacity_each
 *                                records: segment j holds device j's valid paths,
 *                                counts[j] of them (counts: host, [n]) -- one
 *                                in-place all-gather of the record buffers
 * p->flags: SUMMARY is implied; DENSE / TRACE / CIR are not supported here. */
extern "C" int hrt_multi_run_gathered(hrt_multi *m, const HrtRunParams *p,
                                      HrtPairSummary *const *pair_dev, HrtBounceSummary *const *bounce_dev,
                                      HrtPathRecord *const *paths_dev, uint64_t paths_capacity_each, uint64_t *counts)
{
  if (!m || !p || !pair_dev || !bounce_dev) return HRT_E_ARG;
  if (p->flags & (HRT_FLAG_DENSE | HRT_FLAG_TRACE | HRT_FLAG_CIR | HRT_FLAG_RAYSINFO)) return mfail(m, HRT_E_ARG, "hrt_multi_run_gathered: summary and path-list results only");
  if (paths_dev && (!counts || !paths_capacity_each)) return mfail(m, HRT_E_ARG, "paths_dev needs counts and paths_capacity_each");
  const int n = m->n;
  int rc = ensure_nccl(m);
  if (rc) return rc;
  const size_t R = p->num_rx, T = p->num_tx, B = p->num_bounces;
  const size_t pw = R * T * B * 6, bw = T * B * 4;          /* 64-bit words per table */
  if (m->cap_tables < pw) {
    free_tables(m);
    m->d_pair.assign(n, nullptr); m->d_bounce.assign(n, nullptr);
    m->d_pair_all.assign(n, nullptr); m->d_bounce_all.assign(n, nullptr);
    for (int i = 0; i < n; ++i) {
      cudaSetDevice(m->devices[i]);
      if (cudaMalloc(&m->d_pair[i], pw * 8) != cudaSuccess || cudaMalloc(&m->d_bounce[i], (bw ? bw : 1) * 8) != cudaSuccess ||
          cudaMalloc(&m->d_pair_all[i], (size_t)n * pw * 8) != cudaSuccess ||
          cudaMalloc(&m->d_bounce_all[i], (size_t)n * (bw ? bw : 1) * 8) != cudaSuccess)
        return mfail(m, HRT_E_NOMEM, "device %d: out of memory for the gather buffers", m->devices[i]);
    }
    m->cap_tables = pw;
  }
  const uint64_t blk = p->shard_block ? p->shard_block : (1u << 16);
  std::vector<HrtRunParams> q(n, *p);
  std::vector<uint64_t> found(n, 0);
  for (int i = 0; i < n; ++i) {
    q[i].shard_rank = (uint32_t)i; q[i].shard_world = (uint32_t)n; q[i].shard_block = blk;
    q[i].stream = (void *)m->stream[i];
    q[i].flags = (p->flags & (HRT_FLAG_BRUTE_FORCE | HRT_FLAG_HOST_DIRS | HRT_FLAG_COUNT)) | HRT_FLAG_SUMMARY | HRT_FLAG_SUMMARY_DEV;
    q[i].pair_summary = (HrtPairSummary *)m->d_pair[i]; q[i].bounce_summary = (HrtBounceSummary *)m->d_bounce[i];
    if (i) { q[i].los = nullptr; q[i].rays_los = nullptr; }
    q[i].scat = nullptr; q[i].rays_scat = nullptr;
    if (paths_dev) {
      q[i].flags |= HRT_FLAG_PATHLIST | HRT_FLAG_PATHLIST_DEV;
      /* own records straight into own segment of the gathered buffer (in-place all-gather) */
      q[i].paths = paths_dev[i] + (size_t)i * paths_capacity_each; q[i].paths_capacity = paths_capacity_each; q[i].paths_count = &found[i];
    }
  }
  rc = for_each_device(m, [&](int i) {
    cudaSetDevice(m->devices[i]);
    cudaMemsetAsync(m->d_pair[i], 0, pw * 8, m->stream[i]); cudaMemsetAsync(m->d_bounce[i], 0, bw * 8, m->stream[i]);
    return hrt_run(m->ctx[i], &q[i]);
  });
  if (rc) return rc;
  /* exchange: every device receives every device's tables (and records) over NVLink */
  ncclResult_t r = m->nccl.GroupStart();
  for (int i = 0; i < n && r == ncclSuccess; ++i) {
    r = m->nccl.AllGather(m->d_pair[i], m->d_pair_all[i], pw, ncclUint64, m->comm[i], m->stream[i]);
    if (r == ncclSuccess && bw) r = m->nccl.AllGather(m->d_bounce[i], m->d_bounce_all[i], bw, ncclUint64, m->comm[i], m->stream[i]);
    if (r == ncclSuccess && paths_dev)
      r = m->nccl.AllGather(paths_dev[i] + (size_t)i * paths_capacity_each, paths_dev[i], (size_t)paths_capacity_each * sizeof(HrtPathRecord),
                            ncclChar, m->comm[i], m->stream[i]);
  }
  if (r == ncclSuccess) r = m->nccl.GroupEnd(); else m->nccl.GroupEnd();
  if (r != ncclSuccess) return mfail(m, HRT_E_CUDA, "ncclAllGather: %s", m->nccl.GetErrorString(r));
  for (int i = 0; i < n; ++i) {
    cudaSetDevice(m->devices[i]);
    k_reduce_tables<<<(unsigned)((pw + 255) / 256), 256, 0, m->stream[i]>>>((unsigned long long *)pair_dev[i], (const unsigned long long *)m->d_pair_all[i], pw, n, 6u, 4u);
    if (bw) k_reduce_tables<<<(unsigned)((bw + 255) / 256), 256, 0, m->stream[i]>>>((unsigned long long *)bounce_dev[i], (const unsigned long long *)m->d_bounce_all[i], bw, n, 0u, 0u);
  }
  for (int i = 0; i < n; ++i) {
    cudaSetDevice(m->devices[i]);
    const cudaError_t e = cudaStreamSynchronize(m->stream[i]);
    if (e != cudaSuccess) return mfail(m, HRT_E_CUDA, "device %d: %s", m->devices[i], cudaGetErrorString(e));
  }
  if (counts) for (int i = 0; i < n; ++i) counts[i] = found[i];
  return HRT_OK;
}
