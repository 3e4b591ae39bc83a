/* hrt_build_kernels.cuh -- device side of the scene pipeline (included by
 * hrt_cuda.cu, one translation unit): triangle set-up (reference
 * src/compute_paths.c:208-224), the binned-SAH BVH builder, the Morton/Karras
 * builder kept as HRT_BVH_LBVH=1, node emission with octant copies, and the
 * moving-mesh kernels of hrt_scene_advance.  Host drivers: hrt_cuda.cu. */
#pragma once

/* ---------------------------------------------------------- build kernels */

__device__ __forceinline__ unsigned enc_f(float f)
{ unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__host__ __device__ __forceinline__ float dec_f(unsigned u)
{
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  float f; memcpy(&f, &u, 4); return f;
}

/* verts: all meshes' vertices concatenated; idx: 3 per triangle, already offset */
__global__ void k_tri_setup(const float *verts, const uint32_t *idx, uint32_t n,
                            float4 *recs, float *boxes, unsigned *bounds /* lo xyz, hi xyz */)
{
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const uint32_t ia = idx[3 * g], ib = idx[3 * g + 1], ic = idx[3 * g + 2];
  const V3 a = v3(verts[3 * ia], verts[3 * ia + 1], verts[3 * ia + 2]);
  const V3 b = v3(verts[3 * ib], verts[3 * ib + 1], verts[3 * ib + 2]);
  const V3 c = v3(verts[3 * ic], verts[3 * ic + 1], verts[3 * ic + 2]);
  const HrtTriSetup s = hrt_tri_setup(a, b, c);
  recs[3 * g] = s.q0; recs[3 * g + 1] = s.q1; recs[3 * g + 2] = s.q2;
  float *bx = boxes + 6 * (size_t)g;
  bx[0] = s.lo.x; bx[1] = s.lo.y; bx[2] = s.lo.z; bx[3] = s.hi.x; bx[4] = s.hi.y; bx[5] = s.hi.z;
  atomicMin(&bounds[0], enc_f(s.lo.x)); atomicMin(&bounds[1], enc_f(s.lo.y)); atomicMin(&bounds[2], enc_f(s.lo.z));
  atomicMax(&bounds[3], enc_f(s.hi.x)); atomicMax(&bounds[4], enc_f(s.hi.y)); atomicMax(&bounds[5], enc_f(s.hi.z));
}

__global__ void k_morton(const float *boxes, uint32_t n, const unsigned *bounds, uint64_t *keys)
{
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  const V3 slo = v3(dec_f(bounds[0]), dec_f(bounds[1]), dec_f(bounds[2]));
  const V3 shi = v3(dec_f(bounds[3]), dec_f(bounds[4]), dec_f(bounds[5]));
  const V3 inv = v3(1.f / fmaxf(shi.x - slo.x, 1e-30f), 1.f / fmaxf(shi.y - slo.y, 1e-30f),
                    1.f / fmaxf(shi.z - slo.z, 1e-30f));
  const float *bx = boxes + 6 * (size_t)g;
  keys[g] = hrt_morton_key(v3(bx[0], bx[1], bx[2]), v3(bx[3], bx[4], bx[5]), slo, inv, g);
}

/* leaf order: triangle records, ids and leaf boxes follow the sorted keys */
__global__ void k_gather(const uint64_t *keys, uint32_t n, const float4 *recs, const float *boxes,
                         float4 *tris, uint32_t *tri_gid, float *node_box)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const uint32_t g = (uint32_t)(keys[s] & 0xFFFFFFFFull);
  tris[3 * s] = recs[3 * g]; tris[3 * s + 1] = recs[3 * g + 1]; tris[3 * s + 2] = recs[3 * g + 2];
  tri_gid[s] = g;
  float *dst = node_box + 6 * (size_t)(n - 1 + s);
  const float *src = boxes + 6 * (size_t)g;
  for (int k = 0; k < 6; ++k) dst[k] = src[k];
}

__global__ void k_karras(const uint64_t *keys, int n, int *kl, int *kr, int *kfirst, int *klast,
                         int *parent)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int l, r, f, la;
  hrt_karras_node(keys, n, i, &l, &r, &f, &la);
  kl[i] = l; kr[i] = r; kfirst[i] = f; klast[i] = la;
  parent[l] = i; parent[r] = i;
  if (i == 0) parent[0] = -1;   /* written by nobody else: node 0 is never a child */
}

/* bottom-up boxes: the second thread to reach a node merges its children */
__global__ void k_refit(int n, const int *kl, const int *kr, const int *parent, float *node_box,
                        unsigned *arrive)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int cur = parent[n - 1 + s];
  while (cur >= 0) {
    __threadfence();
    if (atomicAdd(&arrive[cur], 1u) == 0u) return;
    const volatile float *a = node_box + 6 * (size_t)kl[cur];
    const volatile float *b = node_box + 6 * (size_t)kr[cur];
    float *o = node_box + 6 * (size_t)cur;
    for (int k = 0; k < 3; ++k) o[k] = fminf(a[k], b[k]);
    for (int k = 3; k < 6; ++k) o[k] = fmaxf(a[k], b[k]);
    cur = parent[cur];
  }
}

__global__ void k_mark(int n, const int *kfirst, const int *klast, int leaf_max, int *used)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  used[i] = (klast[i] - kfirst[i] + 1) > leaf_max ? 1 : 0;
}

__global__ void k_emit(int n, const int *kl, const int *kr, const int *kfirst, const int *klast,
                       const int *used, const int *newidx, const float *node_box, int leaf_max,
                       float pad, float4 *nodes, uint32_t octants, uint32_t num_nodes)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1 || !used[i]) return;
  const int l = kl[i], r = kr[i];
  const float *bl = node_box + 6 * (size_t)l, *br = node_box + 6 * (size_t)r;
  for (uint32_t oct = 0; oct < octants; ++oct)
    hrt_emit_node(nodes + 4 * ((size_t)oct * num_nodes + newidx[i]),
                  hrt_child_ref(l, n, kfirst, klast, newidx, leaf_max),
                  hrt_child_ref(r, n, kfirst, klast, newidx, leaf_max),
                  v3(bl[0], bl[1], bl[2]), v3(bl[3], bl[4], bl[5]),
                  v3(br[0], br[1], br[2]), v3(br[3], br[4], br[5]), pad, oct);
}

/* ---------------------------------------------- binned-SAH builder (default)
 * Top-down, level-synchronous: every level bins the triangles of each open node
 * (16 bins x 3 axes over the node's centroid bounds), picks the plane with the
 * lowest surface-area cost, and partitions the node's slice of the index array.
 * Ranges of <= leaf_max triangles become leaves; the final index array IS the
 * leaf order.  Compared with the Morton/Karras tree this roughly halves the
 * box tests of a shadow query on street scenes (scripts/bvh_lab.py).
 * Tree shape never influences results: the traversal returns the minimum over
 * (t, triangle id) whatever the tree. */
#define SAH_BINS 16
#define SAH_FORCE_MEDIAN_LEVEL 32   /* depth <= 32 + log2(n) < HRT_STACK */

struct SahWork {
  uint32_t start, end;
  int inner;                 /* index of this node in the raw node arrays */
  unsigned cb[6];            /* centroid bounds, order-encoded (enc_f): lo xyz, hi xyz */
  float box[6];              /* the node's own box */
  float cmin[3], cscale[3];
  int axis, bin;             /* split: axis < 0 = halve by position */
  uint32_t nl;
  int child[2];              /* work index on the next level, -1 = leaf */
  unsigned fill[2];
};
struct SahBin { unsigned count, lo[3], hi[3]; };

__device__ __forceinline__ int sah_bin_of(const SahWork &w, int ax, float c)
{
  const int b = (int)((c - w.cmin[ax]) * w.cscale[ax]);
  return b < 0 ? 0 : (b > SAH_BINS - 1 ? SAH_BINS - 1 : b);
}

/* level 0: identity order, everything belongs to the root (work[0], set up by
 * the host), whose centroid bounds are accumulated here */
__global__ void k_sah_init(uint32_t n, const float *boxes, uint32_t *idx, int *work_of, SahWork *work)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  idx[p] = p; work_of[p] = 0;
  const float *b = boxes + 6 * (size_t)p;
  for (int k = 0; k < 3; ++k) {
    const unsigned e = enc_f(0.5f * (b[k] + b[3 + k]));
    atomicMin(&work[0].cb[k], e); atomicMax(&work[0].cb[3 + k], e);
  }
}

/* per open node: bin grid from the centroid bounds, empty bins */
__global__ void k_sah_prep(int count, SahWork *work, SahBin *bins)
{
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= count) return;
  SahWork &W = work[w];
  for (int k = 0; k < 3; ++k) {
    const float lo = dec_f(W.cb[k]), hi = dec_f(W.cb[3 + k]);
    W.cmin[k] = lo;
    W.cscale[k] = hi > lo ? (float)SAH_BINS * (1.f - 1e-6f) / (hi - lo) : 0.f;
  }
  W.fill[0] = W.fill[1] = 0;
  SahBin z; z.count = 0;
  for (int k = 0; k < 3; ++k) { z.lo[k] = 0xFFFFFFFFu; z.hi[k] = 0u; }
  for (int k = 0; k < 3 * SAH_BINS; ++k) bins[(size_t)w * 3 * SAH_BINS + k] = z;
}

__global__ void k_sah_bin(uint32_t n, const uint32_t *idx, const int *work_of, const SahWork *work,
                          const float *boxes, SahBin *bins)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int w = work_of[p];
  if (w < 0) return;
  const SahWork &W = work[w];
  const float *b = boxes + 6 * (size_t)idx[p];
  unsigned lo[3], hi[3];
  for (int k = 0; k < 3; ++k) { lo[k] = enc_f(b[k]); hi[k] = enc_f(b[3 + k]); }
  for (int ax = 0; ax < 3; ++ax) {
    SahBin *bin = &bins[((size_t)w * 3 + ax) * SAH_BINS + sah_bin_of(W, ax, 0.5f * (b[ax] + b[3 + ax]))];
    atomicAdd(&bin->count, 1u);
    for (int k = 0; k < 3; ++k) { atomicMin(&bin->lo[k], lo[k]); atomicMax(&bin->hi[k], hi[k]); }
  }
}

__device__ __forceinline__ float sah_area(const float lo[3], const float hi[3])
{
  const float x = hi[0] - lo[0], y = hi[1] - lo[1], z = hi[2] - lo[2];
  return x * y + y * z + z * x;
}

/* per open node: best plane, children, raw node record */
__global__ void k_sah_split(int count, int level, int leaf_max, SahWork *work, const SahBin *bins, SahWork *next,
                            int *counters, int2 *raw_ref, float *raw_box)
{
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= count) return;
  SahWork &W = work[w];
  const uint32_t cnt = W.end - W.start;
  float best = 3.0e38f; int best_ax = -1, best_k = 0; uint32_t best_nl = 0;
  if (level < SAH_FORCE_MEDIAN_LEVEL) {
    for (int ax = 0; ax < 3; ++ax) {
      const SahBin *B = bins + ((size_t)w * 3 + ax) * SAH_BINS;
      float r_area[SAH_BINS];
      float lo[3] = { 3e38f, 3e38f, 3e38f }, hi[3] = { -3e38f, -3e38f, -3e38f };
      for (int k = SAH_BINS - 1; k >= 1; --k) {
        if (B[k].count)
          for (int j = 0; j < 3; ++j) { lo[j] = fminf(lo[j], dec_f(B[k].lo[j])); hi[j] = fmaxf(hi[j], dec_f(B[k].hi[j])); }
        r_area[k] = sah_area(lo, hi);
      }
      for (int j = 0; j < 3; ++j) { lo[j] = 3e38f; hi[j] = -3e38f; }
      uint32_t nl = 0;
      for (int k = 1; k < SAH_BINS; ++k) {
        if (B[k - 1].count)
          for (int j = 0; j < 3; ++j) { lo[j] = fminf(lo[j], dec_f(B[k - 1].lo[j])); hi[j] = fmaxf(hi[j], dec_f(B[k - 1].hi[j])); }
        nl += B[k - 1].count;
        if (nl == 0 || nl == cnt) continue;
        const float c = sah_area(lo, hi) * (float)nl + r_area[k] * (float)(cnt - nl);
        if (c < best) { best = c; best_ax = ax; best_k = k; best_nl = nl; }
      }
    }
  }
  float cbox[2][6];
  if (best_ax >= 0) {
    const SahBin *B = bins + ((size_t)w * 3 + best_ax) * SAH_BINS;
    for (int s = 0; s < 2; ++s) {
      float lo[3] = { 3e38f, 3e38f, 3e38f }, hi[3] = { -3e38f, -3e38f, -3e38f };
      for (int k = s ? best_k : 0; k < (s ? SAH_BINS : best_k); ++k)
        if (B[k].count)
          for (int j = 0; j < 3; ++j) { lo[j] = fminf(lo[j], dec_f(B[k].lo[j])); hi[j] = fmaxf(hi[j], dec_f(B[k].hi[j])); }
      for (int j = 0; j < 3; ++j) { cbox[s][j] = lo[j]; cbox[s][3 + j] = hi[j]; }
    }
  } else {
    best_nl = cnt / 2;
    for (int s = 0; s < 2; ++s) for (int j = 0; j < 6; ++j) cbox[s][j] = W.box[j];
  }
  W.axis = best_ax; W.bin = best_k; W.nl = best_nl;
  int2 ref;
  for (int s = 0; s < 2; ++s) {
    const uint32_t cs = s ? W.start + best_nl : W.start, ce = s ? W.end : W.start + best_nl;
    int r;
    if (ce - cs <= (uint32_t)leaf_max) {
      r = hrt_leaf_ref(cs, ce - cs);
      W.child[s] = -1;
    } else {
      r = atomicAdd(&counters[0], 1);
      const int nw = atomicAdd(&counters[1], 1);
      SahWork &C = next[nw];
      C.start = cs; C.end = ce; C.inner = r;
      for (int j = 0; j < 3; ++j) { C.cb[j] = 0xFFFFFFFFu; C.cb[3 + j] = 0u; }
      for (int j = 0; j < 6; ++j) C.box[j] = cbox[s][j];
      W.child[s] = nw;
    }
    if (s) ref.y = r; else ref.x = r;
  }
  raw_ref[W.inner] = ref;
  float *rb = raw_box + 12 * (size_t)W.inner;
  for (int j = 0; j < 6; ++j) { rb[j] = cbox[0][j]; rb[6 + j] = cbox[1][j]; }
}

/* per triangle: move to its side of the split; children's centroid bounds */
__global__ void k_sah_part(uint32_t n, const uint32_t *idx, const int *work_of, SahWork *work, SahWork *next,
                           const float *boxes, uint32_t *idx_out, int *work_out)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int w = work_of[p];
  const uint32_t prim = idx[p];
  if (w < 0) { idx_out[p] = prim; work_out[p] = -1; return; }
  SahWork &W = work[w];
  const float *b = boxes + 6 * (size_t)prim;
  int side; uint32_t dest;
  if (W.axis < 0) {
    side = (p - W.start) >= W.nl; dest = p;
  } else {
    side = sah_bin_of(W, W.axis, 0.5f * (b[W.axis] + b[3 + W.axis])) >= W.bin;
    dest = side ? W.start + W.nl + atomicAdd(&W.fill[1], 1u) : W.start + atomicAdd(&W.fill[0], 1u);
  }
  idx_out[dest] = prim;
  const int cw = W.child[side];
  work_out[dest] = cw;
  if (cw >= 0)
    for (int k = 0; k < 3; ++k) {
      const unsigned e = enc_f(0.5f * (b[k] + b[3 + k]));
      atomicMin(&next[cw].cb[k], e); atomicMax(&next[cw].cb[3 + k], e);
    }
}

/* leaf order from the final index array */
__global__ void k_gather_idx(const uint32_t *idx, uint32_t n, const float4 *recs, float4 *tris, uint32_t *tri_gid)
{
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const uint32_t g = idx[s];
  tris[3 * s] = recs[3 * g]; tris[3 * s + 1] = recs[3 * g + 1]; tris[3 * s + 2] = recs[3 * g + 2];
  tri_gid[s] = g;
}

__global__ void k_emit_raw(int num_inner, const int2 *raw_ref, const float *raw_box, float pad, float4 *nodes,
                           uint32_t octants, uint32_t num_nodes)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= num_inner) return;
  const int2 r = raw_ref[i];
  const float *b = raw_box + 12 * (size_t)i;
  for (uint32_t oct = 0; oct < octants; ++oct)
    hrt_emit_node(nodes + 4 * ((size_t)oct * num_nodes + i), r.x, r.y, v3(b[0], b[1], b[2]), v3(b[3], b[4], b[5]),
                  v3(b[6], b[7], b[8]), v3(b[9], b[10], b[11]), pad, oct);
}

/* ---- moving meshes (Mesh.velocity, reference inc/scene.h:21-22) ---- */

__global__ void k_move_verts(float *verts, const uint32_t *vmesh, const float *mesh_vel, size_t nv, float dt)
{
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= nv) return;
  const uint32_t m = vmesh[i];
  for (int k = 0; k < 3; ++k) verts[3 * i + k] = HRT_ADD(verts[3 * i + k], HRT_MUL(mesh_vel[3 * m + k], dt));
}

/* boxes of one build level from the level below (children always have larger
 * node indices and live on the next level) and from the leaf triangles */
__global__ void k_sah_refit(int first, int last, const int2 *raw_ref, float *raw_box, const uint32_t *tri_gid,
                            const float *boxes)
{
  const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= last) return;
  const int2 r = raw_ref[i];
  for (int s = 0; s < 2; ++s) {
    const int ref = s ? r.y : r.x;
    float lo[3] = { 3e38f, 3e38f, 3e38f }, hi[3] = { -3e38f, -3e38f, -3e38f };
    if (ref < 0) {
      const uint32_t code = (uint32_t)~ref, f = code >> 3, cnt = (code & 7u) + 1u;
      for (uint32_t k = 0; k < cnt; ++k) {
        const float *b = boxes + 6 * (size_t)tri_gid[f + k];
        for (int j = 0; j < 3; ++j) { lo[j] = fminf(lo[j], b[j]); hi[j] = fmaxf(hi[j], b[3 + j]); }
      }
    } else {
      const float *c = raw_box + 12 * (size_t)ref;
      for (int j = 0; j < 3; ++j) { lo[j] = fminf(c[j], c[6 + j]); hi[j] = fmaxf(c[3 + j], c[9 + j]); }
    }
    float *o = raw_box + 12 * (size_t)i + 6 * s;
    for (int j = 0; j < 3; ++j) { o[j] = lo[j]; o[3 + j] = hi[j]; }
  }
}

/* ---- binary -> 4-wide collapse (hrt_bvh.cuh, "4-wide nodes") ----
 * bn: the emitted binary nodes (plain copy, boxes padded).  Generic over the
 * builder: only child refs are followed. */
__global__ void k_wide_parent(const float4 *bn, int n, int *parent)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0) parent[0] = -1;                   /* node 0 is the root, nobody's child */
  const int l = __float_as_int(bn[4 * (size_t)i + 1].z), r = __float_as_int(bn[4 * (size_t)i + 3].z);
  if (l >= 0) parent[l] = i;
  if (r >= 0) parent[r] = i;
}

/* even[i] = 1 when binary node i sits at even depth (it becomes a wide node) */
__global__ void k_wide_depth(int n, const int *parent, uint8_t *even)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int d = 0;
  for (int j = parent[i]; j >= 0; j = parent[j]) ++d;
  even[i] = (uint8_t)((d & 1) == 0);
}

/* exclusive prefix count of the flags, one block: wide nodes are numbered in
 * binary index order (deterministic; level order for the SAH builder) */
__global__ void __launch_bounds__(1024) k_wide_scan(int n, const uint8_t *even, uint32_t *widx, uint32_t *total)
{
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + (int)threadIdx.x;
    const uint32_t f = i < n ? even[i] : 0u;
    const unsigned m = __ballot_sync(0xFFFFFFFFu, f != 0u);
    if (lane == 0) warp_sum[warp] = (uint32_t)__popc(m);
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < warp; ++w) before += warp_sum[w];
    if (i < n) widx[i] = base + before + (uint32_t)__popc(m & ((1u << lane) - 1u));
    __syncthreads();
    if (threadIdx.x == 1023) base += before + (uint32_t)__popc(m);
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = base;
}

__global__ void k_wide_emit(const float4 *bn, int n, const uint8_t *even, const uint32_t *widx,
                            float4 *wnodes, size_t oct_stride4, uint32_t octants, uint32_t node_stride4)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !even[i]) return;
  HrtWideChild ch[4];
  const int c = hrt_wide_children(bn, i, widx, ch);
  hrt_wide_emit(wnodes, oct_stride4, octants, widx[i], ch, c, node_stride4);
}
