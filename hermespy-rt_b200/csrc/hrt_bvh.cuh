/* hrt_bvh.cuh -- element functions of the GPU BVH builder (LBVH: Morton sort,
 * Karras radix-tree, bottom-up boxes, leaf cut, padded node emission).
 *
 * Replaces the reference's "for every mesh, for every triangle" loop
 * (src/compute_paths.c:246-255, "TODO BVH") with a structure that returns the
 * same answer.  As with hrt_core.cuh, the functions are __host__ __device__ so
 * that tests/emul can run the identical build serially on the CPU; the
 * product builds on the GPU only (kernels in hrt_cuda.cu).
 */
#pragma once

#include "hrt_core.cuh"

/* ---- triangle set-up: corners -> record + bounds (reference :208-224) ---- */

struct HrtTriSetup {
  float4 q0, q1, q2;      /* record, see hrt_core.cuh */
  V3 lo, hi;              /* exact bounds of the three corners */
};

HRT_HD HrtTriSetup hrt_tri_setup(V3 a, V3 b, V3 c)
{
  HrtTriSetup r;
  const V3 ab = v3_sub(b, a);                 /* :217 / :259 */
  const V3 ac = v3_sub(c, a);                 /* :218 / :260 */
  const V3 n  = v3_normalize(v3_cross(ab, ac)); /* :219-220 */
  r.q0.x = a.x;  r.q0.y = a.y;  r.q0.z = a.z;  r.q0.w = ab.x;
  r.q1.x = ab.y; r.q1.y = ab.z; r.q1.z = ac.x; r.q1.w = ac.y;
  r.q2.x = ac.z; r.q2.y = n.x;  r.q2.z = n.y;  r.q2.w = n.z;
  r.lo = v3(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)));
  r.hi = v3(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)));
  return r;
}

/* ---- Morton keys ---- */

HRT_HD uint32_t hrt_expand10(uint32_t v)   /* 10 bits -> every third bit */
{
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}

/* 64-bit sort key: 30-bit Morton code of the box centre (relative to the scene
 * bounds) in the high word, triangle id in the low word -> all keys distinct. */
HRT_HD uint64_t hrt_morton_key(V3 lo, V3 hi, V3 scene_lo, V3 scene_inv_ext, uint32_t id)
{
  const float cx = (0.5f * (lo.x + hi.x) - scene_lo.x) * scene_inv_ext.x;
  const float cy = (0.5f * (lo.y + hi.y) - scene_lo.y) * scene_inv_ext.y;
  const float cz = (0.5f * (lo.z + hi.z) - scene_lo.z) * scene_inv_ext.z;
  const uint32_t x = (uint32_t)fminf(fmaxf(cx * 1024.f, 0.f), 1023.f);
  const uint32_t y = (uint32_t)fminf(fmaxf(cy * 1024.f, 0.f), 1023.f);
  const uint32_t z = (uint32_t)fminf(fmaxf(cz * 1024.f, 0.f), 1023.f);
  const uint32_t code = (hrt_expand10(x) << 2) | (hrt_expand10(y) << 1) | hrt_expand10(z);
  return ((uint64_t)code << 32) | id;
}

/* ---- Karras 2012 radix tree over sorted distinct keys ---- */

HRT_HD int hrt_clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}

HRT_HD int hrt_delta(const uint64_t *keys, int n, int i, int j)
{
  if (j < 0 || j >= n) return -1;
  return hrt_clz64(keys[i] ^ keys[j]);
}

/* Inner node i (0 <= i < n-1) of the tree over n >= 2 leaves.
 * Children are returned as indices into a combined numbering:
 * [0, n-1) inner nodes, n-1+k for leaf k.  Also the covered leaf range. */
HRT_HD void hrt_karras_node(const uint64_t *keys, int n, int i,
                            int *left, int *right, int *first, int *last)
{
  const int d = (hrt_delta(keys, n, i, i + 1) - hrt_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = hrt_delta(keys, n, i, i - d);
  int lmax = 2;
  while (hrt_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (hrt_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = hrt_delta(keys, n, i, j);
  int s = 0, t = l;
  do {
    t = (t + 1) / 2;
    if (hrt_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  const int gamma = i + s * d + (d < 0 ? d : 0);
  const int lo = i < j ? i : j, hi = i < j ? j : i;
  *left  = (lo == gamma)     ? (n - 1 + gamma)     : gamma;
  *right = (hi == gamma + 1) ? (n - 1 + gamma + 1) : gamma + 1;
  *first = lo; *last = hi;
}

/* ---- node emission ----
 * Karras inner node i becomes a traversal node iff it covers more than
 * leaf_max triangles; smaller subtrees (and single leaves) become leaf refs of
 * their parent.  Boxes are padded by `pad` on every side (conservative
 * culling: see DESIGN.md, "exact hits behind inexact boxes"). */
HRT_HD int hrt_child_ref(int child, int n, const int *first, const int *last,
                         const int *new_index, int leaf_max)
{
  if (child >= n - 1) return hrt_leaf_ref((uint32_t)(child - (n - 1)), 1u);   /* single leaf */
  const int cnt = last[child] - first[child] + 1;
  if (cnt <= leaf_max) return hrt_leaf_ref((uint32_t)first[child], (uint32_t)cnt);
  return new_index[child];
}

/* `oct`: direction octant this copy of the node serves (bit k = component k of
 * the ray direction negative).  On those axes the two planes are stored
 * swapped, so that the first slot always holds the plane a ray of that octant
 * reaches first.  oct = 0 is the plain (lo, hi) layout -- the only one emitted since the
 * kernels traverse the 4-wide nodes collapsed from these (hrt_wide_emit below). */
HRT_HD void hrt_emit_node(float4 *out, int ref_l, int ref_r, V3 llo, V3 lhi, V3 rlo, V3 rhi, float pad,
                          uint32_t oct = 0)
{
  llo = v3(llo.x - pad, llo.y - pad, llo.z - pad); lhi = v3(lhi.x + pad, lhi.y + pad, lhi.z + pad);
  rlo = v3(rlo.x - pad, rlo.y - pad, rlo.z - pad); rhi = v3(rhi.x + pad, rhi.y + pad, rhi.z + pad);
  if (oct & 1u) { float t = llo.x; llo.x = lhi.x; lhi.x = t; t = rlo.x; rlo.x = rhi.x; rhi.x = t; }
  if (oct & 2u) { float t = llo.y; llo.y = lhi.y; lhi.y = t; t = rlo.y; rlo.y = rhi.y; rhi.y = t; }
  if (oct & 4u) { float t = llo.z; llo.z = lhi.z; lhi.z = t; t = rlo.z; rlo.z = rhi.z; rhi.z = t; }
  out[0].x = llo.x; out[0].y = lhi.x; out[0].z = llo.y; out[0].w = lhi.y;
  out[1].x = llo.z; out[1].y = lhi.z; out[1].z = hrt_int_as_float(ref_l); out[1].w = 0.f;
  out[2].x = rlo.x; out[2].y = rhi.x; out[2].z = rlo.y; out[2].w = rhi.y;
  out[3].x = rlo.z; out[3].y = rhi.z; out[3].z = hrt_int_as_float(ref_r); out[3].w = 0.f;
}

/* Padding rule: `ulps` fp32 epsilons of the largest coordinate magnitude any
 * ray origin or vertex can have. */
HRT_HD float hrt_box_pad(float max_abs_coord, float ulps)
{
  return ulps * FLT_EPSILON * fmaxf(max_abs_coord, 1.0f);
}

/* ---- 4-wide nodes (BVH4) ----
 * The traversal kernels walk a 4-wide tree collapsed from the binary one: a
 * binary node at even depth becomes a wide node whose children are its
 * grandchildren (or its children where those are leaves).  One visit tests four
 * boxes with one set of loads and NO ordering decision: there is one node copy
 * per ray-direction octant (as for the binary nodes), and inside a copy the
 * children are stored front-to-back for that octant (by the box centre along the
 * octant's diagonal), with the near / far plane of every axis pre-selected.
 *
 * Wide node, 7 x float4 = 112 B, structure of arrays over the four children:
 *   w0 near.x[4]  w1 far.x[4]  w2 near.y[4]  w3 far.y[4]  w4 near.z[4]  w5 far.z[4]
 *   w6 ref[4]     >= 0 wide node index, < 0 leaf (as hrt_leaf_ref), empty slot:
 *                 HRT_WIDE_EMPTY with a box no ray can enter (near +3e38, far -3e38)
 * Pure traversal order, as with every other tree shape: the result is the
 * minimum over (t, triangle id). */
#define HRT_WIDE_F4 7u          /* float4s per wide node */

struct HrtWideChild { int ref; V3 lo, hi; };

/* writes the `octants` copies of wide node `w`; copies are `oct_stride4` float4s apart */
HRT_HD void hrt_wide_emit(float4 *nodes, size_t oct_stride4, uint32_t octants, uint32_t w,
                          const HrtWideChild *ch, int n, uint32_t node_stride4 = HRT_WIDE_F4)
{
  for (uint32_t oct = 0; oct < octants; ++oct) {
    /* front-to-back order for this octant: ascending centre along (+-1, +-1, +-1) */
    int order[4] = { 0, 1, 2, 3 };
    float key[4];
    for (int k = 0; k < 4; ++k) {
      if (k >= n) { key[k] = 3.0e38f; continue; }
      const float cx = ch[k].lo.x + ch[k].hi.x, cy = ch[k].lo.y + ch[k].hi.y, cz = ch[k].lo.z + ch[k].hi.z;
      key[k] = ((oct & 1u) ? -cx : cx) + ((oct & 2u) ? -cy : cy) + ((oct & 4u) ? -cz : cz);
    }
    for (int i = 1; i < 4; ++i)
      for (int j = i; j > 0 && key[order[j]] < key[order[j - 1]]; --j) { const int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
    float v[7][4];
    for (int s = 0; s < 4; ++s) {
      const int k = order[s];
      if (k >= n) {
        v[0][s] = v[2][s] = v[4][s] = 3.0e38f; v[1][s] = v[3][s] = v[5][s] = -3.0e38f;
        if (oct & 1u) { v[0][s] = -3.0e38f; v[1][s] = 3.0e38f; }
        if (oct & 2u) { v[2][s] = -3.0e38f; v[3][s] = 3.0e38f; }
        if (oct & 4u) { v[4][s] = -3.0e38f; v[5][s] = 3.0e38f; }
        v[6][s] = hrt_int_as_float(HRT_WIDE_EMPTY);
        continue;
      }
      v[0][s] = (oct & 1u) ? ch[k].hi.x : ch[k].lo.x; v[1][s] = (oct & 1u) ? ch[k].lo.x : ch[k].hi.x;
      v[2][s] = (oct & 2u) ? ch[k].hi.y : ch[k].lo.y; v[3][s] = (oct & 2u) ? ch[k].lo.y : ch[k].hi.y;
      v[4][s] = (oct & 4u) ? ch[k].hi.z : ch[k].lo.z; v[5][s] = (oct & 4u) ? ch[k].lo.z : ch[k].hi.z;
      v[6][s] = hrt_int_as_float(ch[k].ref);
    }
    float4 *dst = nodes + (size_t)oct * oct_stride4 + (size_t)w * node_stride4;
    for (int q = 0; q < 7; ++q) { dst[q].x = v[q][0]; dst[q].y = v[q][1]; dst[q].z = v[q][2]; dst[q].w = v[q][3]; }
  }
}

/* children of wide node = children of binary node `i` (plain octant-0 copy of
 * the emitted binary nodes, boxes already padded), inner children replaced by
 * their own two children.  widx maps an (even-depth) binary node to its wide
 * index.  Returns the number of children written to ch[0..3]. */
HRT_HD int hrt_wide_children(const float4 *bn, int i, const uint32_t *widx, HrtWideChild *ch)
{
  int n = 0;
  for (int s = 0; s < 2; ++s) {
    const float4 a = bn[4 * (size_t)i + 2 * s], b = bn[4 * (size_t)i + 2 * s + 1];
    const int ref = hrt_float_as_int(b.z);
    if (ref < 0) {
      ch[n].ref = ref; ch[n].lo = v3(a.x, a.z, b.x); ch[n].hi = v3(a.y, a.w, b.y); ++n;
      continue;
    }
    for (int g = 0; g < 2; ++g) {
      const float4 c = bn[4 * (size_t)ref + 2 * g], e = bn[4 * (size_t)ref + 2 * g + 1];
      const int gref = hrt_float_as_int(e.z);
      ch[n].ref = gref < 0 ? gref : (int)widx[gref];
      ch[n].lo = v3(c.x, c.z, e.x); ch[n].hi = v3(c.y, c.w, e.y); ++n;
    }
  }
  return n;
}
