/* hrt_rxmap.cuh -- receiver maps: the shadow queries of k_scatter without a
 * tree walk.
 *
 * Every shadow ray of the reference's per-receiver loop (src/compute_paths.c:
 * 670-683) runs from a hit point o towards a receiver R and on to infinity: it
 * lies on a line through R.  Which triangles a line through R can touch depends
 * on its direction only.  For every receiver the directions around R are
 * therefore divided into the cells of a cube map (6 faces x G x G), and each
 * cell gets the list of triangles some line through R inside that cell could
 * touch -- a conservative rasterisation of the scene as seen from R, every
 * triangle at every depth, not only the visible ones.  A shadow query then is
 * two cell look-ups (direction +d: beyond the receiver, -d: between receiver
 * and hit point, and behind it) and the exact Moeller-Trumbore test of the
 * listed triangles, a handful instead of ~24 box and ~4 triangle tests.
 *
 * As with the BVH this only decides WHICH triangles get the exact test
 * (hrt_mt_test, the reference's arithmetic): the result is the minimum over
 * (t, triangle id) of the accepted candidates.  The lists are conservative:
 *   - a triangle is left out of a cell only if it is separated from the cell's
 *     pyramid (apex R) by one of the pyramid's four side planes by more than
 *     `pad` metres, or if the pyramid lies outside the plane through R and one
 *     of the triangle's edges by more than the angle pad subtends at that edge;
 *   - cells are enlarged by HRT_RXMAP_EPS in cube-map coordinates (the query
 *     computes its cell in fp32);
 *   - pad covers the distance by which the fp32 direction normalize(R - o)
 *     misses R (<= 1e-7 |R - o|) and the reference's epsilon windows.
 * The same documented exception as for the BVH applies: "phantom hits" of rays
 * within ~1e-5 rad of a triangle's plane.
 *
 * The element functions are __host__ __device__ so that tests/emul builds the
 * same maps serially and checks conservativeness against the oracle on the CPU.
 */
#pragma once

#include "hrt_core.cuh"

#define HRT_RXMAP_EPS 2e-4f          /* cell enlargement in cube-map coordinates ([-1, 1] per face) */
#define HRT_RXMAP_BLOCK 8u           /* cells per side of a build block */

/* linear cell index of direction w (any length, not all zero): face-major,
 * then row j, then column i.  Face = 2 * major axis + (negative ? 1 : 0); the
 * two remaining components, in x < y < z order, divided by |major|, are the
 * cell coordinates (a, b) in [-1, 1]. */
HRT_HD uint32_t hrt_rxmap_cell(V3 w, uint32_t G)
{
  const float ax = fabsf(w.x), ay = fabsf(w.y), az = fabsf(w.z);
  uint32_t face; float m, a, b;
  if (ax >= ay && ax >= az) { face = w.x < 0.f ? 1u : 0u; m = ax; a = w.y; b = w.z; }
  else if (ay >= az)        { face = w.y < 0.f ? 3u : 2u; m = ay; a = w.x; b = w.z; }
  else                      { face = w.z < 0.f ? 5u : 4u; m = az; a = w.x; b = w.y; }
  const float half = 0.5f * (float)G, inv = half / m;
  const float fi = fminf(fmaxf(HRT_FMA(a, inv, half), 0.f), (float)(G - 1u));
  const float fj = fminf(fmaxf(HRT_FMA(b, inv, half), 0.f), (float)(G - 1u));
  return (face * G + (uint32_t)fj) * G + (uint32_t)fi;
}

/* Both cells of a line through the apex at once: cell of direction w and of -w.
 * The opposite direction lies on the opposite face (face ^ 1) with both cell
 * coordinates mirrored: G - 1 - i, G - 1 - j.  (For a coordinate exactly on a
 * cell boundary the mirrored index is the neighbour of the exact one; such a
 * direction belongs to both enlarged cells, HRT_RXMAP_EPS.)  1/|major| may be an
 * approximate reciprocal: 2^-22 relative is 1e-4 of a cell, inside the same margin. */
HRT_HD void hrt_rxmap_cells2(V3 w, uint32_t G, uint32_t *cell_pos, uint32_t *cell_neg)
{
  const float ax = fabsf(w.x), ay = fabsf(w.y), az = fabsf(w.z);
  uint32_t face; float m, a, b;
  if (ax >= ay && ax >= az) { face = w.x < 0.f ? 1u : 0u; m = ax; a = w.y; b = w.z; }
  else if (ay >= az)        { face = w.y < 0.f ? 3u : 2u; m = ay; a = w.x; b = w.z; }
  else                      { face = w.z < 0.f ? 5u : 4u; m = az; a = w.x; b = w.y; }
  const float half = 0.5f * (float)G;
#if defined(__CUDA_ARCH__)
  float inv; asm("rcp.approx.f32 %0, %1;" : "=f"(inv) : "f"(m));
  inv *= half;
#else
  const float inv = half / m;
#endif
  const float top = (float)(G - 1u);
  const uint32_t i = (uint32_t)fminf(fmaxf(HRT_FMA(a, inv, half), 0.f), top);
  const uint32_t j = (uint32_t)fminf(fmaxf(HRT_FMA(b, inv, half), 0.f), top);
  *cell_pos = (face * G + j) * G + i;
  *cell_neg = ((face ^ 1u) * G + (G - 1u - j)) * G + (G - 1u - i);
}

/* direction of cube-map point (a, b) on `face` */
HRT_HD V3 hrt_rxmap_dir(uint32_t face, float a, float b)
{
  const float s = (face & 1u) ? -1.f : 1.f;
  switch (face >> 1) {
    case 0:  return v3(s, a, b);
    case 1:  return v3(a, s, b);
    default: return v3(a, b, s);
  }
}

/* the pyramid of the cell range [i0, i1) x [j0, j1) of `face`: four inward side
 * plane normals (planes through the apex), unit length, and four unit corner
 * directions */
struct HrtPyramid { V3 m[4]; V3 c[4]; };

/* 1 / sqrt(x) to ~1 ulp: the approximate reciprocal square root and one Newton step on the device (the build kernel
 * spent a seventh of its instructions on the IEEE square roots and divisions of its pyramids) */
HRT_HD float hrt_rxmap_rsqrt(float x)
{
#if defined(__CUDA_ARCH__)
  float r; asm("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r * (1.5f - 0.5f * x * r * r);
#else
  return 1.f / sqrtf(x);
#endif
}

HRT_HD HrtPyramid hrt_rxmap_pyramid(uint32_t face, uint32_t G, uint32_t i0, uint32_t i1, uint32_t j0, uint32_t j1)
{
  const float g = 2.f / (float)G;
  const float a0 = (float)i0 * g - 1.f - HRT_RXMAP_EPS, a1 = (float)i1 * g - 1.f + HRT_RXMAP_EPS;
  const float b0 = (float)j0 * g - 1.f - HRT_RXMAP_EPS, b1 = (float)j1 * g - 1.f + HRT_RXMAP_EPS;
  HrtPyramid p;
  p.c[0] = hrt_rxmap_dir(face, a0, b0); p.c[1] = hrt_rxmap_dir(face, a1, b0);
  p.c[2] = hrt_rxmap_dir(face, a1, b1); p.c[3] = hrt_rxmap_dir(face, a0, b1);
  const V3 mid = hrt_rxmap_dir(face, 0.5f * (a0 + a1), 0.5f * (b0 + b1));
  for (int k = 0; k < 4; ++k) {
    const V3 u = p.c[k], v = p.c[(k + 1) & 3];
    V3 n = v3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
    const float l = hrt_rxmap_rsqrt(n.x * n.x + n.y * n.y + n.z * n.z);
    const float sgn = (n.x * mid.x + n.y * mid.y + n.z * mid.z) < 0.f ? -l : l;
    p.m[k] = v3(n.x * sgn, n.y * sgn, n.z * sgn);
  }
  for (int k = 0; k < 4; ++k) {
    const V3 u = p.c[k];
    const float l = hrt_rxmap_rsqrt(u.x * u.x + u.y * u.y + u.z * u.z);
    p.c[k] = v3(u.x * l, u.y * l, u.z * l);
  }
  return p;
}

HRT_HD float hrt_dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

/* May a line through the apex inside the pyramid touch the triangle (va, vb, vc:
 * corners relative to the apex) or anything within `pad` of it?  false only
 * when provably not (see the file header). */
HRT_HD bool hrt_rxmap_overlap(const HrtPyramid &p, V3 va, V3 vb, V3 vc, float pad)
{
  /* (a) a side plane of the pyramid separates the padded triangle from it */
  for (int k = 0; k < 4; ++k)
    if (hrt_dot3(va, p.m[k]) < -pad && hrt_dot3(vb, p.m[k]) < -pad && hrt_dot3(vc, p.m[k]) < -pad) return false;
  /* (b) the pyramid lies outside the plane through the apex and one edge.  Only
   * when the apex is well off the triangle's plane (else its image on the sphere
   * of directions degenerates and "inside" has no sign). */
  const V3 e1 = v3(vb.x - va.x, vb.y - va.y, vb.z - va.z), e2 = v3(vc.x - va.x, vc.y - va.y, vc.z - va.z);
  const V3 n = v3(e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x);
  const float nl = sqrtf(hrt_dot3(n, n));
  const float h = hrt_dot3(va, n);                       /* distance of the apex from the plane, times nl */
  if (!(fabsf(h) > 16.f * pad * nl)) return true;
  const V3 vs[3] = { va, vb, vc };
  for (int k = 0; k < 3; ++k) {
    const V3 u = vs[k], v = vs[(k + 1) % 3], w = vs[(k + 2) % 3];
    const V3 e = v3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);   /* normal of the plane (apex, edge) */
    const V3 ed = v3(v.x - u.x, v.y - u.y, v.z - u.z);
    const float sgn = hrt_dot3(w, e) < 0.f ? -1.f : 1.f;
    /* |e| / |edge| = distance of the apex from the edge's line: the pad subtends
     * pad |edge| / |e| there (twice that for safety); corners are unit vectors */
    const float lim = -2.f * pad * sqrtf(hrt_dot3(ed, ed));
    bool out = true;
    for (int c = 0; c < 4; ++c) out = out && (sgn * hrt_dot3(p.c[c], e) < lim);
    if (out) return false;
  }
  return true;
}

/* ---- "sure" cells ----
 * Cell word: offset << 8 | HRT_RXMAP_SURE | length (<= HRT_RXMAP_MAX_LIST).  A cell is "sure" when its list has
 * exactly ONE triangle and every line through the apex inside the (enlarged) cell, followed AWAY from the apex,
 * meets that triangle well inside it and at a non-grazing angle.  The far side of a shadow query (beyond the
 * receiver) then needs no exact test: the reference's test is certain to accept the triangle, no other
 * triangle is listed, and of the hit only its existence, its triangle and "t > 1" matter to the caller (t is
 * at least the distance to the receiver).  "Well inside": all four corner rays of the cell meet the plane in
 * front of the apex with cosine >= c_min >= 0.01 to the normal, at barycentric coordinates (u, v) with
 *   u, v >= m,  u + v <= 1 - m,   m = 0.01 + 8 x (e_fp + e_pad)
 * (the image of the cell on the plane is the convex hull of the four points).  e_fp bounds the fp32 error of
 * the reference's u and v (src/compute_paths.c:264-271) for origins up to `reach` metres from the triangle:
 * u = (s . p) / det with p = d x e2, s = o - a; rounding leaves |err(s . p)| <= 3.6e-7 |s| |e2| + 4e-6 |e2|
 * (three products and two sums at 6e-8 relative each, plus the subtraction that formed s at coordinates of
 * ~100 m) and |det| = |d . n^| |e1 x e2| >= c_min |e1| |e2| sin(phi), so err(u) <= (3.6e-7 reach + 4e-6) /
 * (c_min h_min), h_min the triangle's smallest height; the same for v.  e_pad = pad / (c_min h_min) is the
 * shift of the hit point when the ray misses the apex by pad (the fp32 direction normalize(R - o), see the
 * file header).  Nothing is "sure" when the apex is within 64 pad of the plane or m >= 0.25. */
#define HRT_RXMAP_SURE 0x80u
#define HRT_RXMAP_MAX_LIST 127u

HRT_HD bool hrt_rxmap_sure(const HrtPyramid &p, V3 va, V3 vb, V3 vc, float pad, float reach)
{
  const V3 e1 = v3(vb.x - va.x, vb.y - va.y, vb.z - va.z), e2 = v3(vc.x - va.x, vc.y - va.y, vc.z - va.z);
  const V3 e3 = v3(vc.x - vb.x, vc.y - vb.y, vc.z - vb.z);
  const V3 n = v3(e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x);
  const float nl2 = hrt_dot3(n, n), nl = sqrtf(nl2);
  if (!(nl > 1e-3f)) return false;                          /* |det| = |d . n^| nl stays far above FLT_EPSILON */
  float h = hrt_dot3(va, n) / nl, sg = 1.f;                 /* distance of the plane from the apex */
  if (h < 0.f) { h = -h; sg = -1.f; }
  const float lmax = sqrtf(fmaxf(hrt_dot3(e1, e1), fmaxf(hrt_dot3(e2, e2), hrt_dot3(e3, e3))));
  const float hmin = nl / lmax;
  float cmin = 2.f;
  for (int k = 0; k < 4; ++k) cmin = fminf(cmin, sg * hrt_dot3(p.c[k], n) / nl);
  if (!(cmin >= 0.01f) || !(h > 64.f * pad)) return false;
  const float m = 0.01f + 8.f * ((3.6e-7f * reach + 4e-6f) + pad) / (cmin * hmin);
  if (!(m < 0.25f)) return false;
  for (int k = 0; k < 4; ++k) {
    const float c = sg * hrt_dot3(p.c[k], n) / nl;
    const float tq = h / c;
    const V3 w = v3(p.c[k].x * tq - va.x, p.c[k].y * tq - va.y, p.c[k].z * tq - va.z);
    const V3 we2 = v3(w.y * e2.z - w.z * e2.y, w.z * e2.x - w.x * e2.z, w.x * e2.y - w.y * e2.x);
    const V3 e1w = v3(e1.y * w.z - e1.z * w.y, e1.z * w.x - e1.x * w.z, e1.x * w.y - e1.y * w.x);
    const float u = hrt_dot3(we2, n) / nl2, v = hrt_dot3(e1w, n) / nl2;
    if (!(u >= m && v >= m && u + v <= 1.f - m)) return false;
  }
  return true;
}

/* corners of triangle record (q0, q1, q2) relative to apex r */
HRT_HD void hrt_rxmap_corners(float4 q0, float4 q1, float4 q2, V3 r, V3 *va, V3 *vb, V3 *vc)
{
  const V3 a = v3(q0.x - r.x, q0.y - r.y, q0.z - r.z);
  *va = a;
  *vb = v3(a.x + q0.w, a.y + q1.x, a.z + q1.y);
  *vc = v3(a.x + q1.z, a.y + q1.w, a.z + q2.x);
}

/* ---- depth bounds of a list item ----
 * Every item of a cell's list carries conservative bounds [d_lo, d_hi] on the
 * distance from the receiver at which a line inside the cell can touch the
 * triangle, quantised to 8 bits each with a per-receiver step.  A query walking
 * from the hit point o (at distance D from the receiver) towards the receiver
 *   - skips items that lie entirely behind o (d_lo > D): they could only give t < 0;
 *   - visits the rest nearest-to-o first (lists are sorted by d_hi, descending) and
 *     stops as soon as the best hit so far is closer to o than anything left.
 * Item word: slot | q_hi << 16 | q_lo << 24.  Bounds: the triangle as a whole
 * (exact point-triangle distance, farthest corner), tightened by the distances at
 * which the cell's corner rays meet the triangle's plane when all four do so at a
 * non-grazing angle. */
HRT_HD float hrt_point_tri_dist2(V3 a, V3 b, V3 c)     /* squared distance of the origin from triangle (a, b, c) (Ericson, RTCD 5.1.5) */
{
  const V3 ab = v3(b.x - a.x, b.y - a.y, b.z - a.z), ac = v3(c.x - a.x, c.y - a.y, c.z - a.z);
  const V3 ap = v3(-a.x, -a.y, -a.z);
  const float d1 = hrt_dot3(ab, ap), d2 = hrt_dot3(ac, ap);
  if (d1 <= 0.f && d2 <= 0.f) return hrt_dot3(a, a);
  const V3 bp = v3(-b.x, -b.y, -b.z);
  const float d3 = hrt_dot3(ab, bp), d4 = hrt_dot3(ac, bp);
  if (d3 >= 0.f && d4 <= d3) return hrt_dot3(b, b);
  const float vc = d1 * d4 - d3 * d2;
  if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { const float v = d1 / (d1 - d3); const V3 q = v3(a.x + v * ab.x, a.y + v * ab.y, a.z + v * ab.z); return hrt_dot3(q, q); }
  const V3 cp = v3(-c.x, -c.y, -c.z);
  const float d5 = hrt_dot3(ab, cp), d6 = hrt_dot3(ac, cp);
  if (d6 >= 0.f && d5 <= d6) return hrt_dot3(c, c);
  const float vb = d5 * d2 - d1 * d6;
  if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { const float w = d2 / (d2 - d6); const V3 q = v3(a.x + w * ac.x, a.y + w * ac.y, a.z + w * ac.z); return hrt_dot3(q, q); }
  const float va = d3 * d6 - d5 * d4;
  if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
    const float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    const V3 q = v3(b.x + w * (c.x - b.x), b.y + w * (c.y - b.y), b.z + w * (c.z - b.z));
    return hrt_dot3(q, q);
  }
  const float den = 1.f / (va + vb + vc), v = vb * den, w = vc * den;
  const V3 q = v3(a.x + ab.x * v + ac.x * w, a.y + ab.y * v + ac.y * w, a.z + ab.z * v + ac.z * w);
  return hrt_dot3(q, q);
}

/* conservative [d_lo, d_hi] for triangle (va, vb, vc: corners relative to the apex) inside pyramid p
 * of a G-cell map; pad as in hrt_rxmap_overlap */
HRT_HD void hrt_rxmap_depth(const HrtPyramid &p, V3 va, V3 vb, V3 vc, float pad, uint32_t G, float *d_lo, float *d_hi)
{
  float lo = sqrtf(fmaxf(hrt_point_tri_dist2(va, vb, vc), 0.f)) * 0.9999f - pad;
  float hi = sqrtf(fmaxf(hrt_dot3(va, va), fmaxf(hrt_dot3(vb, vb), hrt_dot3(vc, vc)))) * 1.0001f + pad;
  const V3 e1 = v3(vb.x - va.x, vb.y - va.y, vb.z - va.z), e2 = v3(vc.x - va.x, vc.y - va.y, vc.z - va.z);
  V3 n = v3(e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x);
  const float nl = sqrtf(hrt_dot3(n, n));
  if (nl > 0.f) {
    n = v3(n.x / nl, n.y / nl, n.z / nl);
    float h = hrt_dot3(va, n);                       /* signed distance of the plane from the apex */
    if (h < 0.f) { h = -h; n = v3(-n.x, -n.y, -n.z); }
    float cmin = 2.f, cmax = -2.f;
    for (int k = 0; k < 4; ++k) { const float c = hrt_dot3(p.c[k], n); cmin = fminf(cmin, c); cmax = fmaxf(cmax, c); }
    if (cmin > 0.05f && h > 8.f * pad) {
      /* all four corner rays meet the plane in front of the apex, none at a grazing angle: inside the
       * pyramid the plane is met between h / (largest cosine) and h / (smallest cosine); the largest
       * cosine inside the cell exceeds the corners' by at most the cell's angular diameter (< 3 / G) */
      const float far_ = (h + pad) / cmin * 1.0001f;
      const float near_ = (h - pad) / fminf(1.f, cmax + 3.f / (float)G) * 0.9999f;
      hi = fminf(hi, far_); lo = fmaxf(lo, near_);
    }
  }
  *d_lo = fmaxf(lo, 0.f); *d_hi = hi;
}

/* item word from a slot and its bounds; inv_step = 255 / (largest distance in the receiver's map) */
HRT_HD uint32_t hrt_rxmap_item(uint32_t slot, float d_lo, float d_hi, float inv_step)
{
  const float ql = floorf(fminf(fmaxf(d_lo * inv_step, 0.f), 255.f));
  const float qh = ceilf(fminf(fmaxf(d_hi * inv_step, 0.f), 255.f));
  return slot | ((uint32_t)qh << 16) | ((uint32_t)ql << 24);
}

/* 255 / (distance from the apex to the farthest corner of the scene's bounding box, plus a metre) */
HRT_HD float hrt_rxmap_inv_step(V3 apex, V3 lo, V3 hi)
{
  const float dx = fmaxf(fabsf(lo.x - apex.x), fabsf(hi.x - apex.x)), dy = fmaxf(fabsf(lo.y - apex.y), fabsf(hi.y - apex.y));
  const float dz = fmaxf(fabsf(lo.z - apex.z), fabsf(hi.z - apex.z));
  return 255.f / (sqrtf(dx * dx + dy * dy + dz * dz) + 1.f);
}

/* a cell's items by d_hi, descending (insertion sort: the lists are a handful of items) */
HRT_HD void hrt_rxmap_sort_items(uint32_t *it, uint32_t n)
{
  for (uint32_t i = 1; i < n; ++i) {
    const uint32_t w = it[i], key = (w >> 16) & 255u;
    uint32_t j = i;
    while (j > 0 && ((it[j - 1] >> 16) & 255u) < key) { it[j] = it[j - 1]; --j; }
    it[j] = w;
  }
}

/* Query-side thresholds in quantised units.  Margins: one quantisation step is already in the
 * floor / ceil of the item bounds; on top 5 cm + 0.5 % of the distance cover the fp32 error of a
 * computed t (up to ~1e-2 m for rays within 1e-3 rad of a plane; phantom hits excepted as always). */
struct HrtMapDepth { float dist, inv_step; int q_behind; };
HRT_HD HrtMapDepth hrt_rxmap_query_depth(float dist, float inv_step)
{
  HrtMapDepth m; m.dist = dist; m.inv_step = inv_step;
  const float q = fminf(HRT_FMA(dist, 1.005f, 0.05f) * inv_step, 300.f);        /* q_lo above this: entirely behind the hit point */
#if defined(__CUDA_ARCH__)
  m.q_behind = __float2int_ru(q);
#else
  m.q_behind = (int)ceilf(q);
#endif
  return m;
}
/* after a hit at distance t from o (between o and the receiver): items whose q_hi is below the result
 * lie entirely farther from o than that hit */
HRT_HD int hrt_rxmap_stop(const HrtMapDepth &m, float t)
{
  const float s = HRT_FMA(m.dist - t, 0.995f, -0.05f);           /* distance of the hit from the receiver, less the margin */
  const float q = fminf(s * m.inv_step, 300.f);                  /* s <= 0: q <= 0, nothing is below it */
#if defined(__CUDA_ARCH__)
  return __float2int_rd(q);
#else
  return (int)floorf(q);
#endif
}
