/* hrt_core.cuh -- per-ray arithmetic of compute_paths(), written once as
 * __host__ __device__ code.
 *
 * The product only ever runs it on the GPU (kernels in hrt_cuda.cu).  The host
 * instantiation exists for one purpose: tests/emul/ compiles it into a serial
 * CPU driver so the kernel logic can be checked against the oracle in the
 * GPU-less authoring container.  No product entry point reaches a host path.
 *
 * Exactness contract: everything that decides WHICH
 * triangle is hit, and the whole delay/direction chain, uses separately
 * rounded fp32 + - * / sqrt in the reference's operation order
 * (inc/vec3.h:10-43, src/compute_paths.c:259-276, :650-659).  Those operations
 * are spelled with the HRT_* macros below, which map to the __f*_rn intrinsics
 * on the device -- nvcc never contracts those into FMA, whatever -fmad says.
 * FMA is used only where a conservative answer suffices (ray/box culling).
 */
#pragma once

#include <float.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define HRT_HD __host__ __device__ __forceinline__
#else
#define HRT_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define HRT_MUL(a, b) __fmul_rn((a), (b))
#define HRT_ADD(a, b) __fadd_rn((a), (b))
#define HRT_SUB(a, b) __fsub_rn((a), (b))
#define HRT_DIV(a, b) __fdiv_rn((a), (b))
#define HRT_SQRT(a)   __fsqrt_rn((a))
#define HRT_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
#define HRT_MUL(a, b) ((a) * (b))
#define HRT_ADD(a, b) ((a) + (b))
#define HRT_SUB(a, b) ((a) - (b))
#define HRT_DIV(a, b) ((a) / (b))
#define HRT_SQRT(a)   sqrtf((a))
#define HRT_FMA(a, b, c) fmaf((a), (b), (c))
#endif

#define HRT_PI    3.14159265358979323846f  /* reference src/compute_paths.c:18 */
#define HRT_C0    299792458.0f             /* reference src/compute_paths.c:19 */
#define HRT_EPS   FLT_EPSILON              /* reference :251, :263-275 */
#define HRT_T_MAX 1e9f                     /* reference :251 */
#define HRT_NONE  0xFFFFFFFFu              /* query found nothing */
#define HRT_IDLE  0xFFFFFFFEu              /* ray already dead, not traced */

#if !defined(__CUDACC__)
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
#endif

struct V3 { float x, y, z; };

HRT_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
/* reference inc/vec3.h:10-43, same operand order */
HRT_HD V3 v3_sub(V3 a, V3 b) { return v3(HRT_SUB(a.x, b.x), HRT_SUB(a.y, b.y), HRT_SUB(a.z, b.z)); }
HRT_HD V3 v3_add(V3 a, V3 b) { return v3(HRT_ADD(a.x, b.x), HRT_ADD(a.y, b.y), HRT_ADD(a.z, b.z)); }
HRT_HD V3 v3_scale(V3 a, float s) { return v3(HRT_MUL(a.x, s), HRT_MUL(a.y, s), HRT_MUL(a.z, s)); }
HRT_HD float v3_dot(V3 a, V3 b)
{ return HRT_ADD(HRT_ADD(HRT_MUL(a.x, b.x), HRT_MUL(a.y, b.y)), HRT_MUL(a.z, b.z)); }
HRT_HD V3 v3_cross(V3 a, V3 b)
{
  return v3(HRT_SUB(HRT_MUL(a.y, b.z), HRT_MUL(a.z, b.y)),
            HRT_SUB(HRT_MUL(a.z, b.x), HRT_MUL(a.x, b.z)),
            HRT_SUB(HRT_MUL(a.x, b.y), HRT_MUL(a.y, b.x)));
}
HRT_HD V3 v3_normalize(V3 a)
{
  float len = HRT_SQRT(v3_dot(a, a));
  return v3(HRT_DIV(a.x, len), HRT_DIV(a.y, len), HRT_DIV(a.z, len));
}

/* -------------------------------------------------------------------------
 * Scene records in HBM / shared memory.
 *
 * Triangle record, 48 B = 3 x float4, in BVH leaf order:
 *   q0 = (a.x, a.y, a.z, ab.x)   a  = first corner
 *   q1 = (ab.y, ab.z, ac.x, ac.y) ab = b - a, ac = c - a  (the reference forms
 *   q2 = (ac.z, n.x, n.y, n.z)        them per test, :259-260; same rounding)
 * n is the unit normal of reference :208-224.
 *
 * BVH2 node, 64 B = 4 x float4, children's boxes stored in the parent, one
 * 32-byte record per child (a child's box and ref are fetched with two loads):
 *   n0 = (L.lo.x, L.hi.x, L.lo.y, L.hi.y)   n1 = (L.lo.z, L.hi.z, L.ref, -)
 *   n2 = (R.lo.x, R.hi.x, R.lo.y, R.hi.y)   n3 = (R.lo.z, R.hi.z, R.ref, -)
 * ref >= 0: inner node index.  ref < 0: leaf, ~ref = (first_slot << 3) | (count-1).
 * ------------------------------------------------------------------------- */
#define HRT_LEAF_MAX_CAP 8

HRT_HD int   hrt_leaf_ref(uint32_t first, uint32_t count) { return ~(int)((first << 3) | (count - 1u)); }
HRT_HD float hrt_int_as_float(int v)
{
#if defined(__CUDA_ARCH__)
  return __int_as_float(v);
#else
  float f; memcpy(&f, &v, 4); return f;
#endif
}
HRT_HD int hrt_float_as_int(float f)
{
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  int v; memcpy(&v, &f, 4); return v;
#endif
}

/* Work counters for the roofline accounting: box tests and
 * Moeller-Trumbore tests by the stage they reach (A: det, B: u, C: v, D: t).
 * HrtNoCount compiles to nothing; HrtCount is used by the instrumented kernel
 * variants only. */
struct HrtNoCount {
  HRT_HD void box(uint32_t) {}
  HRT_HD void tri(int) {}
};
struct HrtCount {
  uint32_t c[5];
  HRT_HD void box(uint32_t n) { c[0] += n; }
  HRT_HD void tri(int stage) { c[1 + stage] += 1u; }
};

struct HrtHit {
  float    t;      /* distance, reference *t                           */
  uint32_t gid;    /* triangle id in (mesh, face) order; HRT_NONE: miss */
  uint32_t slot;   /* position of that triangle in BVH leaf order       */
};

/* One Moeller-Trumbore test with the reference's accept/reject decisions
 * (src/compute_paths.c:259-276), against the running best (t, gid).
 *
 * The quotients u, v, t are only formed when a division-free bound cannot
 * already prove the reference's test rejects: |N/det| compared against
 * thresholds that sit 1e-5 away from the reference's, while IEEE division is
 * monotonic and off by at most 2^-24 relative -- so the shortcut never changes
 * a decision, it only skips divisions. */
/* TIE_ONLY_T: accept on t <= best and leave the (t == best) tie to the caller
 * (who then looks up the triangle id): saves fetching the id of every candidate. */
template <class Cnt, bool TIE_ONLY_T = false>
HRT_HD bool hrt_mt_test(float4 q0, float4 q1, float4 q2, V3 o, V3 d,
                        float best, uint32_t best_gid, uint32_t gid, float *t_out, Cnt &cnt)
{
  cnt.tri(0);
  const V3 a  = v3(q0.x, q0.y, q0.z);
  const V3 ab = v3(q0.w, q1.x, q1.y);
  const V3 ac = v3(q1.z, q1.w, q2.x);
  const V3 pv = v3_cross(d, ac);                         /* :261 */
  const float det = v3_dot(ab, pv);                      /* :262 */
  if (det > -HRT_EPS && det < HRT_EPS) return false;     /* :263 (NaN passes, as there) */
  cnt.tri(1);
  /* division-free bounds on the quotients N/det: N*det against multiples of
   * det^2 (|det| >= 1.2e-7 here, so neither product under- or overflows) */
  const float ad  = HRT_MUL(det, det);
  const float lo  = HRT_MUL(ad, -1e-5f), hi = HRT_MUL(ad, 1.00001f);
  const V3 sv = v3_sub(o, a);                            /* :264 */
  const float nu = v3_dot(sv, pv);
  const float snu = HRT_MUL(nu, det);
  if (snu < lo || snu > hi) return false;                /* u clearly outside */
  /* clearly inside (1e-5 away from both ends, the quotient being off by 2^-24
   * at most): the reference's u test passes, no need to form u yet */
  const float in_lo = HRT_MUL(ad, 1e-5f), in_hi = HRT_MUL(ad, 0.99999f);
  const bool u_sure = snu > in_lo && snu < in_hi;
  float u = 0.f;
  if (!u_sure) {
    u = HRT_DIV(nu, det);                                /* :265 */
    if ((u < 0.f && -u > HRT_EPS) || (u > 1.f && HRT_SUB(u, 1.f) > HRT_EPS)) return false; /* :266 */
  }
  cnt.tri(2);
  const V3 qv = v3_cross(sv, ab);                        /* :269 */
  const float nv = v3_dot(d, qv);
  const float snv = HRT_MUL(nv, det);
  if (snv < lo || snv > hi) return false;                /* v clearly outside */
  if (!(u_sure && snv > in_lo && HRT_ADD(snu, snv) < in_hi)) {
    if (u_sure) u = HRT_DIV(nu, det);
    const float v = HRT_DIV(nv, det);                    /* :270 */
    const float uv = HRT_ADD(u, v);
    if ((v < 0.f && -v > HRT_EPS) || (uv > 1.f && HRT_SUB(uv, 1.f) > HRT_EPS)) return false; /* :271 */
  }
  cnt.tri(3);
  const float nt = v3_dot(ac, qv);
  const float snt = HRT_MUL(nt, det);
  if (!(snt > 0.f) && nt == nt) return false;            /* t <= 0 */
  if (snt > HRT_MUL(HRT_MUL(best, ad), 1.00001f)) return false;   /* clearly behind the best hit so far */
  const float t = HRT_DIV(nt, det);                      /* :274 */
  if (!(t > HRT_EPS)) return false;                      /* :275 */
  if (TIE_ONLY_T) { if (t <= best) { *t_out = t; return true; } return false; }
  if (t < best || (t == best && gid < best_gid)) { *t_out = t; return true; }
  return false;
}

/* The triangle a shadow ray starts on is a candidate of every one of its queries, and nearly always a
 * miss by "t <= 0" (the ray leaves the surface; it hits only when the receiver lies behind it).  That
 * decision needs, of the ray's direction, only the sign of det: t's numerator nt = ac . ((o - a) x ab)
 * (:269, :274) depends on the origin alone.  hrt_mt_self_nt forms it once per hit point with the very
 * operations of hrt_mt_test; hrt_mt_self_miss then answers most queries from stage A (p = d x ac, det)
 * alone: true = hrt_mt_test is certain to return false (its det and t <= 0 exits, which no earlier exit
 * can turn into a hit); false = undecided, run hrt_mt_test. */
HRT_HD float hrt_mt_self_nt(float4 q0, float4 q1, float4 q2, V3 o)
{
  const V3 a  = v3(q0.x, q0.y, q0.z);
  const V3 ab = v3(q0.w, q1.x, q1.y);
  const V3 ac = v3(q1.z, q1.w, q2.x);
  const V3 sv = v3_sub(o, a);
  const V3 qv = v3_cross(sv, ab);
  return v3_dot(ac, qv);
}

template <class Cnt>
HRT_HD bool hrt_mt_self_miss(float4 q0, float4 q1, float4 q2, V3 d, float nt, Cnt &cnt)
{
  const V3 ab = v3(q0.w, q1.x, q1.y);
  const V3 ac = v3(q1.z, q1.w, q2.x);
  const V3 pv = v3_cross(d, ac);
  const float det = v3_dot(ab, pv);
  if (det > -HRT_EPS && det < HRT_EPS) { cnt.tri(0); return true; }
  const float snt = HRT_MUL(nt, det);
  if (!(snt > 0.f) && nt == nt) { cnt.tri(0); return true; }
  return false;
}

/* Ray/box culling state.  Conservative by construction: boxes are padded by
 * the builder, the far bound carries slack, NaNs never reject. */
struct HrtRayCull {
  V3 inv;    /* 1/d with tiny components clamped */
  V3 ood;    /* -o * inv */
};

HRT_HD float hrt_safe_inv(float d)
{
  const float tiny = 1e-20f;
  if (fabsf(d) < tiny) d = (hrt_float_as_int(d) < 0) ? -tiny : tiny;
#if defined(__CUDA_ARCH__)
  float r;                                   /* MUFU.RCP, 1 ulp: culling only, covered by the box padding */
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r;
#else
  return 1.0f / d;
#endif
}

HRT_HD HrtRayCull hrt_ray_cull(V3 o, V3 d)
{
  HrtRayCull c;
  c.inv = v3(hrt_safe_inv(d.x), hrt_safe_inv(d.y), hrt_safe_inv(d.z));
  c.ood = v3(-o.x * c.inv.x, -o.y * c.inv.y, -o.z * c.inv.z);
  return c;
}

/* entry distance of the ray into [lo,hi]; returns false when the slab interval
 * is empty within [0, tmax] */
HRT_HD bool hrt_slab(const HrtRayCull &c, float lox, float hix, float loy, float hiy,
                     float loz, float hiz, float tmax, float *t_near)
{
  const float x0 = HRT_FMA(lox, c.inv.x, c.ood.x), x1 = HRT_FMA(hix, c.inv.x, c.ood.x);
  const float y0 = HRT_FMA(loy, c.inv.y, c.ood.y), y1 = HRT_FMA(hiy, c.inv.y, c.ood.y);
  const float z0 = HRT_FMA(loz, c.inv.z, c.ood.z), z1 = HRT_FMA(hiz, c.inv.z, c.ood.z);
  const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.f));
  const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fminf(fmaxf(z0, z1), tmax));
  *t_near = tn;
  return tn <= tf;
}

/* direction octant of a ray: bit k set when component k of 1/d is negative */
HRT_HD uint32_t hrt_octant(const HrtRayCull &c)
{
  return ((uint32_t)hrt_float_as_int(c.inv.x) >> 31) | (((uint32_t)hrt_float_as_int(c.inv.y) >> 31) << 1) |
         (((uint32_t)hrt_float_as_int(c.inv.z) >> 31) << 2);
}

/* Scene view handed to the traversal: how to fetch node / triangle words.
 * `Mem` provides node(i,k) and tri(slot,k) returning float4 -- from shared
 * memory when the scene was staged there, else from global memory. */
struct HrtGlobalMem {
  const float4 *nodes;
  const float4 *tris;
  HRT_HD float4 node(int i, int k) const
  {
#if defined(__CUDA_ARCH__)
    return __ldg(&nodes[4 * i + k]);
#else
    return nodes[4 * i + k];
#endif
  }
  HRT_HD float4 tri(uint32_t s, int k) const
  {
#if defined(__CUDA_ARCH__)
    return __ldg(&tris[3 * s + k]);
#else
    return tris[3 * s + k];
#endif
  }
  /* 4-wide nodes (hrt_bvh.cuh): float4 k of wide node i of the selected octant copy */
  const float4 *wnodes;
  uint32_t wstride = 7u;       /* float4s from one wide node to the next: 7, or 8 (128-byte aligned nodes, scenes read from global memory) */
  HRT_HD float4 wide(int i, int k) const
  {
#if defined(__CUDA_ARCH__)
    return __ldg(&wnodes[(size_t)wstride * (size_t)i + k]);
#else
    return wnodes[(size_t)wstride * (size_t)i + k];
#endif
  }
  HRT_HD void select_wide_octant(uint32_t oct, size_t stride4)
  {
    wnodes += (size_t)oct * stride4;
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+l"(wnodes));
#endif
  }
};

#define HRT_STACK 64
#if defined(__CUDACC__)
struct __align__(8) HrtStackEntry { int ref; float tn; };
#else
struct HrtStackEntry { int ref; float tn; };
#endif

/* Closest hit over the BINARY tree as built (plain node copy) == the reference's
 * loop over every triangle (moeller_trumbore, :237-287): minimum t, ties to the
 * lowest (mesh, face).  The kernels walk the 4-wide tree collapsed from it
 * (hrt_closest_hit_wide); this walk is kept as the independent cross-check of
 * the collapse that tests/emul runs against the oracle.
 * root_ref: inner node index, a leaf ref, or 0 with num_tris == 0. */
template <class Mem, class Gid, class Cnt>
HRT_HD HrtHit hrt_closest_hit(const Mem &mem, const Gid tri_gid, int root_ref,
                              uint32_t num_tris, V3 o, V3 d, Cnt &cnt)
{
  HrtHit h; h.t = HRT_T_MAX; h.gid = HRT_NONE; h.slot = HRT_NONE;
  if (num_tris == 0) return h;
  const HrtRayCull c = hrt_ray_cull(o, d);
  float tmax = HRT_T_MAX * 1.0001f;          /* far bound with slack */
  HrtStackEntry stack[HRT_STACK];
  int sp = 0, cur = root_ref;
  for (;;) {
    while (cur >= 0) {
      const float4 n0 = mem.node(cur, 0), n1 = mem.node(cur, 1), n2 = mem.node(cur, 2), n3 = mem.node(cur, 3);
      float tl, tr;
      const bool hl = hrt_slab(c, n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, tmax, &tl);
      const bool hr = hrt_slab(c, n2.x, n2.y, n2.z, n2.w, n3.x, n3.y, tmax, &tr);
      const int rl = hrt_float_as_int(n1.z), rr = hrt_float_as_int(n3.z);
      cnt.box(2u);
      if (hl && hr) {
        const bool left_first = tl <= tr;
        HrtStackEntry e;
        e.ref = left_first ? rr : rl; e.tn = left_first ? tr : tl;
        stack[sp++] = e;
        cur = left_first ? rl : rr;
      } else if (hl) cur = rl;
      else if (hr) cur = rr;
      else {
        bool got = false;
        while (sp > 0) { const HrtStackEntry e = stack[--sp]; if (e.tn <= tmax) { cur = e.ref; got = true; break; } }
        if (!got) return h;
      }
    }
    const uint32_t code = (uint32_t)~cur, first = code >> 3, ntri = (code & 7u) + 1u;
    for (uint32_t k = 0; k < ntri; ++k) {
      const uint32_t s = first + k;
      float t;
      const uint32_t gid = tri_gid[s];
      if (hrt_mt_test(mem.tri(s, 0), mem.tri(s, 1), mem.tri(s, 2), o, d, h.t, h.gid, gid, &t, cnt)) {
        h.t = t; h.gid = gid; h.slot = s;
        tmax = HRT_FMA(t, 1.0001f, 1e-30f);
      }
    }
    bool got = false;
    while (sp > 0) { const HrtStackEntry e = stack[--sp]; if (e.tn <= tmax) { cur = e.ref; got = true; break; } }
    if (!got) return h;
  }
}

/* Closest hit over the 4-wide tree (hrt_bvh.cuh, "4-wide nodes"): the same
 * minimum over (t, triangle id) as the reference's loop over every triangle.
 * One visit = seven 16-byte loads, four slab tests, no ordering decision (the
 * octant copy stores the children front to back); the children that are hit are
 * pushed in reverse order and the nearest is popped straight away.
 * SORTED: `mem.wnodes` holds 8 octant copies `oct_stride4` float4s apart; else a
 * single plain copy (lo/hi per axis, min/max in the slab test).
 * root_ref: wide node index, a leaf ref, or anything with num_tris == 0. */
#define HRT_WSTACK 96   /* 3 pushes per level, depth <= (32 + log2 n) / 2 + 1 */
#define HRT_WIDE_EMPTY ((int)0x80000000)   /* ref of an unused child slot */
/* SELF: the ray starts on triangle slot `self_slot` (a shadow ray leaving a hit point); self_nt = hrt_mt_self_nt
 * of that triangle and origin: its test is decided from stage A alone whenever possible (hrt_mt_self_miss). */
template <bool SORTED, bool SELF = false, class Mem, class Gid, class Cnt>
HRT_HD HrtHit hrt_closest_hit_wide(const Mem &mem_in, const Gid tri_gid, int root_ref,
                                   uint32_t num_tris, V3 o, V3 d, Cnt &cnt, size_t oct_stride4 = 0,
                                   uint32_t self_slot = HRT_NONE, float self_nt = 0.f)
{
  HrtHit h; h.t = HRT_T_MAX; h.gid = HRT_NONE; h.slot = HRT_NONE;
  if (num_tris == 0) return h;
  HrtRayCull c = hrt_ray_cull(o, d);
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+f"(c.ood.x), "+f"(c.ood.y), "+f"(c.ood.z));
#endif
  Mem mem = mem_in;
  if (SORTED) mem.select_wide_octant(hrt_octant(c), oct_stride4);
  float tmax = HRT_T_MAX * 1.0001f;
  HrtStackEntry stack[HRT_WSTACK];
  int sp = 0, cur = root_ref;
  for (;;) {
    while (cur >= 0) {
      /* axis by axis, so that at most two plane vectors are live next to the running
       * entry / exit distances of the four children */
      float tn[4], tf[4]; bool hit[4]; int ref[4];
      {
        const float4 n4 = mem.wide(cur, 0), f4 = mem.wide(cur, 1);
        if (SORTED) {
          tn[0] = fmaxf(HRT_FMA(n4.x, c.inv.x, c.ood.x), 0.f); tn[1] = fmaxf(HRT_FMA(n4.y, c.inv.x, c.ood.x), 0.f);
          tn[2] = fmaxf(HRT_FMA(n4.z, c.inv.x, c.ood.x), 0.f); tn[3] = fmaxf(HRT_FMA(n4.w, c.inv.x, c.ood.x), 0.f);
          tf[0] = fminf(HRT_FMA(f4.x, c.inv.x, c.ood.x), tmax); tf[1] = fminf(HRT_FMA(f4.y, c.inv.x, c.ood.x), tmax);
          tf[2] = fminf(HRT_FMA(f4.z, c.inv.x, c.ood.x), tmax); tf[3] = fminf(HRT_FMA(f4.w, c.inv.x, c.ood.x), tmax);
        } else {
#define HRT_WIDE_AX0(K, M) { const float a = HRT_FMA(n4.M, c.inv.x, c.ood.x), b = HRT_FMA(f4.M, c.inv.x, c.ood.x); \
                             tn[K] = fmaxf(fminf(a, b), 0.f); tf[K] = fminf(fmaxf(a, b), tmax); }
          HRT_WIDE_AX0(0, x) HRT_WIDE_AX0(1, y) HRT_WIDE_AX0(2, z) HRT_WIDE_AX0(3, w)
#undef HRT_WIDE_AX0
        }
      }
#define HRT_WIDE_AX(K, M, INV, OOD)                                                                  \
      { const float a = HRT_FMA(n4.M, INV, OOD), b = HRT_FMA(f4.M, INV, OOD);                           \
        if (SORTED) { tn[K] = fmaxf(tn[K], a); tf[K] = fminf(tf[K], b); }                              \
        else { tn[K] = fmaxf(tn[K], fminf(a, b)); tf[K] = fminf(tf[K], fmaxf(a, b)); } }
      {
        const float4 n4 = mem.wide(cur, 2), f4 = mem.wide(cur, 3);
        HRT_WIDE_AX(0, x, c.inv.y, c.ood.y) HRT_WIDE_AX(1, y, c.inv.y, c.ood.y)
        HRT_WIDE_AX(2, z, c.inv.y, c.ood.y) HRT_WIDE_AX(3, w, c.inv.y, c.ood.y)
      }
      {
        const float4 n4 = mem.wide(cur, 4), f4 = mem.wide(cur, 5);
        HRT_WIDE_AX(0, x, c.inv.z, c.ood.z) HRT_WIDE_AX(1, y, c.inv.z, c.ood.z)
        HRT_WIDE_AX(2, z, c.inv.z, c.ood.z) HRT_WIDE_AX(3, w, c.inv.z, c.ood.z)
      }
#undef HRT_WIDE_AX
      {
        const float4 rf = mem.wide(cur, 6);
        ref[0] = hrt_float_as_int(rf.x); ref[1] = hrt_float_as_int(rf.y);
        ref[2] = hrt_float_as_int(rf.z); ref[3] = hrt_float_as_int(rf.w);
      }
      for (int k = 0; k < 4; ++k) hit[k] = tn[k] <= tf[k] && (SORTED || ref[k] != HRT_WIDE_EMPTY);
      if (!SORTED) {
        /* single plain copy: its child order suits octant 0 only, so order the four by entry
         * distance here (misses last) -- a 5-exchange network; the octant copies need none of it */
#define HRT_WIDE_CSWAP(A, B)                                                                        \
        if ((hit[B] && !hit[A]) || (hit[A] && hit[B] && tn[B] < tn[A])) {                            \
          const float tt = tn[A]; tn[A] = tn[B]; tn[B] = tt;                                          \
          const int rr = ref[A]; ref[A] = ref[B]; ref[B] = rr;                                        \
          const bool hh = hit[A]; hit[A] = hit[B]; hit[B] = hh;                                       \
        }
        HRT_WIDE_CSWAP(0, 1) HRT_WIDE_CSWAP(2, 3) HRT_WIDE_CSWAP(0, 2) HRT_WIDE_CSWAP(1, 3) HRT_WIDE_CSWAP(1, 2)
#undef HRT_WIDE_CSWAP
      }
      cnt.box((uint32_t)(ref[0] != HRT_WIDE_EMPTY) + (uint32_t)(ref[1] != HRT_WIDE_EMPTY) +
              (uint32_t)(ref[2] != HRT_WIDE_EMPTY) + (uint32_t)(ref[3] != HRT_WIDE_EMPTY));
      /* far to near onto the stack, then take the nearest back off */
      if (hit[3]) { HrtStackEntry e; e.ref = ref[3]; e.tn = tn[3]; stack[sp++] = e; }
      if (hit[2]) { HrtStackEntry e; e.ref = ref[2]; e.tn = tn[2]; stack[sp++] = e; }
      if (hit[1]) { HrtStackEntry e; e.ref = ref[1]; e.tn = tn[1]; stack[sp++] = e; }
      if (hit[0]) { cur = ref[0]; continue; }
      bool got = false;
      while (sp > 0) {
        const HrtStackEntry e = stack[--sp];
        if (e.tn <= tmax) { cur = e.ref; got = true; break; }
      }
      if (!got) return h;
    }
    {
      const uint32_t code = (uint32_t)~cur;
      const uint32_t first = code >> 3, ntri = (code & 7u) + 1u;
      for (uint32_t k = 0; k < ntri; ++k) {
        const uint32_t s = first + k;
        float t;
        if (SELF && s == self_slot && hrt_mt_self_miss(mem.tri(s, 0), mem.tri(s, 1), mem.tri(s, 2), d, self_nt, cnt)) continue;
        const uint32_t gid = tri_gid[s];
        if (hrt_mt_test(mem.tri(s, 0), mem.tri(s, 1), mem.tri(s, 2), o, d, h.t, h.gid, gid, &t, cnt)) {
          h.t = t; h.gid = gid; h.slot = s;
          tmax = HRT_FMA(t, 1.0001f, 1e-30f);
        }
      }
    }
    /* pop, skipping subtrees that start beyond the current best */
    for (;;) {
      if (sp == 0) return h;
      const HrtStackEntry e = stack[--sp];
      if (e.tn <= tmax) { cur = e.ref; break; }
    }
  }
}

/* Brute force over every triangle in leaf order -- debug/validation path of
 * the kernels (HRT_FLAG_BRUTE_FORCE), same decisions by construction. */
template <class Mem, class Gid, class Cnt>
HRT_HD HrtHit hrt_closest_hit_brute(const Mem &mem, const Gid tri_gid,
                                    uint32_t num_tris, V3 o, V3 d, Cnt &cnt)
{
  HrtHit h; h.t = HRT_T_MAX; h.gid = HRT_NONE; h.slot = HRT_NONE;
  for (uint32_t s = 0; s < num_tris; ++s) {
    float t;
    const uint32_t gid = tri_gid[s];
    if (hrt_mt_test(mem.tri(s, 0), mem.tri(s, 1), mem.tri(s, 2), o, d, h.t, h.gid, gid, &t, cnt)) {
      h.t = t; h.gid = gid; h.slot = s;
    }
  }
  return h;
}

/* Folded incidence angle, reference :280-283: acos in double of the fp32 dot
 * product (normal first), rounded to float, then folded into [0, pi/2]. */
HRT_HD float hrt_theta_fold(V3 n, V3 d)
{
  float th = (float)acos((double)v3_dot(n, d));
  if ((double)th > (double)HRT_PI / 2.) th = HRT_SUB(HRT_PI, th);
  return th;
}

/* ------------------------------------------------------------- materials
 * Derived per-material constants, computed on the HOST with the host libm
 * (powf) exactly as the reference does (:171-206) and uploaded; the raw
 * scattering parameters of g_materials ride along. */
struct HrtMaterial {
  float eta_abs2, eta_abs_inv_sqrt;
  float sqrt_re, sqrt_im;
  float inv_re, inv_im;
  float r;            /* 1 - s */
  float s;            /* scattering coefficient */
  float s1_alpha;     /* (float)uint8 */
  float pad_[3];
};
struct HrtMaterialTable { HrtMaterial m[17]; };

HRT_HD void hrt_cdiv(float ar, float ai, float br, float bi, float *cr, float *ci)
{
  const float den = HRT_ADD(HRT_MUL(br, br), HRT_MUL(bi, bi));                 /* :161 */
  *cr = HRT_DIV(HRT_ADD(HRT_MUL(ar, br), HRT_MUL(ai, bi)), den);               /* :162 */
  *ci = HRT_DIV(HRT_SUB(HRT_MUL(ai, br), HRT_MUL(ar, bi)), den);               /* :163 */
}

/* reference refl_coefs :300-344; out = (te_re, te_im, tm_re, tm_im) */
HRT_HD void hrt_refl_coefs(const HrtMaterial &m, float th, float out[4])
{
  const float s1 = sinf(th);                                                   /* :310 */
  if (HRT_MUL(m.eta_abs_inv_sqrt, s1) > 1.f - HRT_EPS) {                       /* :311 */
    out[0] = out[2] = 1.f; out[1] = out[3] = 0.f; return;
  }
  const float s1sq = HRT_MUL(s1, s1);                                          /* :318 */
  const float c2r = HRT_SQRT(HRT_ADD(1.f, HRT_MUL(HRT_DIV(m.inv_re, m.eta_abs2), s1sq))); /* :319 */
  const float c2i = HRT_SQRT(HRT_SUB(1.f, HRT_MUL(HRT_DIV(m.inv_im, m.eta_abs2), s1sq))); /* :320 */
  const float pr = HRT_SUB(HRT_MUL(m.sqrt_re, c2r), HRT_MUL(m.sqrt_im, c2i));  /* :323 */
  const float pi = HRT_ADD(HRT_MUL(m.sqrt_re, c2i), HRT_MUL(m.sqrt_im, c2r));  /* :324 */
  const float c1 = cosf(th);                                                   /* :325 */
  hrt_cdiv(HRT_SUB(c1, pr), -pi, HRT_ADD(c1, pr), pi, &out[0], &out[1]);       /* :326 */
  const float qr = HRT_MUL(m.sqrt_re, c1), qi = HRT_MUL(m.sqrt_im, c1);        /* :331 */
  hrt_cdiv(HRT_SUB(qr, c2r), HRT_SUB(qi, c2i), HRT_ADD(qr, c2r), HRT_ADD(qi, c2i),
           &out[2], &out[3]);                                                  /* :333 */
  out[0] = HRT_MUL(out[0], m.r); out[1] = HRT_MUL(out[1], m.r);                /* :340 */
  out[2] = HRT_MUL(out[2], m.r); out[3] = HRT_MUL(out[3], m.r);
}

/* reference scat_coefs :359-415 */
HRT_HD void hrt_scat_coefs(const HrtMaterial &m, float th_s, float th_i, float out[4])
{
  const float cs = cosf(th_s), ci = cosf(th_i), si = sinf(th_i);               /* :372 */
  const float lobe = HRT_MUL(m.s, expf(HRT_MUL(-m.s1_alpha, fabsf(HRT_SUB(th_s, th_i))))); /* :378 */
  const float rough = HRT_DIV(1.0f, HRT_ADD(1.0f, m.s1_alpha));                /* :382 */
  const float spec = HRT_MUL(rough, cs);                                       /* :383 */
  const float diff = HRT_MUL(HRT_SUB(1.0f, rough), cs);                        /* :384 */
  float te_r = HRT_MUL(lobe, HRT_ADD(spec, diff));                             /* :388 */
  float tm_r = HRT_MUL(lobe, HRT_ADD(HRT_MUL(spec, ci), diff));                /* :390 */
  const float ph = HRT_MUL(HRT_MUL(m.s1_alpha, si), 0.1f);                     /* :394 */
  const float sp = sinf(ph);
  float te_i = HRT_MUL(te_r, sp);                                              /* :395 */
  float tm_i = HRT_MUL(tm_r, sp);                                              /* :396 */
  const float nrm = HRT_SQRT(HRT_ADD(HRT_ADD(HRT_ADD(HRT_MUL(te_r, te_r), HRT_MUL(te_i, te_i)),
                                             HRT_MUL(tm_r, tm_r)), HRT_MUL(tm_i, tm_i))); /* :399 */
  if (nrm > 1e-6f) {                                                           /* :401 */
    te_r = HRT_DIV(te_r, nrm); te_i = HRT_DIV(te_i, nrm);
    tm_r = HRT_DIV(tm_r, nrm); tm_i = HRT_DIV(tm_i, nrm);
  }
  out[0] = te_r; out[1] = te_i; out[2] = tm_r; out[3] = tm_i;
}

/* Per-hit constants of the scattering model (material only): hoisted out of
 * the receiver loop by the kernels. */
struct HrtScatConst { float rough, one_minus_rough, neg_alpha, alpha, s; };
HRT_HD HrtScatConst hrt_scat_const(const HrtMaterial &m)
{
  HrtScatConst c;
  c.rough = HRT_DIV(1.0f, HRT_ADD(1.0f, m.s1_alpha));                          /* :382 */
  c.one_minus_rough = HRT_SUB(1.0f, c.rough);
  c.neg_alpha = -m.s1_alpha; c.alpha = m.s1_alpha; c.s = m.s;
  return c;
}

/* ----------------------------------------------------- launch directions
 * Fibonacci sphere, reference :444-451.  fp32 index math, double trig rounded
 * to fp32.  CUDA's double acos/sin/cos are within 2 ulp, glibc's within 1, so
 * after rounding to fp32 the two can differ only when the double value sits
 * within ~2^-50 (relative) of an fp32 rounding boundary.  `*ambiguous` is set
 * when a value is within 2^-46 of one; the host then recomputes those few rays
 * (about 1 in 5e5) with its own libm, which makes launch directions bit-exact
 * for every ray (hrt_cuda.cu, fix_ambiguous_dirs). */
HRT_HD float hrt_round_checked(double x, bool *ambiguous)
{
  const float f = (float)x;
  const double k = 1.4210854715202004e-14;  /* 2^-46 */
  if ((float)(x * (1.0 - k)) != f || (float)(x * (1.0 + k)) != f) *ambiguous = true;
  return f;
}

HRT_HD V3 hrt_launch_dir(uint64_t path, uint64_t num_paths, bool *ambiguous)
{
  const float k = HRT_ADD((float)path, .5f);                                   /* :444 */
  const float arg = HRT_SUB(1.f, HRT_DIV(HRT_MUL(2.f, k), (float)num_paths));
  const float phi = hrt_round_checked(acos((double)arg), ambiguous);           /* :445 */
  const float th = HRT_MUL(HRT_MUL(HRT_PI, HRT_ADD(1.f, HRT_SQRT(5.f))), k);   /* :446 */
  const double sphi = sin((double)phi);
  V3 d;
  d.x = hrt_round_checked(cos((double)th) * sphi, ambiguous);                  /* :448 */
  d.y = hrt_round_checked(sin((double)th) * sphi, ambiguous);                  /* :449 */
  d.z = hrt_round_checked(cos((double)phi), ambiguous);                        /* :450 */
  return d;
}

/* ------------------------------------------------------------ one bounce
 * Everything the reference does to a ray that hit something (:621-659),
 * except the closest-hit query itself. */
struct HrtRunConst {
  float fsl_k;   /* 4*pi*f/c, reference :484 */
  float dop_k;   /* f/c,      reference :488 */
};

struct HrtRayState {
  V3 o, d;
  float te_r, te_i, tm_r, tm_i;
  float tau;
};

HRT_HD void hrt_bounce_update(HrtRayState &s, const HrtMaterial &m, const HrtRunConst &k,
                              float t, V3 n, float theta)
{
  float rc[4];
  hrt_refl_coefs(m, theta, rc);                                                /* :623 */
  float fsl = HRT_MUL(k.fsl_k, t);                                             /* :627 */
  fsl = HRT_MUL(fsl, fsl);                                                     /* :628 */
  if (fsl > 1.f) {                                                             /* :629 */
    rc[0] = HRT_DIV(rc[0], fsl); rc[1] = HRT_DIV(rc[1], fsl);
    rc[2] = HRT_DIV(rc[2], fsl); rc[3] = HRT_DIV(rc[3], fsl);
  }
  const float a = HRT_SUB(HRT_MUL(s.te_r, rc[0]), HRT_MUL(s.te_i, rc[1]));     /* :636 */
  const float b = HRT_ADD(HRT_MUL(s.te_r, rc[1]), HRT_MUL(s.te_i, rc[0]));
  const float c = HRT_SUB(HRT_MUL(s.tm_r, rc[2]), HRT_MUL(s.tm_i, rc[3]));
  const float e = HRT_ADD(HRT_MUL(s.tm_r, rc[3]), HRT_MUL(s.tm_i, rc[2]));
  s.te_r = a; s.te_i = b; s.tm_r = c; s.tm_i = e;
  s.tau = HRT_ADD(s.tau, HRT_DIV(t, HRT_C0));                                  /* :645 */
  V3 step = v3_scale(s.d, t);                                                  /* :650 */
  s.o = v3_add(step, s.o);                                                     /* :651 */
  const float dn = v3_dot(s.d, n);                                             /* :654 */
  step = v3_scale(n, HRT_MUL(2.f, dn));                                        /* :655 */
  s.d = v3_sub(s.d, step);                                                     /* :656 */
  step = v3_scale(s.d, 1e-4f);                                                 /* :658 */
  s.o = v3_add(s.o, step);                                                     /* :659 */
}

/* ------------------------------------------------------- scatter to one RX
 * reference :676-722 after the shadow query has been resolved. */
struct HrtScatterOut {
  float te_r, te_i, tm_r, tm_i, tau, dfreq;   /* dfreq: amount SUBTRACTED from freq_shift */
  V3 dir_rx;
};

/* Correctly rounded quotients from one shared reciprocal (Markstein): r = refined
 * 1/y, then q = x r corrected twice with the exact residual fma(-y, q, x).  Equal
 * to IEEE x / y bit for bit -- 4e8 random (x, y) pairs incl. reciprocals off by an
 * ulp and tiny numerators: 0 mismatches already after ONE correction -- at a third
 * of the instructions of three separate divisions.  (+-0 numerators give +0.) */
struct HrtRecip { float y, r; };
HRT_HD HrtRecip hrt_recip(float y)
{
  HrtRecip k; k.y = y;
#if defined(__CUDA_ARCH__)
  float r0; asm("rcp.approx.f32 %0, %1;" : "=f"(r0) : "f"(y));
  k.r = HRT_FMA(r0, HRT_FMA(-y, r0, 1.f), r0);
#else
  k.r = 1.0f / y;
#endif
  return k;
}
HRT_HD float hrt_div_by(float x, const HrtRecip &k)
{
#if defined(__CUDA_ARCH__)
  float q = x * k.r;
  q = HRT_FMA(HRT_FMA(-k.y, q, x), k.r, q);
  q = HRT_FMA(HRT_FMA(-k.y, q, x), k.r, q);
  return q;
#else
  return x / k.y;
#endif
}
/* x / c0 (delays, :645, :709): the same with the correctly rounded constant reciprocal */
HRT_HD float hrt_div_c0(float x)
{
#if defined(__CUDA_ARCH__)
  const float rc = 1.0f / HRT_C0;
  float q = x * rc;
  q = HRT_FMA(HRT_FMA(-HRT_C0, q, x), rc, q);
  q = HRT_FMA(HRT_FMA(-HRT_C0, q, x), rc, q);
  return q;
#else
  return x / HRT_C0;
#endif
}

/* geometry of the shadow ray: direction (unit) and distance to the receiver */
HRT_HD V3 hrt_shadow_dir(V3 o, V3 rx, float *dist)
{
  const V3 dv = v3_sub(rx, o);                                                 /* :676 */
  const float len = HRT_SQRT(v3_dot(dv, dv));                                  /* :677 */
  *dist = len;
  if (!(len > 1e-30f && len < 1e30f)) return v3(HRT_DIV(dv.x, len), HRT_DIV(dv.y, len), HRT_DIV(dv.z, len));
  const HrtRecip k = hrt_recip(len);
  return v3(hrt_div_by(dv.x, k), hrt_div_by(dv.y, k), hrt_div_by(dv.z, k));    /* :678 */
}

HRT_HD HrtScatterOut hrt_scatter_path(const HrtRayState &s, const HrtMaterial &m,
                                      const HrtRunConst &k, V3 n, V3 mesh_vel,
                                      V3 sd, float dist, float theta_i)
{
  HrtScatterOut r;
  const float th_s = acosf(v3_dot(sd, n));                                     /* :694 */
  float sc[4];
  hrt_scat_coefs(m, th_s, theta_i, sc);                                        /* :696 */
  r.te_r = HRT_SUB(HRT_MUL(s.te_r, sc[0]), HRT_MUL(s.te_i, sc[1]));            /* :698 */
  r.te_i = HRT_ADD(HRT_MUL(s.te_r, sc[1]), HRT_MUL(s.te_i, sc[0]));
  r.tm_r = HRT_SUB(HRT_MUL(s.tm_r, sc[2]), HRT_MUL(s.tm_i, sc[3]));
  r.tm_i = HRT_ADD(HRT_MUL(s.tm_r, sc[3]), HRT_MUL(s.tm_i, sc[2]));
  r.dir_rx = v3(-sd.x, -sd.y, -sd.z);                                          /* :707 */
  r.tau = HRT_ADD(s.tau, HRT_DIV(dist, HRT_C0));                               /* :709 */
  float l2 = HRT_MUL(k.fsl_k, dist);                                           /* :711 */
  l2 = HRT_MUL(l2, l2);
  if (l2 > 1.f) {                                                              /* :713 */
    r.te_r = HRT_DIV(r.te_r, l2); r.te_i = HRT_DIV(r.te_i, l2);
    r.tm_r = HRT_DIV(r.tm_r, l2); r.tm_i = HRT_DIV(r.tm_i, l2);
  }
  const V3 dd = v3_sub(sd, s.d);                                               /* :720 */
  r.dfreq = HRT_MUL(v3_dot(dd, mesh_vel), k.dop_k);                            /* :721 */
  return r;
}

/* The same path for the kernels' receiver loop: identical formulas for the
 * delay, direction and Doppler terms (bit-exact); the gains apply the two
 * normalisations (|scat| and the free-space loss, :401-405 and :713-718) as one
 * reciprocal-multiply instead of eight divisions -- a few ulp, inside the
 * 1e-4 gain tolerance that libm differences impose anyway. */
HRT_HD HrtScatterOut hrt_scatter_path_fast(const HrtRayState &s, const HrtScatConst &m,
                                           const HrtRunConst &k, V3 n, V3 mesh_vel,
                                           V3 sd, float dist, float theta_i)
{
  HrtScatterOut r;
  const float th_s = acosf(v3_dot(sd, n));                                     /* :694 */
  float si, ci;
#if defined(__CUDA_ARCH__)
  sincosf(theta_i, &si, &ci);
#else
  si = sinf(theta_i); ci = cosf(theta_i);
#endif
  const float cs = cosf(th_s);                                                 /* :372 */
  const float lobe = HRT_MUL(m.s, expf(HRT_MUL(m.neg_alpha, fabsf(HRT_SUB(th_s, theta_i))))); /* :378 */
  const float spec = HRT_MUL(m.rough, cs);                                     /* :383 */
  const float diff = HRT_MUL(m.one_minus_rough, cs);                           /* :384 */
  const float te_r = HRT_MUL(lobe, HRT_ADD(spec, diff));                       /* :388 */
  const float tm_r = HRT_MUL(lobe, HRT_ADD(HRT_MUL(spec, ci), diff));          /* :390 */
  const float sp = sinf(HRT_MUL(HRT_MUL(m.alpha, si), 0.1f));                  /* :394 */
  const float te_i = HRT_MUL(te_r, sp), tm_i = HRT_MUL(tm_r, sp);              /* :395 */
  const float nrm = HRT_SQRT(HRT_ADD(HRT_ADD(HRT_ADD(HRT_MUL(te_r, te_r), HRT_MUL(te_i, te_i)),
                                             HRT_MUL(tm_r, tm_r)), HRT_MUL(tm_i, tm_i))); /* :399 */
  float l2 = HRT_MUL(k.fsl_k, dist);                                           /* :711 */
  l2 = HRT_MUL(l2, l2);
  float den = nrm > 1e-6f ? nrm : 1.f;                                         /* :401 */
  if (l2 > 1.f) den = HRT_MUL(den, l2);                                        /* :713 */
  const float sc = HRT_DIV(1.f, den);
  r.te_r = HRT_MUL(HRT_SUB(HRT_MUL(s.te_r, te_r), HRT_MUL(s.te_i, te_i)), sc); /* :698 */
  r.te_i = HRT_MUL(HRT_ADD(HRT_MUL(s.te_r, te_i), HRT_MUL(s.te_i, te_r)), sc);
  r.tm_r = HRT_MUL(HRT_SUB(HRT_MUL(s.tm_r, tm_r), HRT_MUL(s.tm_i, tm_i)), sc);
  r.tm_i = HRT_MUL(HRT_ADD(HRT_MUL(s.tm_r, tm_i), HRT_MUL(s.tm_i, tm_r)), sc);
  r.dir_rx = v3(-sd.x, -sd.y, -sd.z);                                          /* :707 */
  r.tau = HRT_ADD(s.tau, hrt_div_c0(dist));                                    /* :709 */
  const V3 dd = v3_sub(sd, s.d);                                               /* :720 */
  r.dfreq = HRT_MUL(v3_dot(dd, mesh_vel), k.dop_k);                            /* :721 */
  return r;
}

/* ------------------------------------------- closed form of a scatter path
 * The reference L2-normalises its four scattering coefficients (:399-405)
 * whenever their norm exceeds 1e-6.  All four carry the common factor
 * A = s * exp(-alpha |theta_s - theta_i|) * cos(theta_s), so after the
 * normalisation only its SIGN is left:
 *   (te_r, te_i, tm_r, tm_i) = sign(A) * (1, p, g, g p) / sqrt((1 + g^2)(1 + p^2)),
 *   g = rough * cos(theta_i) + (1 - rough),  p = sin(0.1 alpha sin(theta_i)),
 * with cos(theta_i) = |n_i . d| and sin(theta_i) = sqrt(1 - (n_i . d)^2) for the
 * folded incidence angle (:280-283).  No acos / acosf / expf / cosf / sinf per
 * path; the result agrees with the reference's fp32 evaluation to ~1e-6
 * relative (tests/emul: emul_scatter_cf_vs_exact), 100x inside the 1e-4 gain
 * tolerance.  The closed form is used only when a cheap estimate proves the
 * norm test passes with 1 % to spare; everything else -- norm near or below
 * 1e-6, s = 0 (metal), alpha = 0, |n_i . d| > 1 (NaN in the reference) --
 * takes the libm path above (hrt_scatter_path_fast). */
HRT_HD float hrt_sqrt_fast(float v)
{
#if defined(__CUDA_ARCH__)
  float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v)); return r;
#else
  return sqrtf(v);
#endif
}
HRT_HD float hrt_rsqrt_fast(float v)
{
#if defined(__CUDA_ARCH__)
  float r; asm("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v)); return r;
#else
  return 1.0f / sqrtf(v);
#endif
}
HRT_HD float hrt_rcp_fast(float v)
{
#if defined(__CUDA_ARCH__)
  float r; asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(v)); return r;
#else
  return 1.0f / v;
#endif
}
HRT_HD float hrt_exp2_fast(float v)
{
#if defined(__CUDA_ARCH__)
  float r; asm("ex2.approx.f32 %0, %1;" : "=f"(r) : "f"(v)); return r;
#else
  return exp2f(v);
#endif
}
/* acos to 7e-5 rad on [-1, 1] (Abramowitz-Stegun 4.4.45): for the norm ESTIMATE only */
HRT_HD float hrt_acos_est(float x)
{
  const float a = fminf(fabsf(x), 1.f);
  float p = HRT_FMA(-0.0187293f, a, 0.0742610f);
  p = HRT_FMA(p, a, -0.2121144f);
  p = HRT_FMA(p, a, 1.5707288f);
  const float r = hrt_sqrt_fast(1.f - a) * p;
  return x < 0.f ? HRT_PI - r : r;
}

struct HrtScatCf {
  float rough, omr;       /* 1/(1+alpha), 1 - rough (:382-384)              */
  float ph_k;             /* 0.1 alpha (:394)                                */
  float lobe_lb;          /* s exp(-alpha pi) <= lobe factor                 */
  float s, alpha_l2e;     /* for the estimate: s, alpha log2(e)              */
  float ok;               /* 1: closed form allowed for this material, else 0 */
};
HRT_HD HrtScatCf hrt_scat_cf(const HrtMaterial &m)
{
  HrtScatCf c;
  c.rough = HRT_DIV(1.0f, HRT_ADD(1.0f, m.s1_alpha));
  c.omr = HRT_SUB(1.0f, c.rough);
  c.ph_k = 0.1f * m.s1_alpha;
  c.s = m.s; c.alpha_l2e = m.s1_alpha * 1.4426950408889634f;
  c.lobe_lb = m.s * hrt_exp2_fast(-c.alpha_l2e * HRT_PI) * 0.98f;
  c.ok = (m.s > 0.f && m.s1_alpha >= 1.f && c.ph_k <= 0.5f) ? 1.f : 0.f;
  return c;
}

/* (g, p, 1/sqrt((1+g^2)(1+p^2))) from cos and sin of the incidence angle */
HRT_HD void hrt_scat_unit(const HrtScatCf &m, float ci, float si, float *g, float *p, float *inv_n)
{
  *g = HRT_FMA(m.rough, ci, m.omr);
  const float ph = m.ph_k * si, q = ph * ph;              /* sin(ph), 0 <= ph <= 0.5: |rel err| < 1e-8 */
  *p = ph * HRT_FMA(q, HRT_FMA(q, HRT_FMA(q, -1.f / 5040.f, 1.f / 120.f), -1.f / 6.f), 1.f);
  *inv_n = hrt_rsqrt_fast(HRT_FMA(*g, *g, 1.f) * HRT_FMA(*p, *p, 1.f));
}

/* Gains of one scatter path in closed form.  xs = d . n (scattering direction
 * against the surface normal), (ci, si) the carried incidence angle, th_i_est
 * any 1e-3-accurate value of that angle.  Returns false when the closed form
 * must not be used (caller falls back to hrt_scatter_path_fast); the delay,
 * direction and Doppler terms are formed by the caller exactly as before. */
HRT_HD bool hrt_scatter_gains_cf(const HrtRayState &s, const HrtScatCf &m, float fsl_k, float xs, float dist,
                                 float ci, float si, float *te_r, float *te_i, float *tm_r, float *tm_i)
{
  if (!(m.ok != 0.f) || !(ci <= 1.f)) return false;
  float g, p, inv_n;
  hrt_scat_unit(m, ci, si, &g, &p, &inv_n);
  const float ax = fabsf(xs);
  /* norm = lobe |cos theta_s| sqrt((1+g^2)(1+p^2)) >= lobe_lb |xs|: clearly above 1e-6? */
  if (!(m.lobe_lb * ax > 1.02e-6f)) {
    /* estimate with approximate angles: 1e-3 relative, 2 % margin */
    const float dth = fabsf(hrt_acos_est(xs) - hrt_acos_est(ci));
    const float est = m.s * hrt_exp2_fast(-m.alpha_l2e * dth) * ax;
    if (!(est * inv_n * 0.98f > 1.02e-6f * inv_n * inv_n)) return false;     /* est / inv_n > 1.02e-6 / 0.98 */
  }
  float k = xs < 0.f ? -inv_n : inv_n;
  float l2 = HRT_MUL(fsl_k, dist);                                             /* :711 */
  l2 = HRT_MUL(l2, l2);
  if (l2 > 1.f) k *= hrt_rcp_fast(l2);                                         /* :713 */
  const float gk = g * k;
  *te_r = HRT_FMA(-s.te_i, p, s.te_r) * k;                                     /* :698 */
  *te_i = HRT_FMA(s.te_r, p, s.te_i) * k;
  *tm_r = HRT_FMA(-s.tm_i, p, s.tm_r) * gk;
  *tm_i = HRT_FMA(s.tm_r, p, s.tm_i) * gk;
  return true;
}

/* cos / sin of the folded incidence angle from the fp32 dot product the
 * reference hands to acos (:281): |x| and sqrt(1 - x^2) */
HRT_HD void hrt_fold_cos_sin(float x, float *ci, float *si)
{
  *ci = fabsf(x);
  *si = hrt_sqrt_fast(fmaxf(HRT_FMA(-x, x, 1.f), 0.f));
}

/* One scatter path as the kernels form it: closed-form gains when allowed, the
 * libm formulas otherwise.  cx_i: fp32 dot product n.d of the shadow hit whose
 * angle is carried (what the reference feeds to acos, :281), or 2 = none yet:
 * the primary incidence angle theta_p with (ci_p, si_p) = (cosf, sinf)(theta_p). */
#define HRT_CX_PRIMARY 2.f
HRT_HD HrtScatterOut hrt_scatter_path_auto(const HrtRayState &s, const HrtScatConst &m, const HrtScatCf &mcf,
                                           const HrtRunConst &k, V3 n, V3 mesh_vel, V3 sd, float dist,
                                           float cx_i, float theta_p, float ci_p, float si_p)
{
  HrtScatterOut r;
  float ci = ci_p, si = si_p;
  if (cx_i != HRT_CX_PRIMARY) hrt_fold_cos_sin(cx_i, &ci, &si);
  const float xs = v3_dot(sd, n);                                              /* :694, argument of acosf */
  if (hrt_scatter_gains_cf(s, mcf, k.fsl_k, xs, dist, ci, si, &r.te_r, &r.te_i, &r.tm_r, &r.tm_i)) {
    r.dir_rx = v3(-sd.x, -sd.y, -sd.z);                                        /* :707 */
    r.tau = HRT_ADD(s.tau, hrt_div_c0(dist));                                  /* :709 */
    r.dfreq = HRT_MUL(v3_dot(v3_sub(sd, s.d), mesh_vel), k.dop_k);             /* :720-721 */
    return r;
  }
  /* norm test not clear-cut, s = 0, alpha = 0, |n.d| > 1 ...: the reference's formulas with libm */
  float theta_i = theta_p;
  if (cx_i != HRT_CX_PRIMARY) {
    theta_i = (float)acos((double)cx_i);                                       /* :281-283 */
    if ((double)theta_i > (double)HRT_PI / 2.) theta_i = HRT_SUB(HRT_PI, theta_i);
  }
  return hrt_scatter_path_fast(s, m, k, n, mesh_vel, sd, dist, theta_i);
}

/* ------------------------------------------------------------------ LoS
 * reference :520-577 for one (rx, tx) pair, query result passed in. */
struct HrtLosOut {
  V3 dir_tx, dir_rx;
  float a, tau, freq;
  int state;      /* 0 blocked, 1 clear, 2 co-located */
};

HRT_HD HrtLosOut hrt_los_finish(V3 d, bool hit, float t_hit, V3 tx_vel0, V3 rx_vel0,
                                const HrtRunConst &k, float f_over_c)
{
  HrtLosOut r;
  r.dir_tx = r.dir_rx = v3(0.f, 0.f, 0.f); r.a = 0.f; r.tau = 0.f; r.freq = 0.f;
  if (hit && t_hit <= 1.f) { r.state = 0; return r; }                          /* :548 */
  const float len = HRT_SQRT(v3_dot(d, d));                                    /* :558 */
  const V3 u = v3(HRT_DIV(d.x, len), HRT_DIV(d.y, len), HRT_DIV(d.z, len));    /* :560 */
  r.dir_tx = u; r.dir_rx = v3(-u.x, -u.y, -u.z);
  const float fsl = HRT_MUL(k.fsl_k, len);                                     /* :564 */
  r.a = fsl > 1.f ? HRT_DIV(1.f, fsl) : 1.f;
  r.tau = HRT_DIV(len, HRT_C0);                                                /* :571 */
  r.freq = HRT_MUL(HRT_SUB(v3_dot(tx_vel0, u), v3_dot(rx_vel0, u)), f_over_c); /* :573-574 */
  r.state = 1;
  return r;
}

/* 64-bit mixer for the order-independent checksums of summary mode */
HRT_HD uint64_t hrt_mix64(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
