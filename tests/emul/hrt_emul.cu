/* hrt_emul.cu -- TEST-ONLY serial CPU instantiation of the kernel logic.
 *
 * The authoring container has no GPU.  hrt_core.cuh / hrt_bvh.cuh are written
 * as __host__ __device__ code, so this driver runs the very same functions the
 * kernels call (BVH build steps, traversal, Moeller-Trumbore with the
 * division-free shortcuts, bounce, scatter, LoS) one ray at a time on the CPU,
 * and tests/test_emul_vs_oracle.py compares the result with the oracle.  It
 * mirrors the launch logic of hrt_cuda.cu; it is NOT linked into, loaded by or
 * reachable from libhermespy_rt.so.
 */
#include <algorithm>
#include <vector>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/hrt_cuda.h"
#include "../../hermespy-rt_b200/csrc/hrt_bvh.cuh"
#include "../../hermespy-rt_b200/csrc/hrt_rxmap.cuh"
#include "../../hermespy-rt_b200/csrc/hrt_ext.cuh"

struct EmulScene {
  std::vector<float4> tris, nodes;
  std::vector<uint32_t> gid, mesh_of, mesh_mat;
  std::vector<V3> mesh_vel;
  uint32_t n = 0; int root = 0; uint32_t num_nodes = 0;
  std::vector<float4> wnodes;          /* 4-wide nodes, 8 octant copies */
  uint32_t num_wide = 0; int wroot = 0;
  float pad = 0.f;                     /* node padding (hrt_box_pad) */
};

/* binary nodes (octant-0 copy) -> 4-wide nodes, as hrt_cuda.cu's collapse kernels do:
 * even-depth binary nodes become wide nodes, numbered in binary index order */
static void collapse_wide(EmulScene &E)
{
  E.wroot = E.root; E.num_wide = 0;
  if (E.num_nodes == 0) return;
  std::vector<int> depth(E.num_nodes, -1);
  std::vector<int> todo(1, 0); depth[0] = 0;
  while (!todo.empty()) {
    const int i = todo.back(); todo.pop_back();
    for (int s = 0; s < 2; ++s) {
      const int ref = hrt_float_as_int(E.nodes[4 * (size_t)i + 2 * s + 1].z);
      if (ref >= 0) { depth[ref] = depth[i] + 1; todo.push_back(ref); }
    }
  }
  std::vector<uint32_t> widx(E.num_nodes, 0);
  for (uint32_t i = 0; i < E.num_nodes; ++i) if (depth[i] >= 0 && !(depth[i] & 1)) widx[i] = E.num_wide++;
  E.wnodes.assign((size_t)E.num_wide * HRT_WIDE_F4 * 8, float4());
  for (uint32_t i = 0; i < E.num_nodes; ++i) {
    if (depth[i] < 0 || (depth[i] & 1)) continue;
    HrtWideChild ch[4];
    const int n = hrt_wide_children(E.nodes.data(), (int)i, widx.data(), ch);
    hrt_wide_emit(E.wnodes.data(), (size_t)E.num_wide * HRT_WIDE_F4, 8u, widx[i], ch, n);
  }
  E.wroot = 0;
}

static void build(const Scene *sc, EmulScene &E, int leaf_max, float pad_ulps, float extra_abs)
{
  std::vector<V3> A, B, C;
  float max_abs = extra_abs;
  for (uint32_t m = 0; m < sc->num_meshes; ++m) {
    const Mesh *me = &sc->meshes[m];
    for (uint32_t f = 0; f < me->num_triangles; ++f) {
      const Vec3 a = me->vs[me->is[3 * f]], b = me->vs[me->is[3 * f + 1]], c = me->vs[me->is[3 * f + 2]];
      A.push_back(v3(a.x, a.y, a.z)); B.push_back(v3(b.x, b.y, b.z)); C.push_back(v3(c.x, c.y, c.z));
      E.mesh_of.push_back(m);
    }
    for (uint32_t v = 0; v < me->num_vertices; ++v)
      max_abs = fmaxf(max_abs, fmaxf(fabsf(me->vs[v].x), fmaxf(fabsf(me->vs[v].y), fabsf(me->vs[v].z))));
    E.mesh_mat.push_back(me->material_index);
    E.mesh_vel.push_back(v3(me->velocity.x, me->velocity.y, me->velocity.z));
  }
  const int n = (int)A.size();
  E.n = n;
  E.pad = hrt_box_pad(max_abs, pad_ulps);
  if (n == 0) return;
  std::vector<HrtTriSetup> S(n);
  V3 slo = v3(1e30f, 1e30f, 1e30f), shi = v3(-1e30f, -1e30f, -1e30f);
  for (int g = 0; g < n; ++g) {
    S[g] = hrt_tri_setup(A[g], B[g], C[g]);
    slo = v3(fminf(slo.x, S[g].lo.x), fminf(slo.y, S[g].lo.y), fminf(slo.z, S[g].lo.z));
    shi = v3(fmaxf(shi.x, S[g].hi.x), fmaxf(shi.y, S[g].hi.y), fmaxf(shi.z, S[g].hi.z));
  }
  const V3 inv = v3(1.f / fmaxf(shi.x - slo.x, 1e-30f), 1.f / fmaxf(shi.y - slo.y, 1e-30f), 1.f / fmaxf(shi.z - slo.z, 1e-30f));
  std::vector<uint64_t> keys(n);
  for (int g = 0; g < n; ++g) keys[g] = hrt_morton_key(S[g].lo, S[g].hi, slo, inv, g);
  std::sort(keys.begin(), keys.end());
  E.tris.resize(3 * (size_t)n); E.gid.resize(n);
  std::vector<V3> blo(2 * (size_t)n), bhi(2 * (size_t)n);
  for (int s = 0; s < n; ++s) {
    const uint32_t g = (uint32_t)(keys[s] & 0xFFFFFFFFull);
    E.tris[3 * s] = S[g].q0; E.tris[3 * s + 1] = S[g].q1; E.tris[3 * s + 2] = S[g].q2;
    E.gid[s] = g;
    blo[n - 1 + s] = S[g].lo; bhi[n - 1 + s] = S[g].hi;
  }
  if (n <= leaf_max) { E.root = E.wroot = hrt_leaf_ref(0u, (uint32_t)n); return; }
  std::vector<int> kl(n - 1), kr(n - 1), kf(n - 1), kla(n - 1), parent(2 * (size_t)n, -1), arrive(n - 1, 0);
  for (int i = 0; i < n - 1; ++i) {
    hrt_karras_node(keys.data(), n, i, &kl[i], &kr[i], &kf[i], &kla[i]);
    parent[kl[i]] = i; parent[kr[i]] = i;
  }
  parent[0] = -1;
  for (int s = 0; s < n; ++s) {           /* same arrival protocol as k_refit */
    int cur = parent[n - 1 + s];
    while (cur >= 0) {
      if (arrive[cur]++ == 0) break;
      const int a = kl[cur], b = kr[cur];
      blo[cur] = v3(fminf(blo[a].x, blo[b].x), fminf(blo[a].y, blo[b].y), fminf(blo[a].z, blo[b].z));
      bhi[cur] = v3(fmaxf(bhi[a].x, bhi[b].x), fmaxf(bhi[a].y, bhi[b].y), fmaxf(bhi[a].z, bhi[b].z));
      cur = parent[cur];
    }
  }
  std::vector<int> used(n - 1), newidx(n - 1);
  int count = 0;
  for (int i = 0; i < n - 1; ++i) { used[i] = (kla[i] - kf[i] + 1) > leaf_max; newidx[i] = count; count += used[i]; }
  E.num_nodes = count; E.nodes.resize(4 * (size_t)count); E.root = 0;   /* one plain copy */
  const float pad = hrt_box_pad(max_abs, pad_ulps);
  for (int i = 0; i < n - 1; ++i) {
    if (!used[i]) continue;
    for (uint32_t oct = 0; oct < 1; ++oct)
      hrt_emit_node(&E.nodes[4 * ((size_t)oct * count + newidx[i])],
                    hrt_child_ref(kl[i], n, kf.data(), kla.data(), newidx.data(), leaf_max),
                    hrt_child_ref(kr[i], n, kf.data(), kla.data(), newidx.data(), leaf_max),
                    blo[kl[i]], bhi[kl[i]], blo[kr[i]], bhi[kr[i]], pad, oct);
  }
  collapse_wide(E);
}

static HrtHit query(const EmulScene &E, V3 o, V3 d, int brute, uint32_t self_slot = HRT_NONE, float self_nt = 0.f)
{
  HrtGlobalMem m; m.nodes = E.nodes.data(); m.tris = E.tris.data(); m.wnodes = E.wnodes.data();
  HrtNoCount nc;
  if (brute == 1) return hrt_closest_hit_brute(m, E.gid.data(), E.n, o, d, nc);
  if (brute == 2) return hrt_closest_hit(m, E.gid.data(), E.root, E.n, o, d, nc);                        /* binary tree as built */
  if (brute == 4) return hrt_closest_hit_wide<false>(m, E.gid.data(), E.wroot, E.n, o, d, nc);        /* 4-wide, plain copy (octant 0) */
  if (self_slot != HRT_NONE)                                                                          /* shadow rays, as k_scatter */
    return hrt_closest_hit_wide<true, true>(m, E.gid.data(), E.wroot, E.n, o, d, nc, (size_t)E.num_wide * HRT_WIDE_F4, self_slot, self_nt);
  return hrt_closest_hit_wide<true>(m, E.gid.data(), E.wroot, E.n, o, d, nc, (size_t)E.num_wide * HRT_WIDE_F4);  /* 4-wide: what the kernels run */
}

/* receiver maps (hrt_rxmap.cuh), built serially with the same two-level scheme and
 * element functions as the kernels k_rxmap_* */
struct EmulRxMap { uint32_t G = 0; std::vector<uint32_t> start, count; std::vector<uint8_t> sure; std::vector<uint32_t> items; std::vector<float> inv_step; };

static void build_rxmap(const EmulScene &E, const Vec3 *rx, size_t R, uint32_t G, float pad, EmulRxMap &M)
{
  M.G = G; M.start.assign(R * 6 * G * G, 0); M.count.assign(R * 6 * G * G, 0); M.sure.assign(R * 6 * G * G, 0); M.items.clear(); M.inv_step.assign(R, 0.f);
  const uint32_t nb = G / HRT_RXMAP_BLOCK;
  std::vector<uint32_t> cand;
  /* scene bounds: the per-receiver quantisation step covers the farthest corner of the bounding box */
  V3 lo = v3(3e38f, 3e38f, 3e38f), hi = v3(-3e38f, -3e38f, -3e38f);
  for (uint32_t s = 0; s < E.n; ++s) {
    const float4 q0 = E.tris[3 * s], q1 = E.tris[3 * s + 1], q2 = E.tris[3 * s + 2];
    const V3 c[3] = { v3(q0.x, q0.y, q0.z), v3(q0.x + q0.w, q0.y + q1.x, q0.z + q1.y), v3(q0.x + q1.z, q0.y + q1.w, q0.z + q2.x) };
    for (int k = 0; k < 3; ++k) {
      lo = v3(fminf(lo.x, c[k].x), fminf(lo.y, c[k].y), fminf(lo.z, c[k].z));
      hi = v3(fmaxf(hi.x, c[k].x), fmaxf(hi.y, c[k].y), fmaxf(hi.z, c[k].z));
    }
  }
  for (size_t r = 0; r < R; ++r) {
    const V3 apex = v3(rx[r].x, rx[r].y, rx[r].z);
    M.inv_step[r] = hrt_rxmap_inv_step(apex, lo, hi);
    for (uint32_t f = 0; f < 6; ++f)
      for (uint32_t bj = 0; bj < nb; ++bj)
        for (uint32_t bi = 0; bi < nb; ++bi) {
          const HrtPyramid bp = hrt_rxmap_pyramid(f, G, bi * HRT_RXMAP_BLOCK, (bi + 1) * HRT_RXMAP_BLOCK, bj * HRT_RXMAP_BLOCK, (bj + 1) * HRT_RXMAP_BLOCK);
          cand.clear();
          for (uint32_t s = 0; s < E.n; ++s) {
            V3 va, vb, vc; hrt_rxmap_corners(E.tris[3 * s], E.tris[3 * s + 1], E.tris[3 * s + 2], apex, &va, &vb, &vc);
            if (hrt_rxmap_overlap(bp, va, vb, vc, pad)) cand.push_back(s);
          }
          for (uint32_t cj = 0; cj < HRT_RXMAP_BLOCK; ++cj)
            for (uint32_t ci = 0; ci < HRT_RXMAP_BLOCK; ++ci) {
              const uint32_t i = bi * HRT_RXMAP_BLOCK + ci, j = bj * HRT_RXMAP_BLOCK + cj;
              const HrtPyramid cp = hrt_rxmap_pyramid(f, G, i, i + 1, j, j + 1);
              const size_t cell = ((r * 6 + f) * G + j) * G + i;
              M.start[cell] = (uint32_t)M.items.size();
              for (uint32_t s : cand) {
                V3 va, vb, vc; hrt_rxmap_corners(E.tris[3 * s], E.tris[3 * s + 1], E.tris[3 * s + 2], apex, &va, &vb, &vc);
                if (!hrt_rxmap_overlap(cp, va, vb, vc, pad)) continue;
                float dl, dh; hrt_rxmap_depth(cp, va, vb, vc, pad, G, &dl, &dh);
                M.items.push_back(hrt_rxmap_item(s, dl, dh, M.inv_step[r]));
              }
              M.count[cell] = (uint32_t)M.items.size() - M.start[cell];
              hrt_rxmap_sort_items(&M.items[M.start[cell]], M.count[cell]);
              if (M.count[cell] == 1) {
                const uint32_t s = M.items[M.start[cell]] & 0xFFFFu;
                V3 va, vb, vc; hrt_rxmap_corners(E.tris[3 * s], E.tris[3 * s + 1], E.tris[3 * s + 2], apex, &va, &vb, &vc);
                M.sure[cell] = hrt_rxmap_sure(cp, va, vb, vc, pad, 510.f / M.inv_step[r]) ? 1 : 0;
              }
            }
        }
  }
}

static unsigned long long *g_dbg = nullptr;   /* lab counters: tests on the origin's own plane / elsewhere, per side; accepted per side */
/* shadow query through the receiver map, exactly as query_map in hrt_run_kernels.cuh */
static HrtHit query_map(const EmulScene &E, const EmulRxMap &M, size_t r, V3 o, V3 d, float dist, uint32_t self_slot, float self_nt,
                        unsigned long long *tests = nullptr, bool *took_sure = nullptr)
{
  if (took_sure) *took_sure = false;
  HrtHit h; h.t = HRT_T_MAX; h.gid = HRT_NONE; h.slot = HRT_NONE;
  HrtNoCount nc;
  uint32_t c_pos, c_neg;
  hrt_rxmap_cells2(d, M.G, &c_pos, &c_neg);          /* (-d) first, (+d) only if nothing in front of the receiver */
  const HrtMapDepth md = hrt_rxmap_query_depth(dist, M.inv_step[r]);
  int q_stop = -1;
  const bool self_out = self_slot != HRT_NONE &&
                        hrt_mt_self_miss(E.tris[3 * self_slot], E.tris[3 * self_slot + 1], E.tris[3 * self_slot + 2], d, self_nt, nc);
  for (int side = 0; side < 2; ++side) {
    const size_t cell = r * 6 * M.G * M.G + (side ? c_pos : c_neg);
    if (side == 1 && M.sure[cell] && h.gid == HRT_NONE && dist > 1.001f) {     /* "sure" cell: no test, t not computed */
      const uint32_t s = M.items[M.start[cell]] & 0xFFFFu;
      h.gid = E.gid[s]; h.slot = s; h.t = dist;
      if (took_sure) *took_sure = true;
      break;
    }
    for (uint32_t k = 0; k < M.count[cell]; ++k) {
      const uint32_t w = M.items[M.start[cell] + k], s = w & 0xFFFFu;
      if (side == 0) {
        if ((int)((w >> 16) & 255u) < q_stop) break;           /* everything left is farther from o than the best hit */
        if ((int)(w >> 24) > md.q_behind) continue;            /* entirely behind o */
        if (s == self_slot && self_out) continue;
      }
      float t;
      if (tests) ++*tests;
      if (g_dbg) {
        const float4 q0 = E.tris[3 * s], q2 = E.tris[3 * s + 2];
        const float pd = fabsf((o.x - q0.x) * q2.y + (o.y - q0.y) * q2.z + (o.z - q0.z) * q2.w);
        g_dbg[side * 2 + (pd < 3e-4f ? 0 : 1)]++;
      }
      if (hrt_mt_test<HrtNoCount, true>(E.tris[3 * s], E.tris[3 * s + 1], E.tris[3 * s + 2], o, d, h.t, 0u, 0u, &t, nc)) {
        if (g_dbg) g_dbg[4 + side]++;
        if (t < h.t || E.gid[s] < h.gid) { h.t = t; h.gid = E.gid[s]; h.slot = s; }
        if (side == 0) q_stop = hrt_rxmap_stop(md, h.t);
      }
    }
    if (h.t < dist * 0.999f) break;
  }
  return h;
}

static V3 nrm(const EmulScene &E, uint32_t slot) { const float4 q = E.tris[3 * slot + 2]; return v3(q.y, q.z, q.w); }
static V3 tov(Vec3 a) { return v3(a.x, a.y, a.z); }

extern "C" int emul_closest_hits(const Scene *sc, const Ray *rays, size_t n, int leaf_max, float pad_ulps,
                                 int brute, uint32_t *tri, float *t, float *theta)
{
  EmulScene E;
  float ma = 0.f;
  for (size_t i = 0; i < n; ++i) ma = fmaxf(ma, fmaxf(fabsf(rays[i].o.x), fmaxf(fabsf(rays[i].o.y), fabsf(rays[i].o.z))));
  build(sc, E, leaf_max, pad_ulps, ma);
  for (size_t i = 0; i < n; ++i) {
    const HrtHit h = query(E, tov(rays[i].o), tov(rays[i].d), brute);
    const bool hit = h.gid != HRT_NONE;
    tri[i] = h.gid; t[i] = hit ? h.t : -1.f;
    theta[i] = hit ? hrt_theta_fold(nrm(E, h.slot), tov(rays[i].d)) : 0.f;
  }
  return 0;
}

/* Dense compute_paths with the product's output conventions: gains/tau/dir of
 * untouched slots are 0, freq_shift carries the Doppler base everywhere.
 * RaysInfo is not emulated (pure host bookkeeping in hrt_cuda.cu). */
extern "C" int emul_compute_paths(const Scene *sc, const Vec3 *rx_pos, const Vec3 *tx_pos,
                                  const Vec3 *rx_vel, const Vec3 *tx_vel, float f_ghz,
                                  size_t R, size_t T, size_t P, size_t B,
                                  ChannelInfo *los, ChannelInfo *scat,
                                  uint32_t *tr_hit, uint8_t *tr_state,
                                  int leaf_max, float pad_ulps, int brute_and_mode)
{
  const int brute = brute_and_mode & 0xFF;
  const bool closed_form = (brute_and_mode & 0x100) != 0;   /* gains as k_scatter forms them */
  const bool use_map = (brute_and_mode & 0x200) != 0;       /* shadow queries through receiver maps */
  EmulScene E;
  float ma = 0.f;
  for (size_t i = 0; i < R; ++i) ma = fmaxf(ma, fmaxf(fabsf(rx_pos[i].x), fmaxf(fabsf(rx_pos[i].y), fabsf(rx_pos[i].z))));
  for (size_t i = 0; i < T; ++i) ma = fmaxf(ma, fmaxf(fabsf(tx_pos[i].x), fmaxf(fabsf(tx_pos[i].y), fabsf(tx_pos[i].z))));
  build(sc, E, leaf_max, pad_ulps, ma);
  HrtMaterialTable mats; memset(&mats, 0, sizeof mats);
  for (uint32_t m = 0; m < sc->num_meshes; ++m)
    hrt_materials_derive(sc->meshes[m].material_index, f_ghz, (HrtMaterialDerived *)&mats.m[sc->meshes[m].material_index]);
  HrtRunConst k;
  const float f_hz = (float)((double)f_ghz * 1e9);
  k.fsl_k = 4.f * HRT_PI * f_hz / HRT_C0; k.dop_k = f_hz / HRT_C0;
  EmulRxMap M;
  if (use_map) {
    const char *g = getenv("EMUL_RXMAP_G");
    build_rxmap(E, rx_pos, R, g ? (uint32_t)atoi(g) : 64u, 4.f * E.pad, M);   /* 4 x the node padding, as hrt_cuda.cu */
  }

  for (size_t r = 0, q = 0; r < R; ++r)
    for (size_t t = 0; t < T; ++t, ++q) {
      los->a_te_im[q] = los->a_tm_im[q] = 0.f;
      const V3 o = tov(tx_pos[t]);
      const V3 d = v3_sub(tov(rx_pos[r]), o);
      HrtLosOut res;
      if (v3_dot(d, d) < HRT_EPS) { res.dir_rx = v3(1, 0, 0); res.dir_tx = v3(-1, 0, 0); res.a = 1.f; res.tau = 0.f; res.freq = 0.f; res.state = 2; }
      else { const HrtHit h = query(E, o, d, brute); res = hrt_los_finish(d, h.gid != HRT_NONE, h.t, tov(tx_vel[0]), tov(rx_vel[0]), k, k.dop_k); }
      if (res.state == 0) { los->a_te_re[q] = los->a_tm_re[q] = los->tau[q] = 0.f; continue; }
      los->directions_tx[q] = {res.dir_tx.x, res.dir_tx.y, res.dir_tx.z};
      los->directions_rx[q] = {res.dir_rx.x, res.dir_rx.y, res.dir_rx.z};
      los->a_te_re[q] = los->a_tm_re[q] = res.a; los->tau[q] = res.tau; los->freq_shift[q] = res.freq;
    }

  std::vector<HrtRayState> st(T * P);
  std::vector<uint8_t> alive(T * P, 1);
  for (size_t p = 0; p < P; ++p) {
    bool amb = false;
    const V3 d = hrt_launch_dir(p, P, &amb);
    for (size_t t = 0; t < T; ++t) {
      HrtRayState &s = st[t * P + p];
      s.o = tov(tx_pos[t]); s.d = d; s.te_r = 1.f; s.te_i = 0.f; s.tm_r = 1.f; s.tm_i = 0.f; s.tau = 0.f;
      for (size_t b = 0; b < B; ++b) {
        tr_hit[(t * B + b) * P + p] = HRT_IDLE;
        const size_t j = (t * B + b) % T;
        const size_t src = (j % B == 0) ? j / B : t;
        float base = v3_dot(tov(tx_vel[src]), d); base = HRT_MUL(base, k.dop_k);
        for (size_t r = 0; r < R; ++r) {
          const size_t so = ((r * T + t) * B + b) * P + p;
          scat->a_te_re[so] = scat->a_te_im[so] = scat->a_tm_re[so] = scat->a_tm_im[so] = scat->tau[so] = 0.f;
          scat->freq_shift[so] = base;
          scat->directions_rx[so] = {0.f, 0.f, 0.f};
          tr_state[so] = 0;
        }
      }
    }
  }
  for (size_t b = 0; b < B; ++b)
    for (size_t t = 0; t < T; ++t)
      for (size_t p = 0; p < P; ++p) {
        const size_t i = t * P + p;
        if (!alive[i]) continue;
        HrtRayState &s = st[i];
        const HrtHit h = query(E, s.o, s.d, brute);
        tr_hit[(t * B + b) * P + p] = h.gid;
        if (h.gid == HRT_NONE) { alive[i] = 0; continue; }
        const V3 n = nrm(E, h.slot);
        const float theta = hrt_theta_fold(n, s.d);
        const uint32_t mesh = E.mesh_of[h.gid];
        const HrtMaterial &mat = mats.m[E.mesh_mat[mesh]];
        hrt_bounce_update(s, mat, k, h.t, n, theta);
        float carry = theta, cx_carry = HRT_CX_PRIMARY;
        const float ci_p = cosf(theta), si_p = sinf(theta);
        const float self_nt = hrt_mt_self_nt(E.tris[3 * h.slot], E.tris[3 * h.slot + 1], E.tris[3 * h.slot + 2], s.o);   /* as k_scatter */
        for (size_t r = 0; r < R; ++r) {
          const size_t so = ((r * T + t) * B + b) * P + p;
          float dist;
          const V3 sd = hrt_shadow_dir(s.o, tov(rx_pos[r]), &dist);
          const HrtHit sh = use_map ? query_map(E, M, r, s.o, sd, dist, h.slot, self_nt)
                                    : query(E, s.o, sd, brute, brute == 0 ? h.slot : HRT_NONE, self_nt);
          if (sh.gid != HRT_NONE) { carry = hrt_theta_fold(nrm(E, sh.slot), sd); cx_carry = v3_dot(nrm(E, sh.slot), sd); }
          if (sh.gid != HRT_NONE && sh.t <= 1.f) { tr_state[so] = 2; continue; }
          /* closed_form: what k_scatter runs (hrt_scatter_path_auto); else the reference's formulas line by line */
          const HrtScatterOut o = closed_form
            ? hrt_scatter_path_auto(s, hrt_scat_const(mat), hrt_scat_cf(mat), k, n, E.mesh_vel[mesh], sd, dist, cx_carry, theta, ci_p, si_p)
            : hrt_scatter_path(s, mat, k, n, E.mesh_vel[mesh], sd, dist, carry);
          scat->a_te_re[so] = o.te_r; scat->a_te_im[so] = o.te_i; scat->a_tm_re[so] = o.tm_r; scat->a_tm_im[so] = o.tm_i;
          scat->tau[so] = o.tau;
          scat->freq_shift[so] = HRT_SUB(scat->freq_shift[so], o.dfreq);
          scat->directions_rx[so] = {o.dir_rx.x, o.dir_rx.y, o.dir_rx.z};
          tr_state[so] = 1;
        }
      }
  return 0;
}

/* work counters for BVH-quality experiments on the CPU: average box / triangle
 * tests per closest-hit query over a batch of rays */
extern "C" int emul_count_work(const Scene *sc, const Ray *rays, size_t n, int leaf_max, double out[6])
{
  EmulScene E; build(sc, E, leaf_max, 64.f, 0.f);
  HrtGlobalMem m; m.nodes = E.nodes.data(); m.tris = E.tris.data(); m.wnodes = E.wnodes.data();
  HrtCount c;
  unsigned long long tot[5] = {0, 0, 0, 0, 0};
  const bool binary = getenv("EMUL_COUNT_BINARY") != nullptr;
  for (size_t i = 0; i < n; ++i) {
    for (int k = 0; k < 5; ++k) c.c[k] = 0;
    if (binary) hrt_closest_hit(m, E.gid.data(), E.root, E.n, tov(rays[i].o), tov(rays[i].d), c);
    else hrt_closest_hit_wide<true>(m, E.gid.data(), E.wroot, E.n, tov(rays[i].o), tov(rays[i].d), c, (size_t)E.num_wide * HRT_WIDE_F4);
    for (int k = 0; k < 5; ++k) tot[k] += c.c[k];
  }
  for (int k = 0; k < 5; ++k) out[k] = (double)tot[k] / (double)n;
  out[5] = E.num_nodes;
  return 0;
}

extern "C" int emul_bvh_stats(const Scene *sc, int leaf_max, uint32_t *num_nodes, uint32_t *num_tris)
{
  EmulScene E; build(sc, E, leaf_max, 64.f, 0.f);
  *num_nodes = E.num_nodes; *num_tris = E.n;
  return 0;
}

/* The kernels' one-reciprocal form of a scatter path (hrt_scatter_path_fast)
 * against the division-for-division form that follows the reference line by line
 * (hrt_scatter_path), same libm on both sides: worst complex relative deviation of
 * the gains over n random (state, material, geometry) samples; tau, direction and
 * Doppler term must be identical.  Returns the number of non-identical exact words. */
extern "C" int emul_scatter_fast_vs_exact(size_t n, uint32_t seed, float f_ghz, double *worst_rel)
{
  HrtMaterialTable mats;
  memset(&mats, 0, sizeof mats);
  for (uint32_t i = 0; i < NUM_G_MATERIALS; ++i) hrt_materials_derive(i, f_ghz, (HrtMaterialDerived *)&mats.m[i]);
  uint64_t st = 0x9E3779B97F4A7C15ull ^ seed;
  auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (float)((st >> 40) * (1.0 / 16777216.0)); };
  HrtRunConst k;
  const float f_hz = (float)((double)f_ghz * 1e9);
  k.fsl_k = 4.f * HRT_PI * f_hz / HRT_C0; k.dop_k = f_hz / HRT_C0;
  int bad = 0; double worst = 0.0;
  for (size_t i = 0; i < n; ++i) {
    HrtRayState s;
    s.o = v3(rnd() * 100.f - 50.f, rnd() * 100.f - 50.f, rnd() * 20.f);
    s.d = v3_normalize(v3(rnd() - .5f, rnd() - .5f, rnd() - .5f));
    s.te_r = rnd() - .5f; s.te_i = rnd() - .5f; s.tm_r = rnd() - .5f; s.tm_i = rnd() - .5f;
    s.tau = rnd() * 1e-6f;
    const HrtMaterial &m = mats.m[1 + (size_t)(rnd() * 15.99f)];
    const V3 nrm = v3_normalize(v3(rnd() - .5f, rnd() - .5f, rnd() - .5f));
    const V3 mv = v3(rnd() * 30.f - 15.f, rnd() * 30.f - 15.f, 0.f);
    float dist;
    const V3 sd = hrt_shadow_dir(s.o, v3(rnd() * 100.f - 50.f, rnd() * 100.f - 50.f, 1.5f), &dist);
    if (i % 7 == 0) dist = rnd() * 1e-3f;                 /* free-space factor below 1: no division by it */
    const float theta_i = rnd() * 1.5707f;
    const HrtScatterOut a = hrt_scatter_path(s, m, k, nrm, mv, sd, dist, theta_i);
    const HrtScatterOut b = hrt_scatter_path_fast(s, hrt_scat_const(m), k, nrm, mv, sd, dist, theta_i);
    if (memcmp(&a.tau, &b.tau, 4) || memcmp(&a.dfreq, &b.dfreq, 4) || memcmp(&a.dir_rx, &b.dir_rx, 12)) ++bad;
    const double e_te = hypot((double)a.te_r - b.te_r, (double)a.te_i - b.te_i), m_te = hypot((double)a.te_r, (double)a.te_i);
    const double e_tm = hypot((double)a.tm_r - b.tm_r, (double)a.tm_i - b.tm_i), m_tm = hypot((double)a.tm_r, (double)a.tm_i);
    if (m_te > 0 && e_te / m_te > worst) worst = e_te / m_te;
    if (m_tm > 0 && e_tm / m_tm > worst) worst = e_tm / m_tm;
  }
  *worst_rel = worst;
  return bad;
}

/* The closed form of the normalised scattering vector (hrt_scatter_path_auto, what
 * k_scatter runs) against the reference's formulas line by line (hrt_scatter_path),
 * same libm for the latter's transcendentals: worst complex relative deviation of the
 * gains over n random samples -- all 17 materials, incidence angle given as the dot
 * product the reference feeds to acos (or as a primary angle), scattering directions
 * down to grazing.  tau, direction and Doppler term must be identical.  *n_closed =
 * samples that took the closed form.  Returns the number of non-identical exact words. */
extern "C" int emul_scatter_cf_vs_exact(size_t n, uint32_t seed, float f_ghz, double *worst_rel, size_t *n_closed)
{
  HrtMaterialTable mats;
  memset(&mats, 0, sizeof mats);
  for (uint32_t i = 0; i < NUM_G_MATERIALS; ++i) hrt_materials_derive(i, f_ghz, (HrtMaterialDerived *)&mats.m[i]);
  uint64_t st = 0x9E3779B97F4A7C15ull ^ seed;
  auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (float)((st >> 40) * (1.0 / 16777216.0)); };
  HrtRunConst k;
  const float f_hz = (float)((double)f_ghz * 1e9);
  k.fsl_k = 4.f * HRT_PI * f_hz / HRT_C0; k.dop_k = f_hz / HRT_C0;
  int bad = 0; double worst = 0.0; size_t closed = 0;
  for (size_t i = 0; i < n; ++i) {
    HrtRayState s;
    s.o = v3(rnd() * 100.f - 50.f, rnd() * 100.f - 50.f, rnd() * 20.f);
    s.d = v3_normalize(v3(rnd() - .5f, rnd() - .5f, rnd() - .5f));
    s.te_r = rnd() - .5f; s.te_i = rnd() - .5f; s.tm_r = rnd() - .5f; s.tm_i = rnd() - .5f;
    s.tau = rnd() * 1e-6f;
    const HrtMaterial &m = mats.m[(size_t)(rnd() * 16.99f)];
    float dist;
    const V3 rxp = v3(rnd() * 100.f - 50.f, rnd() * 100.f - 50.f, 1.5f);
    const V3 sd = hrt_shadow_dir(s.o, rxp, &dist);
    /* surface normal: random, or nearly perpendicular to the scattering direction (grazing, |cos theta_s| down to 1e-7) */
    V3 nrm = v3_normalize(v3(rnd() - .5f, rnd() - .5f, rnd() - .5f));
    if (i % 3 == 0) {
      const V3 perp = v3_normalize(v3_cross(sd, nrm));
      const float eps = powf(10.f, -7.f * rnd()) * (rnd() < .5f ? -1.f : 1.f);
      nrm = v3_normalize(v3_add(perp, v3_scale(sd, eps)));
    }
    const V3 mv = v3(rnd() * 30.f - 15.f, rnd() * 30.f - 15.f, 0.f);
    if (i % 7 == 0) dist = rnd() * 1e-3f;
    /* incidence: a shadow-hit dot product in [-1, 1] (dense near 0 and +-1), or the primary angle */
    float cx = HRT_CX_PRIMARY, theta_p = rnd() * 1.5707f, theta_i = theta_p;
    if (i % 5 != 0) {
      const float u = rnd();
      cx = (i % 2) ? (2.f * u - 1.f) : ((i % 4) ? powf(10.f, -7.f * u) : 1.f - powf(10.f, -7.f * u));
      if (rnd() < .5f) cx = -cx;
      theta_i = (float)acos((double)cx);
      if ((double)theta_i > (double)HRT_PI / 2.) theta_i = HRT_SUB(HRT_PI, theta_i);
    }
    const HrtScatterOut a = hrt_scatter_path(s, m, k, nrm, mv, sd, dist, theta_i);
    const HrtScatCf mcf = hrt_scat_cf(m);
    float t0, t1, t2, t3, ci = cosf(theta_p), si = sinf(theta_p);
    if (cx != HRT_CX_PRIMARY) hrt_fold_cos_sin(cx, &ci, &si);
    if (hrt_scatter_gains_cf(s, mcf, k.fsl_k, v3_dot(sd, nrm), dist, ci, si, &t0, &t1, &t2, &t3)) ++closed;
    const HrtScatterOut b = hrt_scatter_path_auto(s, hrt_scat_const(m), mcf, k, nrm, mv, sd, dist, cx, theta_p, cosf(theta_p), sinf(theta_p));
    if (memcmp(&a.tau, &b.tau, 4) || memcmp(&a.dfreq, &b.dfreq, 4) || memcmp(&a.dir_rx, &b.dir_rx, 12)) ++bad;
    const double e_te = hypot((double)a.te_r - b.te_r, (double)a.te_i - b.te_i), m_te = hypot((double)a.te_r, (double)a.te_i);
    const double e_tm = hypot((double)a.tm_r - b.tm_r, (double)a.tm_i - b.tm_i), m_tm = hypot((double)a.tm_r, (double)a.tm_i);
    if (m_te > 0 && e_te / m_te > worst) worst = e_te / m_te;
    if (m_tm > 0 && e_tm / m_tm > worst) worst = e_tm / m_tm;
    if ((m_te == 0 && e_te != 0) || (m_tm == 0 && e_tm != 0)) ++bad;
  }
  *worst_rel = worst; *n_closed = closed;
  return bad;
}

/* Receiver maps against the brute-force loop over every triangle (same
 * hrt_mt_test): for every origin x receiver the shadow query through the map must
 * return the same (triangle, t).  Returns the number of differing queries;
 * *avg_tests = triangle tests per query through the map. */
extern "C" long emul_rxmap_vs_brute(const Scene *sc, const Vec3 *rx, size_t R, const Vec3 *origins, size_t n,
                                    uint32_t G, double *avg_tests, double *avg_list, const uint32_t *origin_gid)
{
  EmulScene E;
  float ma = 0.f;
  for (size_t i = 0; i < R; ++i) ma = fmaxf(ma, fmaxf(fabsf(rx[i].x), fmaxf(fabsf(rx[i].y), fabsf(rx[i].z))));
  for (size_t i = 0; i < n; ++i) ma = fmaxf(ma, fmaxf(fabsf(origins[i].x), fmaxf(fabsf(origins[i].y), fabsf(origins[i].z))));
  build(sc, E, 2, 64.f, ma);
  EmulRxMap M;
  build_rxmap(E, rx, R, G, 4.f * E.pad, M);
  unsigned long long tests = 0, n_sure = 0; long bad = 0;
  static unsigned long long dbg[6]; if (getenv("EMUL_RXMAP_VERBOSE")) { memset(dbg, 0, sizeof dbg); g_dbg = dbg; }
  HrtNoCount nc;
  HrtGlobalMem m; m.nodes = E.nodes.data(); m.tris = E.tris.data(); m.wnodes = E.wnodes.data();
  for (size_t i = 0; i < n; ++i) {
    /* "own" triangle of the origin, for the early-out of the triangle a ray starts on (valid for any triangle:
     * the one the caller names, else the one whose plane is nearest) */
    uint32_t own = 0; float own_pd = 3e38f;
    if (origin_gid && origin_gid[i] != HRT_NONE) {
      for (uint32_t s = 0; s < E.n; ++s) if (E.gid[s] == origin_gid[i]) own = s;      /* the triangle the origin was placed on */
    } else
    for (uint32_t s = 0; s < E.n; ++s) {
      const float4 q0 = E.tris[3 * s], q2 = E.tris[3 * s + 2];
      const float pd = fabsf((origins[i].x - q0.x) * q2.y + (origins[i].y - q0.y) * q2.z + (origins[i].z - q0.z) * q2.w);
      if (pd < own_pd) { own_pd = pd; own = s; }
    }
    const float own_nt = hrt_mt_self_nt(E.tris[3 * own], E.tris[3 * own + 1], E.tris[3 * own + 2], tov(origins[i]));
    for (size_t r = 0; r < R; ++r) {
      float dist;
      const V3 o = tov(origins[i]);
      const V3 sd = hrt_shadow_dir(o, tov(rx[r]), &dist);
      if (!(dist > 0.f)) continue;
      bool sure = false;
      const HrtHit a = query_map(E, M, r, o, sd, dist, own, own_nt, &tests, &sure);
      const HrtHit b = hrt_closest_hit_brute(m, E.gid.data(), E.n, o, sd, nc);
      n_sure += sure;
      if (getenv("EMUL_RXMAP_VERBOSE")) {
        static unsigned long long cat[6]; static unsigned long long seen = 0;
        uint32_t cp_, cn_; hrt_rxmap_cells2(sd, M.G, &cp_, &cn_);
        const size_t cell = r * 6 * M.G * M.G + cp_;
        const bool near_hit = b.gid != HRT_NONE && b.t < dist * 0.999f;
        if (near_hit) ++cat[0];
        else if (M.count[cell] == 0) ++cat[1];
        else if (M.count[cell] == 1 && M.sure[cell]) ++cat[2];
        else if (M.count[cell] == 1) ++cat[3];
        else if (M.count[cell] == 2) ++cat[4];
        else ++cat[5];
        if (++seen == n * R) fprintf(stderr, "far side: near hit %llu, empty %llu, sure %llu, single not sure %llu, two %llu, more %llu\n", cat[0], cat[1], cat[2], cat[3], cat[4], cat[5]);
      }
      /* a "sure" answer names the triangle and promises t >= dist > 1 without computing t */
      if (sure ? (a.gid != b.gid || !(b.t > 1.f) || !(b.t >= dist * 0.999f)) : (a.gid != b.gid || (a.gid != HRT_NONE && memcmp(&a.t, &b.t, 4)))) ++bad;
    }
  }
  *avg_tests = (double)tests / (double)(n * R);
  *avg_list = (double)M.items.size() / (double)M.start.size();
  if (g_dbg) { fprintf(stderr, "near side: own-plane tests %llu, other %llu, accepted %llu; far side: own-plane %llu, other %llu, accepted %llu\n", dbg[0], dbg[1], dbg[4], dbg[2], dbg[3], dbg[5]); g_dbg = nullptr; }
  if (const char *e = getenv("EMUL_RXMAP_VERBOSE")) { (void)e; fprintf(stderr, "rxmap: %llu of %zu queries answered by a sure cell\n", n_sure, n * R); }
  return bad;
}

/* Opt-in extensions (hrt_ext.cuh), fp32 as the kernels evaluate them, for comparison with the
 * oracle's double-precision statement (tests/test_extensions.py). */
extern "C" void emul_ext_refr_coefs(uint32_t material, float f_ghz, float theta1, float out[4], float dir_in[3], float nrm[3], float dir_out[3], int *ok)
{
  HrtMaterial m; memset(&m, 0, sizeof m);
  hrt_materials_derive(material, f_ghz, (HrtMaterialDerived *)&m);
  hrt_refr_coefs(m, theta1, out);
  V3 t;
  *ok = hrt_refract_dir(m, v3(dir_in[0], dir_in[1], dir_in[2]), v3(nrm[0], nrm[1], nrm[2]), &t) ? 1 : 0;
  dir_out[0] = t.x; dir_out[1] = t.y; dir_out[2] = t.z;
}
extern "C" float emul_ext_pattern_pi(float s1, float s2, float s3, int a1, int a3, const float ki[3], const float ks[3], const float n[3])
{ return hrt_scat_pattern_pi(s1, s2, s3, a1, a3, v3(ki[0], ki[1], ki[2]), v3(ks[0], ks[1], ks[2]), v3(n[0], n[1], n[2])); }
extern "C" float emul_ext_lobe_norm(int alpha, float cos_i, float sin_i) { return hrt_lobe_norm(alpha, cos_i, sin_i); }

/* "Sure" cells (hrt_rxmap_sure) against the exact test: random receivers, triangles and cells; for every cell
 * the predicate accepts, rays from random origins on the far side of the receiver (up to `reach` away), aimed at
 * the receiver with the kernels' own fp32 direction and continuing through the cell, must be accepted by
 * hrt_mt_test with t >= the distance to the receiver.  Returns the number of violations; *n_sure = cells accepted,
 * *n_rays = rays tested. */
extern "C" long emul_sure_check(size_t n, uint32_t seed, float scale, unsigned long long *n_sure, unsigned long long *n_rays)
{
  uint64_t st = seed * 0x9E3779B97F4A7C15ull + 12345u;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (float)((st >> 11) & 0xFFFFFF) / 16777216.f; };
  long bad = 0; *n_sure = 0; *n_rays = 0;
  HrtNoCount nc;
  for (size_t i = 0; i < n; ++i) {
    const V3 apex = v3((rnd() - 0.5f) * scale, (rnd() - 0.5f) * scale, rnd() * 0.1f * scale);
    const uint32_t G = 32u << (uint32_t)(rnd() * 3.99f);                 /* 32 .. 256 */
    const uint32_t face = (uint32_t)(rnd() * 5.99f), ci = (uint32_t)(rnd() * (G - 0.01f)), cj = (uint32_t)(rnd() * (G - 0.01f));
    const HrtPyramid cp = hrt_rxmap_pyramid(face, G, ci, ci + 1, cj, cj + 1);
    /* a triangle around the cell's centre ray at a random distance, random orientation and size */
    const V3 mid = v3((cp.c[0].x + cp.c[2].x) * 0.5f, (cp.c[0].y + cp.c[2].y) * 0.5f, (cp.c[0].z + cp.c[2].z) * 0.5f);
    const float dist = 0.05f * scale * (0.02f + rnd());
    const V3 ctr = v3(apex.x + mid.x * dist, apex.y + mid.y * dist, apex.z + mid.z * dist);
    const float size = dist * (4.f + 200.f * rnd()) / (float)G;
    V3 p[3];
    for (int k = 0; k < 3; ++k) p[k] = v3(ctr.x + (rnd() - 0.5f) * size, ctr.y + (rnd() - 0.5f) * size, ctr.z + (rnd() - 0.5f) * size);
    const HrtTriSetup S = hrt_tri_setup(p[0], p[1], p[2]);
    V3 va, vb, vc; hrt_rxmap_corners(S.q0, S.q1, S.q2, apex, &va, &vb, &vc);
    const float max_abs = scale;                                        /* coordinates stay within +-scale */
    const float pad = 4.f * hrt_box_pad(max_abs, 64.f), reach = 2.f * scale * 1.8f;
    if (!hrt_rxmap_sure(cp, va, vb, vc, pad, reach)) continue;
    ++*n_sure;
    for (int r = 0; r < 64; ++r) {
      /* a direction inside the cell (bilinear mix of the corner directions), an origin on the other side of the apex */
      const float a = rnd(), b = rnd();
      V3 w = v3((cp.c[0].x * (1 - a) + cp.c[1].x * a) * (1 - b) + (cp.c[3].x * (1 - a) + cp.c[2].x * a) * b,
                (cp.c[0].y * (1 - a) + cp.c[1].y * a) * (1 - b) + (cp.c[3].y * (1 - a) + cp.c[2].y * a) * b,
                (cp.c[0].z * (1 - a) + cp.c[1].z * a) * (1 - b) + (cp.c[3].z * (1 - a) + cp.c[2].z * a) * b);
      const float back = 1.01f + rnd() * rnd() * scale * 0.9f;
      const V3 o = v3(apex.x - w.x * back, apex.y - w.y * back, apex.z - w.z * back);
      float dd;
      const V3 d = hrt_shadow_dir(o, apex, &dd);
      if (!(dd > 1.001f)) continue;
      uint32_t c_pos, c_neg; hrt_rxmap_cells2(d, G, &c_pos, &c_neg);
      if (c_pos != (face * G + cj) * G + ci) continue;                  /* rounded into a neighbouring cell: not this cell's query */
      ++*n_rays;
      float t = 0.f;
      const bool hit = hrt_mt_test<HrtNoCount, true>(S.q0, S.q1, S.q2, o, d, HRT_T_MAX, 0u, 0u, &t, nc);
      if (!hit || !(t >= dd * 0.999f) || !(t > 1.f)) ++bad;
    }
  }
  return bad;
}
