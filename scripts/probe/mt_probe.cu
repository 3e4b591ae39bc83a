// GPU-vs-host probe of one Moeller-Trumbore evaluation (debug aid, not product)
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
#include "../../hermespy-rt_b200/csrc/hrt_core.cuh"

struct Dump { float pv[3], det, sv[3], nu, snu, ad, qv[3], nv, snv, nt, snt, u, v, t; int accepted; float tout; };

__host__ __device__ void probe(V3 a, V3 b, V3 c, V3 o, V3 d, Dump *D)
{
  const V3 ab = v3_sub(b, a), ac = v3_sub(c, a);
  const V3 pv = v3_cross(d, ac);
  const float det = v3_dot(ab, pv);
  const V3 sv = v3_sub(o, a);
  const float nu = v3_dot(sv, pv);
  const V3 qv = v3_cross(sv, ab);
  const float nv = v3_dot(d, qv), nt = v3_dot(ac, qv);
  D->pv[0] = pv.x; D->pv[1] = pv.y; D->pv[2] = pv.z; D->det = det;
  D->sv[0] = sv.x; D->sv[1] = sv.y; D->sv[2] = sv.z; D->nu = nu; D->snu = HRT_MUL(nu, det); D->ad = HRT_MUL(det, det);
  D->qv[0] = qv.x; D->qv[1] = qv.y; D->qv[2] = qv.z; D->nv = nv; D->snv = HRT_MUL(nv, det); D->nt = nt; D->snt = HRT_MUL(nt, det);
  D->u = HRT_DIV(nu, det); D->v = HRT_DIV(nv, det); D->t = HRT_DIV(nt, det);
  float4 q0, q1, q2;
  q0.x = a.x; q0.y = a.y; q0.z = a.z; q0.w = ab.x; q1.x = ab.y; q1.y = ab.z; q1.z = ac.x; q1.w = ac.y; q2.x = ac.z; q2.y = q2.z = q2.w = 0.f;
  HrtNoCount nc; float t = -1.f;
  D->accepted = hrt_mt_test(q0, q1, q2, o, d, HRT_T_MAX, HRT_NONE, 0u, &t, nc) ? 1 : 0;
  D->tout = t;
}
__global__ void k(V3 a, V3 b, V3 c, V3 o, V3 d, Dump *D) { probe(a, b, c, o, d, D); }

static void show(const char *w, const Dump &D)
{
  printf("%s pv %a %a %a det %a\n   sv %a %a %a nu %a snu %a ad %a\n   qv %a %a %a nv %a snv %a nt %a snt %a\n   u %a v %a t %a accepted %d t %g\n", w,
         D.pv[0], D.pv[1], D.pv[2], D.det, D.sv[0], D.sv[1], D.sv[2], D.nu, D.snu, D.ad, D.qv[0], D.qv[1], D.qv[2], D.nv, D.snv, D.nt, D.snt, D.u, D.v, D.t, D.accepted, D.tout);
}
int main()
{
  const V3 a = v3(-0.66000003f, 4.69999981f, 1.5f), b = v3(-2.20000005f, 4.69999981f, 0.75f), c = v3(-2.20000005f, 6.49999952f, 0.7500003f);
  const V3 o = v3(-8.600656509399414f, -5.7860188484191895f, -2.3672056198120117f), d = v3(0.5171605944633484f, 0.8179910778999329f, 0.25186413526535034f);
  Dump h, g, *dg; probe(a, b, c, o, d, &h);
  cudaMalloc(&dg, sizeof(Dump)); k<<<1, 1>>>(a, b, c, o, d, dg); cudaMemcpy(&g, dg, sizeof g, cudaMemcpyDeviceToHost);
  show("host", h); show("gpu ", g);
  printf("%s\n", memcmp(&h, &g, sizeof h) ? "DIFFERENT" : "identical");
  return 0;
}
