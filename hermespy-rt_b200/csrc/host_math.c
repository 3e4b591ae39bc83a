/* host_math.c -- the two pieces of arithmetic that stay on the host CPU because
 * their results must come from the host libm the reference itself uses.
 *
 *  hrt_materials_derive : reference precompute_materials + csqrtf
 *                         (src/compute_paths.c:136-151, :171-206).  17 table rows
 *                         per call; powf comes from glibc exactly as there.
 *  hrt_host_launch_dir  : reference Fibonacci launch direction (:444-451) for
 *                         the few rays (about 1 in 5e5) whose GPU result sits on
 *                         an fp32 rounding boundary (hrt_core.cuh,
 *                         hrt_launch_dir).  This is not a CPU fallback of the
 *                         path: no ray is ever traced on the host.
 *
 * Build: gcc -O3 -ffp-contract=off, no -march (no FMA), like the reference
 * (GNUmakefile:2,14).
 */
#include "../../include/hrt_cuda.h"

#include <float.h>

#define HM_PI 3.14159265358979323846f   /* reference src/compute_paths.c:18 */

static void hm_csqrt(float re, float im, float mag, float *o_re, float *o_im)
{
  *o_re = sqrtf((re + mag) / 2.f);
  if (fabsf(im) < FLT_EPSILON && re >= -FLT_EPSILON) { *o_im = 0.f; return; }
  float v = sqrtf((mag - re) / 2.f);
  *o_im = im < 0.f ? -v : v;
}

void hrt_materials_derive(uint32_t index, float f_ghz, HrtMaterialDerived *o)
{
  const Material *m = &g_materials[index < NUM_G_MATERIALS ? index : 0];
  float eta_re = m->a * powf(f_ghz, m->b);
  float eta_im = (m->c * powf(f_ghz, m->d)) / (0.0556325027352135f * f_ghz);
  float abs2 = eta_re * eta_re + eta_im * eta_im;
  float mag = sqrtf(abs2);
  memset(o, 0, sizeof *o);
  o->eta_abs2 = abs2;
  o->eta_abs_inv_sqrt = 1.f / sqrtf(mag);
  hm_csqrt(eta_re, eta_im, mag, &o->sqrt_re, &o->sqrt_im);
  o->inv_re = eta_re / abs2;
  o->inv_im = -eta_im / abs2;
  o->r = 1.f - m->s;
  o->s = m->s;
  o->s1_alpha = (float)m->s1_alpha;
}

void hrt_host_launch_dir(uint64_t path, uint64_t num_paths, float out[3])
{
  float k = (float)path + .5f;
  float phi = (float)acos((double)(1.f - 2.f * k / (float)num_paths));
  float th = HM_PI * (1.f + sqrtf(5.f)) * k;
  out[0] = (float)(cos((double)th) * sin((double)phi));
  out[1] = (float)(sin((double)th) * sin((double)phi));
  out[2] = (float)cos((double)phi);
}
