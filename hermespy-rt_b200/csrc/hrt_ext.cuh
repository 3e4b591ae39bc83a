/* hrt_ext.cuh -- OPT-IN EXTENSIONS next to the path.
 *
 * The reference leaves two things as TODOs right inside compute_paths():
 *   - "Spawn refraction rays with gains as per eqs. (31c)-(31d) from ITU-R
 *     P.2040-3" (src/compute_paths.c:587, :726-728), and
 *   - "Incorporate s2 and s3" into scat_coefs (:414); Material carries s1, s2,
 *     s3, s1_alpha, s3_alpha for it (inc/scene.h:47-65) and the two binomial
 *     tables of :22-120 are the constants of the directive-lobe normalisation.
 * There is no reference behaviour to match, so nothing here runs unless the
 * caller sets HRT_FLAG_EXT_LOBES / HRT_FLAG_EXT_REFRACT; the oracle carries the
 * same definitions in double precision (oracle/hrt_oracle.c, "extensions") and
 * the tests compare against those.
 *
 * (1) Transmission coefficients, ITU-R P.2040-3 eqs. (31c), (31d), with the
 *     complex relative permittivity eta = eta' - j eta'' of eq. (9b):
 *       T_TE = 2 cos(th1) / (cos(th1) + sqrt(eta - sin^2 th1))
 *       T_TM = 2 sqrt(eta) cos(th1) / (eta cos(th1) + sqrt(eta - sin^2 th1))
 *     and the refracted direction by Snell's law with n = Re sqrt(eta).
 * (2) Scattering pattern of the three-lobe model the Material fields describe
 *     (Degli-Esposti et al., IEEE TAP 55(1), 2007, eqs. (5)-(10); the same
 *     model Sionna RT calls Lambertian / Directive / Backscattering pattern):
 *       f_s = s1 f_dir + s2 f_lamb + s3 f_back,
 *       f_lamb = cos(th_s) / pi,
 *       f_dir  = ((1 + k_r . k_s) / 2)^a1 / F_a1(th_i),
 *       f_back = ((1 - k_i . k_s) / 2)^a3 / F_a3(th_i),
 *       F_a(th_i) = 2^-a sum_{k=0..a} C(a,k) I_k,
 *       I_k = 2 pi / (k+1) * { 1                                        k even
 *                            { cos(th_i) sum_{w=0..(k-1)/2} C(2w,w) (sin(th_i)/2)^(2w)   k odd
 *     (k_i incident, k_r specular, k_s scattered unit directions; every lobe
 *     integrates to 1 over the hemisphere above the surface).  With
 *     HRT_FLAG_EXT_LOBES the scattered amplitude of a path is
 *       s sqrt(pi f_s) * (1, p, g, g p) / sqrt((1 + g^2)(1 + p^2))
 *     i.e. the reference's polarisation mix (its normalised scat_coefs vector,
 *     hrt_scat_unit) scaled by the pattern -- instead of the reference's
 *     placeholder lobe, which its own normalisation (:399-405) cancels.
 */
#pragma once

#include "hrt_core.cuh"

#define HRT_EXT_ALPHA_MAX 19     /* the reference's tables stop at alpha = 19 (:22-120) */

/* complex square root, principal branch */
HRT_HD void hrt_csqrt(float re, float im, float *o_re, float *o_im)
{
  const float mag = sqrtf(re * re + im * im);
  const float r = sqrtf(fmaxf(0.5f * (mag + re), 0.f));
  const float i = sqrtf(fmaxf(0.5f * (mag - re), 0.f));
  *o_re = r; *o_im = im < 0.f ? -i : i;
}

/* eta' and eta'' (both >= 0; eta = eta' - j eta'') back from the derived constants */
HRT_HD void hrt_eta_of(const HrtMaterial &m, float *eta_re, float *eta_im_pos)
{
  *eta_re = m.inv_re * m.eta_abs2;
  *eta_im_pos = -m.inv_im * m.eta_abs2;
}

/* out = (T_TE re, im, T_TM re, im), eqs. (31c), (31d) */
HRT_HD void hrt_refr_coefs(const HrtMaterial &m, float theta1, float out[4])
{
  float er, ei;
  hrt_eta_of(m, &er, &ei);
  const float c1 = cosf(theta1), s1 = sinf(theta1);
  float qr, qi;                                         /* sqrt(eta - sin^2 th1), eta = er - j ei */
  hrt_csqrt(er - s1 * s1, -ei, &qr, &qi);
  /* T_TE = 2 c1 / (c1 + q) */
  {
    const float dr = c1 + qr, di = qi, den = dr * dr + di * di;
    out[0] = 2.f * c1 * dr / den; out[1] = -2.f * c1 * di / den;
  }
  /* T_TM = 2 sqrt(eta) c1 / (eta c1 + q) */
  {
    float sr, si;
    hrt_csqrt(er, -ei, &sr, &si);
    const float nr = 2.f * c1 * sr, ni = 2.f * c1 * si;
    const float dr = er * c1 + qr, di = -ei * c1 + qi, den = dr * dr + di * di;
    out[2] = (nr * dr + ni * di) / den; out[3] = (ni * dr - nr * di) / den;
  }
}

/* refracted unit direction (Snell, n = Re sqrt(eta)); d: incident unit direction, nrm: unit normal of
 * either orientation.  Returns false on total internal reflection (cannot happen entering a denser medium). */
HRT_HD bool hrt_refract_dir(const HrtMaterial &m, V3 d, V3 nrm, V3 *out)
{
  float er, ei, sr, si;
  hrt_eta_of(m, &er, &ei);
  hrt_csqrt(er, -ei, &sr, &si);
  const float n = fmaxf(sr, 1e-6f);
  float c1 = -(d.x * nrm.x + d.y * nrm.y + d.z * nrm.z);
  if (c1 < 0.f) { nrm = v3(-nrm.x, -nrm.y, -nrm.z); c1 = -c1; }     /* normal against the incident ray */
  const float k = 1.f / n, s2sq = k * k * fmaxf(1.f - c1 * c1, 0.f);
  if (s2sq > 1.f) return false;
  const float c2 = sqrtf(1.f - s2sq), f = k * c1 - c2;
  V3 t = v3(k * d.x + f * nrm.x, k * d.y + f * nrm.y, k * d.z + f * nrm.z);
  const float l = 1.f / sqrtf(t.x * t.x + t.y * t.y + t.z * t.z);
  *out = v3(t.x * l, t.y * l, t.z * l);
  return true;
}

/* F_a(th_i) of the directive / backscatter lobes; the binomials are generated on the fly
 * (exact in fp32 up to C(19, 9) = 92378 and C(18, 9) = 48620) */
HRT_HD float hrt_lobe_norm(int alpha, float cos_i, float sin_i)
{
  if (alpha > HRT_EXT_ALPHA_MAX) alpha = HRT_EXT_ALPHA_MAX;
  const float q = 0.25f * sin_i * sin_i;          /* (sin/2)^2 */
  float sum = 0.f, c_ak = 1.f;                    /* C(alpha, k) */
  float odd_series = 0.f, c_2ww = 1.f, qw = 1.f;  /* sum_{w <= (k-1)/2} C(2w,w) q^w, extended as k grows */
  int w_done = -1;
  for (int k = 0; k <= alpha; ++k) {
    float ik = 6.283185307179586f / (float)(k + 1);
    if (k & 1) {
      const int wmax = (k - 1) / 2;
      while (w_done < wmax) {
        ++w_done;
        if (w_done > 0) { c_2ww = c_2ww * (float)(2 * w_done) * (float)(2 * w_done - 1) / ((float)w_done * (float)w_done); qw *= q; }
        odd_series += c_2ww * qw;
      }
      ik *= cos_i * odd_series;
    }
    sum += c_ak * ik;
    c_ak = c_ak * (float)(alpha - k) / (float)(k + 1);
  }
  return ldexpf(sum, -alpha);
}

HRT_HD float hrt_powi(float x, int n) { float r = 1.f; for (int k = 0; k < n; ++k) r *= x; return r; }

/* pi * f_s for the three-lobe model; ki: incident direction (towards the surface), ks: scattered
 * direction (away from it), n: unit normal on the incidence side.  s1 + s2 + s3 is taken as given. */
HRT_HD float hrt_scat_pattern_pi(float s1, float s2, float s3, int a1, int a3, V3 ki, V3 ks, V3 n)
{
  float cos_i = -(ki.x * n.x + ki.y * n.y + ki.z * n.z);
  if (cos_i < 0.f) { n = v3(-n.x, -n.y, -n.z); cos_i = -cos_i; }
  cos_i = fminf(cos_i, 1.f);
  const float sin_i = sqrtf(fmaxf(1.f - cos_i * cos_i, 0.f));
  const float cos_s = ks.x * n.x + ks.y * n.y + ks.z * n.z;
  if (!(cos_s > 0.f)) return 0.f;                       /* scattered into the surface */
  const float two_c = 2.f * cos_i;
  const V3 kr = v3(ki.x + two_c * n.x, ki.y + two_c * n.y, ki.z + two_c * n.z);      /* specular direction */
  const float dir = 0.5f * (1.f + (kr.x * ks.x + kr.y * ks.y + kr.z * ks.z));
  const float back = 0.5f * (1.f - (ki.x * ks.x + ki.y * ks.y + ki.z * ks.z));
  const float pi = 3.14159265358979f;
  float f = s2 * cos_s;
  if (s1 != 0.f) f += s1 * pi * hrt_powi(fmaxf(dir, 0.f), a1) / hrt_lobe_norm(a1, cos_i, sin_i);
  if (s3 != 0.f) f += s3 * pi * hrt_powi(fmaxf(back, 0.f), a3) / hrt_lobe_norm(a3, cos_i, sin_i);
  return f;
}

/* raw scattering parameters of g_materials the extension needs (uploaded next to HrtMaterialTable) */
struct HrtExtMaterial { float s1, s2, s3; int a1, a3; };
struct HrtExtTable { HrtExtMaterial m[17]; };

/* One scatter path with the extended pattern (HRT_FLAG_EXT_LOBES): same delay, direction and
 * Doppler term as the reference path; gains = state * s sqrt(pi f_s) * unit mix / free-space loss. */
HRT_HD HrtScatterOut hrt_scatter_path_ext(const HrtRayState &s, const HrtScatCf &mcf, const HrtExtMaterial &e,
                                          const HrtRunConst &k, V3 n, V3 mesh_vel, V3 sd, float dist,
                                          V3 k_inc, float ci, float si)
{
  HrtScatterOut r;
  float g, p, inv_n;
  hrt_scat_unit(mcf, ci, si, &g, &p, &inv_n);
  const float amp = mcf.s * sqrtf(fmaxf(hrt_scat_pattern_pi(e.s1, e.s2, e.s3, e.a1, e.a3, k_inc, sd, n), 0.f));
  float kk = amp * inv_n;
  float l2 = HRT_MUL(k.fsl_k, dist);                                           /* :711 */
  l2 = HRT_MUL(l2, l2);
  if (l2 > 1.f) kk /= l2;                                                      /* :713 */
  const float gk = g * kk;
  r.te_r = HRT_FMA(-s.te_i, p, s.te_r) * kk;
  r.te_i = HRT_FMA(s.te_r, p, s.te_i) * kk;
  r.tm_r = HRT_FMA(-s.tm_i, p, s.tm_r) * gk;
  r.tm_i = HRT_FMA(s.tm_r, p, s.tm_i) * gk;
  r.dir_rx = v3(-sd.x, -sd.y, -sd.z);                                          /* :707 */
  r.tau = HRT_ADD(s.tau, HRT_DIV(dist, HRT_C0));                               /* :709 */
  r.dfreq = HRT_MUL(v3_dot(v3_sub(sd, s.d), mesh_vel), k.dop_k);               /* :720-721 */
  return r;
}
