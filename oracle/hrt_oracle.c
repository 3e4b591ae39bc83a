/* hrt_oracle.c -- CPU restatement of hermespy-rt's compute_paths().
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it, and only as the checker
 * or as the timed CPU baseline.  The product library (libhermespy_rt.so) never
 * links or calls this file.
 *
 * Parity pinning: the reference ships no golden vectors (its test/test.py only
 * asserts array shapes), so this restatement is pinned by executing the
 * reference itself: oracle/Makefile compiles /root/reference/src/*.c unmodified
 * into oracle/_ref/libhrt_ref.so, tests/test_oracle_vs_ref.py compares every
 * reference-written output word bit-for-bit, and tests/golden/ holds vectors
 * generated from that build (tests/golden/make_golden.py).
 *
 * What this file adds over the reference: a trace (hit triangle, hit distance,
 * incidence angle per (tx,bounce,path); shadow-query state per output slot), so
 * that "hit indices bit-exact" is testable -- the reference API does not expose
 * them.  The arithmetic (operation order, float/double promotion) follows
 * /root/reference/src/compute_paths.c line by line as cited below; build with
 * -ffp-contract=off and no -march so no FMA is formed (reference:
 * GNUmakefile:2,14, plain gcc -O3).
 */
#include "../include/hermespy_rt.h"

#include <float.h>
#include <stdio.h>

#define ORC_PI   3.14159265358979323846f   /* reference src/compute_paths.c:18 */
#define ORC_C0   299792458.0f              /* reference src/compute_paths.c:19 */
#define ORC_EPS  FLT_EPSILON
#define ORC_NONE 0xFFFFFFFFu               /* closest-hit query found nothing */
#define ORC_IDLE 0xFFFFFFFEu               /* ray was already dead, not traced */

/* Flattened copy of the scene in (mesh, face) order; tri id = running index. */
typedef struct {
  uint32_t num_tris;
  Vec3 *a, *b, *c;       /* the three corners of each triangle */
  Vec3 *n;               /* unit normal, reference :208-224 */
  uint32_t *mesh_of;     /* owning mesh */
  uint32_t *face_of;     /* index inside the mesh */
} OrcTris;

/* Per-material derived constants, reference :125-132 and :171-206. */
typedef struct {
  float eta_re, eta_im, eta_abs, eta_abs2, eta_abs_inv_sqrt;
  float sqrt_re, sqrt_im, inv_re, inv_im, inv_sqrt_re, inv_sqrt_im, r;
} OrcMat;

/* Optional trace, all arrays caller-allocated (NULL = not wanted). */
typedef struct {
  uint32_t *hit_tri;     /* [T][B][P] global triangle id / ORC_NONE / ORC_IDLE */
  float    *hit_t;       /* [T][B][P] distance of the primary hit            */
  float    *hit_theta;   /* [T][B][P] folded incidence angle                 */
  uint8_t  *slot_state;  /* [R][T][B][P] 0 untouched, 1 path, 2 occluded     */
  uint32_t *shadow_tri;  /* [R][T][B][P] closest triangle of the shadow query */
  float    *theta_used;  /* [R][T][B][P] incidence angle given to scat_coefs */
} OrcTrace;

/* ---- scene flattening + normals (reference :208-224) --------------------- */

static int orc_flatten(const Scene *sc, OrcTris *ft)
{
  uint32_t total = 0;
  for (uint32_t m = 0; m < sc->num_meshes; ++m) total += sc->meshes[m].num_triangles;
  ft->num_tris = total;
  size_t n = total ? total : 1;
  ft->a = malloc(n * sizeof(Vec3)); ft->b = malloc(n * sizeof(Vec3));
  ft->c = malloc(n * sizeof(Vec3)); ft->n = malloc(n * sizeof(Vec3));
  ft->mesh_of = malloc(n * sizeof(uint32_t));
  ft->face_of = malloc(n * sizeof(uint32_t));
  if (!ft->a || !ft->b || !ft->c || !ft->n || !ft->mesh_of || !ft->face_of) return -1;
  uint32_t g = 0;
  for (uint32_t m = 0; m < sc->num_meshes; ++m) {
    const Mesh *me = &sc->meshes[m];
    for (uint32_t f = 0; f < me->num_triangles; ++f, ++g) {
      ft->a[g] = me->vs[me->is[3 * f + 0]];
      ft->b[g] = me->vs[me->is[3 * f + 1]];
      ft->c[g] = me->vs[me->is[3 * f + 2]];
      Vec3 ab = vec3_sub(&ft->b[g], &ft->a[g]);       /* :217 */
      Vec3 ac = vec3_sub(&ft->c[g], &ft->a[g]);       /* :218 */
      Vec3 nn = vec3_cross(&ab, &ac);                 /* :219 */
      ft->n[g] = vec3_normalize(&nn);                 /* :220 */
      ft->mesh_of[g] = m;
      ft->face_of[g] = f;
    }
  }
  return 0;
}

static void orc_free_tris(OrcTris *ft)
{
  free(ft->a); free(ft->b); free(ft->c); free(ft->n);
  free(ft->mesh_of); free(ft->face_of);
}

/* ---- material table (reference :136-151 csqrtf, :171-206) ---------------- */

static void orc_csqrt(float re, float im, float mag, float *o_re, float *o_im)
{
  *o_re = sqrtf((re + mag) / 2.f);                              /* :144 */
  if (fabsf(im) < ORC_EPS && re >= -ORC_EPS) { *o_im = 0.f; return; } /* :145 */
  float v = sqrtf((mag - re) / 2.f);                            /* :148 */
  *o_im = (im < 0.f) ? -v : v;                                  /* :149 */
}

void orc_material(uint32_t index, float f_ghz, OrcMat *o)
{
  const Material *m = &g_materials[index];
  o->eta_re   = m->a * powf(f_ghz, m->b);                                   /* :184 */
  o->eta_im   = (m->c * powf(f_ghz, m->d)) / (0.0556325027352135f * f_ghz); /* :186 */
  o->eta_abs2 = o->eta_re * o->eta_re + o->eta_im * o->eta_im;              /* :188 */
  o->eta_abs  = sqrtf(o->eta_abs2);                                         /* :190 */
  o->eta_abs_inv_sqrt = 1.f / sqrtf(o->eta_abs);                            /* :191 */
  orc_csqrt(o->eta_re, o->eta_im, o->eta_abs, &o->sqrt_re, &o->sqrt_im);    /* :193 */
  o->inv_re = o->eta_re / o->eta_abs2;                                      /* :197 */
  o->inv_im = -o->eta_im / o->eta_abs2;                                     /* :198 */
  orc_csqrt(o->inv_re, o->inv_im, 1.f / o->eta_abs,
            &o->inv_sqrt_re, &o->inv_sqrt_im);                              /* :200 */
  o->r = 1.f - m->s;                                                        /* :204 */
}

/* ---- extensions ----------------------------------
 * NOT reference behaviour: the reference has these as TODOs (:587, :726-728,
 * :414).  Double-precision statement of the definitions in
 * hermespy-rt_b200/csrc/hrt_ext.cuh, which the opt-in HRT_FLAG_EXT_* modes of
 * the product are tested against.  Active only through oracle_compute_paths_ext. */
#define ORC_EXT_LOBES   1u
#define ORC_EXT_REFRACT 2u
typedef struct {
  uint32_t path; uint16_t tx, bounce;
  float o[3], d[3];
  float t_te_re, t_te_im, t_tm_re, t_tm_im;
} OrcRefractRecord;
static unsigned g_ext = 0;
static OrcRefractRecord *g_refr = NULL;
static size_t g_refr_cap = 0, g_refr_n = 0;

static void ext_csqrt(double re, double im, double *o_re, double *o_im)
{
  const double mag = sqrt(re * re + im * im);
  *o_re = sqrt(fmax(0.5 * (mag + re), 0.0));
  const double i = sqrt(fmax(0.5 * (mag - re), 0.0));
  *o_im = im < 0.0 ? -i : i;
}

/* ITU-R P.2040-3 eqs. (31c), (31d), eta = eta' - j eta'' (eq. 9b) */
void oracle_ext_refr_coefs(uint32_t material, float f_ghz, double theta1, double out[4])
{
  OrcMat m; orc_material(material, f_ghz, &m);
  const double er = m.eta_re, ei = m.eta_im;
  const double c1 = cos(theta1), s1 = sin(theta1);
  double qr, qi, sr, si;
  ext_csqrt(er - s1 * s1, -ei, &qr, &qi);
  ext_csqrt(er, -ei, &sr, &si);
  { const double dr = c1 + qr, di = qi, den = dr * dr + di * di;
    out[0] = 2.0 * c1 * dr / den; out[1] = -2.0 * c1 * di / den; }
  { const double nr = 2.0 * c1 * sr, ni = 2.0 * c1 * si, dr = er * c1 + qr, di = -ei * c1 + qi, den = dr * dr + di * di;
    out[2] = (nr * dr + ni * di) / den; out[3] = (ni * dr - nr * di) / den; }
}

int oracle_ext_refract_dir(uint32_t material, float f_ghz, const double d[3], const double n_in[3], double out[3])
{
  OrcMat m; orc_material(material, f_ghz, &m);
  double sr, si; ext_csqrt(m.eta_re, -(double)m.eta_im, &sr, &si);
  const double n = fmax(sr, 1e-6);
  double nn[3] = { n_in[0], n_in[1], n_in[2] };
  double c1 = -(d[0] * nn[0] + d[1] * nn[1] + d[2] * nn[2]);
  if (c1 < 0) { for (int k = 0; k < 3; ++k) nn[k] = -nn[k]; c1 = -c1; }
  const double k = 1.0 / n, s2sq = k * k * fmax(1.0 - c1 * c1, 0.0);
  if (s2sq > 1.0) return 0;
  const double c2 = sqrt(1.0 - s2sq), f = k * c1 - c2;
  double t[3], l = 0;
  for (int q = 0; q < 3; ++q) { t[q] = k * d[q] + f * nn[q]; l += t[q] * t[q]; }
  l = sqrt(l);
  for (int q = 0; q < 3; ++q) out[q] = t[q] / l;
  return 1;
}

static double ext_binom(int n, int k) { double r = 1; for (int i = 1; i <= k; ++i) r = r * (n - k + i) / i; return floor(r + 0.5); }

double oracle_ext_lobe_norm(int alpha, double cos_i, double sin_i)
{
  double sum = 0;
  for (int k = 0; k <= alpha; ++k) {
    double ik = 2.0 * 3.14159265358979323846 / (k + 1);
    if (k & 1) {
      double ser = 0;
      for (int w = 0; w <= (k - 1) / 2; ++w) ser += ext_binom(2 * w, w) * pow(sin_i / 2.0, 2.0 * w);
      ik *= cos_i * ser;
    }
    sum += ext_binom(alpha, k) * ik;
  }
  return sum / pow(2.0, alpha);
}

/* pi * f_s of the three-lobe model (Degli-Esposti 2007); ki towards the surface, ks away, n unit normal */
double oracle_ext_pattern_pi(double s1, double s2, double s3, int a1, int a3,
                             const double ki[3], const double ks[3], const double n_in[3])
{
  const double pi = 3.14159265358979323846;
  double n[3] = { n_in[0], n_in[1], n_in[2] };
  double cos_i = -(ki[0] * n[0] + ki[1] * n[1] + ki[2] * n[2]);
  if (cos_i < 0) { for (int k = 0; k < 3; ++k) n[k] = -n[k]; cos_i = -cos_i; }
  if (cos_i > 1) cos_i = 1;
  const double sin_i = sqrt(fmax(1.0 - cos_i * cos_i, 0.0));
  const double cos_s = ks[0] * n[0] + ks[1] * n[1] + ks[2] * n[2];
  if (!(cos_s > 0)) return 0;
  double kr[3], dir = 0, back = 0;
  for (int k = 0; k < 3; ++k) { kr[k] = ki[k] + 2.0 * cos_i * n[k]; dir += kr[k] * ks[k]; back += ki[k] * ks[k]; }
  dir = 0.5 * (1.0 + dir); back = 0.5 * (1.0 - back);
  double f = s2 * cos_s;
  if (s1 != 0) f += s1 * pi * pow(fmax(dir, 0.0), a1) / oracle_ext_lobe_norm(a1, cos_i, sin_i);
  if (s3 != 0) f += s3 * pi * pow(fmax(back, 0.0), a3) / oracle_ext_lobe_norm(a3, cos_i, sin_i);
  return f;
}

/* scattering coefficients of EXT_LOBES: s sqrt(pi f_s) times the reference's polarisation mix
 * (1, p, g, g p) / sqrt((1 + g^2)(1 + p^2)) of scat_coefs (:382-396) */
static void ext_scat(uint32_t material, const Vec3 *k_inc, const Vec3 *k_s, const Vec3 *nrm, float theta_i, float sc[4])
{
  const Material *m = &g_materials[material];
  const double rough = 1.0 / (1.0 + m->s1_alpha), ci = cos((double)theta_i), si = sin((double)theta_i);
  const double g = rough * ci + (1.0 - rough), p = sin(m->s1_alpha * si * 0.1);
  const double ki[3] = { k_inc->x, k_inc->y, k_inc->z }, ks[3] = { k_s->x, k_s->y, k_s->z }, nn[3] = { nrm->x, nrm->y, nrm->z };
  const double amp = m->s * sqrt(fmax(oracle_ext_pattern_pi(m->s1, m->s2, m->s3, m->s1_alpha, m->s3_alpha, ki, ks, nn), 0.0));
  const double k = amp / sqrt((1.0 + g * g) * (1.0 + p * p));
  sc[0] = (float)k; sc[1] = (float)(k * p); sc[2] = (float)(k * g); sc[3] = (float)(k * g * p);
}

/* ---- closest hit (reference moeller_trumbore :237-287) ------------------- */

/* Returns 1 when something was hit.  *theta is only written on a hit, which is
 * what produces the carry-over across receivers. */
int orc_closest_hit(const OrcTris *ft, const Ray *ray,
                    float *t_out, uint32_t *tri_out, float *theta)
{
  float best = 1e9;                                    /* :251 */
  int found = 0;
  for (uint32_t g = 0; g < ft->num_tris; ++g) {
    Vec3 ab = vec3_sub(&ft->b[g], &ft->a[g]);          /* :259 */
    Vec3 ac = vec3_sub(&ft->c[g], &ft->a[g]);          /* :260 */
    Vec3 pv = vec3_cross(&ray->d, &ac);                /* :261 */
    float det = vec3_dot(&ab, &pv);                    /* :262 */
    if (det > -ORC_EPS && det < ORC_EPS) continue;     /* :263 */
    Vec3 sv = vec3_sub(&ray->o, &ft->a[g]);            /* :264 */
    float u = vec3_dot(&sv, &pv) / det;                /* :265 */
    if (((double)u < 0. && fabs((double)u) > ORC_EPS) ||
        ((double)u > 1. && fabs((double)u - 1.) > ORC_EPS)) continue;   /* :266-268 */
    Vec3 qv = vec3_cross(&sv, &ab);                    /* :269 */
    float v = vec3_dot(&ray->d, &qv) / det;            /* :270 */
    float uv = u + v;
    if (((double)v < 0. && fabs((double)v) > ORC_EPS) ||
        ((double)uv > 1. && fabs((double)uv - 1.) > ORC_EPS)) continue; /* :271-273 */
    float tt = vec3_dot(&ac, &qv) / det;               /* :274 */
    if (tt > ORC_EPS && tt < best) {                   /* :275 strict <: first wins */
      best = tt;
      found = 1;
      *t_out = tt;
      *tri_out = g;
      float th = (float)acos((double)vec3_dot(&ft->n[g], &ray->d));     /* :281 */
      if ((double)th > (double)ORC_PI / 2.) th = ORC_PI - th;           /* :282-283 */
      *theta = th;
    }
  }
  return found;
}

/* ---- reflection coefficients (reference refl_coefs :300-344) ------------- */

static void orc_cdiv(float ar, float ai, float br, float bi, float *cr, float *ci)
{
  float den = br * br + bi * bi;               /* :161 */
  *cr = (ar * br + ai * bi) / den;             /* :162 */
  *ci = (ai * br - ar * bi) / den;             /* :163 */
}

void orc_refl(const OrcMat *m, float th, float out[4])
{
  float s1 = sinf(th);                                             /* :310 */
  if (m->eta_abs_inv_sqrt * s1 > 1.f - ORC_EPS) {                  /* :311 */
    out[0] = out[2] = 1.f; out[1] = out[3] = 0.f; return;
  }
  float s1sq = s1 * s1;                                            /* :318 */
  float c2r = sqrtf(1.f + m->inv_re / m->eta_abs2 * s1sq);         /* :319 */
  float c2i = sqrtf(1.f - m->inv_im / m->eta_abs2 * s1sq);         /* :320 */
  float pr = m->sqrt_re * c2r - m->sqrt_im * c2i;                  /* :323 */
  float pi = m->sqrt_re * c2i + m->sqrt_im * c2r;                  /* :324 */
  float c1 = cosf(th);                                             /* :325 */
  orc_cdiv(c1 - pr, -pi, c1 + pr, pi, &out[0], &out[1]);           /* :326 */
  float qr = m->sqrt_re * c1, qi = m->sqrt_im * c1;                /* :331-332 */
  orc_cdiv(qr - c2r, qi - c2i, qr + c2r, qi + c2i, &out[2], &out[3]); /* :333 */
  for (int k = 0; k < 4; ++k) out[k] *= m->r;                      /* :340-343 */
}

/* ---- scattering coefficients (reference scat_coefs :359-415) ------------- */

void orc_scat(uint32_t material, float th_s, float th_i, float out[4])
{
  const Material *m = &g_materials[material];
  float cs = cosf(th_s), ci = cosf(th_i), si = sinf(th_i);          /* :372-374 */
  float lobe = m->s * expf(-m->s1_alpha * fabsf(th_s - th_i));      /* :378-379 */
  float rough = 1.0f / (1.0f + m->s1_alpha);                        /* :382 */
  float spec = rough * cs;                                          /* :383 */
  float diff = (1.0f - rough) * cs;                                 /* :384 */
  float te_r = lobe * (spec + diff);                                /* :388 */
  float tm_r = lobe * (spec * ci + diff);                           /* :390 */
  float ph = m->s1_alpha * si * 0.1f;                               /* :394 */
  float te_i = te_r * sinf(ph);                                     /* :395 */
  float tm_i = tm_r * sinf(ph);                                     /* :396 */
  float nrm = sqrtf(te_r * te_r + te_i * te_i + tm_r * tm_r + tm_i * tm_i); /* :399 */
  if (nrm > 1e-6f) { te_r /= nrm; te_i /= nrm; tm_r /= nrm; tm_i /= nrm; } /* :401 */
  out[0] = te_r; out[1] = te_i; out[2] = tm_r; out[3] = tm_i;
}

/* ---- Fibonacci launch direction (reference :443-451) --------------------- */

Vec3 orc_launch_dir(size_t path, size_t num_paths)
{
  float k = (float)path + .5f;                                       /* :444 */
  float phi = (float)acos((double)(1.f - 2.f * k / (float)num_paths)); /* :445 */
  float th = ORC_PI * (1.f + sqrtf(5.f)) * k;                        /* :446 */
  Vec3 d;
  d.x = (float)(cos((double)th) * sin((double)phi));                 /* :448 */
  d.y = (float)(sin((double)th) * sin((double)phi));                 /* :449 */
  d.z = (float)cos((double)phi);                                     /* :450 */
  return d;
}

/* ---- the whole function (reference compute_paths :419-757) --------------- */

/* Same arguments and output semantics as the reference, including its quirks
 * (SURVEY appendix A).  Unlike the reference the scene is not mutated.  Only
 * output words the reference writes are written here.  Returns 0, or -1 when
 * out of memory. */
int oracle_compute_paths(
    const Scene *scene,
    const Vec3 *rx_pos, const Vec3 *tx_pos, const Vec3 *rx_vel, const Vec3 *tx_vel,
    float f_ghz, size_t R, size_t T, size_t P, size_t B,
    ChannelInfo *los, RaysInfo *rlos, ChannelInfo *scat, RaysInfo *rscat,
    OrcTrace *trace)
{
  OrcTris ft;
  if (orc_flatten(scene, &ft)) return -1;

  OrcMat mats[NUM_G_MATERIALS];
  memset(mats, 0, sizeof mats);
  for (uint32_t m = 0; m < scene->num_meshes; ++m)       /* :176-180 */
    orc_material(scene->meshes[m].material_index, f_ghz,
                 &mats[scene->meshes[m].material_index]);

  const size_t TP = T * P;
  Ray   *rays = malloc(TP * sizeof(Ray));
  float *g_te_r = malloc(TP * sizeof(float)), *g_te_i = calloc(TP, sizeof(float));
  float *g_tm_r = malloc(TP * sizeof(float)), *g_tm_i = calloc(TP, sizeof(float));
  float *delay = calloc(TP, sizeof(float));
  uint8_t *alive = malloc(TP / 8 + 1);
  if (!rays || !g_te_r || !g_te_i || !g_tm_r || !g_tm_i || !delay || !alive) return -1;

  for (size_t p = 0; p < P; ++p) {                       /* :443-456 */
    Vec3 d = orc_launch_dir(p, P);
    for (size_t t = 0; t < T; ++t) { rays[t * P + p].o = tx_pos[t]; rays[t * P + p].d = d; }
  }
  for (size_t i = 0; i < TP; ++i) g_te_r[i] = g_tm_r[i] = 1.f;   /* :464 */
  for (size_t i = 0; i < TP / 8 + 1; ++i) rscat->rays_active[i] = alive[i] = 0xff; /* :470 */

  float f_hz = f_ghz * 1e9;                              /* :483 (double product) */
  float fsl_k = 4.f * ORC_PI * f_hz / ORC_C0;            /* :484 */
  float dop_k = f_hz / ORC_C0;                           /* :488 */

  /* Doppler base value; the index algebra is the reference's (:494-508),
   * correct for T == 1 only (appendix A-8). */
  for (size_t t = 0; t < T; ++t)
    for (size_t p = 0; p < P; ++p) {
      float v = vec3_dot(&tx_vel[t], &rays[t * P + p].d);
      v *= dop_k;
      scat->freq_shift[t * P * B + p] = v;
    }
  for (size_t b = 1; b < B; ++b)
    memcpy(scat->freq_shift + TP * b, scat->freq_shift, TP * sizeof(float));
  for (size_t r = 1; r < R; ++r)
    memcpy(scat->freq_shift + TP * B * r, scat->freq_shift, TP * B * sizeof(float));

  /* ---- line of sight, reference :514-577 ---- */
  for (size_t i = 0; i < R * T; ++i) los->a_te_im[i] = los->a_tm_im[i] = 0.f;
  float theta = 0.f;  /* reference leaves it uninitialised; never read before set */
  for (size_t r = 0, k = 0; r < R; ++r)
    for (size_t t = 0; t < T; ++t, ++k) {
      Ray *lr = &rlos->rays[k];
      lr->o = tx_pos[t];
      lr->d = vec3_sub(&rx_pos[r], &lr->o);                        /* :528 */
      if (vec3_dot(&lr->d, &lr->d) < ORC_EPS) {                    /* :531-544 */
        los->directions_rx[k] = (Vec3){1.f, 0.f, 0.f};
        los->directions_tx[k] = (Vec3){-1.f, 0.f, 0.f};
        los->a_te_re[k] = los->a_tm_re[k] = 1.f;
        los->tau[k] = 0.f;
        los->freq_shift[k] = 0.f;
        rlos->rays_active[k / 8] |= 1 << (k % 8);
        continue;
      }
      float tt = -1.f; uint32_t tri = ORC_NONE;
      int hit = orc_closest_hit(&ft, lr, &tt, &tri, &theta);       /* :547 */
      if (hit && tt <= 1.f) {                                      /* :548-554 */
        los->a_te_re[k] = los->a_tm_re[k] = los->tau[k] = 0.f;
        rlos->rays_active[k / 8] &= ~(1 << (k % 8));
        continue;
      }
      float len = sqrtf(vec3_dot(&lr->d, &lr->d));                 /* :558 */
      Vec3 u = { lr->d.x / len, lr->d.y / len, lr->d.z / len };    /* :560 */
      los->directions_tx[k] = u;
      los->directions_rx[k] = (Vec3){ -u.x, -u.y, -u.z };
      float fsl = fsl_k * len;                                     /* :564 */
      los->a_te_re[k] = los->a_tm_re[k] = (fsl > 1.f) ? 1.f / fsl : 1.f;
      los->tau[k] = len / ORC_C0;                                  /* :571 */
      float fs = vec3_dot(&tx_vel[0], &u) - vec3_dot(&rx_vel[0], &u);  /* :573, index 0 */
      fs *= f_hz / ORC_C0;                                         /* :574 */
      los->freq_shift[k] = fs;
      rlos->rays_active[k / 8] |= 1 << (k % 8);
    }

  /* ---- bounces + per-receiver scatter, reference :589-745 ---- */
  memcpy(rscat->rays, rays, TP * sizeof(Ray));                     /* :589 */
  if (trace && trace->slot_state) memset(trace->slot_state, 0, R * T * B * P);

  for (size_t b = 0; b < B; ++b) {
    for (size_t t = 0; t < T; ++t) {
      for (size_t p = 0; p < P; ++p) {
        const size_t i = t * P + p;             /* ray index == bit index */
        const size_t tr = (t * B + b) * P + p;  /* trace index */
        if (!(alive[i >> 3] & (1u << (i & 7)))) {                  /* :604 */
          if (trace && trace->hit_tri) trace->hit_tri[tr] = ORC_IDLE;
          continue;
        }
        float tt = -1.f; uint32_t tri = ORC_NONE;
        int hit = orc_closest_hit(&ft, &rays[i], &tt, &tri, &theta);   /* :615 */
        if (trace && trace->hit_tri) {
          trace->hit_tri[tr] = hit ? tri : ORC_NONE;
          if (trace->hit_t) trace->hit_t[tr] = hit ? tt : -1.f;
          if (trace->hit_theta) trace->hit_theta[tr] = hit ? theta : 0.f;
        }
        if (!hit) { alive[i >> 3] &= ~(1u << (i & 7)); continue; } /* :616-620 */

        const uint32_t mesh = ft.mesh_of[tri];
        const uint32_t mat = scene->meshes[mesh].material_index;  /* :622 */
        const Vec3 d_in = rays[i].d;                               /* incident direction (extensions only) */
        if ((g_ext & ORC_EXT_REFRACT) && g_refr) {
          /* the refraction ray the reference's TODO (:587, :726-728) would spawn here: state gains
           * times (31c)/(31d), free-space loss of this segment as for the reflected ray (:627-634) */
          double T4[4], dd[3] = { d_in.x, d_in.y, d_in.z }, nn[3] = { ft.n[tri].x, ft.n[tri].y, ft.n[tri].z }, dt[3];
          oracle_ext_refr_coefs(mat, f_ghz, (double)theta, T4);
          if (oracle_ext_refract_dir(mat, f_ghz, dd, nn, dt)) {
            if (g_refr_n < g_refr_cap) {
              OrcRefractRecord *q = &g_refr[g_refr_n];
              double l = (double)(fsl_k * tt); l *= l; if (!(l > 1.0)) l = 1.0;
              q->path = (uint32_t)p; q->tx = (uint16_t)t; q->bounce = (uint16_t)b;
              for (int c = 0; c < 3; ++c) {
                const double hp = (&rays[i].o.x)[c] + (double)(&d_in.x)[c] * tt;
                q->o[c] = (float)(hp + 1e-4 * dt[c]); q->d[c] = (float)dt[c];
              }
              q->t_te_re = (float)((g_te_r[i] * T4[0] - g_te_i[i] * T4[1]) / l); q->t_te_im = (float)((g_te_r[i] * T4[1] + g_te_i[i] * T4[0]) / l);
              q->t_tm_re = (float)((g_tm_r[i] * T4[2] - g_tm_i[i] * T4[3]) / l); q->t_tm_im = (float)((g_tm_r[i] * T4[3] + g_tm_i[i] * T4[2]) / l);
            }
            ++g_refr_n;
          }
        }
        float rc[4];
        orc_refl(&mats[mat], theta, rc);                           /* :623 */
        float fsl = fsl_k * tt;  fsl *= fsl;                       /* :627-628 */
        if (fsl > 1.f) for (int q = 0; q < 4; ++q) rc[q] /= fsl;   /* :629-634 */
        float n_te_r = g_te_r[i] * rc[0] - g_te_i[i] * rc[1];      /* :636-639 */
        float n_te_i = g_te_r[i] * rc[1] + g_te_i[i] * rc[0];
        float n_tm_r = g_tm_r[i] * rc[2] - g_tm_i[i] * rc[3];
        float n_tm_i = g_tm_r[i] * rc[3] + g_tm_i[i] * rc[2];
        g_te_r[i] = n_te_r; g_te_i[i] = n_te_i; g_tm_r[i] = n_tm_r; g_tm_i[i] = n_tm_i;
        delay[i] += tt / ORC_C0;                                   /* :645 */

        Ray *ry = &rays[i];
        Vec3 step = vec3_scale(&ry->d, tt);                        /* :650 */
        ry->o = vec3_add(&step, &ry->o);                           /* :651 */
        Vec3 nrm = ft.n[tri];                                      /* :653 */
        float dn = vec3_dot(&ry->d, &nrm);                         /* :654 */
        step = vec3_scale(&nrm, 2.f * dn);                         /* :655 */
        ry->d = vec3_sub(&ry->d, &step);                           /* :656 */
        step = vec3_scale(&ry->d, 1e-4f);                          /* :658 */
        ry->o = vec3_add(&ry->o, &step);                           /* :659 */

        const Vec3 *mv = &scene->meshes[mesh].velocity;            /* :662 */
        Vec3 zero = vec3_sub(&ry->d, &ry->d);                      /* :663, same object */
        scat->freq_shift[i] += vec3_dot(&zero, mv) * dop_k;        /* :664 */

        Ray sh; sh.o = ry->o;                                      /* :671 */
        for (size_t r = 0; r < R; ++r) {
          const size_t s = ((r * T + t) * B + b) * P + p;          /* :674 */
          sh.d = vec3_sub(&rx_pos[r], &sh.o);                      /* :676 */
          float dist = sqrtf(vec3_dot(&sh.d, &sh.d));              /* :677 */
          sh.d = vec3_normalize(&sh.d);                            /* :678 */
          float st = -1.f; uint32_t stri = ORC_NONE;
          int shit = orc_closest_hit(&ft, &sh, &st, &stri, &theta);  /* :682 */
          if (trace && trace->shadow_tri) trace->shadow_tri[s] = shit ? stri : ORC_NONE;
          if (trace && trace->theta_used) trace->theta_used[s] = theta;
          if (shit && st <= 1.f) {                                 /* :683-691 */
            scat->a_te_re[s] = scat->a_te_im[s] = scat->a_tm_re[s] = scat->a_tm_im[s]
                             = scat->tau[s] = 0.f;
            if (trace && trace->slot_state) trace->slot_state[s] = 2;
            continue;
          }
          float th_s = acosf(vec3_dot(&sh.d, &nrm));               /* :694 */
          float sc[4];
          if (g_ext & ORC_EXT_LOBES) ext_scat(mat, &d_in, &sh.d, &nrm, theta, sc);
          else orc_scat(mat, th_s, theta, sc);                     /* :696 */
          scat->a_te_re[s] = g_te_r[i] * sc[0] - g_te_i[i] * sc[1];   /* :698-705 */
          scat->a_te_im[s] = g_te_r[i] * sc[1] + g_te_i[i] * sc[0];
          scat->a_tm_re[s] = g_tm_r[i] * sc[2] - g_tm_i[i] * sc[3];
          scat->a_tm_im[s] = g_tm_r[i] * sc[3] + g_tm_i[i] * sc[2];
          scat->directions_rx[s] = (Vec3){ -sh.d.x, -sh.d.y, -sh.d.z };  /* :707 */
          scat->tau[s] = delay[i] + dist / ORC_C0;                 /* :709 */
          float l2 = fsl_k * dist;  l2 *= l2;                      /* :711-712 */
          if (l2 > 1.f) {                                          /* :713-718 */
            scat->a_te_re[s] /= l2; scat->a_te_im[s] /= l2;
            scat->a_tm_re[s] /= l2; scat->a_tm_im[s] /= l2;
          }
          Vec3 dd = vec3_sub(&sh.d, &ry->d);                       /* :720 */
          float fs = vec3_dot(&dd, mv) * dop_k;                    /* :721 */
          scat->freq_shift[s] -= fs;                               /* :722 */
          if (trace && trace->slot_state) trace->slot_state[s] = 1;
        }
      }
      /* RaysInfo rows, reference :732-743 (stride B, TX-0 mask: appendix A-9) */
      size_t row = (t * B + (b + 1));
      memcpy(rscat->rays + row * P, rays + t * P, P * sizeof(Ray));
      memcpy(rscat->rays_active + row * (P / 8 + 1), alive, P / 8 + 1);
    }
  }

  free(rays); free(g_te_r); free(g_te_i); free(g_tm_r); free(g_tm_i);
  free(delay); free(alive);
  orc_free_tris(&ft);
  return 0;
}

/* compute_paths with the opt-in extensions (ext: ORC_EXT_LOBES | ORC_EXT_REFRACT); refraction rays are
 * appended to refr[0..cap), *n_refr = number spawned */
int oracle_compute_paths_ext(
    const Scene *scene,
    const Vec3 *rx_pos, const Vec3 *tx_pos, const Vec3 *rx_vel, const Vec3 *tx_vel,
    float f_ghz, size_t R, size_t T, size_t P, size_t B,
    ChannelInfo *los, RaysInfo *rlos, ChannelInfo *scat, RaysInfo *rscat,
    OrcTrace *trace, unsigned ext, OrcRefractRecord *refr, size_t cap, size_t *n_refr)
{
  g_ext = ext; g_refr = refr; g_refr_cap = cap; g_refr_n = 0;
  const int rc = oracle_compute_paths(scene, rx_pos, tx_pos, rx_vel, tx_vel, f_ghz, R, T, P, B, los, rlos, scat, rscat, trace);
  if (n_refr) *n_refr = g_refr_n;
  g_ext = 0; g_refr = NULL; g_refr_cap = 0;
  return rc;
}

/* Stand-alone closest hit over a scene for unit tests: `n` rays in, per ray
 * the global triangle id (ORC_NONE on a miss), distance and folded angle. */
int oracle_closest_hits(const Scene *scene, const Ray *rays, size_t n,
                        uint32_t *tri, float *t, float *theta)
{
  OrcTris ft;
  if (orc_flatten(scene, &ft)) return -1;
  for (size_t i = 0; i < n; ++i) {
    float tt = -1.f, th = 0.f; uint32_t g = ORC_NONE;
    int hit = orc_closest_hit(&ft, &rays[i], &tt, &g, &th);
    tri[i] = hit ? g : ORC_NONE; t[i] = hit ? tt : -1.f; theta[i] = hit ? th : 0.f;
  }
  orc_free_tris(&ft);
  return 0;
}

/* Launch directions for `n` consecutive path indices starting at `first`. */
void oracle_launch_dirs(size_t first, size_t n, size_t num_paths, Vec3 *out)
{
  for (size_t i = 0; i < n; ++i) out[i] = orc_launch_dir(first + i, num_paths);
}

/* Unit normals of every triangle in (mesh, face) order -- what the reference's
 * precompute_normals (:208-224) stores in Mesh.ns; `out` holds sum(num_triangles). */
int oracle_normals(const Scene *scene, Vec3 *out)
{
  OrcTris ft;
  if (orc_flatten(scene, &ft)) return -1;
  memcpy(out, ft.n, (size_t)ft.num_tris * sizeof(Vec3));
  orc_free_tris(&ft);
  return 0;
}
