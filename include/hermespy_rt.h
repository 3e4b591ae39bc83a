/* hermespy_rt.h -- public C API of the B200-native compute_paths() library.
 *
 * This single header declares every type and entry point of the reference's
 * drop-in boundary.  The struct layouts, field order, argument order and units
 * are the reference's ABI and must not change:
 *
 *   Vec3, vec3_* helpers ......... reference inc/vec3.h:6-43
 *   Ray .......................... reference inc/ray.h:6-9
 *   Mesh, Scene, Material ........ reference inc/scene.h:10-66
 *   free_mesh / free_scene ....... reference inc/scene.h:72-86
 *   scene_save / scene_load ...... reference inc/scene.h:95,105 (src/scene.c:7-83)
 *   g_materials, MaterialIndex,
 *   get_material_index ........... reference inc/materials.h:9-33 (src/materials.c)
 *   ChannelInfo, RaysInfo ........ reference inc/compute_paths.h:13-30
 *   compute_paths ................ reference inc/compute_paths.h:59-74
 *
 * The shim headers compute_paths.h / scene.h / materials.h / vec3.h / ray.h /
 * common.h next to this file only include it, so code written against the
 * reference's inc/ directory compiles unchanged with -I<repo>/include.
 *
 * Everything computes on the GPU (sm_100a).  There is no CPU fallback: without
 * a usable CUDA device compute_paths() prints a diagnostic and exit(8)s, which
 * is the reference's error convention (inc/common.h:20-25).
 */
#ifndef HERMESPY_RT_H
#define HERMESPY_RT_H

#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Direction markers used in the reference prototypes (inc/common.h:4-5). */
#ifndef IN
#define IN
#endif
#ifndef OUT
#define OUT
#endif

/* ------------------------------------------------------------------ vectors */

typedef struct { float x, y, z; } Vec3;          /* 12 bytes, no padding */
typedef struct { Vec3 o; Vec3 d; } Ray;          /* origin, direction    */

/* The helpers keep the reference's rounding order (inc/vec3.h:10-43): every
 * product and sum is a separate fp32 operation, dot = (x*x' + y*y') + z*z'. */
static inline Vec3 vec3_sub(const Vec3 *a, const Vec3 *b)
{ Vec3 r; r.x = a->x - b->x; r.y = a->y - b->y; r.z = a->z - b->z; return r; }
static inline Vec3 vec3_add(const Vec3 *a, const Vec3 *b)
{ Vec3 r; r.x = a->x + b->x; r.y = a->y + b->y; r.z = a->z + b->z; return r; }
static inline Vec3 vec3_cross(const Vec3 *a, const Vec3 *b)
{
  Vec3 r;
  r.x = a->y * b->z - a->z * b->y;
  r.y = a->z * b->x - a->x * b->z;
  r.z = a->x * b->y - a->y * b->x;
  return r;
}
static inline float vec3_dot(const Vec3 *a, const Vec3 *b)
{ return a->x * b->x + a->y * b->y + a->z * b->z; }
static inline Vec3 vec3_scale(const Vec3 *a, float s)
{ Vec3 r; r.x = a->x * s; r.y = a->y * s; r.z = a->z * s; return r; }
static inline Vec3 vec3_normalize(const Vec3 *a)
{
  float len = sqrtf(a->x * a->x + a->y * a->y + a->z * a->z);
  Vec3 r; r.x = a->x / len; r.y = a->y / len; r.z = a->z / len; return r;
}

/* -------------------------------------------------------------------- scene */

typedef struct {
  uint32_t  num_vertices;
  Vec3     *vs;              /* [num_vertices] */
  uint32_t  num_triangles;
  uint32_t *is;              /* [3*num_triangles], indices into vs */
  uint32_t  material_index;  /* index into g_materials */
  Vec3      velocity;        /* m/s, global frame */
  Vec3     *ns;              /* [num_triangles] unit normals; not in the file.
                                compute_paths() (re)allocates it like the
                                reference does (src/compute_paths.c:212). */
} Mesh;

typedef struct {
  uint32_t num_meshes;
  Mesh    *meshes;
} Scene;

typedef struct {
  uint32_t    name_sz;
  const char *name;          /* not NUL-terminated by contract */
  float a, b, c, d;          /* ITU-R P.2040-3 table 3 */
  float s;                   /* scattering coefficient, [0,1] */
  float s1, s2, s3;          /* lobe ratios (unused by the path) */
  uint8_t s1_alpha;          /* directive lobe width */
  uint8_t s3_alpha;          /* backward lobe width (unused by the path) */
} Material;

static inline void free_mesh(Mesh *mesh)
{ free(mesh->vs); free(mesh->is); free(mesh->ns); }
static inline void free_scene(Scene *scene)
{
  for (uint32_t m = 0; m < scene->num_meshes; ++m) free_mesh(&scene->meshes[m]);
  free(scene->meshes);
}

/* .hrt I/O.  Errors: perror + exit(8), as the reference. */
void  scene_save(IN Scene *scene, IN const char *filepath);
Scene scene_load(IN const char *filepath);

/* ---------------------------------------------------------------- materials */

#define NUM_G_MATERIALS 17
extern Material g_materials[NUM_G_MATERIALS];

typedef enum {
  MATERIAL_AIR = 0, MATERIAL_CONCRETE, MATERIAL_BRICK, MATERIAL_PLASTERBOARD,
  MATERIAL_WOOD, MATERIAL_GLASS1, MATERIAL_GLASS2, MATERIAL_CEILING_BOARD1,
  MATERIAL_CEILING_BOARD2, MATERIAL_CHIPBOARD, MATERIAL_PLYWOOD, MATERIAL_MARBLE,
  MATERIAL_FLOORBOARD, MATERIAL_METAL, MATERIAL_VERY_DRY_GROUND,
  MATERIAL_MEDIUM_DRY_GROUND, MATERIAL_WET_GROUND
} MaterialIndex;

MaterialIndex get_material_index(const char *name);

/* ------------------------------------------------------------ compute_paths */

typedef struct {
  uint32_t num_rays;         /* per rx-tx pair: 1 (LoS) or B*P (scatter) */
  Vec3  *directions_rx;      /* (num_rx, num_tx, num_rays) */
  Vec3  *directions_tx;      /* LoS only; never written for scatter */
  float *a_te_re, *a_te_im;  /* (num_rx, num_tx, num_rays) */
  float *a_tm_re, *a_tm_im;
  float *tau;                /* s  */
  float *freq_shift;         /* Hz */
} ChannelInfo;

typedef struct {
  uint32_t num_bounces, num_rays;
  Ray     *rays;             /* (num_tx, num_bounces, num_paths) */
  uint8_t *rays_active;      /* bitmask (num_tx, num_bounces, num_paths/8+1) */
} RaysInfo;

/* All output arrays are allocated by the caller (sizes: INTEGRATION.md).
 * carrier_frequency_GHz is in GHz; positions in m; velocities in m/s. */
void compute_paths(
    IN Scene *scene,
    IN Vec3 *rx_pos, IN Vec3 *tx_pos, IN Vec3 *rx_vel, IN Vec3 *tx_vel,
    IN float carrier_frequency_GHz,
    IN size_t num_rx, IN size_t num_tx, IN size_t num_rays, IN size_t num_bounces,
    OUT ChannelInfo *chanInfo_los,  OUT RaysInfo *raysInfo_los,
    OUT ChannelInfo *chanInfo_scat, OUT RaysInfo *raysInfo_scat);

/* Extension (no reference counterpart; it is the reduction a caller of
 * compute_paths() performs next, and the reference's own TODO for large runs,
 * SURVEY section 8 row f2): the channel impulse response per (rx, tx) of the
 * SAME path set -- LoS + scatter paths of all bounces -- accumulated on the GPU
 * without materialising per-path records:
 *   cir[((rx * num_tx + tx) * num_bins + bin) * 4 + k],
 *   k = 0..3: sum of a_te_re, a_te_im, a_tm_re, a_tm_im over the paths with
 *   bin = floor((tau - tau0_s) / dt_s) in [0, num_bins).
 * `cir` is caller-allocated (num_rx * num_tx * num_bins * 4 floats) and
 * overwritten.  Returns the number of paths outside the delay window. */
size_t compute_cir(
    IN Scene *scene,
    IN Vec3 *rx_pos, IN Vec3 *tx_pos, IN Vec3 *rx_vel, IN Vec3 *tx_vel,
    IN float carrier_frequency_GHz,
    IN size_t num_rx, IN size_t num_tx, IN size_t num_rays, IN size_t num_bounces,
    IN float tau0_s, IN float dt_s, IN size_t num_bins, OUT float *cir);

#ifdef __cplusplus
}
#endif
#endif /* HERMESPY_RT_H */
