/* scene.c -- .hrt scene files (host C).
 *
 * Replaces reference src/scene.c:7-83.  File layout (native
 * little-endian, no padding):
 *   "HRT" | u32 num_meshes | per mesh: u32 nv | nv*3 f32 | u32 nt | nt*3 u32 |
 *   u32 material_index | 3 f32 velocity
 * Failure convention is the reference's: message on stderr, exit(8).  On top of
 * the reference's checks (1..1000 meshes, short reads) the loader rejects
 * vertex indices >= nv and material indices >= NUM_G_MATERIALS, which the
 * reference would read out of bounds.
 */
#include "../../include/hermespy_rt.h"

#include <errno.h>
#include <stdio.h>

#define HRT_MAX_MESHES 1000u   /* reference src/scene.c:54 */

static void die(const char *what, const char *path)
{
  if (errno) fprintf(stderr, "hermespy_rt: %s (%s): %s\n", what, path, strerror(errno));
  else       fprintf(stderr, "hermespy_rt: %s (%s)\n", what, path);
  exit(8);
}

static void put(const void *p, size_t sz, size_t n, FILE *f, const char *path)
{ if (fwrite(p, sz, n, f) != n) die("short write", path); }

static void get(void *p, size_t sz, size_t n, FILE *f, const char *path)
{ if (fread(p, sz, n, f) != n) { errno = 0; die("truncated scene file", path); } }

void scene_save(Scene *scene, const char *filepath)
{
  FILE *f = fopen(filepath, "wb");
  if (!f) die("cannot open scene file for writing", filepath);
  put("HRT", 1, 3, f, filepath);
  put(&scene->num_meshes, 4, 1, f, filepath);
  for (uint32_t m = 0; m < scene->num_meshes; ++m) {
    const Mesh *me = &scene->meshes[m];
    put(&me->num_vertices, 4, 1, f, filepath);
    put(me->vs, sizeof(Vec3), me->num_vertices, f, filepath);
    put(&me->num_triangles, 4, 1, f, filepath);
    put(me->is, 4, (size_t)3 * me->num_triangles, f, filepath);
    put(&me->material_index, 4, 1, f, filepath);
    put(&me->velocity, sizeof(Vec3), 1, f, filepath);
  }
  if (fclose(f)) die("cannot close scene file", filepath);
}

Scene scene_load(const char *filepath)
{
  Scene sc; sc.num_meshes = 0; sc.meshes = NULL;
  errno = 0;
  FILE *f = fopen(filepath, "rb");
  if (!f) die("cannot open scene file", filepath);

  char magic[3];
  get(magic, 1, 3, f, filepath);
  if (memcmp(magic, "HRT", 3)) { errno = 0; die("not an HRT file", filepath); }
  get(&sc.num_meshes, 4, 1, f, filepath);
  if (sc.num_meshes == 0 || sc.num_meshes > HRT_MAX_MESHES) {
    errno = 0; die("mesh count outside 1..1000", filepath);
  }
  sc.meshes = (Mesh *)calloc(sc.num_meshes, sizeof(Mesh));
  if (!sc.meshes) die("out of memory", filepath);

  for (uint32_t m = 0; m < sc.num_meshes; ++m) {
    Mesh *me = &sc.meshes[m];
    get(&me->num_vertices, 4, 1, f, filepath);
    me->vs = (Vec3 *)malloc((size_t)(me->num_vertices ? me->num_vertices : 1) * sizeof(Vec3));
    if (!me->vs) die("out of memory", filepath);
    get(me->vs, sizeof(Vec3), me->num_vertices, f, filepath);
    get(&me->num_triangles, 4, 1, f, filepath);
    size_t ni = (size_t)3 * me->num_triangles;
    me->is = (uint32_t *)malloc((ni ? ni : 1) * sizeof(uint32_t));
    if (!me->is) die("out of memory", filepath);
    get(me->is, 4, ni, f, filepath);
    get(&me->material_index, 4, 1, f, filepath);
    get(&me->velocity, sizeof(Vec3), 1, f, filepath);
    me->ns = NULL;   /* not stored in the file; compute_paths() fills it */
    for (size_t k = 0; k < ni; ++k)
      if (me->is[k] >= me->num_vertices) { errno = 0; die("vertex index out of range", filepath); }
    if (me->material_index >= NUM_G_MATERIALS) { errno = 0; die("material index out of range", filepath); }
  }
  fclose(f);
  return sc;
}
