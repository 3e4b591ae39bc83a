/* materials.c -- radio material table and name lookup (host C).
 *
 * Replaces reference src/materials.c:3-122.  The numbers are ITU-R P.2040-3
 * table 3 (a, b, c, d) plus the reference's scattering parameters; they are
 * data the path's results depend on, so they are identical by necessity.
 * Row layout: {name_sz, name, a, b, c, d, s, s1, s2, s3, s1_alpha, s3_alpha}.
 */
#include "../../include/hermespy_rt.h"

#define MAT(nm, a, b, c, d, s, s1, s2, s3, al1, al3) \
  { (uint32_t)(sizeof(nm) - 1), nm, a, b, c, d, s, s1, s2, s3, al1, al3 }

Material g_materials[NUM_G_MATERIALS] = {
  /*  0 */ MAT("air",               1.f,    0.f,   0.f,       0.001f,  0.1f, 0.5f,  0.3f,  0.2f,  2, 2),
  /*  1 */ MAT("concrete",          5.24f,  0.f,   0.0462f,   0.7822f, 0.5f, 0.33f, 0.34f, 0.33f, 4, 4),
  /*  2 */ MAT("brick",             3.91f,  0.f,   0.0238f,   0.16f,   0.4f, 0.4f,  0.3f,  0.3f,  3, 3),
  /*  3 */ MAT("plasterboard",      2.73f,  0.f,   0.0085f,   0.9395f, 0.3f, 0.4f,  0.4f,  0.2f,  3, 3),
  /*  4 */ MAT("wood",              1.99f,  0.f,   0.0047f,   1.0718f, 0.2f, 0.5f,  0.3f,  0.2f,  2, 2),
  /*  5 */ MAT("glass",             6.31f,  0.f,   0.0036f,   1.3394f, 0.3f, 0.4f,  0.4f,  0.2f,  3, 3),
  /*  6 */ MAT("glass",             5.79f,  0.f,   0.0004f,   1.658f,  0.3f, 0.4f,  0.4f,  0.2f,  3, 3),
  /*  7 */ MAT("ceiling board",     1.48f,  0.f,   0.0011f,   1.0750f, 0.2f, 0.5f,  0.3f,  0.2f,  2, 2),
  /*  8 */ MAT("ceiling board",     1.52f,  0.f,   0.0029f,   1.029f,  0.2f, 0.5f,  0.3f,  0.2f,  2, 2),
  /*  9 */ MAT("chipboard",         2.58f,  0.f,   0.0217f,   0.7800f, 0.4f, 0.4f,  0.3f,  0.3f,  3, 3),
  /* 10 */ MAT("plywood",           2.71f,  0.f,   0.33f,     0.f,     0.3f, 0.5f,  0.3f,  0.2f,  3, 3),
  /* 11 */ MAT("marble",            7.074f, 0.f,   0.0055f,   0.9262f, 0.3f, 0.4f,  0.4f,  0.2f,  3, 3),
  /* 12 */ MAT("floorboard",        3.66f,  0.f,   0.0044f,   1.3515f, 0.3f, 0.4f,  0.4f,  0.2f,  3, 3),
  /* 13 */ MAT("metal",             1.f,    0.f,   10000000.f,0.f,     0.f,  0.f,   1.f,   0.f,   1, 1),
  /* 14 */ MAT("very dry ground",   3.f,    0.f,   0.00015f,  2.52f,   0.4f, 0.3f,  0.4f,  0.3f,  4, 4),
  /* 15 */ MAT("medium dry ground", 15.f,  -0.1f,  0.035f,    1.63f,   0.5f, 0.33f, 0.34f, 0.33f, 4, 4),
  /* 16 */ MAT("wet ground",        30.f,  -0.4f,  0.15f,     1.30f,   0.5f, 0.33f, 0.34f, 0.33f, 4, 4),
};

/* Lookup keys are the reference's (src/materials.c:98-116): lower case with
 * underscores, glass/ceiling board numbered.  Unknown names map to air, as in
 * the reference (src/materials.c:121). */
static const char *const k_lookup_names[NUM_G_MATERIALS] = {
  "air", "concrete", "brick", "plasterboard", "wood", "glass1", "glass2",
  "ceiling_board1", "ceiling_board2", "chipboard", "plywood", "marble",
  "floorboard", "metal", "very_dry_ground", "medium_dry_ground", "wet_ground",
};

MaterialIndex get_material_index(const char *name)
{
  for (int i = 0; i < NUM_G_MATERIALS; ++i)
    if (!strcmp(name, k_lookup_names[i])) return (MaterialIndex)i;
  return MATERIAL_AIR;
}
