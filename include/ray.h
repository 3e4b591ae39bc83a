/* Shim: the reference splits its API over inc/ray.h; here every declaration
 * lives in hermespy_rt.h so that the ABI is stated once. */
#include "hermespy_rt.h"
