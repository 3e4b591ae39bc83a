/* Shim for the reference's inc/common.h:4-25 (IN/OUT markers and the
 * perror-free-exit convention used by its I/O code). */
#ifndef HRT_COMMON_SHIM_H
#define HRT_COMMON_SHIM_H
#include "hermespy_rt.h"
#include <stdio.h>

/* free every pointer argument (may be none) */
#define FREE_POINTERS(...)                                              \
  do {                                                                  \
    void *hrt_p_[] = { NULL, ##__VA_ARGS__ };                           \
    for (size_t hrt_i_ = 1; hrt_i_ < sizeof hrt_p_ / sizeof *hrt_p_; ++hrt_i_) \
      free(hrt_p_[hrt_i_]);                                             \
  } while (0)

/* report errno-style failure, release the listed pointers, leave with rc */
#define PERROR_CLEANUP_EXIT(msg, rc, ...)                               \
  do { perror(msg); FREE_POINTERS(__VA_ARGS__); exit(rc); } while (0)
#endif
