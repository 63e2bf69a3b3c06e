/* CPU oracle (C restatement of the JaxDEM step path).  TEST INFRASTRUCTURE ONLY:
 * see oracle/__init__.py and oracle_impl.h. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REAL float
#define INT int32_t
#define UINT uint32_t
#define INT_MAX_V INT32_MAX
#define INT_MIN_V INT32_MIN
#define SFX _f32
#define FLOOR floorf
#define CEIL ceilf
#define FMOD fmodf
#define RINT rintf
#define SQRT sqrtf
#define POW powf
#define LOG logf
#include "oracle_impl.h"
#undef REAL
#undef INT
#undef UINT
#undef INT_MAX_V
#undef INT_MIN_V
#undef SFX
#undef FLOOR
#undef CEIL
#undef FMOD
#undef RINT
#undef SQRT
#undef POW
#undef LOG

#define REAL double
#define INT int64_t
#define UINT uint64_t
#define INT_MAX_V INT64_MAX
#define INT_MIN_V INT64_MIN
#define SFX _f64
#define FLOOR floor
#define CEIL ceil
#define FMOD fmod
#define RINT rint
#define SQRT sqrt
#define POW pow
#define LOG log
#include "oracle_impl.h"

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}
