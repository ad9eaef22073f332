/* TEST INFRASTRUCTURE ONLY (oracle).  Stand-in for Torch3's <general.h>, which the
 * reference includes everywhere but does not vendor.
 * Only what the hot-path sources use: `real` (= float: the reference links Torch's
 * *_opt_float build, /root/reference/cmake/FindTorch3.cmake:32), error(), warning(),
 * message(), INF, REAL_EPSILON, min/max. */
#ifndef ORACLE_SHIM_GENERAL_H
#define ORACLE_SHIM_GENERAL_H
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <cstdarg>
#include <ctime>
#include <algorithm>
namespace Torch {
typedef float real;
inline void error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    fprintf(stderr, "ERROR: "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n");
    va_end(ap); fflush(stderr); exit(-1);
}
inline void warning(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    fprintf(stderr, "WARNING: "); vfprintf(stderr, fmt, ap); fprintf(stderr, "\n");
    va_end(ap);
}
inline void message(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vfprintf(stderr, fmt, ap); fprintf(stderr, "\n");
    va_end(ap);
}
}
#define INF FLT_MAX
#define REAL_EPSILON FLT_EPSILON
using Torch::real;
using std::min;
using std::max;
#endif
