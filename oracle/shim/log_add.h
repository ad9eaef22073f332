/* TEST INFRASTRUCTURE ONLY (oracle).  Stand-in for Torch3's <log_add.h>.
 * LOG_ZERO = -FLT_MAX is an [ext] assumption: it is used only
 * as a dead-token sentinel, the product uses the same constant. */
#ifndef ORACLE_SHIM_LOG_ADD_H
#define ORACLE_SHIM_LOG_ADD_H
#include "general.h"
#define LOG_2_PI 1.83787706640934548355
#define LOG_ZERO (-INF)
#define LOG_ONE 0
namespace Torch {
inline real logAdd(real x, real y) {
    if (x < y) { real t = x; x = y; y = t; }
    real diff = y - x;
    if (diff < -18.42) return x;
    return x + log(1.0 + exp(diff));
}
}
#endif
