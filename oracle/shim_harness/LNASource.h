/* TEST INFRASTRUCTURE ONLY.  LNA (posterior) files are outside the GMM decode path. */
#ifndef ORACLE_SHIM_HARNESS_LNASOURCE_H
#define ORACLE_SHIM_HARNESS_LNASOURCE_H
#include "HTKSource.h"
namespace Tracter {
class LNASource : public FrameStore, public ISource {
public:
    void Open(const char*, TimeType = -1, TimeType = -1) { fprintf(stderr, "LNASource: not available in the oracle build\n"); exit(-1); }
};
}
#endif
