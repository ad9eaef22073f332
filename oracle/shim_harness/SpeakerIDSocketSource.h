/* TEST INFRASTRUCTURE ONLY.  The speaker-ID chain is only built when the environment names a host (FrontEnd.h:71-79). */
#ifndef ORACLE_SHIM_HARNESS_SPEAKERID_H
#define ORACLE_SHIM_HARNESS_SPEAKERID_H
#include "HTKSource.h"
namespace Tracter {
class SpeakerIDSocketSource : public FrameStore {
public:
    void Open(const char*) {}
};
}
#endif
