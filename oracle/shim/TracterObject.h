/* TEST INFRASTRUCTURE ONLY (oracle).  Stand-in for Tracter's <TracterObject.h>:
 * WFSTDecoderLite uses only mObjectName and GetEnv (WFSTDecoderLite.cpp:48,73,117). */
#ifndef ORACLE_SHIM_TRACTEROBJECT_H
#define ORACLE_SHIM_TRACTEROBJECT_H
#include <cstdlib>
namespace Tracter {
class Object {
public:
    virtual ~Object() throw() {}
protected:
    const char* mObjectName;
    int GetEnv(const char* name, int dflt) {
        const char* e = getenv(name);
        return e ? atoi(e) : dflt;
    }
};
}
#endif
