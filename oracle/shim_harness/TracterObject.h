/* TEST INFRASTRUCTURE ONLY (oracle harness build).  Stand-in for Tracter's <TracterObject.h> with what FrontEnd.h and the
 * decoders use of it: mObjectName and the GetEnv overloads. */
#ifndef ORACLE_SHIM_HARNESS_TRACTEROBJECT_H
#define ORACLE_SHIM_HARNESS_TRACTEROBJECT_H
#include <cstdlib>
namespace Tracter {
typedef long IndexType;
typedef long long TimeType;            /* nanoseconds */
#define ONEe9 1000000000LL
#define ORACLE_FRAME_PERIOD_NS 10000000LL   /* 10 ms frames: HTK's default sample period (100 000 x 100 ns) */
class Object {
public:
    virtual ~Object() throw() {}
protected:
    const char* mObjectName;
    int GetEnv(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
    const char* GetEnv(const char* name, const char* dflt) { const char* e = getenv(name); return e ? e : dflt; }
};
/* a stage of Tracter's processing graph: only its identity matters here */
template <class T> class Component { public: virtual ~Component() {} };
class ISource { public: virtual ~ISource() {} virtual void Open(const char* name, TimeType begin = -1, TimeType end = -1) = 0; };
}
#endif
