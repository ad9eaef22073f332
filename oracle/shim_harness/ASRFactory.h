/* TEST INFRASTRUCTURE ONLY.  Tracter::ASRFactory is only reached with FRONTEND_FACTORY (live audio front-ends); the harness
 * oracle reads HTK feature files. */
#ifndef ORACLE_SHIM_HARNESS_ASRFACTORY_H
#define ORACLE_SHIM_HARNESS_ASRFACTORY_H
#include <cassert>
#include <cstdio>
#include "TracterObject.h"
namespace Tracter {
class ASRFactory {
public:
    Component<float>* CreateSource(ISource*& s) { s = 0; return 0; }
    Component<float>* CreateFrontend(Component<float>* c) { return c; }
};
}
#endif
