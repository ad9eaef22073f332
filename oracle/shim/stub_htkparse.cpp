/* TEST INFRASTRUCTURE ONLY (oracle).  The bison/flex MMF parser is generated code that
 * cannot be built here (no bison); models reach the oracle through readBinary (JMBI). */
#include "htkparse.h"
HTKDef htk_def;
int htkparse(void*) { return 1; }
void cleanHTKDef() {}
