// Case shim only: the reference's umbrella header includes "staticFuncs.h"
// (reference include/SequenceAlignment.h:259) but ships "StaticFuncs.h".
// Needed on case-sensitive file systems; contains no reference code.
#include "StaticFuncs.h"
