// Stand-in for <Rmath.h>; the window kernels themselves never call nmath.
#ifndef GB_REF_SHIM_RMATH_H
#define GB_REF_SHIM_RMATH_H
#endif
