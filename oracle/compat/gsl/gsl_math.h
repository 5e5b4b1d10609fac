/* Stub of <gsl/gsl_math.h> for building the reference hot path without GSL.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref build). GSL is an un-vendored, un-pinned
 * system dependency of the reference (Makefile:84); the hot-path translation units
 * only need the libm constants this header pulls in (globals.h:21). */
#ifndef TOY_COMPAT_GSL_MATH_H
#define TOY_COMPAT_GSL_MATH_H
#include <math.h>
#include <float.h>
#include <limits.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846264338328
#endif
#ifndef M_SQRT2
#define M_SQRT2 1.41421356237309504880168872421
#endif
#endif
