/* Stub of <gsl/gsl_const_cgsm.h>: the five CGS constants globals.h:66-71 names.
 * Values are the CODATA figures GSL 2.x publishes; none of them is read on the
 * SPH/WVT hot path. TEST INFRASTRUCTURE ONLY. */
#ifndef TOY_COMPAT_GSL_CONST_CGSM_H
#define TOY_COMPAT_GSL_CONST_CGSM_H
#define GSL_CONST_CGSM_SPEED_OF_LIGHT (2.99792458e10)
#define GSL_CONST_CGSM_GRAVITATIONAL_CONSTANT (6.673e-8)
#define GSL_CONST_CGSM_BOLTZMANN (1.3806504e-16)
#define GSL_CONST_CGSM_MASS_PROTON (1.67262158e-24)
#define GSL_CONST_CGSM_MASS_ELECTRON (9.10938188e-28)
#endif
