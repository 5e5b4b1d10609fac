/* Empty stand-in for <gsl/gsl_sf_gamma.h>: the reference includes it but calls nothing from it.
 * TEST INFRASTRUCTURE ONLY. */
