/* Empty stub of <gsl/gsl_rng.h> (globals.h:24); the hot path draws no random numbers.
 * TEST INFRASTRUCTURE ONLY. */
