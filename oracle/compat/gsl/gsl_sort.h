/* Empty stand-in for <gsl/gsl_sort.h>: the reference includes it but calls nothing from it.
 * TEST INFRASTRUCTURE ONLY. */
