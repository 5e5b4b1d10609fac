/* Stub of <gsl/gsl_heapsort.h> (sort.c:9). The single GSL routine on the hot path,
 * gsl_heapsort_index (sort.c:192), is supplied by oracle/ref_harness.c.
 * TEST INFRASTRUCTURE ONLY. */
#ifndef TOY_COMPAT_GSL_HEAPSORT_H
#define TOY_COMPAT_GSL_HEAPSORT_H
#include <stddef.h>
typedef int (*gsl_comparison_fn_t)(const void *, const void *);
int gsl_heapsort_index(size_t *p, const void *array, size_t count, size_t size,
                       gsl_comparison_fn_t compare);
#endif
