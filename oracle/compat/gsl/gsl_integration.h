/* Stand-in for <gsl/gsl_integration.h> (setup.c:1, temperature.c:1, velocities.c:1) so that the
 * WHOLE reference driver can be built where GSL is absent.  TEST INFRASTRUCTURE ONLY.
 * Same call signatures and tolerance contract as GSL's QUADPACK routines; implemented in
 * oracle/compat/gsl_compat.c from the published QUADPACK algorithm (Piessens et al. 1983) with
 * Gauss-Kronrod tables computed by tools/make_gk_tables.py.  Not bit-identical to GSL: the
 * set-up stages it serves are outside the hot path and their parity is unpinned. */
#ifndef TOY_COMPAT_GSL_INTEGRATION_H
#define TOY_COMPAT_GSL_INTEGRATION_H
#include <stddef.h>
#include "gsl_errno.h"

typedef struct {
    double (*function)(double x, void *params);
    void *params;
} gsl_function;
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)

typedef struct {
    size_t limit;
    double *a, *b, *r, *e;
} gsl_integration_workspace;

enum { GSL_INTEG_GAUSS15 = 1, GSL_INTEG_GAUSS21 = 2, GSL_INTEG_GAUSS31 = 3,
       GSL_INTEG_GAUSS41 = 4, GSL_INTEG_GAUSS51 = 5, GSL_INTEG_GAUSS61 = 6 };

gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace *w);
int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *w, double *result,
                        double *abserr);
int gsl_integration_qags(const gsl_function *f, double a, double b, double epsabs,
                         double epsrel, size_t limit, gsl_integration_workspace *w,
                         double *result, double *abserr);
#endif
