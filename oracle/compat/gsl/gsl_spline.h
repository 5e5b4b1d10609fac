/* Stand-in for <gsl/gsl_spline.h> (setup.c:2, temperature.c:2, velocities.c:2): natural cubic
 * spline (gsl_interp_cspline), value and second derivative.  TEST INFRASTRUCTURE ONLY; see
 * gsl_integration.h. */
#ifndef TOY_COMPAT_GSL_SPLINE_H
#define TOY_COMPAT_GSL_SPLINE_H
#include <stddef.h>
#include "gsl_errno.h"

typedef struct { size_t cache; } gsl_interp_accel;
typedef struct { const char *name; } gsl_interp_type;
extern const gsl_interp_type *gsl_interp_cspline;

typedef struct {
    size_t size;
    double *x, *y, *c;       /* c = second derivatives / 2 at the knots */
} gsl_spline;

gsl_interp_accel *gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel *a);
gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size);
int gsl_spline_init(gsl_spline *s, const double xa[], const double ya[], size_t size);
void gsl_spline_free(gsl_spline *s);
double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a);
double gsl_spline_eval_deriv2(const gsl_spline *s, double x, gsl_interp_accel *a);
#endif
