/* Stand-in for <gsl/gsl_errno.h> (velocities.c:223,271).  TEST INFRASTRUCTURE ONLY. */
#ifndef TOY_COMPAT_GSL_ERRNO_H
#define TOY_COMPAT_GSL_ERRNO_H
typedef void gsl_error_handler_t(const char *reason, const char *file, int line, int gsl_errno);
gsl_error_handler_t *gsl_set_error_handler_off(void);
gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *new_handler);
enum { GSL_SUCCESS = 0, GSL_EMAXITER = 11, GSL_EROUND = 18 };
#endif
