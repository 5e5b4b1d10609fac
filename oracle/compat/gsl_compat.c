/*
 * oracle/compat/gsl_compat.c -- the handful of GSL routines the reference DRIVER calls outside
 * the hot path (setup.c, temperature.c, velocities.c), so that the whole program can be built
 * and run here, once with its own tree.o/sph.o/wvt_relax.o/peano.o and once with
 * toycluster_b200/host/gpu_shim.c + libtoygpu.so in their place.
 *
 * TEST INFRASTRUCTURE ONLY.  GSL is absent from this image and un-pinned in the reference.
 * These are independent implementations of the published algorithms behind the same
 * signatures -- QUADPACK's adaptive Gauss-Kronrod QAG, QAGS as QAG(21) + Wynn epsilon
 * extrapolation, natural cubic splines -- accurate to the tolerances the callers ask for, NOT
 * bit-identical to GSL.  Both driver variants link this same file, so every stage outside the
 * hot path computes identically in both and differences in the output can only come from the
 * hot path.
 */
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include "gsl/gsl_integration.h"
#include "gsl/gsl_spline.h"
#include "gk_tables.h"

/* ------------------------------------------------------------------ error handler */

static gsl_error_handler_t *Handler = NULL;

gsl_error_handler_t *gsl_set_error_handler_off(void)
{
    gsl_error_handler_t *old = Handler;
    Handler = NULL;
    return old;
}

gsl_error_handler_t *gsl_set_error_handler(gsl_error_handler_t *h)
{
    gsl_error_handler_t *old = Handler;
    Handler = h;
    return old;
}

/* ------------------------------------------------------------------ Gauss-Kronrod */

/* One application of the (2n+1)-point Kronrod rule on [a, b]; QUADPACK's error heuristic. */
static double gk_rule(const gsl_function *f, double a, double b, int n, const double *xgk,
                      const double *wgk, const double *wg, double *abserr, double *resabs,
                      double *resasc)
{
    const double c = 0.5 * (a + b), h = 0.5 * (b - a), ah = fabs(h);
    double fv1[31], fv2[31];
    const double fc = GSL_FN_EVAL(f, c);
    /* the Gauss nodes are xgk[1], xgk[3], ...; n is even here, so the centre is not one */
    double resk = fc * wgk[n], resg = 0;
    double rabs = fabs(resk);

    for (int j = 0; j < n; j++) {
        const double dx = h * xgk[j];
        const double f1 = GSL_FN_EVAL(f, c - dx), f2 = GSL_FN_EVAL(f, c + dx);
        fv1[j] = f1;
        fv2[j] = f2;
        resk += wgk[j] * (f1 + f2);
        rabs += wgk[j] * (fabs(f1) + fabs(f2));
        if (j % 2 == 1) resg += wg[j / 2] * (f1 + f2);
    }
    const double mean = 0.5 * resk;
    double rasc = wgk[n] * fabs(fc - mean);
    for (int j = 0; j < n; j++) rasc += wgk[j] * (fabs(fv1[j] - mean) + fabs(fv2[j] - mean));

    double err = fabs((resk - resg) * h);
    rabs *= ah;
    rasc *= ah;
    if (rasc != 0 && err != 0) {
        const double s = pow(200 * err / rasc, 1.5);
        err = s < 1 ? rasc * s : rasc;
    }
    if (rabs > DBL_MIN / (50 * DBL_EPSILON)) {
        const double m = 50 * DBL_EPSILON * rabs;
        if (m > err) err = m;
    }
    *abserr = err;
    *resabs = rabs;
    *resasc = rasc;
    return resk * h;
}

static double apply_rule(const gsl_function *f, double a, double b, int key, double *abserr,
                         double *resabs, double *resasc)
{
    if (key <= GSL_INTEG_GAUSS21) return gk_rule(f, a, b, 10, xgk21, wgk21, wg21, abserr, resabs, resasc);
    if (key <= GSL_INTEG_GAUSS41) return gk_rule(f, a, b, 20, xgk41, wgk41, wg41, abserr, resabs, resasc);
    return gk_rule(f, a, b, 30, xgk61, wgk61, wg61, abserr, resabs, resasc);
}

gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n)
{
    gsl_integration_workspace *w = malloc(sizeof *w);
    w->limit = n;
    w->a = malloc(4 * n * sizeof(double));
    w->b = w->a + n;
    w->r = w->b + n;
    w->e = w->r + n;
    return w;
}

void gsl_integration_workspace_free(gsl_integration_workspace *w)
{
    if (!w) return;
    free(w->a);
    free(w);
}

/* Wynn's epsilon algorithm on the sequence s[0..m-1] (m <= 64):
 *   eps_{-1}^(j) = 0, eps_0^(j) = s_j, eps_{k+1}^(j) = eps_{k-1}^(j+1) + 1/(eps_k^(j+1) - eps_k^(j)).
 * The even columns converge to the limit; returns the last entry of the highest even column
 * that could be formed and, as the error estimate, its distance to the previous even column. */
static double wynn(const double *s, int m, double *err)
{
    double prev[65], cur[65], next[65];
    double best = s[m - 1], last_even = s[m - 1];
    *err = m > 1 ? fabs(s[m - 1] - s[m - 2]) : HUGE_VAL;
    for (int j = 0; j <= m; j++) prev[j] = 0;
    for (int j = 0; j < m; j++) cur[j] = s[j];
    int len = m;
    for (int col = 1; len > 1; col++) {
        for (int j = 0; j + 1 < len; j++) {
            const double d = cur[j + 1] - cur[j];
            if (d == 0 || !isfinite(1 / d)) return best;      /* converged to round-off */
            next[j] = prev[j + 1] + 1 / d;
        }
        len--;
        memcpy(prev, cur, sizeof(double) * (len + 1));
        memcpy(cur, next, sizeof(double) * len);
        if (col % 2 == 0) {
            best = cur[len - 1];
            *err = fabs(best - last_even);
            last_even = best;
        }
    }
    return best;
}

static int adaptive(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                    size_t limit, int key, int extrapolate, gsl_integration_workspace *w,
                    double *result, double *abserr)
{
    if (limit > w->limit) limit = w->limit;
    double resabs, resasc, err;
    double res = apply_rule(f, a, b, key, &err, &resabs, &resasc);
    size_t n = 1;
    w->a[0] = a; w->b[0] = b; w->r[0] = res; w->e[0] = err;
    double area = res, errsum = err;
    double tol = fmax(epsabs, epsrel * fabs(area));
    int status = GSL_SUCCESS;
    double seq[64];
    int nseq = 0;
    double ext_best = res, ext_err = HUGE_VAL;

    if ((err <= 100 * DBL_EPSILON * resabs && err > tol) || limit == 1) {
        *result = res; *abserr = err;
        return err > tol ? GSL_EROUND : GSL_SUCCESS;
    }
    while (errsum > tol) {
        if (n >= limit) { status = GSL_EMAXITER; break; }
        size_t worst = 0;                       /* interval with the largest error */
        for (size_t i = 1; i < n; i++) if (w->e[i] > w->e[worst]) worst = i;
        const double a1 = w->a[worst], b2 = w->b[worst], m = 0.5 * (a1 + b2);
        if (!(m > a1 && m < b2)) { status = GSL_EROUND; break; }   /* interval cannot shrink */
        double e1, e2, ra, rs;
        const double r1 = apply_rule(f, a1, m, key, &e1, &ra, &rs);
        const double r2 = apply_rule(f, m, b2, key, &e2, &ra, &rs);
        area += r1 + r2 - w->r[worst];
        errsum += e1 + e2 - w->e[worst];
        w->a[worst] = a1; w->b[worst] = m; w->r[worst] = r1; w->e[worst] = e1;
        w->a[n] = m; w->b[n] = b2; w->r[n] = r2; w->e[n] = e2;
        n++;
        tol = fmax(epsabs, epsrel * fabs(area));
        if (extrapolate && (n & (n - 1)) == 0 && nseq < 64) {      /* after 2, 4, 8, ... pieces */
            seq[nseq++] = area;
            if (nseq >= 5) {
                double ee;
                const double v = wynn(seq, nseq, &ee);
                ee = fmax(ee, 10 * DBL_EPSILON * fabs(v));
                if (ee < ext_err) { ext_err = ee; ext_best = v; }
                if (ext_err <= fmax(epsabs, epsrel * fabs(ext_best)) && ext_err < errsum) break;
            }
        }
    }
    double sum = 0;                              /* re-sum for round-off */
    for (size_t i = 0; i < n; i++) sum += w->r[i];
    if (extrapolate && ext_err < errsum) { *result = ext_best; *abserr = ext_err; }
    else { *result = sum; *abserr = errsum; }
    if (status != GSL_SUCCESS && Handler) Handler("integration did not converge", __FILE__, __LINE__, status);
    return status;
}

int gsl_integration_qag(const gsl_function *f, double a, double b, double epsabs, double epsrel,
                        size_t limit, int key, gsl_integration_workspace *w, double *result,
                        double *abserr)
{
    return adaptive(f, a, b, epsabs, epsrel, limit, key, 0, w, result, abserr);
}

int gsl_integration_qags(const gsl_function *f, double a, double b, double epsabs,
                         double epsrel, size_t limit, gsl_integration_workspace *w,
                         double *result, double *abserr)
{
    return adaptive(f, a, b, epsabs, epsrel, limit, GSL_INTEG_GAUSS21, 1, w, result, abserr);
}

/* ------------------------------------------------------------------ natural cubic spline */

static const gsl_interp_type Cspline = {"cspline"};
const gsl_interp_type *gsl_interp_cspline = &Cspline;

gsl_interp_accel *gsl_interp_accel_alloc(void) { return calloc(1, sizeof(gsl_interp_accel)); }
void gsl_interp_accel_free(gsl_interp_accel *a) { free(a); }

gsl_spline *gsl_spline_alloc(const gsl_interp_type *T, size_t size)
{
    (void)T;
    gsl_spline *s = malloc(sizeof *s);
    s->size = size;
    s->x = malloc(3 * size * sizeof(double));
    s->y = s->x + size;
    s->c = s->y + size;
    return s;
}

void gsl_spline_free(gsl_spline *s)
{
    if (!s) return;
    free(s->x);
    free(s);
}

/* y'' continuous, y''(x_0) = y''(x_{n-1}) = 0; tridiagonal system by the Thomas algorithm. */
int gsl_spline_init(gsl_spline *s, const double xa[], const double ya[], size_t size)
{
    const size_t n = size;
    memcpy(s->x, xa, n * sizeof(double));
    memcpy(s->y, ya, n * sizeof(double));
    double *c = s->c;
    for (size_t i = 0; i < n; i++) c[i] = 0;
    if (n < 3) return GSL_SUCCESS;
    double *diag = malloc(2 * n * sizeof(double)), *rhs = diag + n;
    for (size_t i = 1; i + 1 < n; i++) {
        const double h0 = xa[i] - xa[i - 1], h1 = xa[i + 1] - xa[i];
        diag[i] = 2 * (h0 + h1);
        rhs[i] = 3 * ((ya[i + 1] - ya[i]) / h1 - (ya[i] - ya[i - 1]) / h0);
    }
    for (size_t i = 2; i + 1 < n; i++) {           /* forward elimination */
        const double h = xa[i] - xa[i - 1];
        const double m = h / diag[i - 1];
        diag[i] -= m * h;
        rhs[i] -= m * rhs[i - 1];
    }
    for (size_t i = n - 2; i >= 1; i--) {          /* back substitution */
        const double h = xa[i + 1] - xa[i];
        c[i] = (rhs[i] - h * c[i + 1]) / diag[i];
    }
    free(diag);
    return GSL_SUCCESS;
}

static size_t locate(const gsl_spline *s, double x, gsl_interp_accel *a)
{
    const size_t n = s->size;
    size_t lo = 0, hi = n - 1;
    if (a && a->cache + 1 < n && x >= s->x[a->cache] && x < s->x[a->cache + 1]) return a->cache;
    while (hi - lo > 1) {
        const size_t m = (lo + hi) / 2;
        if (s->x[m] > x) hi = m; else lo = m;
    }
    if (a) a->cache = lo;
    return lo;
}

double gsl_spline_eval(const gsl_spline *s, double x, gsl_interp_accel *a)
{
    const size_t i = locate(s, x, a);
    const double h = s->x[i + 1] - s->x[i], d = x - s->x[i];
    const double dy = s->y[i + 1] - s->y[i];
    const double ci = s->c[i], cn = s->c[i + 1];
    const double b = dy / h - h * (cn + 2 * ci) / 3;
    const double e = (cn - ci) / (3 * h);
    return s->y[i] + d * (b + d * (ci + d * e));
}

double gsl_spline_eval_deriv2(const gsl_spline *s, double x, gsl_interp_accel *a)
{
    const size_t i = locate(s, x, a);
    const double h = s->x[i + 1] - s->x[i], d = x - s->x[i];
    const double ci = s->c[i], cn = s->c[i + 1];
    return 2 * ci + 6 * d * (cn - ci) / (3 * h);
}
