/*
 * oracle/ref_harness.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Thin C harness linked against the UNMODIFIED reference hot-path sources
 * (peano.c sort.c tree.c sph.c wvt_relax.c aux.c, compiled where they lie under
 * /root/reference/src by oracle/Makefile) to form oracle/_ref/libtoyref.so.
 * It only supplies what the full driver would have supplied around the path:
 *
 *   - the global state Setup() fills (Param, Halo[], P, SphP; setup.c:244-250),
 *   - Gas_density_profile (setup.c:598-615, default build: no DOUBLE_BETA_COOL_CORES),
 *   - gsl_heapsort_index (the one GSL routine on the path, sort.c:192; GSL itself is
 *     absent here, so this is the published sift-down heapsort on an index array),
 *   - link-time hooks so a test can observe the monolithic WVT loop
 *     (wvt_relax.c:61-218) once per iteration without editing it.  wvt_relax.c is
 *     compiled with -DFind_sph_quantities=refhook_find_sph -Dprintf=refhook_printf
 *     -DMalloc_info=refhook_malloc, nothing else.
 *
 * Nothing under toycluster_b200/ may link or load this file's product; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 */
#include "globals.h"
#include "tree.h"
#include <setjmp.h>

/* gsl_heapsort_index (sort.c:192): compat/gsl_heapsort.c */

/* ------------------------------------------------- setup.c:598-615 (default) */

double Gas_density_profile(const double r, const double rho0, const double beta,
                           const double rc, const double rcut, const bool Is_Cuspy)
{
    (void)Is_Cuspy;
    double rho = rho0 * pow(1 + p2(r / rc), -3.0 / 2.0 * beta)
                 / (1 + p3(r / rcut) * (r / rcut));
    return rho;
}

/* ------------------------------------------------------------------ hooks */

static jmp_buf Stop_env;
static int Iter_seen = 0;
static int Iter_limit = 1 << 30;
static int (*Iter_cb)(int) = NULL;
static double Time_density = 0;

static float *Wvt_buf[4] = {NULL, NULL, NULL, NULL}; /* hsml, delta[0..2] */
static int Wvt_nbuf = 0;
static int Wvt_leaked = 0; /* loop left by longjmp: wvt_relax.c:220 never ran */

static char *Log_buf = NULL;
static size_t Log_len = 0, Log_cap = 0;
static int Log_echo = 0;

int refhook_printf(const char *fmt, ...)
{
    char line[1024];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(line, sizeof line, fmt, ap);
    va_end(ap);
    if (n < 0)
        return n;
    if ((size_t)n >= sizeof line)
        n = sizeof line - 1;
    if (Log_len + n + 1 > Log_cap) {
        Log_cap = 2 * (Log_len + n + 1) + 4096;
        Log_buf = realloc(Log_buf, Log_cap);
    }
    memcpy(Log_buf + Log_len, line, n);
    Log_len += n;
    Log_buf[Log_len] = 0;
    if (Log_echo)
        fwrite(line, 1, n, stdout);
    return n;
}

void *refhook_malloc(const char *func, const char *file, const int line, size_t size)
{
    void *ptr = Malloc_info(func, file, line, size);
    if (!strcmp(func, "Regularise_sph_particles") && Wvt_nbuf < 4)
        Wvt_buf[Wvt_nbuf++] = ptr;
    return ptr;
}

#ifndef SHIM_BUILD
/* Called by wvt_relax.c:67 in place of Find_sph_quantities(). At entry the scratch
 * arrays hsml[]/delta[][] still hold the previous iteration's values in the particle
 * order of the previous sort, and P.Pos has been moved (wvt_relax.c:193-213). */
void refhook_find_sph(void)
{
    const int it = Iter_seen++;

    if (Iter_cb != NULL && Iter_cb(it) != 0)
        longjmp(Stop_env, 1);

    if (it >= Iter_limit)
        longjmp(Stop_env, 1);

    double t0 = omp_get_wtime();
    Find_sph_quantities();
    Time_density += omp_get_wtime() - t0;
}
#endif

/* ------------------------------------------------------------------ API */

int ref_sizeof_P(void) { return (int)sizeof(struct ParticleData); }
int ref_sizeof_SphP(void) { return (int)sizeof(struct GasParticleData); }
void *ref_P(void) { return P; }
void *ref_SphP(void) { return SphP; }
int ref_nthreads(void) { return Omp.NThreads; }

/* halo_tab: nhalos rows of 9 doubles {dcom x,y,z, rho0, beta, rcore, rcut, cuspy, mass_gas} */
int ref_setup(int n_gas, double boxsize, double mpart, double mtotal, int nhalos,
              const double *halo_tab, int nthreads)
{
    if (P != NULL || n_gas <= 0 || nhalos > MAXHALOS)
        return -1; /* one problem size per loaded copy (peano.c:53-61, tree.c:343-346) */

    if (nthreads > 0)
        omp_set_num_threads(nthreads);

    #pragma omp parallel
    {   /* main.c:15-26 */
        Omp.ThreadID = omp_get_thread_num();
        Omp.NThreads = omp_get_num_threads();
        Omp.Seed[2] = 14041981 * (Omp.ThreadID + 1);
    }

    memset(&Param, 0, sizeof Param);
    Param.Ntotal = n_gas;
    Param.Npart[0] = n_gas;
    Param.Mpart[0] = mpart;
    Param.Mtotal = mtotal;
    Param.Boxsize = boxsize;
    Param.Nhalos = nhalos;

    for (int i = 0; i < nhalos; i++) {
        const double *h = halo_tab + 9 * i;
        memset(&Halo[i], 0, sizeof Halo[i]);
        Halo[i].D_CoM[0] = h[0];
        Halo[i].D_CoM[1] = h[1];
        Halo[i].D_CoM[2] = h[2];
        Halo[i].Rho0 = h[3];
        Halo[i].Beta = h[4];
        Halo[i].Rcore = h[5];
        Halo[i].Rcut = h[6];
        Halo[i].Have_Cuspy = (int)h[7];
        Halo[i].Mass[0] = h[8];
    }

    P = calloc(n_gas, sizeof *P);          /* setup.c:244-246 */
    SphP = calloc(n_gas, sizeof *SphP);    /* setup.c:248-250 */

    return (P && SphP) ? 0 : -2;
}

/* pos: n x 3 floats; hsml may be NULL (cold start, SphP.Hsml = 0). IDs are 0..n-1 so the
 * permutation applied by peano.c:85-126 can be read back from P[].ID. */
void ref_load(const float *pos, const float *hsml)
{
    const int n = Param.Npart[0];
    for (int i = 0; i < n; i++) {
        memset(&P[i], 0, sizeof P[i]);
        memset(&SphP[i], 0, sizeof SphP[i]);
        P[i].Pos[0] = pos[3 * i];
        P[i].Pos[1] = pos[3 * i + 1];
        P[i].Pos[2] = pos[3 * i + 2];
        P[i].ID = i;
        SphP[i].ID = (float)i;
        if (hsml)
            SphP[i].Hsml = hsml[i];
    }
}

void ref_set_apot(const float *apot) /* n x 3 */
{
    const int n = Param.Npart[0];
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++)
            SphP[i].Apot[k] = apot[3 * i + k];
}

/* out arrays may be NULL. All in the current (sorted) particle order. */
void ref_read(float *pos, int *id, float *hsml, float *rho, float *varhsml,
              float *rho_model, float *bfld, unsigned long long *key_hi,
              unsigned long long *key_lo, int *tree_parent)
{
    const int n = Param.Npart[0];
    for (int i = 0; i < n; i++) {
        if (pos) {
            pos[3 * i] = P[i].Pos[0];
            pos[3 * i + 1] = P[i].Pos[1];
            pos[3 * i + 2] = P[i].Pos[2];
        }
        if (id) id[i] = P[i].ID;
        if (hsml) hsml[i] = SphP[i].Hsml;
        if (rho) rho[i] = SphP[i].Rho;
        if (varhsml) varhsml[i] = SphP[i].VarHsmlFac;
        if (rho_model) rho_model[i] = SphP[i].Rho_Model;
        if (bfld) {
            bfld[3 * i] = SphP[i].Bfld[0];
            bfld[3 * i + 1] = SphP[i].Bfld[1];
            bfld[3 * i + 2] = SphP[i].Bfld[2];
        }
        if (key_hi) key_hi[i] = (unsigned long long)(P[i].Key >> 64);
        if (key_lo) key_lo[i] = (unsigned long long)P[i].Key;
        if (tree_parent) tree_parent[i] = P[i].Tree_Parent;
    }
}

#ifdef SHIM_BUILD
/* libtoyshim.so: the same harness, but the three operators come from
 * toycluster_b200/host/gpu_shim.c + libtoygpu.so instead of sph.c / wvt_relax.c. */
void toyshim_set_max_iters(int n);
#else
void ref_peano_key(double x, double y, double z, unsigned long long *hi,
                   unsigned long long *lo, int reversed)
{
    peanoKey k = reversed ? Reversed_Peano_Key(x, y, z) : Peano_Key(x, y, z);
    *hi = (unsigned long long)(k >> 64);
    *lo = (unsigned long long)k;
}
#endif

#ifndef SHIM_BUILD
void ref_sort(void) { Sort_Particles_By_Peano_Key(); }
void ref_build_tree(void) { Build_Tree(); }
int ref_find_ngb_tree(int ipart, float hsml, int *list) { return Find_ngb_tree(ipart, hsml, list); }
int ref_find_ngb_simple(int ipart, float hsml, int *list) { return Find_ngb_simple(ipart, hsml, list); }
float ref_guess_hsml(int ipart) { return Guess_hsml(ipart, DESNNGB); }
#endif
float ref_global_density_model(int ipart) { return Global_density_model(ipart); }
void ref_find_sph_quantities(void) { Find_sph_quantities(); }
void ref_bfld_from_rotA(void) { Bfld_from_rotA_SPH(); }

/* Run the WVT loop. Stops (by longjmp out of the hook) before the density pass of
 * iteration `max_iters`, i.e. after max_iters complete iterations, or earlier when the
 * reference's own termination logic fires or the callback returns non-zero.
 * Returns the number of density passes the loop started. */
int ref_regularise(int max_iters, int (*cb)(int), int echo)
{
    if (Wvt_leaked)
        for (int i = 0; i < Wvt_nbuf; i++)
            free(Wvt_buf[i]);
    Wvt_leaked = 0;

    Iter_seen = 0;
    Iter_limit = max_iters;
    Iter_cb = cb;
    Wvt_nbuf = 0;
    Log_len = 0;
    Log_echo = echo;
    Time_density = 0;

#ifdef SHIM_BUILD
    toyshim_set_max_iters(max_iters);
    Regularise_sph_particles();
    return max_iters;
#else
    if (setjmp(Stop_env) == 0)
        Regularise_sph_particles();
    else
        Wvt_leaked = 1;
#endif

    Iter_cb = NULL;
    return Iter_seen;
}

/* Scratch arrays of the last/ongoing WVT loop: which = 0 hsml, 1..3 delta x,y,z.
 * Valid inside the callback and after a ref_regularise() that was stopped by the hook. */
const float *ref_wvt_scratch(int which)
{
    return (which >= 0 && which < Wvt_nbuf) ? Wvt_buf[which] : NULL;
}

const char *ref_log(void) { return Log_buf ? Log_buf : ""; }
double ref_time_density(void) { return Time_density; }
