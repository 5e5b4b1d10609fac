/*
 * oracle/toy_oracle.c -- CPU restatement of Toycluster's SPH-density + WVT hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load this; the product (toycluster_b200/) never does and has no CPU fallback.
 *
 * Plain C99 on explicit arrays (no globals), written from the reference's behaviour:
 *   peano.c:128-203, 211-284   Hilbert keys             -> to_peano_key, to_reversed_key
 *   peano.c:46-126 + sort.c    key sort + reorder       -> to_sort (ties by index, see below)
 *   tree.c:124-357             the sequential octree build, node centres included -> tree_build
 *   tree.c:25-111              the depth-first walk, its open test and the particle predicate
 *                              -> find_ngb / to_find_ngb (ascending, first 2360)
 *   tree.c:113-121             Guess_hsml -> to_guess_hsml
 *   sph.c:13-214, 426-440      density / hsml solve     -> to_find_sph_quantities
 *   wvt_relax.c:61-256         one WVT iteration + the control loop -> to_wvt_iteration,
 *                              to_regularise
 *   sph.c:216-300              rot(A)                   -> to_bfld_from_rotA
 * Pinned against the reference itself (oracle/_ref, the unmodified sources) by
 * tests/test_oracle_vs_ref.py and against the committed fixtures in tests/golden/.
 *
 * Deliberate difference: gsl_heapsort_index (sort.c:192) is unstable, so bit-identical keys
 * (bit-identical positions) come out in heap order; here, and on the GPU, ties are broken by
 * the previous index.  Fixtures assert there are no duplicate keys.
 *
 * Compile with -ffp-contract=off: the reference is built -std=c99, so float expressions such
 * as tree.c:88 are evaluated without FMA.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <omp.h>

#define DESNNGB 295     /* globals.h:48 */
#define NNGBDEV 0.05    /* globals.h:49 */
#define NGBMAX 2360     /* globals.h:50 */
#define NUMITER 64      /* wvt_relax.c:7 */
#define SQRT3 1.73205080756887719       /* globals.h:62, the literal */
#define FOURPITHIRD 4.18879032135009765 /* globals.h:63, the literal */

typedef unsigned __int128 u128;

typedef struct {
    int n;
    double box, mpart, mtotal;
    int nhalos;
    const double *halos; /* rows of 9: dcom[3], rho0, beta, rcore, rcut, cuspy, mass_gas */
} to_sys;

/* ------------------------------------------------------------------ Peano-Hilbert keys */

static void transpose_axes(double x, double y, double z, uint64_t X[3])
{
    /* peano.c:134-177: axes {y,z,x} scaled by 2^63, Skilling's inverse-undo over bit planes
     * 63..1, then Gray encoding. */
    const double scale = 9223372036854775808.0;
    uint64_t a = (uint64_t)(y * scale), b = (uint64_t)(z * scale), c = (uint64_t)(x * scale);

    for (int plane = 63; plane >= 1; plane--) {
        const uint64_t q = (uint64_t)1 << plane, low = q - 1;
        if (a & q) a ^= low;
        if (b & q) a ^= low; else { uint64_t t = (a ^ b) & low; a ^= t; b ^= t; }
        if (c & q) a ^= low; else { uint64_t t = (a ^ c) & low; a ^= t; c ^= t; }
    }
    b ^= a;
    c ^= b;
    uint64_t g = c;
    for (int s = 1; s < 64; s <<= 1) g ^= g >> s;
    const uint64_t t = c ^ g;
    X[0] = a ^ t;
    X[1] = b ^ t;
    X[2] = g;
}

static u128 peano_key(double x, double y, double z)
{
    /* peano.c:181-202: planes 63..21 interleaved MSB first into 128 bits, then << 2; the
     * plane-63 triplet falls off the top, leaving planes 62..21. */
    uint64_t X[3];
    transpose_axes(x, y, z, X);
    u128 key = 0;
    for (int plane = 62; plane >= 21; plane--)
        key = (key << 3) | (((X[0] >> plane) & 1) << 2) | (((X[1] >> plane) & 1) << 1) |
              ((X[2] >> plane) & 1);
    return key << 2;
}

static u128 reversed_key(double x, double y, double z)
{
    /* peano.c:264-283: planes 20..62 pushed in that order (so plane 62, tree level 1, ends
     * lowest), then a zero level-0 triplet. */
    uint64_t X[3];
    transpose_axes(x, y, z, X);
    u128 key = 0;
    for (int plane = 20; plane <= 62; plane++)
        key = (key << 3) | (((X[0] >> plane) & 1) << 2) | (((X[1] >> plane) & 1) << 1) |
              ((X[2] >> plane) & 1);
    return key << 3;
}

void to_peano_key(double x, double y, double z, int reversed, uint64_t *hi, uint64_t *lo)
{
    const u128 k = reversed ? reversed_key(x, y, z) : peano_key(x, y, z);
    *hi = (uint64_t)(k >> 64);
    *lo = (uint64_t)k;
}

/* ------------------------------------------------------------------ sort */

typedef struct { u128 key; int idx; } keyed;

static int cmp_keyed(const void *a, const void *b)
{
    const keyed *p = a, *q = b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    return (p->idx > q->idx) - (p->idx < q->idx);
}

/* peano.c:46-81: keys of pos/box (double divide of the float coordinate), ascending
 * permutation.  perm[k] = previous index of the particle now at k. Returns #duplicate keys. */
int to_sort(int n, const float *pos, double box, int *perm, uint64_t *key_hi, uint64_t *key_lo)
{
    keyed *v = malloc((size_t)n * sizeof *v);
    #pragma omp parallel for
    for (int i = 0; i < n; i++) {
        v[i].key = peano_key(pos[3 * i] / box, pos[3 * i + 1] / box, pos[3 * i + 2] / box);
        v[i].idx = i;
    }
    qsort(v, n, sizeof *v, cmp_keyed);
    int dup = 0;
    for (int k = 0; k < n; k++) {
        perm[k] = v[k].idx;
        if (key_hi) key_hi[k] = (uint64_t)(v[k].key >> 64);
        if (key_lo) key_lo[k] = (uint64_t)v[k].key;
        if (k && v[k].key == v[k - 1].key) dup++;
    }
    free(v);
    return dup;
}

/* ------------------------------------------------------------------ neighbour search */

/* tree.c:67-88, all in float, no FMA: |d| per axis, one wrap by the box, r^2 < h^2. */
static inline int near_f32(const float *pi, const float *pj, float h, float box, float boxhalf)
{
    float dx = fabsf(pi[0] - pj[0]), dy = fabsf(pi[1] - pj[1]), dz = fabsf(pi[2] - pj[2]);
    if (dx > boxhalf) dx -= box;
    if (dy > boxhalf) dy -= box;
    if (dz > boxhalf) dz -= box;
    return dx * dx + dy * dy + dz * dz < h * h;
}

/* The reference's search index, restated in full: the sequentially built octree of
 * tree.c:124-271 (nodes in depth-first order; `node + 1` is the first child, DNext skips a
 * finished subtree, DNext < 0 marks a leaf whose first particle is -DNext-1) and the walk of
 * tree.c:25-111.  The walk is part of the contract, not only its particle predicate: a node's
 * centre is placed by comparing the position of the particle that creates it with the centre
 * of the parent (tree.c:298-310), while membership is decided by key triplets.  A particle
 * lying exactly on a centre plane of its parent cell (or within the float rounding of a deep
 * centre) is a member of the upper cell but `Pos > centre` is false, so the node -- and every
 * descendant, whose centres derive from it -- is displaced by one cell size, and the open
 * test of tree.c:56-58 then prunes particles that are within reach.  ~8 nodes per 1e6
 * particles in the merger workloads; the neighbour sets, and through them rho / hsml / the
 * displacements, are the reference's only if this is reproduced. */
typedef struct { int level, triplet, npart, dnext; float pos[3], size; } tnode;
typedef struct {
    int n, nn;
    double box;
    tnode *T;
    int *parent_of;     /* P[i].Tree_Parent */
} tree;

static void tree_free(tree *t)
{
    free(t->T);
    free(t->parent_of);
    free(t);
}

/* tree.c:284-317 */
static void new_node(tree *t, const float *pos, int ip, int par, u128 key3, int lvl)
{
    tnode *N = &t->T[t->nn++];
    const tnode *Q = &t->T[par];
    N->level = lvl;
    N->triplet = (int)(key3 & 7);
    N->npart = 1;
    N->dnext = -ip - 1;
    const float size = t->box / (1 << lvl);                        /* tree.c:304 */
    N->size = size;
    for (int d = 0; d < 3; d++) {
        const int sign = -1 + 2 * (pos[3 * ip + d] > Q->pos[d]);   /* tree.c:298-302 */
        N->pos[d] = Q->pos[d] + sign * size * 0.5;                 /* tree.c:308-310 */
    }
    t->parent_of[ip] = par;
}

static tree *tree_build(int n, const float *pos, double box)
{
    tree *t = malloc(sizeof *t);
    const int maxn = (int)(n * 0.7) + 64;        /* tree.c:3,341 */
    t->n = n;
    t->box = box;
    t->nn = 0;
    t->T = calloc(maxn, sizeof *t->T);
    t->parent_of = malloc((size_t)n * sizeof(int));
    tnode *T = t->T;

    new_node(t, pos, 0, 0, 0, 0);                /* tree.c:131: root made from particle 0 */
    T[0].pos[0] = T[0].pos[1] = T[0].pos[2] = box / 2;             /* tree.c:133 */
    int last_parent = 0;
    u128 last_key = reversed_key((float)(pos[0] / box), (float)(pos[1] / box),
                                 (float)(pos[2] / box)) >> 3;   /* tree.c:137-143 (float!) */

    for (int ip = 1; ip < n; ip++) {
        u128 key = reversed_key(pos[3 * ip] / box, pos[3 * ip + 1] / box, pos[3 * ip + 2] / box);
        int node = 0, lvl = 0, parent = 0, new_branch = 1;

        while (lvl < 42) {
            if ((int)(key & 7) == T[node].triplet) {              /* inside: descend */
                if (T[node].npart == 1) {                        /* refine (tree.c:163-171) */
                    T[node].dnext = 0;
                    new_node(t, pos, ip - 1, node, last_key, lvl + 1);
                    last_key >>= 3;
                }
                T[node].npart++;
                new_branch &= node != last_parent;
                parent = node;
                node++;
                lvl++;
                key >>= 3;
            } else {                                              /* skip to the sibling */
                if (T[node].dnext == 0 || node == t->nn - 1) break;
                node += T[node].dnext > 1 ? T[node].dnext : 1;
            }
        }
        if (lvl > 41) { t->parent_of[ip] = parent; continue; }    /* tree.c:194-199 */

        if (new_branch) {                                          /* tree.c:201-226 */
            int c = 0;
            if (T[node].npart <= 8) c = node;
            else if (T[last_parent].npart <= 8) c = last_parent;
            if (c != 0) {
                T[c].dnext = -ip + T[c].npart - 1;
                memset(&T[c + 1], 0, (size_t)(t->nn - c - 1) * sizeof *T);
                t->nn = c + 1;
                for (int j = ip - T[c].npart; j < ip; j++) t->parent_of[j] = c;
            }
        }
        if (T[node].dnext == 0) T[node].dnext = t->nn - node;
        new_node(t, pos, ip, parent, key, lvl);
        last_key = key >> 3;
        last_parent = parent;
        if (t->nn >= maxn - 2) break;                              /* tree.c:287-292 would exit */
    }

    T[0].dnext = 0;                                                /* tree.c:238-266 */
    int stack[43] = {0}, lowest = 0;
    for (int i = 1; i < t->nn; i++) {
        const int lvl = T[i].level;
        while (lvl <= lowest) {
            const int node = stack[lowest];
            if (node > 0) T[node].dnext = i - node;
            stack[lowest] = 0;
            lowest--;
        }
        if (T[i].dnext == 0) { stack[lvl] = i; lowest = lvl; }
    }
    return t;
}

/* tree.c:37-58: |d| per axis in float, one wrap, against 0.5*sqrt3*Size + hsml. */
static inline int node_open(const tnode *N, const float *pi, float h, float box, float boxhalf)
{
    float dx = fabsf(pi[0] - N->pos[0]), dy = fabsf(pi[1] - N->pos[1]), dz = fabsf(pi[2] - N->pos[2]);
    if (dx > boxhalf) dx -= box;
    if (dy > boxhalf) dy -= box;
    if (dz > boxhalf) dz -= box;
    const float dl = 0.5 * SQRT3 * N->size + h;
    return dx * dx + dy * dy + dz * dz < dl * dl;
}

/* Find_ngb_tree (tree.c:25-111): depth-first walk from node 1, ascending particle index,
 * returns as soon as the list holds NGBMAX entries. */
static int find_ngb(const tree *t, const float *pos, int i, float h, int *list)
{
    const float box = (float)t->box, boxhalf = (float)(t->box * 0.5);
    const float *pi = pos + 3 * i;
    const tnode *T = t->T;
    int cnt = 0, node = 1;
    if (t->nn < 2) {   /* a one-particle "tree": tree.c would read past NNodes; nothing to find */
        return 0;
    }
    for (;;) {
        if (node_open(&T[node], pi, h, box, boxhalf)) {
            if (T[node].dnext < 0) {
                const int first = -(T[node].dnext + 1), last = first + T[node].npart;
                for (int j = first; j < last; j++) {
                    if (near_f32(pi, pos + 3 * j, h, box, boxhalf)) list[cnt++] = j;
                    if (cnt == NGBMAX) return cnt;
                }
            }
            node++;
            if (node >= t->nn) break;
            continue;
        }
        node += T[node].dnext > 1 ? T[node].dnext : 1;
        if (node >= t->nn) break;
    }
    return cnt;
}

/* The same contract without the tree: every j with near_f32, ascending, first NGBMAX
 * (wvt_relax.c:296-340, Find_ngb_simple). */
int to_find_ngb_simple(int n, const float *pos, double box, int i, float h, int *list)
{
    const float boxf = (float)box, boxhalf = (float)(box * 0.5);
    int cnt = 0;
    for (int j = 0; j < n && cnt < NGBMAX; j++)
        if (near_f32(pos + 3 * i, pos + 3 * j, h, boxf, boxhalf)) list[cnt++] = j;
    return cnt;
}

int to_find_ngb(int n, const float *pos, double box, int i, float h, int *list)
{
    tree *t = tree_build(n, pos, box);
    const int cnt = find_ngb(t, pos, i, h, list);
    tree_free(t);
    return cnt;
}

/* Number of nodes of the tree whose centre is not the centre of their key cell (diagnostic
 * for tests): a node is displaced when the sign test of tree.c:298-302 disagreed with the
 * cell the key put the particle in, or when an ancestor is displaced. */
int to_tree_displaced(int n, const float *pos, double box, int *node_first, int *node_count, int cap)
{
    tree *t = tree_build(n, pos, box);
    int found = 0;
    for (int k = 1; k < t->nn; k++) {
        const tnode *N = &t->T[k];
        if (N->dnext >= 0) continue;                 /* leaves carry the particles */
        const int first = -(N->dnext + 1);
        int out = 0;
        for (int j = first; j < first + N->npart && !out; j++)
            for (int d = 0; d < 3; d++)
                if (fabsf(pos[3 * j + d] - N->pos[d]) > 0.5001f * N->size) out = 1;
        if (out) {
            if (found < cap) { node_first[found] = first; node_count[found] = N->npart; }
            found++;
        }
    }
    tree_free(t);
    return found;
}

/* ------------------------------------------------------------------ Guess_hsml (tree.c) */

/* 2*Guess_hsml (sph.c:26, tree.c:113-121) for every particle of a SORTED position array. */
static void guess_from_tree(const tree *t, float *out)
{
    for (int i = 0; i < t->n; i++) {
        const tnode *N = &t->T[t->parent_of[i]];
        const float numdens = N->npart / (N->size * N->size * N->size);   /* tree.c:117 */
        const float s = pow(FOURPITHIRD / numdens, 1. / 3.);              /* tree.c:118 */
        out[i] = 2 * (2 * s);
    }
}

void to_guess_hsml(int n, const float *pos, double box, float *out)
{
    tree *t = tree_build(n, pos, box);
    guess_from_tree(t, out);
    tree_free(t);
}

/* ------------------------------------------------------------------ kernels */

static inline float wc6(float r, float h)          /* sph.c:426-432 */
{
    const double u = r / h;
    const double t = 1 - u;
    return 1365.0 / (64 * M_PI) / (h * h * h) * t * t * t * t * t * t * t * t *
           (1 + 8 * u + 25 * u * u + 32 * u * u * u);
}

static inline float wc6_deriv(float r, float h)    /* sph.c:434-440 */
{
    const float u = r / h;
    const double t = 1 - u;
    return 1365.0 / (64 * M_PI) / (h * h * h * h) * -22.0 * t * t * t * t * t * t * t * u *
           (16 * u * u + 7 * u + 1);
}

static inline double wc6_unnormalised(float r, float h)   /* wvt_relax.c:275-281 */
{
    const double u = r / h;
    const double t = 1 - u;
    return 1365.0 / (64 * M_PI) * t * t * t * t * t * t * t * t *
           (1 + 8 * u + 25 * u * u + 32 * u * u * u);
}

/* ------------------------------------------------------------------ density */

/* sph.c:80-214 on a frozen neighbour list. */
static int find_hsml(const to_sys *s, const float *pos, int i, const int *list, int cnt,
                     float *drho_out, float *hsml_io, float *rho_out, long *evals, long *iters)
{
    const double boxhalf = 0.5 * s->box, box = s->box;
    double upper = *hsml_io * SQRT3, lower = 0, hsml = *hsml_io, rho = 0, drho = 0;
    int it = 0, done = 0;

    for (;;) {
        const double pi[3] = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
        double wk_ngb = 0;
        rho = drho = 0;
        it++;
        for (int k = 0; k < cnt; k++) {
            const int j = list[k];
            double d[3];
            for (int a = 0; a < 3; a++) {
                d[a] = pi[a] - pos[3 * j + a];
                if (d[a] > boxhalf) d[a] -= box;
                if (d[a] < -boxhalf) d[a] += box;
            }
            const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            if (r2 > hsml * hsml) continue;
            const double r = sqrt(r2);
            const double wk = wc6(r, hsml), dwk = wc6_deriv(r, hsml);
            wk_ngb += FOURPITHIRD * wk * (hsml * hsml * hsml);
            rho += s->mpart * wk;
            drho += -s->mpart * (3 / hsml * wk + r / hsml * dwk);
        }
        *evals += cnt;
        if (it > 128) break;
        const double dev = fabs(wk_ngb - DESNNGB);
        if (dev < NNGBDEV) { done = 1; break; }
        if (fabs(upper - lower) < 1e-4) { hsml *= 1.26; break; }
        if (dev < 0.5 * DESNNGB) {
            const double omega = 1 + drho * hsml / (3 * rho);
            double fac = 1 - (wk_ngb - DESNNGB) / (3 * wk_ngb * omega);
            fac = fmin(1.24, fac);
            fac = fmax(1 / 1.24, fac);
            hsml *= fac;
        } else {
            if (wk_ngb > DESNNGB) upper = hsml;
            if (wk_ngb < DESNNGB) lower = hsml;
            hsml = pow(0.5 * (lower * lower * lower + upper * upper * upper), 1.0 / 3.0);
        }
    }
    *iters += it;
    *hsml_io = (float)hsml;
    *rho_out = (float)rho;
    if (done) {
        *drho_out = (float)drho;
        const double bias = -0.0116 * pow(DESNNGB * 0.01, -2.236) * s->mpart * wc6(0, hsml);
        *rho_out += bias;
    }
    return done;
}

/* The loop of sph.c:19-72 on an already SORTED particle set. hsml: in = warm start (0 =>
 * 2*Guess_hsml), out = solved. stats (may be NULL): pair_evals, searches, hsml_iters. */
int to_density(const to_sys *s, const float *pos, float *hsml, float *rho, float *varhsml,
               long *stats)
{
    const int n = s->n;
    tree *g = tree_build(n, pos, s->box);
    float *guess = NULL;
    for (int i = 0; i < n; i++)
        if (hsml[i] == 0) {
            guess = malloc((size_t)n * sizeof(float));
            guess_from_tree(g, guess);
            break;
        }
    long evals = 0, searches = 0, iters = 0;
    int bad = 0;
    #pragma omp parallel reduction(+ : evals, searches, iters, bad)
    {
        int *list = malloc(NGBMAX * sizeof(int));
        #pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < n; i++) {
            float h = hsml[i] == 0 ? guess[i] : hsml[i];
            float drho = 0, r = 0;
            int done = 0;
            for (int guard = 0; guard < 4096 && !done; guard++) {
                const int cnt = find_ngb(g, pos, i, h, list);
                searches++;
                if (cnt == NGBMAX) { h /= 1.24; continue; }
                if (cnt < DESNNGB) { h *= 1.23; continue; }
                done = find_hsml(s, pos, i, list, cnt, &drho, &h, &r, &evals, &iters);
            }
            if (!done) bad++;
            hsml[i] = h;
            rho[i] = r;
            varhsml[i] = 1.0 / (1 + h / (3 * r) * drho);          /* sph.c:66 */
        }
        free(list);
    }
    if (stats) { stats[0] += evals; stats[1] += searches; stats[2] += iters; }
    free(guess);
    tree_free(g);
    return bad;
}

/* ------------------------------------------------------------------ density model */

static float density_model(const to_sys *s, const float *p)       /* wvt_relax.c:227-256 */
{
    const double boxhalf = s->box * 0.5;
    double rho = 0;
    for (int i = 0; i < s->nhalos; i++) {
        const double *h = s->halos + 9 * i;
        if (h[8] == 0) continue;
        const double dx = p[0] - h[0] - boxhalf, dy = p[1] - h[1] - boxhalf, dz = p[2] - h[2] - boxhalf;
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        const double q = r / h[5], c = r / h[6];
        const double rho_i = h[3] * pow(1 + q * q, -3.0 / 2.0 * h[4]) / (1 + c * c * c * c);   /* setup.c:601 */
        rho = fmax(rho_i, rho);
    }
    return rho;
}

void to_density_model(const to_sys *s, const float *pos, float *out)
{
    for (int i = 0; i < s->n; i++) out[i] = density_model(s, pos + 3 * i);
}

/* ------------------------------------------------------------------ one WVT iteration */

/* wvt_relax.c:66-214 for an already SORTED particle set whose density pass has been done
 * (rho = SphP.Rho).  Outputs the scratch arrays and moves pos in place. */
void to_wvt_displace(const to_sys *s, float *pos, const float *rho, double step,
                     double *err_max, double *err_mean, float *rho_model, float *hsml_wvt,
                     float *delta /* n x 3 */, long *stats)
{
    const int n = s->n;
    const double box = s->box, boxinv = 1 / box;
    double emax = 0, emean = 0, vsum = 0;

    for (int i = 0; i < n; i++) {                                  /* wvt_relax.c:73-87 */
        const float rm = density_model(s, pos + 3 * i);
        const float err = fabs(rho[i] - rm) / rm;
        emax = fmax(err, emax);
        emean += err;
    }
    *err_max = emax;
    *err_mean = emean / n;

    for (int i = 0; i < n; i++) {                                  /* wvt_relax.c:108-118 */
        const float rm = density_model(s, pos + 3 * i);
        rho_model[i] = rm;
        hsml_wvt[i] = pow(DESNNGB * s->mpart / rm / FOURPITHIRD, 1. / 3.);
        vsum += hsml_wvt[i] * hsml_wvt[i] * hsml_wvt[i];
    }
    const float norm = pow(DESNNGB / vsum / FOURPITHIRD, 1.0 / 3.0);
    for (int i = 0; i < n; i++) hsml_wvt[i] *= norm;

    tree *g = tree_build(n, pos, box);
    long pairs = 0, searches = 0;
    #pragma omp parallel reduction(+ : pairs, searches)
    {
        int *list = malloc(NGBMAX * sizeof(int));
        #pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < n; i++) {                              /* wvt_relax.c:128-171 */
            float d[3] = {0, 0, 0};
            const int cnt = find_ngb(g, pos, i, hsml_wvt[i] * box, list);
            searches++;
            for (int k = 0; k < cnt; k++) {
                const int j = list[k];
                if (j == i) continue;
                float dx = (pos[3 * i] - pos[3 * j]) * boxinv;
                float dy = (pos[3 * i + 1] - pos[3 * j + 1]) * boxinv;
                float dz = (pos[3 * i + 2] - pos[3 * j + 2]) * boxinv;
                dx = dx > 0.5 ? dx - 1 : dx;
                dy = dy > 0.5 ? dy - 1 : dy;
                dz = dz > 0.5 ? dz - 1 : dz;
                dx = dx < -0.5 ? dx + 1 : dx;
                dy = dy < -0.5 ? dy + 1 : dy;
                dz = dz < -0.5 ? dz + 1 : dz;
                const float r2 = dx * dx + dy * dy + dz * dz;
                const float h = 0.5 * (hsml_wvt[i] + hsml_wvt[j]);
                if (r2 > h * h) continue;
                const float r = sqrt(r2);
                const float wk = wc6_unnormalised(r, h);
                d[0] += step * hsml_wvt[i] * wk * dx / r;
                d[1] += step * hsml_wvt[i] * wk * dy / r;
                d[2] += step * hsml_wvt[i] * wk * dz / r;
                pairs++;
            }
            delta[3 * i] = d[0];
            delta[3 * i + 1] = d[1];
            delta[3 * i + 2] = d[2];
        }
        free(list);
    }
    tree_free(g);
    if (stats) { stats[0] += pairs; stats[1] += searches; }

    for (int i = 0; i < n; i++)                                    /* wvt_relax.c:193-213 */
        for (int a = 0; a < 3; a++) {
            float x = pos[3 * i + a];
            x += (float)(delta[3 * i + a] * box);
            while (x < 0) x += box;
            while (x > box) x -= box;
            pos[3 * i + a] = x;
        }
}

/* ------------------------------------------------------------------ rot(A) */

void to_bfld_from_rotA(const to_sys *s, const float *pos, const float *hsml, const float *rho,
                       const float *varhsml, const float *apot, float *bfld)   /* sph.c:216-300 */
{
    const int n = s->n;
    const double boxhalf = s->box / 2, box = s->box;
    tree *g = tree_build(n, pos, box);
    #pragma omp parallel
    {
        int *list = malloc(NGBMAX * sizeof(int));
        #pragma omp for schedule(dynamic, 64)
        for (int i = 0; i < n; i++) {
            const int cnt = find_ngb(g, pos, i, hsml[i], list);
            const double vf = varhsml[i], h = hsml[i], rho_i = rho[i];
            double b[3] = {0, 0, 0};
            for (int k = 0; k < cnt; k++) {
                const int j = list[k];
                if (j == i) continue;
                double d[3];
                for (int a = 0; a < 3; a++) {
                    d[a] = (double)pos[3 * i + a] - pos[3 * j + a];
                    if (d[a] > boxhalf) d[a] -= box;
                    if (d[a] < -boxhalf) d[a] += box;
                }
                const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                if (r2 > h * h) continue;
                const double r = sqrt(r2);
                const double dwk = wc6_deriv(r, h);
                const double wgt = -s->mpart / rho_i * dwk / r * vf;
                const double dA[3] = {(double)apot[3 * i] - apot[3 * j],
                                      (double)apot[3 * i + 1] - apot[3 * j + 1],
                                      (double)apot[3 * i + 2] - apot[3 * j + 2]};
                b[0] += wgt * (d[2] * dA[1] - d[1] * dA[2]);
                b[1] += wgt * (d[0] * dA[2] - d[2] * dA[0]);
                b[2] += wgt * (d[1] * dA[0] - d[0] * dA[1]);
            }
            bfld[3 * i] = (float)b[0];
            bfld[3 * i + 1] = (float)b[1];
            bfld[3 * i + 2] = (float)b[2];
        }
        free(list);
    }
    tree_free(g);
}

/* ------------------------------------------------------------------ reorder helper */

/* peano.c:85-126 moves whole records; here: gather float arrays of `width` columns. */
void to_gather(int n, const int *perm, const float *in, float *out, int width)
{
    for (int k = 0; k < n; k++)
        memcpy(out + (size_t)k * width, in + (size_t)perm[k] * width, width * sizeof(float));
}
