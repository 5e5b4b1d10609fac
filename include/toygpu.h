/*
 * toygpu.h -- C ABI of libtoygpu.so, the B200 (sm_100a) replacement for Toycluster's
 * SPH-density + WVT-relaxation hot path.
 *
 * The reference has no plugin/FFI layer: its de-facto operator API is three global-state
 * procedures called from main() (main.c:52-56):
 *
 *     void Regularise_sph_particles();   wvt_relax.c:25
 *     void Find_sph_quantities();        sph.c:13
 *     void Bfld_from_rotA_SPH();         sph.c:216
 *
 * toycluster_b200/host/gpu_shim.c defines exactly those three symbols on top of the entry
 * points below.  Every function returns 0 on success and a negative
 * TG_E* code on failure; tg_last_error() gives the message.  All pointers are caller-owned
 * HOST memory unless a name says "dev"; the context owns all device memory.  One caller
 * thread per context, as in the reference (global state, non-reentrant).
 */
#ifndef TOYGPU_H
#define TOYGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef TG_CUBIC_SPLINE   /* libtoygpu_m4.so == the reference built with -DSPH_CUBIC_SPLINE */
#define TG_DESNNGB 50    /* globals.h:42 */
#define TG_NGBMAX  400   /* globals.h:44 */
#else
#define TG_DESNNGB 295   /* globals.h:48 */
#define TG_NGBMAX  2360  /* globals.h:50 */
#endif
#define TG_NUMITER 64    /* wvt_relax.c:7 */

enum {
    TG_OK = 0,
    TG_EINVAL = -1,    /* bad argument / call order */
    TG_ECUDA = -2,     /* CUDA runtime error */
    TG_ENOMEM = -3,
    TG_ENOCONV = -4,   /* hsml iteration did not terminate (sph.c:36-64 would spin) */
    TG_ERANGE = -5     /* position outside [0, Boxsize] (peano.c:130-132 asserts) */
};

/* tg_config.flags */
#define TG_WVT_SEQUENTIAL 1u  /* accumulate the displacement in the reference's order and
                                 precision (float += double, ascending neighbour index,
                                 wvt_relax.c:167-169) instead of one FP64 tree sum */

#define TG_EXACT_NEIGHBOURS 2u /* do NOT reproduce the reference octree's displaced nodes:
                                 Find_ngb_tree (tree.c:25-111) places a node's centre by the
                                 sign test of tree.c:298-302, which disagrees with the key cell
                                 for a particle lying exactly on a centre plane; the displaced
                                 subtree is then pruned for targets within reach (~8 nodes,
                                 ~400 affected targets per 1e6 particles).  Default: reproduced
                                 bit for bit.  With this flag every search returns the exact
                                 predicate set (== Find_ngb_simple, wvt_relax.c:296-340) */

#define TG_FAST 4u             /* FP32 kernel arithmetic in the warm-start sweep (tile_fast.cuh):
                                 neighbour sets, the frozen-list rule and the control flow of
                                 Find_hsml stay exactly the reference's; r, u = r/h and the WC6
                                 polynomials are evaluated in float (packed FP32), sums are
                                 float per lane and FP64 across lanes.  rho, hsml, VarHsmlFac and
                                 the displacement then agree with the reference to ~1e-7 except
                                 where a convergence decision (sph.c:161) flips within that
                                 noise: <= 5e-5 for ~1e-4 of the particles (north_star asks for
                                 1e-5 per iteration as a distribution).  Excludes
                                 TG_WVT_SEQUENTIAL. */

/* One row of the table Global_density_model() walks (wvt_relax.c:227-256):
 * Halo[i].{D_CoM, Rho0, Beta, Rcore, Rcut, Have_Cuspy, Mass[0]} (globals.h:128-157). */
typedef struct {
    double dcom[3];
    double rho0, beta, rcore, rcut;
    int cuspy;
    double mass_gas;
} tg_halo;

typedef struct {
    int device;          /* CUDA device ordinal */
    int n_gas;           /* Param.Npart[0] */
    double boxsize;      /* Param.Boxsize */
    double mpart_gas;    /* Param.Mpart[0] */
    double mtotal;       /* Param.Mtotal (wvt_relax.c:53) */
    unsigned flags;
    int rank, nranks;    /* target-particle partition; 0,1 for one GPU */
    void *stream;        /* cudaStream_t to run on (e.g. the caller's NCCL stream); NULL = own */
    int ngpus;           /* > 1: ONE process drives ngpus devices: the context is a
                            group of rank contexts with their own NCCL communicators; every call
                            below then acts on all of them (rank/nranks must be 0/0 or 0/1) */
    const int *devices;  /* ngpus device ordinals, or NULL for 0 .. ngpus-1 */
    double rho0_fac, rc_fac; /* Param.Rho0_Fac, Param.Rc_Fac of a -DDOUBLE_BETA_COOL_CORES build
                            (setup.c:604-612): both > 0 => halos with tg_halo.cuspy get the second,
                            cool-core beta component; both 0 => the default build, which ignores
                            Have_Cuspy in Gas_density_profile */
} tg_config;

typedef struct tg_ctx tg_ctx;

/* Per-iteration observer == the printf at wvt_relax.c:91-92. Return non-zero to stop. */
typedef int (*tg_log_fn)(int it, double err_max, double err_mean, double err_diff,
                         double step, void *user);

/* Counters of the last tg_find_sph_quantities / WVT iteration (bench.py's metric). */
typedef struct {
    unsigned long long pair_evals;   /* sum over Find_hsml iterations of list length
                                        + WVT pairs (reference semantics) */
    unsigned long long gathered;     /* G_step: distinct neighbours gathered */
    unsigned long long searches;     /* neighbour searches (tree.c:25 equivalents) */
    unsigned long long hsml_iters;   /* Find_hsml inner iterations (sph.c:96) */
    unsigned long long kernels;      /* kernel launches */
    double sweep_ms;                 /* device time of the neighbour sweep kernel(s) */
    double step_ms;                  /* device time of the whole step */
    unsigned long long handed_back;  /* targets the tile sweep returned to the generic sweep
                                        (last sweep launch of the step) */
    unsigned long long displaced_nodes;      /* reference-tree nodes displaced by tree.c:298-302
                                                in the current index (superset count) */
    unsigned long long displaced_particles;  /* particles underneath them (carry a path) */
    unsigned long long displaced_overflow;   /* 1: path table full, some left unreproduced */
    unsigned long long handback_why[5];      /* hand-backs of the step by reason: whole tile (cold
                                                or too many candidate runs), hit list full,
                                                displaced-node candidate in reach, a third
                                                search or density list full, no convergence on
                                                the frozen list */
    double index_ms;                 /* device time of keys + sort + index + model pass: the part of
                                        a step every rank repeats for all n */
    double tail_ms;                  /* device time after the sweep: error sums, move, exchange */
} tg_stats;

/* ---- life cycle -------------------------------------------------------------------- */
int tg_create(tg_ctx **out, const tg_config *cfg);
int tg_destroy(tg_ctx *ctx);
const char *tg_last_error(const tg_ctx *ctx);    /* ctx may be NULL (last create error) */

/* One process per GPU (rank / nranks in tg_config): rank 0 obtains a 128-byte NCCL id, the
 * host ships it to the other ranks by whatever channel it has, every rank joins.  From then on
 * the library itself exchanges what the ranks need -- the all-gather of the moved (x, y, z,
 * Hsml) slices, Rho / VarHsmlFac / Bfld, the driver's records, and the reductions of the
 * error statistics and status flags -- on its own stream, and tg_regularise,
 * tg_wvt_begin (which then returns GLOBAL sums) and tg_upload / tg_download work as on one
 * GPU.  tg_upload* read, and tg_download writes, this rank's slice [lo, hi) of the host arrays. */
int tg_comm_id(unsigned char id128[128]);
int tg_comm_init(tg_ctx *ctx, const unsigned char id128[128]);

/* Halo table for Global_density_model (replaces reading Halo[] / Param.Nhalos). */
int tg_set_halos(tg_ctx *ctx, int n, const tg_halo *halos);

/* ---- data in / out ----------------------------------------------------------------- */
/* AoS records exactly as the driver holds them: struct ParticleData (globals.h:161-168,
 * Pos at +0) and struct GasParticleData (globals.h:170-180, Hsml at +8, Apot at +28).
 * Only the gas range [0, n_gas) is read. */
int tg_upload(tg_ctx *ctx, const void *P, size_t p_stride, const void *SphP, size_t s_stride);
/* Page-lock a host range the caller will pass to tg_upload / tg_download again and again (the
 * driver's P and SphP live for the whole process, setup.c:244-250) so the copies are plain DMA.
 * Optional; ranges are released by tg_unpin_host or tg_destroy and must outlive that. */
int tg_pin_host(tg_ctx *ctx, void *ptr, size_t bytes);
int tg_unpin_host(tg_ctx *ctx, void *ptr);
/* Same from plain arrays: pos[n][3]; hsml[n] or NULL (cold start, SphP.Hsml == 0). */
int tg_upload_soa(tg_ctx *ctx, const float *pos, const float *hsml);
int tg_set_apot(tg_ctx *ctx, const float *apot /* [n][3], upload order */);
/* Multi-rank variants: pos/hsml are the FULL arrays but only this rank's slice crosses PCIe.
 * After tg_upload_soa_slice the host all-gathers pos_hsml_dev (tg_get_exchange) and passes
 * the OR of every rank's *cold to tg_set_cold. */
int tg_upload_soa_slice(tg_ctx *ctx, const float *pos, const float *hsml, int *cold);
int tg_set_cold(tg_ctx *ctx, int any_cold);
int tg_download_soa_slice(tg_ctx *ctx, float *pos, float *hsml);

/* Writes the path's post-state back into the driver's records: the gas range of P and
 * SphP permuted into the Peano order of the last density call (peano.c:85-126 moves whole
 * records), with Pos, Key, Tree_Parent, Hsml, Rho, VarHsmlFac, Rho_Model, Bfld updated. */
int tg_download(tg_ctx *ctx, void *P, size_t p_stride, void *SphP, size_t s_stride);
/* SoA read-back in the current (Peano) order; any pointer may be NULL.
 * perm[k] = upload index of the particle now at k. */
int tg_download_soa(tg_ctx *ctx, float *pos, int32_t *perm, float *hsml, float *rho,
                    float *varhsml, float *rho_model, float *bfld);

/* ---- the three operators ----------------------------------------------------------- */
int tg_find_sph_quantities(tg_ctx *ctx);                       /* == sph.c:13       */
int tg_regularise(tg_ctx *ctx, int max_iters, tg_log_fn log, void *user,
                  int *iters_done);                            /* == wvt_relax.c:25 */
int tg_bfld_from_rotA(tg_ctx *ctx);                            /* == sph.c:216      */

/* One pass of wvt_relax.c:66-214 with the step given by the caller (bench.py's "step",
 * and the per-iteration parity tests).  Runs sort + index + density + error + model hsml +
 * displacement + move; err_max/err_mean as at wvt_relax.c:73-87. */
int tg_wvt_iteration(tg_ctx *ctx, double step, double *err_max, double *err_mean);
/* The same pass in two halves, for hosts that own the control flow of wvt_relax.c:89-104
 * (tg_regularise itself, and multi-rank hosts that must all-reduce the error statistics):
 * begin = sort + index + density (+ displacement with step_guess unless TG_WVT_SEQUENTIAL)
 * and the error sum / max / count of this rank's slice; finish = displacement with the final
 * step (a rescale when it was already computed) + move, or with step_final <= 0 the
 * reference's `break`: positions stay, the new Hsml is kept. */
int tg_wvt_begin(tg_ctx *ctx, double step_guess, double *err_sum, double *err_max, int *count);
int tg_wvt_finish(tg_ctx *ctx, double step_final);
/* With a communicator tg_wvt_finish exchanges the moved (x, y, z, Hsml) slices only; Rho and
 * VarHsmlFac of the other ranks' slices are not needed by the next iteration.  This collective
 * brings them in (tg_regularise and tg_wvt_iteration end with it; call it yourself before a
 * download when you drive begin / finish).  No-op on one GPU. */
int tg_sync_results(tg_ctx *ctx);
/* Scratch of the last iteration (wvt_relax.c:36-44), in that iteration's Peano order. */
int tg_wvt_scratch(tg_ctx *ctx, float *hsml_wvt, float *delta /* [n][3] */);

int tg_get_stats(tg_ctx *ctx, tg_stats *out);

/* Make_magnetic_field() == magnetic_field.c:12-131 in one call: vector potential
 * A = max_halos (rho_gas / Rho0)^eta per particle (:33-69), B = rot A (the sweep of
 * tg_bfld_from_rotA), normalisation to bfld_norm / sqrt(3) at the field maximum and the cap at
 * bmax_main (bmax_sub for particles Halo_containing() puts into a halo with index > 1)
 * (:71-131).  Needs the index of the last tg_find_sph_quantities.  The arrays have one entry per
 * row of tg_set_halos: Halo[j].R_Sample[0], Halo[j].R_Sample[1], Halo[j].Is_Stripped. */
typedef struct {
    double bfld_norm, bfld_eta;          /* Param.Bfld_Norm, Param.Bfld_Eta */
    double bmax_main, bmax_sub;          /* BMAX = 18e-6 and 2e-6 (magnetic_field.c:4,113) */
    int sub_first;                       /* Sub.First */
    const double *r_sample_gas;
    const double *r_sample_dm;
    const int *is_stripped;
} tg_bfield;
int tg_make_magnetic_field(tg_ctx *ctx, const tg_bfield *par, double *norm_out, int *n_limited_out);
/* Reassign_particles_to_halos(), the per-particle half (positions.c:264-283):
 * ids[k] = Halo_containing(gas, Pos_k - Boxsize/2) for the current state, npart[j] = particles
 * of halo j (may be NULL).  Uses r_sample_gas, is_stripped and sub_first of `par`.  The index
 * sort of the ids and the record permutation (positions.c:405-443) stay with the caller: the
 * order of equal ids is whatever the reference's unstable gsl_heapsort_index makes of them. */
int tg_halo_ids(tg_ctx *ctx, const tg_bfield *par, int32_t *ids, long long *npart);
/* Apot of the current order, apot[n][3] (what tg_make_magnetic_field or tg_set_apot left). */
int tg_get_apot(tg_ctx *ctx, float *apot);

/* Gadget block writer (io.c:85-133 add_block / fill_write_buffer): the gas part of
 * a block's write buffer straight from the device's SoA arrays -- no AoS round trip and no
 * per-particle host loop.  tg_set_output_order: order[k] = index IN THE CURRENT DEVICE ORDER of
 * the particle the file holds at position k (what sort_particles(), positions.c:405-443, makes
 * of the records after the path; NULL = the device order itself).  tg_fill_block writes
 * n_gas * {3, 1, 1, 3, 1} floats of TG_BLOCK_{POS, RHO, HSML, BFLD, RHOMODEL} to `out`
 * (io.c:141-166: P.Pos, SphP.Rho, SphP.Hsml, SphP.Bfld, SphP.Rho_Model as float).  An order is
 * only valid for the device order it was given in: after an upload or any operator that sorts
 * again tg_fill_block fails (TG_EINVAL) until tg_set_output_order is called again. */
enum { TG_BLOCK_POS = 0, TG_BLOCK_RHO = 1, TG_BLOCK_HSML = 2, TG_BLOCK_BFLD = 3, TG_BLOCK_RHOMODEL = 4 };
int tg_set_output_order(tg_ctx *ctx, const size_t *order /* [n_gas] or NULL */);
int tg_fill_block(tg_ctx *ctx, int block, float *out);

/* ---- test hooks (parity with peano.c / sort.c / tree.c) ----------------------------- */
/* Peano_Key of every uploaded particle, upload order (peano.c:63-71). */
int tg_peano_keys(tg_ctx *ctx, uint64_t *hi, uint64_t *lo);
/* Key build + sort + reorder only (peano.c:46-81); perm as in tg_download_soa. */
int tg_sort(tg_ctx *ctx, int32_t *perm);
/* Find_ngb_tree(i, h) on the current index (tree.c:25): ascending, at most TG_NGBMAX. */
int tg_find_ngb(tg_ctx *ctx, int i, float h, int32_t *list, int *count);
/* 2*Guess_hsml(i) for every particle of the current order (tree.c:113, sph.c:26). */
int tg_guess_hsml(tg_ctx *ctx, float *out);

/* ---- multi-GPU plumbing (one process per GPU) ---------------------------- */
/* Device pointers of the arrays a rank must exchange after computing its slice
 * [lo, hi) of targets; the host does the all-gather (NCCL via torch.distributed). */
typedef struct {
    void *pos_hsml_dev;   /* float4[n]: x, y, z, Hsml of the current order */
    void *rho_dev;        /* float[n]  */
    void *varhsml_dev;    /* float[n]  */
    void *delta_dev;      /* float[3][n] */
    void *err_dev;        /* double[2]: sum err, max err of the local slice */
    int lo, hi;           /* this rank's targets: [rank*chunk, min(n, (rank+1)*chunk)) */
    int chunk;            /* pos_hsml/rho/varhsml are allocated nranks*chunk elements long, so
                             an in-place all-gather of `chunk` elements per rank fills them */
} tg_exchange;
int tg_get_exchange(tg_ctx *ctx, tg_exchange *out);

#ifdef __cplusplus
}
#endif
#endif /* TOYGPU_H */
