// toygpu.cu -- context, step orchestration and the C ABI of include/toygpu.h.
//
// Data layout in HBM (all SoA, n = gas particles):
//   state   posh  float4[n]  (x, y, z, Hsml) of the current order       16 B
//           id    int[n]     upload index of the particle               4 B
//   sort    key_hi/key_lo u64[n], idx int[n] (+ ping-pong copies)
//   sorted  pw    float4[n]  (x, y, z, raw WVT hsml; sign bit of w = displaced-node flag) --
//                            the one array the sweep gathers
//           soa   float[3][n8] x, y, z again, for the packed phase 1 of the tile sweep
//           hsml_in, rho_model float[n], key_hi_s/key_lo_s u64[n]
//   result  hsml_out, rho, varhsml float[n], delta float[3][n], bfld float[n][3]
//   index   6 float arrays of ~n/31 boxes (+ 4 sub-boxes per level-0 box), tile candidate lists
//   defect  dmap int[n], path nodes float4[max(2^20, n/2)] (defect.cuh)
// About 200 B per particle: 2 GB at 10 M, far inside 180 GB.
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/toygpu.h"
#include "common.cuh"
#include "peano.cuh"
#include "radix.cuh"
#include "bvh.cuh"
#include "model.cuh"
#include "guess.cuh"
#include "sph.cuh"
#include "tile.cuh"
#include "tile_fast.cuh"
#include "sph_fast.cuh"
#include "bfield.cuh"
#include "comm.cuh"
#include <thread>

static thread_local std::string g_create_error;

struct tg_ctx {
    tg_config cfg{};
    std::string err;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int n = 0, lo = 0, hi = 0, chunk = 0;
    Box box{};
    double bias_const = 0;
    double step_begin = 0;          // step the last tg_wvt_begin swept with

    // state (current order)
    float4 *posh = nullptr;
    int *id = nullptr;
    float *apot = nullptr;          // [n][3] or null
    float *stage = nullptr;         // [4n] host<->device staging (pos[n][3], hsml[n])
    bool have_apot = false;
    bool any_cold = true;
    bool index_valid = false;
    bool density_stale = false;     // multi-rank: Rho / VarHsmlFac of the other slices not gathered yet
    bool poisoned = false;          // a step failed half-way: ids and positions disagree until the next upload

    // sort
    uint64_t *key_hi = nullptr, *key_lo = nullptr, *key_tmp = nullptr;
    int *idx = nullptr, *idx_tmp = nullptr;
    unsigned *hist = nullptr;
    int ntiles = 0;
    uint64_t *key_hi_s = nullptr;   // points at key_hi or key_tmp after the last pass
    int *idx_s = nullptr;

    // sorted
    float4 *pw = nullptr;
    float *pwp = nullptr;           // pair-interleaved copy of pw (tile_fast.cuh)
    float *soa = nullptr;           // x[n8], y[n8], z[n8] copy of pw for the tile sweep's phase 1
    float *hsml_in = nullptr, *rho_model = nullptr;
    float *rm_state = nullptr, *rm_state_s = nullptr;   // SphP.Rho_Model as the driver sees it (current order)
    int *id_s = nullptr;
    uint64_t *key_lo_s = nullptr;
    float *apot_s = nullptr;

    // results
    float *hsml_out = nullptr, *rho = nullptr, *varh = nullptr, *delta = nullptr, *bfld = nullptr;

    // index
    Bvh bvh{};
    float *bvh_mem = nullptr, *sub_mem = nullptr;
    int bvh_total = 0;

    // cold start
    signed char *cpl = nullptr, *ev_level = nullptr, *ev_count = nullptr;
    int *ev_start = nullptr;
    float *guess = nullptr;

    // displaced reference-tree nodes (defect.cuh)
    DefectTab defect{};

    // scalars / scratch
    Halo *halos = nullptr;
    int nhalos = 0;
    double *partial = nullptr;      // block partials
    int npartial = 0;
    double *scal = nullptr;         // [0] vsum, [1] err sum, [2] err max
    int *flags = nullptr;           // [0] next, [1] status, [2] range_err, [3] n_tied, [4] cold, [5] nwork, [6] next (work list),
                                    // [7] stop (tg_regularise), [8] n_tied scratch, [9] more long tie runs than
                                    // RS_LONG_CAP, [10] long tie runs (k_fix_long_ties)
    int2 *tie_runs = nullptr;       // [RS_LONG_CAP] (first, end) of the long runs
    unsigned long long *counters = nullptr;   // 4
    double *gscratch = nullptr;
    int sweep_blocks = 0;
    int *tile_ng = nullptr, *tile_groups = nullptr, *worklist = nullptr;
    unsigned *tile_mask = nullptr;  // bit matrices of the tiles in flight (tile_fast.cuh), fast_blocks * TF_NB of them
    int tile_blocks = 0, fast_blocks = 0, sweep_fast_blocks = 0;
    bool use_tiles = true;

    // AoS records of the driver, resident on the device (tg_upload / tg_download)
    unsigned char *rawP = nullptr, *rawS = nullptr, *outP = nullptr, *outS = nullptr;
    size_t raw_pstride = 0, raw_sstride = 0;
    bool have_raw = false;
    std::vector<void *> pinned;     // host ranges registered by tg_pin_host

    // multi-GPU (comm.cuh): one communicator per rank context; a GROUP context (tg_config.ngpus
    // > 1) owns one rank context per device and runs every operator on all of them
    HaloExtra *halo_extra = nullptr;   // tg_make_magnetic_field scratch
    int *n_limited = nullptr;
    int *ngb_scratch = nullptr;        // tg_find_ngb
    unsigned long long *halo_counts = nullptr;   // tg_halo_ids
    int *out_order = nullptr;       // tg_set_output_order: file position -> current device index
    bool have_out_order = false;
    bool out_order_stale = false;   // the device order changed (upload / new index) since tg_set_output_order
    ncclComm_t comm = nullptr;
    double *errbuf = nullptr;       // [3 * nranks] gathered (err sum, err max, stop flag)
    double *hpin = nullptr;         // page-locked host words for the per-step read-backs (truly async copies)
    std::vector<tg_ctx *> kids;

    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // The displaced-node kernels (defect.cuh: one FP64-bound pass + latency-bound path walks of
    // a few hundred particles) run on a side stream next to the box hierarchy and the tile walk
    // and are joined before the first kernel that reads the flags / paths (join_defects).
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool defect_pending = false;
    tg_stats stats{};
    unsigned long long launches = 0;
};

static int fail(tg_ctx *c, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_error = buf;
    return code;
}

 // Run `fn` on every rank context of a group, one host thread per device (the operators block on
// their own stream; NCCL calls of different ranks must be in flight together).
template <class F> static int group_run(tg_ctx *g, F fn)
{
    const size_t m = g->kids.size();
    std::vector<int> rc(m, 0);
    std::vector<std::thread> th;
    for (size_t r = 1; r < m; r++) th.emplace_back([&, r] { rc[r] = fn(g->kids[r]); });
    rc[0] = fn(g->kids[0]);
    for (auto &t : th) t.join();
    for (size_t r = 0; r < m; r++)
        if (rc[r]) { g->err = "rank " + std::to_string(r) + ": " + g->kids[r]->err; return rc[r]; }
    return TG_OK;
}
#define TG_GROUP(c, call)                                                             \
    if ((c) && !(c)->kids.empty()) return group_run((c), [&](tg_ctx *k) { return call; })
#define TG_NOGROUP(c, name)                                                           \
    if ((c) && !(c)->kids.empty())                                                    \
        return fail((c), TG_EINVAL, name ": not available on a multi-GPU group context")
#define TG_GROUP0(c, call)                                                            \
    if ((c) && !(c)->kids.empty()) {                                                  \
        tg_ctx *k = (c)->kids[0];                                                     \
        const int rc_ = call;                                                         \
        if (rc_) (c)->err = k->err;                                                   \
        return rc_;                                                                   \
    }

#define CU(call)                                                                      \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess)                                                        \
            return fail(c, TG_ECUDA, "%s failed: %s (%s:%d)", #call,                  \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                  \
    } while (0)

#define LAUNCH_CHECK()                                                                \
    do {                                                                              \
        c->launches++;                                                                \
        cudaError_t e_ = cudaGetLastError();                                          \
        if (e_ != cudaSuccess)                                                        \
            return fail(c, TG_ECUDA, "kernel launch failed: %s (%s:%d)",              \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                  \
    } while (0)

template <class T> static cudaError_t dmalloc(T **p, size_t count)
{
    return cudaMalloc((void **)p, (count ? count : 1) * sizeof(T));
}

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------ collectives (comm.cuh)

#define NC(call)                                                                      \
    do {                                                                              \
        ncclResult_t r_ = (call);                                                     \
        if (r_ != ncclSuccess)                                                        \
            return fail(c, TG_ECUDA, "%s failed: %s (%s:%d)", #call,                  \
                        nccl_api()->GetErrorString(r_), __FILE__, __LINE__);          \
    } while (0)

// In-place all-gather of an array of nranks*chunk records of `rec_bytes` bytes, this rank's
// records already in place.  No-op without a communicator.
static int gather_slices(tg_ctx *c, void *base, size_t rec_bytes)
{
    if (!c->comm) return TG_OK;
    const size_t per_rank = (size_t)c->chunk * rec_bytes;
    NC(nccl_api()->AllGather((const char *)base + (size_t)c->cfg.rank * per_rank, base, per_rank, ncclInt8,
                             c->comm, c->stream));
    return TG_OK;
}

// max over ranks of `count` ints on the device (status / cold flags): every rank must take the
// same branch afterwards, or the next collective would hang.
static int reduce_flags_max(tg_ctx *c, int *dev, int count)
{
    if (!c->comm) return TG_OK;
    NC(nccl_api()->AllReduce(dev, dev, count, ncclInt32, ncclMax, c->comm, c->stream));
    return TG_OK;
}

// ------------------------------------------------------------------ life cycle

extern "C" const char *tg_last_error(const tg_ctx *c)
{
    return c ? c->err.c_str() : g_create_error.c_str();
}

extern "C" int tg_destroy(tg_ctx *c)
{
    if (!c) return TG_OK;
    if (!c->kids.empty()) {
        for (tg_ctx *k : c->kids) tg_destroy(k);
        delete c;
        return TG_OK;
    }
    cudaSetDevice(c->cfg.device);
    if (c->comm) nccl_api()->CommDestroy(c->comm);
    if (c->errbuf) cudaFree(c->errbuf);
    if (c->hpin) cudaFreeHost(c->hpin);
    if (c->halo_extra) cudaFree(c->halo_extra);
    if (c->n_limited) cudaFree(c->n_limited);
    if (c->ngb_scratch) cudaFree(c->ngb_scratch);
    if (c->halo_counts) cudaFree(c->halo_counts);
    if (c->out_order) cudaFree(c->out_order);
    if (c->tile_mask) cudaFree(c->tile_mask);
    if (c->tie_runs) cudaFree(c->tie_runs);
    void *ptrs[] = {c->posh, c->id, c->apot, c->stage, c->key_hi, c->key_lo, c->key_tmp, c->idx, c->idx_tmp,
                    c->hist, c->pw, c->pwp, c->soa, c->hsml_in, c->rho_model, c->rm_state, c->rm_state_s, c->id_s, c->key_lo_s, c->apot_s,
                    c->hsml_out, c->rho, c->varh, c->delta, c->bfld, c->bvh_mem, c->sub_mem, c->cpl,
                    c->ev_level, c->ev_count, c->ev_start, c->guess, c->halos, c->partial, c->scal,
                    c->flags, c->counters, c->gscratch, c->tile_ng, c->tile_groups, c->worklist,
                    c->defect.events, c->defect.big, c->defect.counts, c->defect.nodes, c->defect.dmap, c->defect.boxflag};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (void *p : {(void *)c->rawP, (void *)c->rawS, (void *)c->outP, (void *)c->outS}) if (p) cudaFree(p);
    for (void *h : c->pinned) cudaHostUnregister(h);
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->stream && c->own_stream) cudaStreamDestroy(c->stream);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    delete c;
    return TG_OK;
}

extern "C" int tg_create(tg_ctx **out, const tg_config *cfg)
{
    tg_ctx *c = nullptr;
    if (!out || !cfg) return fail(c, TG_EINVAL, "tg_create: null argument");
    *out = nullptr;
    if (cfg->n_gas <= 0 || !(cfg->boxsize > 0) || !(cfg->mpart_gas > 0))
        return fail(c, TG_EINVAL, "tg_create: n_gas, boxsize and mpart_gas must be positive");
    if ((cfg->flags & TG_FAST) && (cfg->flags & TG_WVT_SEQUENTIAL))
        return fail(c, TG_EINVAL, "tg_create: TG_FAST and TG_WVT_SEQUENTIAL exclude each other");
#ifdef TG_CUBIC_SPLINE
    if (cfg->flags & TG_FAST)
        return fail(c, TG_EINVAL, "tg_create: TG_FAST has WC6 polynomials only; the cubic-spline build has the exact modes");
#endif
    if (cfg->ngpus > 1) {
        // group context: one rank context per device + one communicator each (ncclCommInitAll)
        NcclApi *api = nccl_api();
        if (!api->ok) return fail(c, TG_EINVAL, "tg_create: ngpus = %d needs NCCL: %s", cfg->ngpus, api->why);
        if (cfg->nranks > 1) return fail(c, TG_EINVAL, "tg_create: ngpus and rank/nranks exclude each other");
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < cfg->ngpus)
            return fail(c, TG_EINVAL, "tg_create: ngpus = %d but %d CUDA device(s) visible", cfg->ngpus, ndev);
        tg_ctx *g = new tg_ctx;
        g->cfg = *cfg;
        g->n = cfg->n_gas;
        std::vector<int> devs(cfg->ngpus);
        for (int r = 0; r < cfg->ngpus; r++) devs[r] = cfg->devices ? cfg->devices[r] : r;
        for (int r = 0; r < cfg->ngpus; r++) {
            tg_config kc = *cfg;
            kc.ngpus = 0; kc.devices = nullptr; kc.stream = nullptr;
            kc.device = devs[r]; kc.rank = r; kc.nranks = cfg->ngpus;
            tg_ctx *k = nullptr;
            const int rc = tg_create(&k, &kc);
            if (rc) { tg_destroy(g); return rc; }
            g->kids.push_back(k);
        }
        std::vector<ncclComm_t> comms(cfg->ngpus);
        ncclResult_t nr = api->CommInitAll(comms.data(), cfg->ngpus, devs.data());
        if (nr != ncclSuccess) {
            tg_destroy(g);
            return fail(nullptr, TG_ECUDA, "ncclCommInitAll failed: %s", api->GetErrorString(nr));
        }
        for (int r = 0; r < cfg->ngpus; r++) {
            tg_ctx *k = g->kids[r];
            k->comm = comms[r];
            cudaSetDevice(k->cfg.device);
            if (cudaMalloc((void **)&k->errbuf, sizeof(double) * 3 * cfg->ngpus) != cudaSuccess) {
                tg_destroy(g);
                return fail(nullptr, TG_ENOMEM, "tg_create: device memory");
            }
        }
        *out = g;
        return TG_OK;
    }
    const int nranks = cfg->nranks > 0 ? cfg->nranks : 1;
    if (cfg->rank < 0 || cfg->rank >= nranks)
        return fail(c, TG_EINVAL, "tg_create: rank %d outside [0,%d)", cfg->rank, nranks);

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(c, TG_ECUDA, "tg_create: no CUDA device (%s); libtoygpu has no CPU path",
                    cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(c, TG_EINVAL, "tg_create: device %d of %d", cfg->device, ndev);

    tg_ctx *ctx = new tg_ctx;
    c = ctx;
    c->cfg = *cfg;
    c->cfg.nranks = nranks;
    const int n = c->n = cfg->n_gas;
    // equal-sized slices (the last may be short) so one all-gather of `chunk` elements per
    // rank reassembles an array; exchanged arrays are padded to nranks*chunk
    c->chunk = ((n + nranks - 1) / nranks + 31) / 32 * 32;   // whole 32-target tiles
    c->lo = std::min(n, cfg->rank * c->chunk);
    c->hi = std::min(n, c->lo + c->chunk);
    const size_t npad = (size_t)c->chunk * nranks;

    c->box.box_d = cfg->boxsize;
    c->box.boxhalf_d = 0.5 * cfg->boxsize;            // sph.c:83
    c->box.boxinv_d = 1 / cfg->boxsize;               // wvt_relax.c:29
    c->box.box_f = (float)cfg->boxsize;               // tree.c:27
    c->box.boxhalf_f = (float)(cfg->boxsize * 0.5);   // tree.c:28
    c->box.mpart = cfg->mpart_gas;
    c->bias_const = -0.0116 * pow(TG_DESNNGB * 0.01, -2.236);   // sph.c:206, host libm

#define CUC(call)                                                                     \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            fail(nullptr, TG_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_));  \
            tg_destroy(ctx);                                                          \
            return e_ == cudaErrorMemoryAllocation ? TG_ENOMEM : TG_ECUDA;            \
        }                                                                             \
    } while (0)

    CUC(cudaSetDevice(cfg->device));
    if (cfg->stream) c->stream = (cudaStream_t)cfg->stream;
    else { CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    for (auto &ev : c->ev) CUC(cudaEventCreate(&ev));
    CUC(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    CUC(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CUC(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));

    CUC(dmalloc(&c->posh, npad));
    CUC(dmalloc(&c->id, n));
    CUC(dmalloc(&c->stage, (size_t)4 * n));
    CUC(dmalloc(&c->key_hi, npad));      // (padded: all-gathered in place by sort_keys)
    CUC(dmalloc(&c->key_lo, npad));
    CUC(dmalloc(&c->key_tmp, n));
    CUC(dmalloc(&c->idx, n));
    CUC(dmalloc(&c->idx_tmp, n));
    c->ntiles = cdiv(n, RS_TILE);
    CUC(dmalloc(&c->hist, (size_t)RS_BINS * c->ntiles + RS_BINS));   // + digit totals
    CUC(dmalloc(&c->pw, ((size_t)n + 7) & ~(size_t)7));
    CUC(dmalloc(&c->pwp, 4 * (((size_t)n + 7) & ~(size_t)7) + 8));       // + the ghost pair (tile_fast.cuh)
    {   // ghost pair: finite and far away, h = 0; written once, k_reorder_model stops before it
        const float ghost[8] = {1e18f, 1e18f, 1e18f, 1e18f, 1e18f, 1e18f, 0.f, 0.f};
        CUC(cudaMemcpyAsync(c->pwp + 4 * (((size_t)n + 7) & ~(size_t)7), ghost, sizeof ghost, cudaMemcpyHostToDevice, c->stream));
        CUC(cudaStreamSynchronize(c->stream));
    }
    CUC(dmalloc(&c->soa, 3 * (((size_t)n + 7) & ~(size_t)7)));
    CUC(dmalloc(&c->hsml_in, n));
    CUC(dmalloc(&c->rho_model, n));
    CUC(dmalloc(&c->rm_state, n));
    CUC(dmalloc(&c->rm_state_s, n));
    CUC(cudaMemsetAsync(c->rm_state, 0, sizeof(float) * n, c->stream));
    CUC(dmalloc(&c->id_s, n));
    CUC(dmalloc(&c->key_lo_s, n));
    CUC(dmalloc(&c->hsml_out, n));
    CUC(dmalloc(&c->rho, npad));
    CUC(dmalloc(&c->varh, npad));
    CUC(dmalloc(&c->delta, (size_t)3 * n));
    CUC(dmalloc(&c->bfld, (size_t)3 * npad));
    CUC(cudaMemsetAsync(c->rho, 0, sizeof(float) * n, c->stream));
    CUC(cudaMemsetAsync(c->varh, 0, sizeof(float) * n, c->stream));
    CUC(cudaMemsetAsync(c->rho_model, 0, sizeof(float) * n, c->stream));
    CUC(cudaMemsetAsync(c->delta, 0, sizeof(float) * 3 * n, c->stream));
    CUC(cudaMemsetAsync(c->bfld, 0, sizeof(float) * 3 * npad, c->stream));

    // index levels
    Bvh &t = c->bvh;
    t.n = n;
    int cnt = cdiv(n, 32), lvl = 0, total = 0;
    for (;;) {
        if (lvl >= MAX_LEVELS) { tg_destroy(ctx); return fail(nullptr, TG_EINVAL, "too many index levels"); }
        t.lvl_n[lvl] = cnt;
        t.lvl_off[lvl] = total;
        total += cnt;
        if (cnt <= 32) break;
        cnt = cdiv(cnt, 32);
        lvl++;
    }
    t.top = lvl;
    c->bvh_total = total;
    CUC(dmalloc(&c->bvh_mem, (size_t)6 * total));
    t.cx = c->bvh_mem; t.cy = c->bvh_mem + total; t.cz = c->bvh_mem + 2 * (size_t)total;
    t.hx = c->bvh_mem + 3 * (size_t)total; t.hy = c->bvh_mem + 4 * (size_t)total;
    t.hz = c->bvh_mem + 5 * (size_t)total;
    {
        const size_t ns = (size_t)4 * t.lvl_n[0];
        CUC(dmalloc(&c->sub_mem, 6 * ns));
        t.scx = c->sub_mem; t.scy = c->sub_mem + ns; t.scz = c->sub_mem + 2 * ns;
        t.shx = c->sub_mem + 3 * ns; t.shy = c->sub_mem + 4 * ns; t.shz = c->sub_mem + 5 * ns;
    }

    CUC(dmalloc(&c->cpl, n));
    CUC(dmalloc(&c->ev_level, n));
    CUC(dmalloc(&c->ev_count, n));
    CUC(dmalloc(&c->ev_start, n));
    CUC(dmalloc(&c->guess, n));
    c->defect.cap_events = std::max(4096, n / 4);
    c->defect.cap_nodes = std::max(1 << 20, n / 2);
    CUC(dmalloc(&c->defect.events, c->defect.cap_events));
    CUC(dmalloc(&c->defect.counts, 8));
    CUC(dmalloc(&c->defect.big, DF_BIG_CAP));
    CUC(dmalloc(&c->defect.nodes, c->defect.cap_nodes));
    CUC(dmalloc(&c->defect.dmap, n));
    c->defect.pwp = c->pwp;
    CUC(dmalloc(&c->defect.boxflag, (size_t)n / 32 + 1));
    CUC(cudaMemsetAsync(c->defect.boxflag, 0, (size_t)n / 32 + 1, c->stream));
    CUC(cudaMemsetAsync(c->defect.counts, 0, 8 * sizeof(int), c->stream));

    CUC(dmalloc(&c->halos, MAX_HALOS));
    c->npartial = cdiv(n, RED_THREADS);
    CUC(dmalloc(&c->partial, (size_t)2 * c->npartial));
    CUC(dmalloc(&c->scal, 4));
    CUC(cudaHostAlloc((void **)&c->hpin, sizeof(double) * (2 * nranks + 8), cudaHostAllocDefault));
    CUC(dmalloc(&c->flags, 12));
    CUC(dmalloc(&c->tie_runs, RS_LONG_CAP));
    CUC(dmalloc(&c->counters, 12));
    CUC(cudaMemsetAsync(c->flags, 0, 12 * sizeof(int), c->stream));
    CUC(cudaMemsetAsync(c->counters, 0, 12 * sizeof(unsigned long long), c->stream));

    // sweep grid: every SM full of resident blocks (persistent, work-stealing)
    cudaDeviceProp prop;
    CUC(cudaGetDeviceProperties(&prop, cfg->device));
    const size_t smem = (size_t)SW_WARPS * SW_LCAP * sizeof(double);
    int per_sm = 1;
    {
        auto set = [&](const void *f) {
            return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        };
        CUC(set((const void *)k_sweep<MODE_DENSITY, false>));
        CUC(set((const void *)k_sweep<MODE_DENSITY | MODE_EXACT, false>));
        CUC(set((const void *)k_sweep<MODE_DENSITY | MODE_EXACT, true>));
        CUC(set((const void *)k_sweep<MODE_WVT, false>));
        CUC(set((const void *)k_sweep<MODE_WVT_SEQ, false>));
        CUC(set((const void *)k_sweep<MODE_DENSITY | MODE_WVT, false>));
        CUC(set((const void *)k_sweep<MODE_ROTA, false>));
        CUC(set((const void *)k_sweep<MODE_ROTA, true>));
        CUC(set((const void *)k_sweep<MODE_DENSITY, true>));
        CUC(set((const void *)k_sweep<MODE_WVT, true>));
        CUC(set((const void *)k_sweep<MODE_DENSITY | MODE_WVT, true>));
        CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep<MODE_DENSITY | MODE_WVT, false>,
                                                          SW_WARPS * 32, smem));
        if (per_sm < 1) per_sm = 1;
    }
    c->sweep_blocks = prop.multiProcessorCount * per_sm;
    {
        auto set = [&](const void *f) {
            return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TL_SMEM);
        };
        CUC(set((const void *)k_sweep_tile<MODE_DENSITY>));
        CUC(set((const void *)k_sweep_tile<MODE_DENSITY | MODE_EXACT>));
        CUC(set((const void *)k_sweep_tile<MODE_WVT>));
        CUC(set((const void *)k_sweep_tile<MODE_DENSITY | MODE_WVT>));
        int tile_per_sm = 1;
        CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&tile_per_sm, k_sweep_tile<MODE_DENSITY | MODE_WVT>,
                                                          TL_WARPS * 32, TL_SMEM));
        if (tile_per_sm < 1) tile_per_sm = 1;
        c->tile_blocks = prop.multiProcessorCount * tile_per_sm;
    }
    {
        auto set = [&](const void *f) {
            return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TF_SMEM);
        };
        CUC(set((const void *)k_sweep_tile_fast<MODE_DENSITY>));
        CUC(set((const void *)k_sweep_tile_fast<MODE_DENSITY | MODE_WVT>));
        CUC(set((const void *)k_sweep_tile_fast<MODE_ROTA>));
        int per = 1;
        CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_sweep_tile_fast<MODE_DENSITY | MODE_WVT>,
                                                          TF_WARPS * 32, TF_SMEM));
        if (per < 1) per = 1;
        c->fast_blocks = prop.multiProcessorCount * per;
        CUC(dmalloc(&c->tile_mask, (size_t)c->fast_blocks * TF_NB * TF_MASK_WORDS));
    }
    {
        auto set = [&](const void *f) {
            return cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SF_SMEM);
        };
        CUC(set((const void *)k_sweep_fast<MODE_DENSITY, false>));
        CUC(set((const void *)k_sweep_fast<MODE_DENSITY, true>));
        CUC(set((const void *)k_sweep_fast<MODE_DENSITY | MODE_WVT, false>));
        CUC(set((const void *)k_sweep_fast<MODE_DENSITY | MODE_WVT, true>));
        int per = 1;
        CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_sweep_fast<MODE_DENSITY | MODE_WVT, false>,
                                                          SF_WARPS * 32, SF_SMEM));
        if (per < 1) per = 1;
        c->sweep_fast_blocks = prop.multiProcessorCount * per;
    }
    CUC(dmalloc(&c->tile_ng, (size_t)t.lvl_n[0]));
    CUC(dmalloc(&c->tile_groups, (size_t)t.lvl_n[0] * TL_ENT));
    CUC(dmalloc(&c->worklist, (size_t)n));
    c->use_tiles = getenv("TOYGPU_NO_TILES") == nullptr;
    CUC(dmalloc(&c->gscratch, (size_t)c->sweep_blocks * SW_WARPS * TG_NGBMAX));
    CUC(cudaStreamSynchronize(c->stream));
#undef CUC
    *out = ctx;
    return TG_OK;
}

// One process per GPU: rank 0 calls tg_comm_id, ships the 128 bytes to the other ranks by
// whatever channel the host has (MPI, torch.distributed, a file), every rank calls tg_comm_init.
extern "C" int tg_comm_id(unsigned char *id128)
{
    tg_ctx *c = nullptr;
    NcclApi *api = nccl_api();
    if (!id128) return fail(c, TG_EINVAL, "tg_comm_id: null argument");
    if (!api->ok) return fail(c, TG_EINVAL, "tg_comm_id: NCCL unavailable: %s", api->why);
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    NC(api->GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return TG_OK;
}

extern "C" int tg_comm_init(tg_ctx *c, const unsigned char *id128)
{
    if (!c || !id128) return fail(c, TG_EINVAL, "tg_comm_init: null argument");
    if (!c->kids.empty() || c->comm) return fail(c, TG_EINVAL, "tg_comm_init: context already has a communicator");
    NcclApi *api = nccl_api();
    if (!api->ok) return fail(c, TG_EINVAL, "tg_comm_init: NCCL unavailable: %s", api->why);
    CU(cudaSetDevice(c->cfg.device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NC(api->CommInitRank(&c->comm, c->cfg.nranks, id, c->cfg.rank));
    CU(dmalloc(&c->errbuf, (size_t)3 * c->cfg.nranks));
    return TG_OK;
}

extern "C" int tg_set_halos(tg_ctx *c, int n, const tg_halo *h)
{
    TG_GROUP(c, tg_set_halos(k, n, h));
    if (!c || n < 0 || n > MAX_HALOS || (n && !h)) return fail(c, TG_EINVAL, "tg_set_halos: bad arguments (max %d rows)", MAX_HALOS);
    if (c->kids.empty() && (c->cfg.rho0_fac < 0 || c->cfg.rc_fac < 0 || (c->cfg.rho0_fac > 0) != (c->cfg.rc_fac > 0)))
        return fail(c, TG_EINVAL, "tg_config: rho0_fac and rc_fac must both be positive (cool cores) or both 0");
    CU(cudaSetDevice(c->cfg.device));
    std::vector<Halo> rows(n);
    for (int i = 0; i < n; i++) {
        rows[i].cx = h[i].dcom[0]; rows[i].cy = h[i].dcom[1]; rows[i].cz = h[i].dcom[2];
        rows[i].rho0 = h[i].rho0; rows[i].beta = h[i].beta;
        rows[i].rcore = h[i].rcore; rows[i].rcut = h[i].rcut;
        rows[i].mass_gas = h[i].mass_gas;
        // setup.c:604-612: only a -DDOUBLE_BETA_COOL_CORES build looks at Have_Cuspy; the caller
        // says which build it stands in for by giving (or not giving) the two factors
        const bool cc = h[i].cuspy && c->cfg.rho0_fac > 0 && c->cfg.rc_fac > 0;
        rows[i].rho0_cc = cc ? h[i].rho0 * c->cfg.rho0_fac : 0.0;
        rows[i].rc_cc = cc ? h[i].rcore / c->cfg.rc_fac : 1.0;
    }
    if (n) CU(cudaMemcpyAsync(c->halos, rows.data(), n * sizeof(Halo), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->nhalos = n;
    return TG_OK;
}

// ------------------------------------------------------------------ data in

// Host SoA -> device staging (plain cudaMemcpyAsync, DMA when the host buffer is pinned),
// then one kernel packs (x, y, z, Hsml) and writes the identity ids.
static int upload_common(tg_ctx *c, const float *pos, const float *hsml)
{
    const int n = c->n;
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaMemcpyAsync(c->stage, pos, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    if (hsml)
        CU(cudaMemcpyAsync(c->stage + 3 * (size_t)n, hsml, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->flags + 4, 0, sizeof(int), c->stream));
    k_pack_state<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->stage, hsml ? c->stage + 3 * (size_t)n : nullptr,
                                                     c->posh, c->id, c->flags + 4);
    LAUNCH_CHECK();
    int cold = 0;
    CU(cudaMemcpyAsync(&cold, c->flags + 4, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->any_cold = cold != 0;
    c->index_valid = false;
    c->poisoned = false;
    if (c->have_out_order) c->out_order_stale = true;
    c->have_apot = false;
    c->have_raw = false;
    CU(cudaMemsetAsync(c->rm_state, 0, sizeof(float) * n, c->stream));
    return TG_OK;
}

extern "C" int tg_upload_soa(tg_ctx *c, const float *pos, const float *hsml)
{
    if (!c || !pos) return fail(c, TG_EINVAL, "tg_upload_soa: null argument");
    TG_GROUP(c, tg_upload_soa(k, pos, hsml));
    if (c->comm) {       // every rank ships its own slice over PCIe; NVLink does the rest
        int rc = tg_upload_soa_slice(c, pos, hsml, nullptr);
        if (rc) return rc;
        if ((rc = gather_slices(c, c->posh, sizeof(float4)))) return rc;
        if ((rc = reduce_flags_max(c, c->flags + 4, 1))) return rc;
        int cold = 0;
        CU(cudaMemcpyAsync(&cold, c->flags + 4, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->any_cold = cold != 0;
        return TG_OK;
    }
    return upload_common(c, pos, hsml);
}

// Multi-rank upload: pos/hsml are the FULL host arrays, but only this rank's slice [lo, hi) is
// copied (PCIe carries n/nranks records per rank); the host then re-assembles the state buffer
// with the same all-gather it uses every step and agrees on the cold flag (tg_set_cold).
extern "C" int tg_upload_soa_slice(tg_ctx *c, const float *pos, const float *hsml, int *cold_out)
{
    if (!c || !pos) return fail(c, TG_EINVAL, "tg_upload_soa_slice: null argument");
    TG_NOGROUP(c, "tg_upload_soa_slice");
    const int n = c->n, lo = c->lo, m = c->hi - c->lo;
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaMemsetAsync(c->flags + 4, 0, sizeof(int), c->stream));
    if (m > 0) {
        CU(cudaMemcpyAsync(c->stage, pos + 3 * (size_t)lo, sizeof(float) * 3 * m, cudaMemcpyHostToDevice, c->stream));
        if (hsml)
            CU(cudaMemcpyAsync(c->stage + 3 * (size_t)n, hsml + lo, sizeof(float) * m, cudaMemcpyHostToDevice, c->stream));
        k_pack_state<<<cdiv(m, 256), 256, 0, c->stream>>>(m, c->stage, hsml ? c->stage + 3 * (size_t)n : nullptr,
                                                         c->posh + lo, c->id + lo, c->flags + 4);
        LAUNCH_CHECK();
    }
    k_iota<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->id);
    LAUNCH_CHECK();
    int cold = 0;
    CU(cudaMemcpyAsync(&cold, c->flags + 4, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->any_cold = cold != 0;
    c->index_valid = false;
    c->poisoned = false;
    if (c->have_out_order) c->out_order_stale = true;
    c->have_apot = false;
    c->have_raw = false;
    CU(cudaMemsetAsync(c->rm_state, 0, sizeof(float) * n, c->stream));
    if (cold_out) *cold_out = cold;
    return TG_OK;
}

extern "C" int tg_set_cold(tg_ctx *c, int any_cold)
{
    if (!c) return TG_EINVAL;
    TG_GROUP(c, tg_set_cold(k, any_cold));
    c->any_cold = any_cold != 0;
    return TG_OK;
}

// Read back this rank's slice of (pos, Hsml) into the FULL host arrays.
extern "C" int tg_download_soa_slice(tg_ctx *c, float *pos, float *hsml)
{
    if (!c) return TG_EINVAL;
    TG_GROUP(c, tg_download_soa_slice(k, pos, hsml));
    CU(cudaSetDevice(c->cfg.device));
    const int n = c->n, lo = c->lo, m = c->hi - c->lo;
    if (m > 0 && (pos || hsml)) {
        k_unpack_state<<<cdiv(m, 256), 256, 0, c->stream>>>(m, c->posh + lo, c->stage, c->stage + 3 * (size_t)n);
        LAUNCH_CHECK();
        if (pos) CU(cudaMemcpyAsync(pos + 3 * (size_t)lo, c->stage, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, c->stream));
        if (hsml) CU(cudaMemcpyAsync(hsml + lo, c->stage + 3 * (size_t)n, sizeof(float) * m, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return TG_OK;
}

// ---- the driver's AoS records (globals.h:161-180) ---------------------------------------
// The raw records cross the bus once per call (124 B per particle: 1.24 GB at 10 M) and are
// unpacked, permuted and patched ON THE DEVICE; the host does no per-particle work.

__global__ void k_unpack_records(int n, const unsigned char *__restrict__ rawP, size_t pstride,
                                 const unsigned char *__restrict__ rawS, size_t sstride,
                                 float4 *__restrict__ posh, int *__restrict__ id, float *__restrict__ apot,
                                 float *__restrict__ rm, int *__restrict__ cold_flag)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float *pp = (const float *)(rawP + (size_t)k * pstride);       // Pos @ +0
    const float *sp = (const float *)(rawS + (size_t)k * sstride);
    const float h = sp[2];                                               // Hsml @ +8
    posh[k] = make_float4(pp[0], pp[1], pp[2], h);
    id[k] = k;
    if (apot) { apot[3 * (size_t)k] = sp[7]; apot[3 * (size_t)k + 1] = sp[8]; apot[3 * (size_t)k + 2] = sp[9]; }   // Apot @ +28
    rm[k] = sstride >= 48 ? sp[11] : 0.f;                                // Rho_Model @ +44
    if (h == 0.f) *cold_flag = 1;
}

// out record k = in record id[k], in 4-byte words (peano.c:96-117 moves whole structs).
__global__ void k_gather_records(size_t total_words, int wpr, const int *__restrict__ id,
                                 const uint32_t *__restrict__ in, uint32_t *__restrict__ out)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_words) return;
    const size_t rec = t / wpr;
    const int w = (int)(t - rec * wpr);
    out[t] = in[(size_t)id[rec] * wpr + w];
}

// The fields the path owns, written over the permuted records.
__global__ void k_patch_records(int n, int first, unsigned char *__restrict__ outP, size_t pstride,
                                unsigned char *__restrict__ outS, size_t sstride,
                                const float4 *__restrict__ posh, const uint64_t *__restrict__ key_lo,
                                const uint64_t *__restrict__ key_hi, const float *__restrict__ rho,
                                const float *__restrict__ varh, const float *__restrict__ bfld,
                                const float *__restrict__ rm)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t *p = (uint32_t *)(outP + (size_t)k * pstride);
    float *g = (float *)(outS + (size_t)k * sstride);
    const float4 ph = posh[k];
    ((float *)p)[0] = ph.x; ((float *)p)[1] = ph.y; ((float *)p)[2] = ph.z;    // Pos @ +0
    if (key_lo) {
        const uint64_t lo = key_lo[k], hi = key_hi[k];
        p[8] = (uint32_t)lo; p[9] = (uint32_t)(lo >> 32);                  // Key @ +32 (little endian u128)
        p[10] = (uint32_t)hi; p[11] = (uint32_t)(hi >> 32);
        p[12] = (uint32_t)((first + k) / 32);                              // Tree_Parent @ +48: index group
    }                                                                      // (tree.c's node ids are private)
    g[1] = rho[k];                                                         // Rho @ +4
    g[2] = ph.w;                                                           // Hsml @ +8
    g[3] = varh[k];                                                        // VarHsmlFac @ +12
    g[4] = bfld[3 * (size_t)k]; g[5] = bfld[3 * (size_t)k + 1]; g[6] = bfld[3 * (size_t)k + 2];   // Bfld @ +16
    g[11] = rm[k];                                                         // Rho_Model @ +44
}

extern "C" int tg_pin_host(tg_ctx *c, void *ptr, size_t bytes)
{
    if (!c || !ptr || !bytes) return fail(c, TG_EINVAL, "tg_pin_host: bad arguments");
    TG_GROUP0(c, tg_pin_host(k, ptr, bytes));     // portable registration: valid on every device
    CU(cudaSetDevice(c->cfg.device));
    for (void *h : c->pinned) if (h == ptr) return TG_OK;
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    c->pinned.push_back(ptr);
    return TG_OK;
}

extern "C" int tg_unpin_host(tg_ctx *c, void *ptr)
{
    if (!c || !ptr) return TG_EINVAL;
    TG_GROUP0(c, tg_unpin_host(k, ptr));
    for (size_t k = 0; k < c->pinned.size(); k++)
        if (c->pinned[k] == ptr) {
            cudaHostUnregister(ptr);
            c->pinned.erase(c->pinned.begin() + k);
            return TG_OK;
        }
    return fail(c, TG_EINVAL, "tg_unpin_host: range was not pinned by tg_pin_host");
}

extern "C" int tg_upload(tg_ctx *c, const void *P, size_t p_stride, const void *SphP, size_t s_stride)
{
    if (!c || !P || !SphP || p_stride < 12 || s_stride < 12 || p_stride % 4 || s_stride % 4)
        return fail(c, TG_EINVAL, "tg_upload: bad arguments (strides must be multiples of 4 and hold Pos / Hsml)");
    TG_GROUP(c, tg_upload(k, P, p_stride, SphP, s_stride));
    if (c->cfg.nranks > 1 && !c->comm)
        return fail(c, TG_EINVAL, "tg_upload with nranks > 1 needs a communicator (tg_comm_init or tg_config.ngpus)");
    const int n = c->n;
    CU(cudaSetDevice(c->cfg.device));
    if (!c->rawP || c->raw_pstride != p_stride || c->raw_sstride != s_stride) {
        for (unsigned char **p : {&c->rawP, &c->rawS, &c->outP, &c->outS}) { if (*p) cudaFree(*p); *p = nullptr; }
        const size_t npad = (size_t)c->chunk * c->cfg.nranks;      // room for the in-place all-gather
        CU(dmalloc(&c->rawP, npad * p_stride));
        CU(dmalloc(&c->rawS, npad * s_stride));
        CU(dmalloc(&c->outP, npad * p_stride));
        CU(dmalloc(&c->outS, npad * s_stride));
        c->raw_pstride = p_stride;
        c->raw_sstride = s_stride;
    }
    const bool with_apot = s_stride >= 40;
    if (with_apot && !c->apot) {
        CU(dmalloc(&c->apot, (size_t)3 * n));
        CU(dmalloc(&c->apot_s, (size_t)3 * n));
    }
    // DMA when the host ranges are pinned (tg_pin_host), staged by the driver otherwise
    // (with a communicator: only this rank's slice crosses PCIe, NVLink all-gathers the rest)
    const size_t lo = c->comm ? c->lo : 0, m = c->comm ? c->hi - c->lo : n;
    if (m) {
        CU(cudaMemcpyAsync(c->rawP + lo * p_stride, (const char *)P + lo * p_stride, m * p_stride, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->rawS + lo * s_stride, (const char *)SphP + lo * s_stride, m * s_stride, cudaMemcpyHostToDevice, c->stream));
    }
    int rc;
    if ((rc = gather_slices(c, c->rawP, p_stride))) return rc;
    if ((rc = gather_slices(c, c->rawS, s_stride))) return rc;
    CU(cudaMemsetAsync(c->flags + 4, 0, sizeof(int), c->stream));
    k_unpack_records<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->rawP, p_stride, c->rawS, s_stride, c->posh, c->id,
                                                         with_apot ? c->apot : nullptr, c->rm_state, c->flags + 4);
    LAUNCH_CHECK();
    int cold = 0;
    CU(cudaMemcpyAsync(&cold, c->flags + 4, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->any_cold = cold != 0;
    c->index_valid = false;
    c->poisoned = false;
    if (c->have_out_order) c->out_order_stale = true;
    c->have_apot = with_apot;
    c->have_raw = true;
    return TG_OK;
}

// apot is given in the CURRENT order of the context (upload order right after an upload,
// Peano order after a density call -- the order tg_download_soa reports).
extern "C" int tg_set_apot(tg_ctx *c, const float *apot)
{
    if (!c || !apot) return fail(c, TG_EINVAL, "tg_set_apot: null argument");
    TG_GROUP(c, tg_set_apot(k, apot));
    CU(cudaSetDevice(c->cfg.device));
    if (!c->apot) {
        CU(dmalloc(&c->apot, (size_t)3 * c->n));
        CU(dmalloc(&c->apot_s, (size_t)3 * c->n));
    }
    CU(cudaMemcpyAsync(c->apot, apot, sizeof(float) * 3 * c->n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->have_apot = true;
    return TG_OK;
}

// ------------------------------------------------------------------ step pieces

static int reset_counters(tg_ctx *c)
{
    CU(cudaMemsetAsync(c->counters, 0, 12 * sizeof(unsigned long long), c->stream));
    CU(cudaMemsetAsync(c->flags + 1, 0, sizeof(int), c->stream));   // sweep status
    CU(cudaMemsetAsync(c->flags + 5, 0, sizeof(int), c->stream));   // handed-back count
    c->launches = 0;
    return TG_OK;
}

// The main stream waits for the displaced-node kernels of the current index (no-op when they
// ran in line or were joined already).
static int join_defects(tg_ctx *c)
{
    if (!c->defect_pending) return TG_OK;
    CU(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    c->defect_pending = false;
    return TG_OK;
}

// peano.c:46-81 (keys, index sort) -- leaves key_hi_s / idx_s.
static int sort_keys(tg_ctx *c)
{
    const int n = c->n, T = 256;
    CU(cudaMemsetAsync(c->flags + 2, 0, 2 * sizeof(int), c->stream));
    // (Every rank computes all keys: with the transducer table the kernel takes 0.19 ms at 10 M,
    // less than all-gathering 16 B per particle would.)
    k_peano_keys<<<cdiv(n, T), T, 0, c->stream>>>(n, c->posh, c->box.box_d, c->key_hi, c->key_lo,
                                                 c->idx, c->flags + 2);
    LAUNCH_CHECK();
    uint64_t *kin = c->key_hi, *kout = c->key_tmp;
    int *iin = c->idx, *iout = c->idx_tmp;
    // Only the top bits need radix passes: 8 tree levels beyond log8(n) leave runs of a few
    // keys at most, and k_fix_ties orders every run by the full 128-bit key anyway.
    int levels = 8;
    for (long long m = 1; m < n; m *= 8) levels++;
    int low_bits = 64 - std::min(64, (3 * levels + RS_BITS - 1) / RS_BITS * RS_BITS);
    if (getenv("TOYGPU_FULL_SORT")) low_bits = 0;
    for (int shift = low_bits; shift < 64; shift += RS_BITS) {
        k_radix_hist<<<c->ntiles, RS_THREADS, 0, c->stream>>>(n, kin, shift, c->ntiles, c->hist);
        LAUNCH_CHECK();
        unsigned *bin_total = c->hist + (size_t)RS_BINS * c->ntiles;
        k_radix_scan<<<RS_BINS, 1024, 0, c->stream>>>(c->ntiles, c->hist, bin_total);
        LAUNCH_CHECK();
        k_radix_scatter<<<c->ntiles, RS_THREADS, 0, c->stream>>>(n, kin, iin, kout, iout, shift,
                                                                c->ntiles, c->hist, bin_total);
        LAUNCH_CHECK();
        std::swap(kin, kout);
        std::swap(iin, iout);
    }
    c->key_hi_s = kin;
    c->idx_s = iin;
    CU(cudaMemsetAsync(c->flags + 8, 0, 3 * sizeof(int), c->stream));
    k_fix_ties<<<cdiv(n, T), T, 0, c->stream>>>(n, c->key_hi_s, c->idx_s, c->key_lo, low_bits, c->flags + 8,
                                               c->tie_runs);
    LAUNCH_CHECK();
    k_fix_long_ties<<<32, 1024, 0, c->stream>>>(c->key_hi_s, c->idx_s, c->key_lo, c->flags + 8, c->tie_runs);
    LAUNCH_CHECK();
    return TG_OK;
}

// Sort_Particles_By_Peano_Key + Build_Tree equivalents, plus the position-only model pass.
static int prepare_index(tg_ctx *c)
{
    const int n = c->n, T = 256;
    if (c->nhalos == 0) return fail(c, TG_EINVAL, "tg_set_halos has not been called");
    if (c->poisoned) return fail(c, TG_EINVAL, "the previous step failed half-way: upload the particles again");
    int rc = join_defects(c);            // (a previous index whose flags nobody read)
    if (rc) return rc;
    if (c->have_out_order) c->out_order_stale = true;        // a file order refers to the order it was given in
    if ((rc = sort_keys(c))) return rc;

    k_reorder_model<<<c->npartial, RED_THREADS, 0, c->stream>>>(
        n, c->idx_s, c->posh, c->id, c->key_lo, c->have_apot ? c->apot : nullptr, c->rm_state,
        c->rm_state_s, c->pw, c->pwp, c->soa, c->hsml_in,
        c->id_s, c->rho_model, c->key_lo_s, c->apot_s, c->halos, c->nhalos, c->box.mpart,
        c->box.boxhalf_d, c->partial);
    LAUNCH_CHECK();

    // tree.c:298-310: nodes displaced by the sign test, and what they prune.  Needs the sorted
    // keys and positions only; nothing before the sweep reads its output (pw.w is read through
    // fabsf by the tile walk), so it runs beside the hierarchy build unless the cold start needs
    // cpl on the main stream anyway.
    const bool emulate = !(c->cfg.flags & TG_EXACT_NEIGHBOURS);
    const bool beside = emulate && !c->any_cold && !getenv("TOYGPU_NO_SIDE_STREAM");
    cudaStream_t ds = beside ? c->side : c->stream;
    if (beside) {
        CU(cudaEventRecord(c->ev_fork, c->stream));
        CU(cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    }
    if (c->any_cold || emulate) {
        k_cpl<<<cdiv(n, T), T, 0, ds>>>(n, c->key_hi_s, c->key_lo_s, c->cpl);
        LAUNCH_CHECK();
    }
    CU(cudaMemsetAsync(c->defect.counts, 0, 8 * sizeof(int), ds));
    CU(cudaMemsetAsync(c->defect.boxflag, 0, (size_t)n / 32 + 1, ds));
    if (emulate) {
        k_defect_detect<<<cdiv(n, T), T, 0, ds>>>(n, c->pw, c->box.box_d, c->cpl, c->defect);
        LAUNCH_CHECK();
        k_defect_paths<<<296, 128, 0, ds>>>(n, c->pw, c->box.box_d, c->key_hi_s, c->key_lo_s,
                                                  c->cpl, c->defect);
        LAUNCH_CHECK();
        k_defect_paths_big<<<296, 128, 0, ds>>>(n, c->pw, c->box.box_d, c->key_hi_s,
                                                      c->key_lo_s, c->cpl, c->defect);
        LAUNCH_CHECK();
    }
    if (beside) {
        CU(cudaEventRecord(c->ev_join, c->side));
        c->defect_pending = true;
    }

    k_final_sum<<<1, RED_THREADS, 0, c->stream>>>(c->npartial, c->partial, c->scal);
    LAUNCH_CHECK();

    const Bvh &t = c->bvh;
    float *m = c->bvh_mem;
    const size_t S = c->bvh_total;
    const float pad = 4e-7f * c->box.box_f;
    k_bvh_leaves<<<cdiv((long long)t.lvl_n[0] * 32, T), T, 0, c->stream>>>(
        n, c->pw, t.lvl_n[0], pad, m, m + S, m + 2 * S, m + 3 * S, m + 4 * S, m + 5 * S,
        (float *)t.scx, (float *)t.scy, (float *)t.scz, (float *)t.shx, (float *)t.shy, (float *)t.shz);
    LAUNCH_CHECK();
    for (int l = 1; l <= t.top; l++) {
        const int oc = t.lvl_off[l - 1], op = t.lvl_off[l];
        k_bvh_up<<<cdiv((long long)t.lvl_n[l] * 32, T), T, 0, c->stream>>>(
            t.lvl_n[l - 1], t.lvl_n[l], pad, m + oc, m + S + oc, m + 2 * S + oc, m + 3 * S + oc,
            m + 4 * S + oc, m + 5 * S + oc, m + op, m + S + op, m + 2 * S + op, m + 3 * S + op,
            m + 4 * S + op, m + 5 * S + op);
        LAUNCH_CHECK();
    }

    if (c->any_cold) {   // tree.c:113-121 stand-in for particles with Hsml == 0
        k_collapse_events<<<cdiv(n, T), T, 0, c->stream>>>(n, c->cpl, c->ev_start, c->ev_level, c->ev_count);
        LAUNCH_CHECK();
        k_guess_hsml<<<cdiv(n, T), T, 0, c->stream>>>(n, c->key_hi_s, c->key_lo_s, c->cpl, c->ev_start,
                                                     c->ev_level, c->ev_count, c->box.box_d, c->guess,
                                                     nullptr);
        LAUNCH_CHECK();
    }

    // the sorted order is now the current order
    std::swap(c->id, c->id_s);
    std::swap(c->rm_state, c->rm_state_s);
    if (c->have_apot) std::swap(c->apot, c->apot_s);
    c->index_valid = true;
    return TG_OK;
}

static SweepArgs sweep_args(tg_ctx *c, double step)
{
    SweepArgs a{};
    a.t = c->bvh;
    a.bx = c->box;
    a.pw = c->pw;
    a.soa = c->soa;
    a.pwp = (const float4 *)c->pwp;
    a.hsml_in = c->hsml_in;
    a.guess = c->guess;
    a.hsml_out = c->hsml_out;
    a.rho_out = c->rho;
    a.varh_out = c->varh;
    a.delta = c->delta;
    a.vsum = c->scal;
    a.step = step;
    a.bias_const = c->bias_const;
    a.lo = c->lo;
    a.hi = c->hi;
    a.next = c->flags;
    a.gscratch = c->gscratch;
    a.counters = c->counters;
    a.status = c->flags + 1;
    a.rho_in = c->rho;
    a.varh_in = c->varh;
    a.apot = c->apot;
    a.bfld = c->bfld;
    a.worklist = c->worklist;
    a.nwork = c->flags + 5;
    a.tile_ng = c->tile_ng;
    a.tile_groups = c->tile_groups;
    a.tile_mask = c->tile_mask;
    a.dmap = c->defect.dmap;
    a.dnodes = c->defect.nodes;
    a.boxflag = c->defect.boxflag;
    return a;
}

// Generic (v1) sweep over [lo, hi).
template <int MODE> static int launch_generic(tg_ctx *c, SweepArgs a)
{
    const size_t smem = (size_t)SW_WARPS * SW_LCAP * sizeof(double);
    { const int rc = join_defects(c); if (rc) return rc; }
    CU(cudaMemsetAsync(c->flags, 0, sizeof(int), c->stream));       // work counter
    a.next = c->flags;
    constexpr bool has_fast = MODE == MODE_DENSITY || MODE == (MODE_DENSITY | MODE_WVT);
    if constexpr (has_fast) {
        if (c->cfg.flags & TG_FAST) {       // the cold pass of TG_FAST (sph_fast.cuh)
            k_sweep_fast<MODE, false><<<c->sweep_fast_blocks, SF_WARPS * 32, SF_SMEM, c->stream>>>(a);
            LAUNCH_CHECK();
            return TG_OK;
        }
    }
    k_sweep<MODE, false><<<c->sweep_blocks, SW_WARPS * 32, smem, c->stream>>>(a);
    LAUNCH_CHECK();
    return TG_OK;
}

// Tile (v2) sweep, then the generic kernel on whatever the tiles handed back.
template <int MODE> static int launch_tiled(tg_ctx *c, SweepArgs a)
{
    const size_t smem = (size_t)SW_WARPS * SW_LCAP * sizeof(double);
    const int tile_lo = c->lo / 32, tile_hi = (c->hi + 31) / 32;
    if (tile_hi <= tile_lo) return TG_OK;
    CU(cudaMemsetAsync(c->flags, 0, sizeof(int), c->stream));
    CU(cudaMemsetAsync(c->flags + 5, 0, 2 * sizeof(int), c->stream));
    k_tile_walk<<<cdiv(tile_hi - tile_lo, TW_WARPS), TW_WARPS * 32, 0, c->stream>>>(
        c->bvh, c->box, c->pw, a.hsml_in, c->scal, tile_lo, tile_hi, c->tile_ng, c->tile_groups,
        MODE == MODE_ROTA ? 1 : 0);
    LAUNCH_CHECK();
    { const int rc = join_defects(c); if (rc) return rc; }
    a.next = c->flags;
    constexpr bool has_fast = MODE == MODE_DENSITY || MODE == (MODE_DENSITY | MODE_WVT) || MODE == MODE_ROTA;
    constexpr bool has_exact = MODE != MODE_ROTA;
    bool fast = false;
    if constexpr (has_fast) fast = (c->cfg.flags & TG_FAST) != 0;
    if (fast) {
        if constexpr (has_fast)
            k_sweep_tile_fast<MODE><<<c->fast_blocks, TF_WARPS * 32, TF_SMEM, c->stream>>>(a, tile_lo, tile_hi);
    } else {
        if constexpr (has_exact)
            k_sweep_tile<MODE><<<c->tile_blocks, TL_WARPS * 32, TL_SMEM, c->stream>>>(a, tile_lo, tile_hi);
    }
    LAUNCH_CHECK();
    a.next = c->flags + 6;
    constexpr bool has_fast_generic = MODE == MODE_DENSITY || MODE == (MODE_DENSITY | MODE_WVT);
    if constexpr (has_fast_generic) {
        if (fast) {                          // hand-backs of the fast tile sweep stay in FP32
            k_sweep_fast<MODE, true><<<c->sweep_fast_blocks, SF_WARPS * 32, SF_SMEM, c->stream>>>(a);
            LAUNCH_CHECK();
            return TG_OK;
        }
    }
    k_sweep<MODE, true><<<c->sweep_blocks, SW_WARPS * 32, smem, c->stream>>>(a);
    LAUNCH_CHECK();
    return TG_OK;
}

template <int MODE> static int launch_sweep(tg_ctx *c, const SweepArgs &a)
{
    constexpr bool tileable = MODE == MODE_DENSITY || MODE == (MODE_DENSITY | MODE_EXACT) || MODE == MODE_WVT ||
                              MODE == (MODE_DENSITY | MODE_WVT);
    if constexpr (tileable) {
        if (c->use_tiles && !c->any_cold) return launch_tiled<MODE>(c, a);
    }
    if constexpr (MODE == MODE_ROTA) {       // rot(A) has a tile path in TG_FAST only
        if (c->use_tiles && !c->any_cold && (c->cfg.flags & TG_FAST)) return launch_tiled<MODE>(c, a);
    }
    return launch_generic<MODE>(c, a);
}

static int check_flags(tg_ctx *c, bool swept = true)
{
    int f[10], dc[4];
    { const int rc = join_defects(c); if (rc) return rc; }
    {   // a rank that fails must take the others with it, or their next collective hangs
        const int rc = reduce_flags_max(c, c->flags + 1, 2);
        if (rc) return rc;
    }
    int *hp = (int *)(c->hpin + 2 * c->cfg.nranks);          // page-locked: 10 + 4 ints
    CU(cudaMemcpyAsync(hp, c->flags, sizeof f, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(hp + 10, c->defect.counts, sizeof dc, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    memcpy(f, hp, sizeof f);
    memcpy(dc, hp + 10, sizeof dc);
    c->stats.displaced_nodes = dc[0];
    c->stats.displaced_particles = dc[3];
    c->stats.displaced_overflow = dc[2];
    const int tie = f[9];
    // the index build has already put the ids into the new order while the positions are only
    // rewritten after a successful sweep: after any of the failures below the two disagree
    if (f[2] || f[1] || tie || (swept && dc[2] && !(c->cfg.flags & TG_EXACT_NEIGHBOURS))) {
        c->poisoned = true;
        c->index_valid = false;
    }
    if (f[2]) return fail(c, TG_ERANGE, "particle position outside [0, Boxsize] (peano.c:130-132)");
    if (tie) return fail(c, TG_ERANGE, "more than %d sort cells hold over %d particles each (the longest %d; coincident "
                         "positions?): rerun with TOYGPU_FULL_SORT=1 or remove the duplicates", RS_LONG_CAP, RS_TIE_CAP, tie);
    if (swept && dc[2] && !(c->cfg.flags & TG_EXACT_NEIGHBOURS))
        return fail(c, TG_ERANGE, "more displaced reference-tree nodes than the path table holds (%d events): the "
                         "neighbour sets would silently stop being the reference's; use TG_EXACT_NEIGHBOURS "
                         "for the exact predicate sets", c->defect.cap_events);
    if (f[1]) return fail(c, TG_ENOCONV, "hsml iteration did not terminate (fewer than %d gas particles in reach?)", TG_DESNNGB);
    return TG_OK;
}

static int density_pass(tg_ctx *c)
{
    SweepArgs a = sweep_args(c, 0);
    // the sequential parity mode also forms Find_hsml's kernels operation for operation
    int rc = (c->cfg.flags & TG_WVT_SEQUENTIAL) ? launch_sweep<MODE_DENSITY | MODE_EXACT>(c, a)
                                                : launch_sweep<MODE_DENSITY>(c, a);
    if (rc) return rc;
    c->any_cold = false;
    return TG_OK;
}

static int carry_state(tg_ctx *c)
{
    if (c->hi > c->lo)
        k_carry<<<cdiv(c->hi - c->lo, 256), 256, 0, c->stream>>>(c->lo, c->hi, c->pw, c->hsml_out, c->posh);
    LAUNCH_CHECK();
    return TG_OK;
}

// After a sweep: every rank holds every slice again (collective 2).
static int gather_state(tg_ctx *c, bool with_density)
{
    int rc = gather_slices(c, c->posh, sizeof(float4));
    if (rc || !with_density) return rc;
    if ((rc = gather_slices(c, c->rho, sizeof(float)))) return rc;
    return gather_slices(c, c->varh, sizeof(float));
}

static int displacement_pass(tg_ctx *c, double step)
{
    SweepArgs a = sweep_args(c, step);
    if (c->cfg.flags & TG_WVT_SEQUENTIAL) return launch_sweep<MODE_WVT_SEQ>(c, a);
    return launch_sweep<MODE_WVT>(c, a);
}

static int move_pass(tg_ctx *c, double scale)
{
    // wvt_relax.c:113: SphP.Rho_Model is (only) written by the model-hsml pass of an iteration
    CU(cudaMemcpyAsync(c->rm_state, c->rho_model, sizeof(float) * c->n, cudaMemcpyDeviceToDevice, c->stream));
    if (c->hi > c->lo)
        k_move<<<cdiv(c->hi - c->lo, 256), 256, 0, c->stream>>>(c->lo, c->hi, c->n, c->pw, c->hsml_out, c->delta,
                                                               c->box.box_d, scale, c->posh);
    LAUNCH_CHECK();
    return TG_OK;
}

static int finish_stats(tg_ctx *c, bool have_sweep_events)
{
    unsigned long long h[9];
    int nwork = 0;
    CU(cudaMemcpyAsync(h, c->counters, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(&nwork, c->flags + 5, sizeof nwork, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->stats.handed_back = (unsigned long long)nwork;

    c->stats.pair_evals = h[0];
    c->stats.gathered = h[1];
    c->stats.searches = h[2];
    c->stats.hsml_iters = h[3];
    for (int k = 0; k < 5; k++) c->stats.handback_why[k] = h[4 + k];
    c->stats.kernels = c->launches;
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
    c->stats.step_ms = ms;
    if (have_sweep_events) {
        CU(cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
        c->stats.sweep_ms = ms;
        CU(cudaEventElapsedTime(&ms, c->ev[0], c->ev[2]));
        c->stats.index_ms = ms;
        CU(cudaEventElapsedTime(&ms, c->ev[3], c->ev[1]));
        c->stats.tail_ms = ms;
    }
    return TG_OK;
}

// ------------------------------------------------------------------ operators

extern "C" int tg_find_sph_quantities(tg_ctx *c)
{
    if (!c) return TG_EINVAL;
    TG_GROUP(c, tg_find_sph_quantities(k));
    CU(cudaSetDevice(c->cfg.device));
    int rc = reset_counters(c);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev[0], c->stream));
    if ((rc = prepare_index(c))) return rc;
    CU(cudaEventRecord(c->ev[2], c->stream));
    if ((rc = density_pass(c))) return rc;
    CU(cudaEventRecord(c->ev[3], c->stream));
    if ((rc = carry_state(c))) return rc;
    if ((rc = gather_state(c, true))) return rc;
    c->density_stale = false;
    CU(cudaEventRecord(c->ev[1], c->stream));
    if ((rc = check_flags(c))) return rc;
    return finish_stats(c, true);
}

// First half of one pass of wvt_relax.c:66-214: sort, index, density solve and the error
// statistics of THIS rank's slice.  Unless TG_WVT_SEQUENTIAL is set the displacement is
// computed in the same sweep with `step_guess` (the step as it stands before this
// iteration's `step *= 0.8` decision, wvt_relax.c:100) and rescaled by tg_wvt_finish.
extern "C" int tg_wvt_begin(tg_ctx *c, double step_guess, double *err_sum, double *err_max, int *count)
{
    if (!c) return TG_EINVAL;
    if (!(step_guess > 0)) return fail(c, TG_EINVAL, "tg_wvt_begin: step must be positive");
    if (!c->kids.empty()) {      // every rank returns the same global numbers
        const int rc = group_run(c, [&](tg_ctx *k) {
            double s = 0, m = 0; int cnt = 0;
            const int r = tg_wvt_begin(k, step_guess, &s, &m, &cnt);
            if (r == TG_OK && k == c->kids[0]) {
                if (err_sum) *err_sum = s;
                if (err_max) *err_max = m;
                if (count) *count = cnt;
            }
            return r;
        });
        return rc;
    }
    CU(cudaSetDevice(c->cfg.device));
    int rc = reset_counters(c);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev[0], c->stream));
    if ((rc = prepare_index(c))) return rc;                  // wvt_relax.c:67 -> sph.c:15-17
    CU(cudaEventRecord(c->ev[2], c->stream));
    if (c->cfg.flags & TG_WVT_SEQUENTIAL) {
        if ((rc = density_pass(c))) return rc;               // sph.c:19-72
    } else {   // one sweep: density solve and displacement share the walk over HBM
        SweepArgs a = sweep_args(c, step_guess);
        if ((rc = launch_sweep<MODE_DENSITY | MODE_WVT>(c, a))) return rc;
        c->any_cold = false;
    }
    CU(cudaEventRecord(c->ev[3], c->stream));
    c->step_begin = step_guess;
    // wvt_relax.c:73-87.  Everything the host needs from this half -- the error statistics (of
    // all ranks: SURVEY 8e, collective 3, reduced in rank order), the status flags -- is queued
    // first and read back behind ONE synchronisation (check_flags).
    const int R = c->comm ? c->cfg.nranks : 1;
    {
        const int m = c->hi - c->lo, nb = cdiv(m, RED_THREADS);
        k_error<<<nb, RED_THREADS, 0, c->stream>>>(c->lo, c->hi, c->rho, c->rho_model, c->partial);
        LAUNCH_CHECK();
        k_final_err<<<1, RED_THREADS, 0, c->stream>>>(nb, c->partial, c->scal + 1);
        LAUNCH_CHECK();
    }
    if (c->comm) NC(nccl_api()->AllGather(c->scal + 1, c->errbuf, 2, ncclFloat64, c->comm, c->stream));
    double *h = c->hpin;                                     // page-locked: the copy really is asynchronous
    CU(cudaMemcpyAsync(h, c->comm ? c->errbuf : c->scal + 1, sizeof(double) * 2 * R, cudaMemcpyDeviceToHost, c->stream));
    if ((rc = check_flags(c))) return rc;                    // synchronises
    double esum = 0, emax = 0;
    for (int r = 0; r < R; r++) { esum += h[2 * r]; emax = std::max(emax, h[2 * r + 1]); }
    const int cnt = c->comm ? c->n : c->hi - c->lo;
    if (err_sum) *err_sum = esum;
    if (err_max) *err_max = emax;
    if (count) *count = cnt;
    return TG_OK;
}

// Second half: displacement with the final step and move (wvt_relax.c:108-214), or, with
// step_final <= 0, leave the loop as the reference's `break` does: positions stay, the
// freshly solved Hsml is kept (wvt_relax.c:94-98).
extern "C" int tg_wvt_finish(tg_ctx *c, double step_final)
{
    if (!c) return TG_EINVAL;
    TG_GROUP(c, tg_wvt_finish(k, step_final));
    CU(cudaSetDevice(c->cfg.device));
    int rc;
    if (step_final <= 0) {
        if ((rc = carry_state(c))) return rc;
    } else {
        double scale = 1.0;
        if (c->cfg.flags & TG_WVT_SEQUENTIAL) {
            if ((rc = displacement_pass(c, step_final))) return rc;
        } else if (step_final != c->step_begin) {
            scale = step_final / c->step_begin;              // delta is linear in the step
        }
        if ((rc = move_pass(c, scale))) return rc;
    }
    // The moved slices travel every iteration; Rho / VarHsmlFac of the other ranks' slices are
    // not needed by the next iteration and follow when the operator returns (tg_sync_results).
    if ((rc = gather_state(c, false))) return rc;
    c->density_stale = c->comm != nullptr;
    CU(cudaEventRecord(c->ev[1], c->stream));
    return finish_stats(c, true);
}

// Multi-rank: make Rho / VarHsmlFac of every slice present on every rank (collective; a no-op
// when they already are).  tg_regularise, tg_wvt_iteration and tg_find_sph_quantities end with
// it; hosts that drive tg_wvt_begin / tg_wvt_finish themselves call it before a download.
extern "C" int tg_sync_results(tg_ctx *c)
{
    if (!c) return TG_EINVAL;
    TG_GROUP(c, tg_sync_results(k));
    if (!c->comm || !c->density_stale) return TG_OK;
    CU(cudaSetDevice(c->cfg.device));
    int rc = gather_slices(c, c->rho, sizeof(float));
    if (rc) return rc;
    if ((rc = gather_slices(c, c->varh, sizeof(float)))) return rc;
    CU(cudaStreamSynchronize(c->stream));
    c->density_stale = false;
    return TG_OK;
}

extern "C" int tg_wvt_iteration(tg_ctx *c, double step, double *err_max, double *err_mean)
{
    double sum = 0, mx = 0;
    int cnt = 0;
    int rc = tg_wvt_begin(c, step, &sum, &mx, &cnt);
    if (rc) return rc;
    if (err_max) *err_max = mx;
    if (err_mean) *err_mean = cnt > 0 ? sum / cnt : 0;
    if ((rc = tg_wvt_finish(c, step))) return rc;
    return tg_sync_results(c);
}

extern "C" int tg_regularise(tg_ctx *c, int max_iters, tg_log_fn log, void *user, int *iters_done)
{
    if (!c) return TG_EINVAL;
    if (!c->kids.empty()) {      // the loop runs on every rank; rank 0 talks to the observer
        return group_run(c, [&](tg_ctx *k) {
            int done = 0;
            const bool first = k == c->kids[0];
            const int r = tg_regularise(k, max_iters, first ? log : nullptr, user, &done);
            if (first && iters_done) *iters_done = done;
            return r;
        });
    }
    if (c->cfg.nranks > 1 && !c->comm)
        return fail(c, TG_EINVAL, "tg_regularise needs the global error statistics: with nranks > 1 "
                                  "call tg_comm_init first (or drive tg_wvt_begin / tg_wvt_finish and "
                                  "all-reduce in between)");
    // wvt_relax.c:46-59
    int it = -1, started = 0;
#ifdef TG_CUBIC_SPLINE
    double step = 0.035;                                     // wvt_relax.c:48-49
#else
    double step = 0.0085;
    if (c->cfg.mtotal < 1e5) step /= 2;
#endif
    double errLast = DBL_MAX, errDiff = DBL_MAX, errDiffLast = DBL_MAX;
    int rc;

    for (;;) {
        if (it++ >= TG_NUMITER) break;                       // wvt_relax.c:63
        if (it >= max_iters) break;                          // test / bench cut-off
        double errSum = 0, errMax = 0;
        int cnt = 0;
        if ((rc = tg_wvt_begin(c, step, &errSum, &errMax, &cnt))) return rc;
        started++;
        const double errMean = errSum / cnt;                 // wvt_relax.c:87
        errDiff = (errLast - errMean) / errMean;             // wvt_relax.c:89
        int stop = log ? log(it, errMax, errMean, errDiff, step, user) : 0;
        if (c->comm) {           // the observer lives on one rank: everybody must hear its verdict
            CU(cudaMemcpyAsync(c->flags + 7, &stop, sizeof(int), cudaMemcpyHostToDevice, c->stream));
            if ((rc = reduce_flags_max(c, c->flags + 7, 1))) return rc;
            CU(cudaMemcpyAsync(&stop, c->flags + 7, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            CU(cudaStreamSynchronize(c->stream));
        }

        bool leave = stop != 0;
        if (errDiff < 0.01 && it > 25) leave = true;         // wvt_relax.c:94
        if (errDiff < 0 && errDiffLast < 0 && it > 10) leave = true;   // wvt_relax.c:97
        if (leave) {
            if ((rc = tg_wvt_finish(c, 0))) return rc;
            break;
        }
        if (errDiff < 0.01 && it > 1) step *= 0.8;           // wvt_relax.c:100
        errLast = errMean;
        errDiffLast = errDiff;
        if ((rc = tg_wvt_finish(c, step))) return rc;        // wvt_relax.c:108-214
    }
    if (iters_done) *iters_done = started;
    return tg_sync_results(c);
}

extern "C" int tg_bfld_from_rotA(tg_ctx *c)
{
    if (!c) return TG_EINVAL;
    TG_GROUP(c, tg_bfld_from_rotA(k));
    if (!c->index_valid) return fail(c, TG_EINVAL, "tg_bfld_from_rotA: no index (call tg_find_sph_quantities first, sph.c:229 reuses the tree)");
    if (!c->have_apot) return fail(c, TG_EINVAL, "tg_bfld_from_rotA: Apot not set");
    CU(cudaSetDevice(c->cfg.device));
    int rc = reset_counters(c);
    if (rc) return rc;
    SweepArgs a = sweep_args(c, 0);
    a.hsml_in = c->hsml_out;    // SphP.Hsml as left by the density call
    CU(cudaEventRecord(c->ev[0], c->stream));
    CU(cudaEventRecord(c->ev[2], c->stream));
    if ((rc = launch_sweep<MODE_ROTA>(c, a))) return rc;
    CU(cudaEventRecord(c->ev[3], c->stream));
    if ((rc = gather_slices(c, c->bfld, 3 * sizeof(float)))) return rc;
    CU(cudaEventRecord(c->ev[1], c->stream));
    if ((rc = check_flags(c))) return rc;
    return finish_stats(c, true);
}

extern "C" int tg_make_magnetic_field(tg_ctx *c, const tg_bfield *par, double *norm_out, int *n_limited_out)
{
    if (!c || !par) return TG_EINVAL;
    if (!c->kids.empty()) {
        return group_run(c, [&](tg_ctx *k) {
            double norm = 0; int lim = 0;
            const int r = tg_make_magnetic_field(k, par, &norm, &lim);
            if (r == TG_OK && k == c->kids[0]) {
                if (norm_out) *norm_out = norm;
                if (n_limited_out) *n_limited_out = lim;
            }
            return r;
        });
    }
    if (!c->index_valid) return fail(c, TG_EINVAL, "tg_make_magnetic_field: no index (call tg_find_sph_quantities first)");
    if (c->nhalos == 0) return fail(c, TG_EINVAL, "tg_set_halos has not been called");
    if (c->cfg.nranks > 1 && !c->comm)
        return fail(c, TG_EINVAL, "tg_make_magnetic_field with nranks > 1 needs a communicator (the field maximum is global)");
    CU(cudaSetDevice(c->cfg.device));
    const int n = c->n, T = 256;
    if (!c->apot) {
        CU(dmalloc(&c->apot, (size_t)3 * n));
        CU(dmalloc(&c->apot_s, (size_t)3 * n));
    }
    if (!c->halo_extra) {                       // persistent scratch: nothing to leak on an early return
        CU(dmalloc(&c->halo_extra, MAX_HALOS));
        CU(dmalloc(&c->n_limited, 1));
    }
    std::vector<HaloExtra> ex(c->nhalos);
    for (int j = 0; j < c->nhalos; j++) {
        ex[j].r_sample_gas = par->r_sample_gas ? par->r_sample_gas[j] : 0;
        ex[j].r_sample_dm = par->r_sample_dm ? par->r_sample_dm[j] : 0;
        ex[j].is_stripped = par->is_stripped ? par->is_stripped[j] : 0;
    }
    HaloExtra *dex = c->halo_extra;
    int *dcnt = c->n_limited;
    CU(cudaMemcpyAsync(dex, ex.data(), sizeof(HaloExtra) * c->nhalos, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(dcnt, 0, sizeof(int), c->stream));
    int rc = reset_counters(c);
    if (rc) return rc;
    CU(cudaEventRecord(c->ev[0], c->stream));
    k_vector_potential<<<cdiv(n, T), T, 0, c->stream>>>(n, c->pw, c->halos, c->nhalos, c->box.boxhalf_f,
                                                       par->bfld_eta, c->apot);
    LAUNCH_CHECK();
    c->have_apot = true;
    SweepArgs a = sweep_args(c, 0);
    a.hsml_in = c->hsml_out;    // SphP.Hsml as left by the density call
    CU(cudaEventRecord(c->ev[2], c->stream));
    if ((rc = launch_sweep<MODE_ROTA>(c, a))) return rc;
    CU(cudaEventRecord(c->ev[3], c->stream));
    // every rank needs the whole field: its maximum is global (magnetic_field.c:77-86), and the
    // cap below indexes Halo_containing by particle number
    if ((rc = gather_slices(c, c->bfld, 3 * sizeof(float)))) return rc;
    const int nb = cdiv(n, RED_THREADS);
    k_bfld_max<<<nb, RED_THREADS, 0, c->stream>>>(n, c->bfld, c->partial);
    LAUNCH_CHECK();
    k_bfld_max_final<<<1, RED_THREADS, 0, c->stream>>>(nb, c->partial, c->scal + 3);
    LAUNCH_CHECK();
    double max_b2 = 0;
    CU(cudaMemcpyAsync(&max_b2, c->scal + 3, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if ((rc = check_flags(c))) return rc;        // (synchronises)
    const double norm = par->bfld_norm / sqrt(max_b2) / sqrt(3.0);     // magnetic_field.c:88-90
    k_bfld_normalise<<<cdiv(n, T), T, 0, c->stream>>>(n, c->pw, c->bfld, norm, c->box.boxhalf_f, c->halos, dex,
                                                     c->nhalos, par->sub_first, c->box.box_d,
                                                     par->bmax_main, par->bmax_sub, dcnt);
    LAUNCH_CHECK();
    CU(cudaEventRecord(c->ev[1], c->stream));
    int cnt = 0;
    CU(cudaMemcpyAsync(&cnt, dcnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if ((rc = finish_stats(c, true))) return rc;
    if (norm_out) *norm_out = norm;
    if (n_limited_out) *n_limited_out = cnt;
    return TG_OK;
}

// Reassign_particles_to_halos(), the per-particle half (positions.c:264-283).
// The sort that follows it in the reference (positions.c:405-443) is gsl_heapsort_index on the
// ids, whose order of EQUAL ids decides the particle order of the output file; it stays with
// the caller (the shim runs the reference's own Qsort_Index on these ids).
extern "C" int tg_halo_ids(tg_ctx *c, const tg_bfield *par, int32_t *ids, long long *npart)
{
    if (!c || !par || !ids) return fail(c, TG_EINVAL, "tg_halo_ids: null argument");
    TG_GROUP0(c, tg_halo_ids(k, par, ids, npart));        // the state is replicated on every rank
    if (c->nhalos == 0) return fail(c, TG_EINVAL, "tg_set_halos has not been called");
    CU(cudaSetDevice(c->cfg.device));
    const int n = c->n, T = 256;
    if (!c->halo_extra) {
        CU(dmalloc(&c->halo_extra, MAX_HALOS));
        CU(dmalloc(&c->n_limited, 1));
    }
    if (!c->halo_counts) CU(dmalloc(&c->halo_counts, MAX_HALOS));
    std::vector<HaloExtra> ex(c->nhalos);
    for (int j = 0; j < c->nhalos; j++) {
        ex[j].r_sample_gas = par->r_sample_gas ? par->r_sample_gas[j] : 0;
        ex[j].r_sample_dm = par->r_sample_dm ? par->r_sample_dm[j] : 0;
        ex[j].is_stripped = par->is_stripped ? par->is_stripped[j] : 0;
    }
    CU(cudaMemcpyAsync(c->halo_extra, ex.data(), sizeof(HaloExtra) * c->nhalos, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync(c->halo_counts, 0, sizeof(unsigned long long) * c->nhalos, c->stream));
    k_halo_ids<<<cdiv(n, T), T, 0, c->stream>>>(n, c->posh, c->box.boxhalf_f, c->halos, c->halo_extra, c->nhalos,
                                               par->sub_first, c->box.box_d, c->idx, c->halo_counts);
    LAUNCH_CHECK();
    CU(cudaMemcpyAsync(ids, c->idx, sizeof(int) * n, cudaMemcpyDeviceToHost, c->stream));
    std::vector<unsigned long long> cnt(c->nhalos);
    CU(cudaMemcpyAsync(cnt.data(), c->halo_counts, sizeof(unsigned long long) * c->nhalos, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (npart) for (int j = 0; j < c->nhalos; j++) npart[j] = (long long)cnt[j];
    return TG_OK;
}

extern "C" int tg_get_apot(tg_ctx *c, float *apot)
{
    if (!c || !apot) return TG_EINVAL;
    TG_GROUP0(c, tg_get_apot(k, apot));
    if (!c->have_apot) return fail(c, TG_EINVAL, "tg_get_apot: Apot not set");
    CU(cudaSetDevice(c->cfg.device));
    CU(cudaMemcpyAsync(apot, c->apot, sizeof(float) * 3 * c->n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TG_OK;
}

extern "C" int tg_get_stats(tg_ctx *c, tg_stats *out)
{
    if (!c || !out) return TG_EINVAL;
    if (!c->kids.empty()) {      // counters summed over the ranks, times = the slowest rank
        tg_stats t = c->kids[0]->stats;
        for (size_t r = 1; r < c->kids.size(); r++) {
            const tg_stats &k = c->kids[r]->stats;
            t.pair_evals += k.pair_evals; t.gathered += k.gathered; t.searches += k.searches;
            t.hsml_iters += k.hsml_iters; t.kernels += k.kernels; t.handed_back += k.handed_back;
            for (int q = 0; q < 5; q++) t.handback_why[q] += k.handback_why[q];
            t.sweep_ms = std::max(t.sweep_ms, k.sweep_ms);
            t.index_ms = std::max(t.index_ms, k.index_ms);
            t.tail_ms = std::max(t.tail_ms, k.tail_ms);
            t.step_ms = std::max(t.step_ms, k.step_ms);
        }
        *out = t;
        return TG_OK;
    }
    *out = c->stats;
    return TG_OK;
}

// ------------------------------------------------------------------ Gadget blocks

// io.c:141-166 for the gas range: element k of the block's write buffer is the field of the
// particle order[k] (the file order sort_particles() left, positions.c:405-443), read from the
// SoA state.  One thread per output float; reads are gathers, writes are coalesced.
__global__ void k_fill_block(int n, int block, const int *__restrict__ order, const float4 *__restrict__ posh,
                             const float *__restrict__ rho, const float *__restrict__ bfld,
                             const float *__restrict__ rm, float *__restrict__ out)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int vals = (block == TG_BLOCK_POS || block == TG_BLOCK_BFLD) ? 3 : 1;
    if (t >= (size_t)n * vals) return;
    const int k = (int)(t / vals), comp = (int)(t - (size_t)k * vals);
    const int src = order ? order[k] : k;
    float v;
    switch (block) {
    case TG_BLOCK_POS: { const float4 p = posh[src]; v = comp == 0 ? p.x : (comp == 1 ? p.y : p.z); break; }
    case TG_BLOCK_RHO: v = rho[src]; break;
    case TG_BLOCK_HSML: v = posh[src].w; break;
    case TG_BLOCK_BFLD: v = bfld[3 * (size_t)src + comp]; break;
    default: v = rm[src]; break;
    }
    out[t] = v;
}

__global__ void k_order_from_host(int n, const unsigned long long *__restrict__ in, int *__restrict__ out, int *bad)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const unsigned long long v = in[k];
    if (v >= (unsigned long long)n) { *bad = 1; out[k] = 0; return; }
    out[k] = (int)v;
}

extern "C" int tg_set_output_order(tg_ctx *c, const size_t *order)
{
    if (!c) return TG_EINVAL;
    TG_GROUP0(c, tg_set_output_order(k, order));          // the state is replicated on every rank
    c->out_order_stale = false;
    if (!order) { c->have_out_order = false; return TG_OK; }
    CU(cudaSetDevice(c->cfg.device));
    const int n = c->n;
    if (!c->out_order) CU(dmalloc(&c->out_order, (size_t)n));
    // size_t[n] crosses the bus as it is (through the unsorted-key scratch, dead once the index is built: 8 B per particle) and is
    // narrowed and range-checked on the device
    static_assert(sizeof(size_t) == 8, "size_t");
    unsigned long long *tmp = (unsigned long long *)c->key_lo;
    CU(cudaMemcpyAsync(tmp, order, sizeof(size_t) * n, cudaMemcpyHostToDevice, c->stream));
    int *bad_dev = c->flags + 8;                           // scratch word of the tie fix-up
    CU(cudaMemsetAsync(bad_dev, 0, sizeof(int), c->stream));
    k_order_from_host<<<cdiv(n, 256), 256, 0, c->stream>>>(n, tmp, c->out_order, bad_dev);
    LAUNCH_CHECK();
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, bad_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (bad) return fail(c, TG_EINVAL, "tg_set_output_order: an index is outside [0, n_gas)");
    c->have_out_order = true;
    return TG_OK;
}

extern "C" int tg_fill_block(tg_ctx *c, int block, float *out)
{
    if (!c || !out || block < TG_BLOCK_POS || block > TG_BLOCK_RHOMODEL)
        return fail(c, TG_EINVAL, "tg_fill_block: bad arguments");
    TG_GROUP0(c, tg_fill_block(k, block, out));
    if (c->poisoned) return fail(c, TG_EINVAL, "tg_fill_block: the last step failed; upload again");
    if (c->have_out_order && c->out_order_stale)
        return fail(c, TG_EINVAL, "tg_fill_block: the device order changed since tg_set_output_order "
                                  "(an upload or an operator ran in between); set the file order again");
    CU(cudaSetDevice(c->cfg.device));
    const int n = c->n;
    const int vals = (block == TG_BLOCK_POS || block == TG_BLOCK_BFLD) ? 3 : 1;
    const size_t total = (size_t)n * vals;
    k_fill_block<<<cdiv((long long)total, 256), 256, 0, c->stream>>>(n, block, c->have_out_order ? c->out_order : nullptr,
                                                                    c->posh, c->rho, c->bfld, c->rm_state, c->stage);
    LAUNCH_CHECK();
    CU(cudaMemcpyAsync(out, c->stage, sizeof(float) * total, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TG_OK;
}

// ------------------------------------------------------------------ data out

extern "C" int tg_download_soa(tg_ctx *c, float *pos, int32_t *perm, float *hsml, float *rho,
                               float *varhsml, float *rho_model, float *bfld)
{
    if (!c) return TG_EINVAL;
    TG_GROUP0(c, tg_download_soa(k, pos, perm, hsml, rho, varhsml, rho_model, bfld));   // state is replicated
    CU(cudaSetDevice(c->cfg.device));
    const int n = c->n;
    cudaStream_t st = c->stream;
    if (pos || hsml) {
        k_unpack_state<<<cdiv(n, 256), 256, 0, st>>>(n, c->posh, c->stage, c->stage + 3 * (size_t)n);
        LAUNCH_CHECK();
        if (pos) CU(cudaMemcpyAsync(pos, c->stage, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
        if (hsml) CU(cudaMemcpyAsync(hsml, c->stage + 3 * (size_t)n, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    }
    if (perm) CU(cudaMemcpyAsync(perm, c->id, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    if (rho) CU(cudaMemcpyAsync(rho, c->rho, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    if (varhsml) CU(cudaMemcpyAsync(varhsml, c->varh, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    if (rho_model) CU(cudaMemcpyAsync(rho_model, c->rm_state, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    if (bfld) CU(cudaMemcpyAsync(bfld, c->bfld, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return TG_OK;
}

extern "C" int tg_download(tg_ctx *c, void *P, size_t p_stride, void *SphP, size_t s_stride)
{
    if (!c || !P || !SphP || p_stride < 52 || s_stride < 48)
        return fail(c, TG_EINVAL, "tg_download: records too small for ParticleData / GasParticleData");
    TG_GROUP(c, tg_download(k, P, p_stride, SphP, s_stride));     // every rank writes its own slice
    const int n = c->n;
    if (c->have_raw && p_stride == c->raw_pstride && s_stride == c->raw_sstride) {
        // device path: gather the resident records by the permutation, patch, one DMA each.
        // With a communicator every rank does this for its slice [lo, hi) of the new order (so
        // the copies to the host run in parallel on every GPU's PCIe link) and the slices of
        // the new resident copy are all-gathered.
        CU(cudaSetDevice(c->cfg.device));
        cudaStream_t st = c->stream;
        const int wp = (int)(p_stride / 4), ws = (int)(s_stride / 4);
        const size_t lo = c->comm ? c->lo : 0, m = c->comm ? c->hi - c->lo : n;
        const size_t tp = m * wp, ts = m * ws;
        if (m) {
            k_gather_records<<<cdiv((long long)tp, 256), 256, 0, st>>>(tp, wp, c->id + lo, (const uint32_t *)c->rawP,
                                                                       (uint32_t *)(c->outP + lo * p_stride));
            LAUNCH_CHECK();
            k_gather_records<<<cdiv((long long)ts, 256), 256, 0, st>>>(ts, ws, c->id + lo, (const uint32_t *)c->rawS,
                                                                       (uint32_t *)(c->outS + lo * s_stride));
            LAUNCH_CHECK();
            k_patch_records<<<cdiv((long long)m, 256), 256, 0, st>>>(
                (int)m, (int)lo, c->outP + lo * p_stride, p_stride, c->outS + lo * s_stride, s_stride, c->posh + lo,
                c->index_valid ? c->key_lo_s + lo : nullptr, c->key_hi_s + lo, c->rho + lo, c->varh + lo,
                c->bfld + 3 * lo, c->rm_state + lo);
            LAUNCH_CHECK();
            CU(cudaMemcpyAsync((char *)P + lo * p_stride, c->outP + lo * p_stride, m * p_stride, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync((char *)SphP + lo * s_stride, c->outS + lo * s_stride, m * s_stride, cudaMemcpyDeviceToHost, st));
        }
        int rc;
        if ((rc = gather_slices(c, c->outP, p_stride))) return rc;
        if ((rc = gather_slices(c, c->outS, s_stride))) return rc;
        // the records now ARE in the current order: they become the resident copy
        std::swap(c->rawP, c->outP);
        std::swap(c->rawS, c->outS);
        k_iota<<<cdiv(n, 256), 256, 0, st>>>(n, c->id);
        LAUNCH_CHECK();
        CU(cudaStreamSynchronize(st));
        return TG_OK;
    }
    // host path (state uploaded as plain arrays, records only on the host)
    std::vector<float> pos((size_t)3 * n), hsml(n), rho(n), varh(n), rhom(n), bfld((size_t)3 * n);
    std::vector<int32_t> perm(n);
    int rc = tg_download_soa(c, pos.data(), perm.data(), hsml.data(), rho.data(), varh.data(),
                             rhom.data(), bfld.data());
    if (rc) return rc;
    std::vector<uint64_t> khi(n), klo(n);
    if (c->index_valid) {
        CU(cudaMemcpy(khi.data(), c->key_hi_s, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(klo.data(), c->key_lo_s, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost));
    }
    // whole records travel with the particle (peano.c:96-117)
    // (host side of the AoS boundary: threaded, it is 1.2 GB of records at 10 M particles)
    char *oldP = (char *)malloc((size_t)n * p_stride), *oldS = (char *)malloc((size_t)n * s_stride);
    if (!oldP || !oldS) { free(oldP); free(oldS); return fail(c, TG_ENOMEM, "tg_download: host memory"); }
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n; k++) {
        memcpy(oldP + (size_t)k * p_stride, (const char *)P + (size_t)k * p_stride, p_stride);
        memcpy(oldS + (size_t)k * s_stride, (const char *)SphP + (size_t)k * s_stride, s_stride);
    }
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n; k++) {
        char *p = (char *)P + (size_t)k * p_stride;
        char *s = (char *)SphP + (size_t)k * s_stride;
        memcpy(p, oldP + (size_t)perm[k] * p_stride, p_stride);
        memcpy(s, oldS + (size_t)perm[k] * s_stride, s_stride);
        memcpy(p, &pos[3 * (size_t)k], 12);                     // Pos @ +0
        if (c->index_valid) {
            memcpy(p + 32, &klo[k], 8);                         // Key @ +32 (little endian u128)
            memcpy(p + 40, &khi[k], 8);
            const int parent = k / 32;                          // index group (tree.c's node ids are private)
            memcpy(p + 48, &parent, 4);                         // Tree_Parent @ +48
        }
        float *g = (float *)s;
        g[1] = rho[k];                                          // Rho @ +4
        g[2] = hsml[k];                                         // Hsml @ +8
        g[3] = varh[k];                                         // VarHsmlFac @ +12
        memcpy(s + 16, &bfld[3 * (size_t)k], 12);               // Bfld @ +16
        g[11] = rhom[k];                                        // Rho_Model @ +44
    }
    free(oldP);
    free(oldS);
    return TG_OK;
}

extern "C" int tg_wvt_scratch(tg_ctx *c, float *hsml_wvt, float *delta)
{
    if (!c) return TG_EINVAL;
    TG_NOGROUP(c, "tg_wvt_scratch");
    CU(cudaSetDevice(c->cfg.device));
    const int n = c->n;
    CU(cudaStreamSynchronize(c->stream));
    if (hsml_wvt) {
        std::vector<float4> h(n);
        double vsum = 0;
        CU(cudaMemcpy(h.data(), c->pw, sizeof(float4) * n, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&vsum, c->scal, sizeof(double), cudaMemcpyDeviceToHost));
        const float norm = (float)pow(TG_DESNNGB / vsum / K_FOURPITHIRD, 1.0 / 3.0);   // wvt_relax.c:120
        for (int i = 0; i < n; i++) hsml_wvt[i] = fabsf(h[i].w) * norm;   // sign bit: defect.cuh flag
    }
    if (delta) {
        std::vector<float> d((size_t)3 * n);
        CU(cudaMemcpy(d.data(), c->delta, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++)
            for (int a = 0; a < 3; a++) delta[3 * (size_t)i + a] = d[(size_t)a * n + i];
    }
    return TG_OK;
}

// ------------------------------------------------------------------ test hooks

extern "C" int tg_peano_keys(tg_ctx *c, uint64_t *hi, uint64_t *lo)
{
    if (!c || !hi || !lo) return TG_EINVAL;
    TG_NOGROUP(c, "tg_peano_keys");
    CU(cudaSetDevice(c->cfg.device));
    const int n = c->n;
    CU(cudaMemsetAsync(c->flags + 2, 0, sizeof(int), c->stream));
    k_peano_keys<<<cdiv(n, 256), 256, 0, c->stream>>>(n, c->posh, c->box.box_d, c->key_hi, c->key_lo,
                                                     c->idx, c->flags + 2);
    LAUNCH_CHECK();
    CU(cudaMemcpyAsync(hi, c->key_hi, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(lo, c->key_lo, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, c->stream));
    c->index_valid = false;
    return check_flags(c, false);
}

extern "C" int tg_sort(tg_ctx *c, int32_t *perm)
{
    if (!c) return TG_EINVAL;
    TG_GROUP(c, tg_sort(k, k->cfg.rank == 0 ? perm : nullptr));
    CU(cudaSetDevice(c->cfg.device));
    int rc = prepare_index(c);
    if (rc) return rc;
    // positions and Hsml unchanged: carry them into the new order
    k_carry<<<cdiv(c->n, 256), 256, 0, c->stream>>>(0, c->n, c->pw, c->hsml_in, c->posh);
    LAUNCH_CHECK();
    if ((rc = check_flags(c, false))) return rc;
    if (perm) CU(cudaMemcpy(perm, c->id, sizeof(int) * c->n, cudaMemcpyDeviceToHost));
    return TG_OK;
}

extern "C" int tg_find_ngb(tg_ctx *c, int i, float h, int32_t *list, int *count)
{
    if (!c || !list || !count || i < 0 || i >= c->n) return TG_EINVAL;
    TG_GROUP0(c, tg_find_ngb(k, i, h, list, count));
    if (!c->index_valid) return fail(c, TG_EINVAL, "tg_find_ngb: no index");
    CU(cudaSetDevice(c->cfg.device));
    if (!c->ngb_scratch) CU(dmalloc(&c->ngb_scratch, TG_NGBMAX + 1));
    int *d = c->ngb_scratch;
    { const int rc = join_defects(c); if (rc) return rc; }
    k_find_ngb<<<1, 32, 0, c->stream>>>(c->bvh, c->box, c->pw, i, h, c->defect.dmap, c->defect.nodes, d, d + TG_NGBMAX);
    c->launches++;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(count, d + TG_NGBMAX, sizeof(int), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(list, d, sizeof(int) * (*count), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(c, TG_ECUDA, "tg_find_ngb: %s", cudaGetErrorString(e));
    return TG_OK;
}

extern "C" int tg_guess_hsml(tg_ctx *c, float *out)
{
    if (!c || !out) return TG_EINVAL;
    TG_GROUP0(c, tg_guess_hsml(k, out));
    if (!c->index_valid) return fail(c, TG_EINVAL, "tg_guess_hsml: no index");
    CU(cudaSetDevice(c->cfg.device));
    const int n = c->n, T = 256;
    k_cpl<<<cdiv(n, T), T, 0, c->stream>>>(n, c->key_hi_s, c->key_lo_s, c->cpl);
    LAUNCH_CHECK();
    k_collapse_events<<<cdiv(n, T), T, 0, c->stream>>>(n, c->cpl, c->ev_start, c->ev_level, c->ev_count);
    LAUNCH_CHECK();
    k_guess_hsml<<<cdiv(n, T), T, 0, c->stream>>>(n, c->key_hi_s, c->key_lo_s, c->cpl, c->ev_start,
                                                 c->ev_level, c->ev_count, c->box.box_d, c->guess, nullptr);
    LAUNCH_CHECK();
    CU(cudaMemcpyAsync(out, c->guess, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return TG_OK;
}

// Diagnostics: candidate-box count per tile of the last tiled sweep (negative = handed back).
extern "C" int tg_debug_tile_counts(tg_ctx *c, int *out, int *ntiles)
{
    if (!c || !ntiles) return TG_EINVAL;
    TG_NOGROUP(c, "tg_debug_tile_counts");
    CU(cudaSetDevice(c->cfg.device));
    *ntiles = c->bvh.lvl_n[0];
    if (out) {
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaMemcpy(out, c->tile_ng, sizeof(int) * c->bvh.lvl_n[0], cudaMemcpyDeviceToHost));
    }
    return TG_OK;
}

extern "C" int tg_get_exchange(tg_ctx *c, tg_exchange *out)
{
    if (!c || !out) return TG_EINVAL;
    TG_NOGROUP(c, "tg_get_exchange");
    out->pos_hsml_dev = c->posh;
    out->rho_dev = c->rho;
    out->varhsml_dev = c->varh;
    out->delta_dev = c->delta;
    out->err_dev = c->scal + 1;
    out->lo = c->lo;
    out->hi = c->hi;
    out->chunk = c->chunk;
    return TG_OK;
}
