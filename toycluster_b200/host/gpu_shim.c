/*
 * gpu_shim.c -- the reference's three hot-path entry points on top of libtoygpu.so.
 *
 * Compiled WITH the reference's own headers (globals.h, proto.h) and linked into the
 * unmodified driver in place of tree.o sph.o wvt_relax.o peano.o (keep sort.o: positions.c:409
 * uses Qsort_Index).  It defines exactly the symbols main.c:52-56 and magnetic_field.c:21
 * call, with the reference's conventions:
 *
 *   void Regularise_sph_particles();   wvt_relax.c:25
 *   void Find_sph_quantities();        sph.c:13
 *   void Bfld_from_rotA_SPH();         sph.c:216
 *   float Global_density_model(int);   wvt_relax.c:227 (proto.h:46)
 *
 *   - state lives in the driver's globals P, SphP, Param, Halo[]; only the gas range
 *     [0, Param.Npart[0]) is touched and whole records are permuted into Peano order;
 *   - no return codes: a library failure is reported like Assert() does (aux.c:57-83) --
 *     message on stderr, exit(EXIT_FAILURE);
 *   - the '#NN: Err max=...' line and the banners of wvt_relax.c:31-34,91-92,222 are
 *     printed from the library's per-iteration callback.
 * The device context is created on first use and kept for the process lifetime, like the
 * reference's static Keys/Idx/Tree buffers (peano.c:53-61, tree.c:343-346).
 */
#include "globals.h"
#include "toygpu.h"

static tg_ctx *Ctx = NULL;
static int Shim_max_iters = 1 << 30;
static unsigned Shim_flags = 0;

void toyshim_set_max_iters(int n) { Shim_max_iters = n; }      /* tests only */
void toyshim_set_flags(unsigned f) { Shim_flags = f; }          /* before first use */

static void check(int rc, const char *what)
{
    Assert(rc == TG_OK, "libtoygpu: %s failed (%d): %s", what, rc, tg_last_error(Ctx));
}

static void ensure_context(void)
{
    if (Ctx != NULL)
        return;

    tg_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.device = getenv("TOYGPU_DEVICE") ? atoi(getenv("TOYGPU_DEVICE")) : 0;
    cfg.n_gas = Param.Npart[0];
    cfg.boxsize = Param.Boxsize;
    cfg.mpart_gas = Param.Mpart[0];
    cfg.mtotal = Param.Mtotal;
    cfg.flags = Shim_flags;
    if (getenv("TOYGPU_FLAGS"))      /* e.g. 1 = TG_WVT_SEQUENTIAL, 2 = TG_EXACT_NEIGHBOURS, 4 = TG_FAST */
        cfg.flags |= (unsigned)strtoul(getenv("TOYGPU_FLAGS"), NULL, 0);
    cfg.rank = 0;
    cfg.nranks = 1;
    /* TOYGPU_NGPUS=8: this one process drives 8 devices; the library partitions the targets,
     * exchanges the slices over NVLink (NCCL) and reduces the error statistics itself */
    cfg.ngpus = getenv("TOYGPU_NGPUS") ? atoi(getenv("TOYGPU_NGPUS")) : 0;
    cfg.devices = NULL;
#ifdef DOUBLE_BETA_COOL_CORES                  /* Makefile:15; setup.c:604-612 */
    cfg.rho0_fac = Param.Rho0_Fac;
    cfg.rc_fac = Param.Rc_Fac;
#endif

    int rc = tg_create(&Ctx, &cfg);
    Assert(rc == TG_OK, "libtoygpu: tg_create failed (%d): %s", rc, tg_last_error(NULL));

    tg_halo *rows = Malloc(Param.Nhalos * sizeof *rows);
    for (int i = 0; i < Param.Nhalos; i++) {
        for (int k = 0; k < 3; k++)
            rows[i].dcom[k] = Halo[i].D_CoM[k];
        rows[i].rho0 = Halo[i].Rho0;
        rows[i].beta = Halo[i].Beta;
        rows[i].rcore = Halo[i].Rcore;
        rows[i].rcut = Halo[i].Rcut;
        rows[i].cuspy = Halo[i].Have_Cuspy;
        rows[i].mass_gas = Halo[i].Mass[0];
    }
    check(tg_set_halos(Ctx, Param.Nhalos, rows), "tg_set_halos");
    Free(rows);

    /* P and SphP are allocated once by Setup() (setup.c:244-250) and cross the bus in every
     * operator call: page-lock the gas range so those copies are DMA.  Best effort. */
    if (!getenv("TOYGPU_NO_PIN")) {
        (void)tg_pin_host(Ctx, P, (size_t)Param.Npart[0] * sizeof *P);
        (void)tg_pin_host(Ctx, SphP, (size_t)Param.Npart[0] * sizeof *SphP);
    }
}

void Find_sph_quantities()
{
    ensure_context();
    check(tg_upload(Ctx, P, sizeof *P, SphP, sizeof *SphP), "tg_upload");
    check(tg_find_sph_quantities(Ctx), "tg_find_sph_quantities");
    check(tg_download(Ctx, P, sizeof *P, SphP, sizeof *SphP), "tg_download");
}

static int print_iteration(int it, double err_max, double err_mean, double err_diff,
                           double step, void *user)
{
    (void)user;
    printf("   #%02d: Err max=%3g mean=%03g diff=%03g"
           " step=%g\n", it, err_max, err_mean, err_diff, step);
    return 0;
}

void Regularise_sph_particles()
{
    printf("Starting iterative SPH regularisation \n"
           "   max %d iterations, tree update every %d iterations\n"
           "   stop at  errmax < %g%%   \n\n", TG_NUMITER, 1, 0.01 * 100);
    fflush(stdout);

    ensure_context();
    int done = 0;
    check(tg_upload(Ctx, P, sizeof *P, SphP, sizeof *SphP), "tg_upload");
    check(tg_regularise(Ctx, Shim_max_iters, &print_iteration, NULL, &done), "tg_regularise");
    check(tg_download(Ctx, P, sizeof *P, SphP, sizeof *SphP), "tg_download");

    printf("\ndone\n\n");
    fflush(stdout);
}

void Bfld_from_rotA_SPH()
{
    printf("Constructing B from rot(A)");
    fflush(stdout);

    ensure_context();
    const int n = Param.Npart[0];
    float *buf = Malloc((size_t)3 * n * sizeof *buf);

    /* Apot was set by the driver on the records in their current order
     * (magnetic_field.c:33-69), which is the library's current order too */
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++)
            buf[3 * (size_t)i + k] = SphP[i].Apot[k];
    check(tg_set_apot(Ctx, buf), "tg_set_apot");
    check(tg_bfld_from_rotA(Ctx), "tg_bfld_from_rotA");
    check(tg_download_soa(Ctx, NULL, NULL, NULL, NULL, NULL, NULL, buf), "tg_download_soa");
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++)
            SphP[i].Bfld[k] = buf[3 * (size_t)i + k];
    Free(buf);

    printf(" done \n\n");
    fflush(stdout);
}

#ifdef TOYGPU_SHIM_MAGNETIC_FIELD
/* magnetic_field.c:12-31: with this defined the shim also replaces
 * magnetic_field.o -- vector potential, rot(A), normalisation and cap run on the device in one
 * call, and Apot / Bfld cross the bus once. */
void Make_magnetic_field()
{
    printf("Magnetic field: \n"
           "   B0              = %g G\n"
           "   eta             = %g \n\n", Param.Bfld_Norm, Param.Bfld_Eta);
    printf("Constructing B from rot(A)");
    fflush(stdout);

    ensure_context();
    const int n = Param.Npart[0], nh = Param.Nhalos;
    double *rs_gas = Malloc(nh * sizeof *rs_gas), *rs_dm = Malloc(nh * sizeof *rs_dm);
    int *stripped = Malloc(nh * sizeof *stripped);
    for (int j = 0; j < nh; j++) {
        rs_gas[j] = Halo[j].R_Sample[0];
        rs_dm[j] = Halo[j].R_Sample[1];
        stripped[j] = Halo[j].Is_Stripped;
    }
    tg_bfield par;
    memset(&par, 0, sizeof par);
    par.bfld_norm = Param.Bfld_Norm;
    par.bfld_eta = Param.Bfld_Eta;
    par.bmax_main = 18e-6;                   /* BMAX, magnetic_field.c:4 */
    par.bmax_sub = 2e-6;                     /* magnetic_field.c:113 */
    par.sub_first = Sub.First;
    par.r_sample_gas = rs_gas;
    par.r_sample_dm = rs_dm;
    par.is_stripped = stripped;
    double norm = 0;
    int limited = 0;
    check(tg_make_magnetic_field(Ctx, &par, &norm, &limited), "tg_make_magnetic_field");

    float *buf = Malloc((size_t)3 * n * sizeof *buf);
    check(tg_get_apot(Ctx, buf), "tg_get_apot");
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++)
            SphP[i].Apot[k] = buf[3 * (size_t)i + k];
    check(tg_download_soa(Ctx, NULL, NULL, NULL, NULL, NULL, NULL, buf), "tg_download_soa");
    for (int i = 0; i < n; i++)
        for (int k = 0; k < 3; k++)
            SphP[i].Bfld[k] = buf[3 * (size_t)i + k];
    Free(buf); Free(rs_gas); Free(rs_dm); Free(stripped);

    printf(" done \n\n");
    printf("Bfld Norm = %g \n", norm);
    printf("Bfld of %d particles limited to %g G\n", limited, 18e-6);
    fflush(stdout);
}
#endif

#ifdef TOYGPU_SHIM_REASSIGN
/* positions.c:264-324.  With this defined the shim also replaces
 * Reassign_particles_to_halos (positions.c is then compiled with
 * -DReassign_particles_to_halos=<unused name>, see oracle/Makefile): the per-particle
 * Halo_containing -- an FP64 pow per halo and particle, serial in the reference -- runs on the
 * device over the positions it already holds.  The index sort of the ids is the reference's own
 * Qsort_Index (sort.c:185-195 -> gsl_heapsort_index): its order of equal ids is what the Gadget
 * file's particle order is made of, so it is not replaced; the permutation and the bookkeeping
 * below restate positions.c:285-323, 405-443. */
int compare_int(const void *a, const void *b);        /* positions.c:391 */

void Reassign_particles_to_halos()
{
    ensure_context();
    const int n = Param.Npart[0], nh = Param.Nhalos;
    int *haloID = Malloc((size_t)n * sizeof *haloID);
    long long *npart = Malloc(nh * sizeof *npart);
    double *rs_gas = Malloc(nh * sizeof *rs_gas);
    int *stripped = Malloc(nh * sizeof *stripped);
    for (int j = 0; j < nh; j++) {
        rs_gas[j] = Halo[j].R_Sample[0];
        stripped[j] = Halo[j].Is_Stripped;
    }
    tg_bfield par;
    memset(&par, 0, sizeof par);
    par.sub_first = Sub.First;
    par.r_sample_gas = rs_gas;
    par.is_stripped = stripped;
    /* the device state is the one Find_sph_quantities / Make_magnetic_field left: same order
     * and positions as P[] */
    check(tg_halo_ids(Ctx, &par, haloID, npart), "tg_halo_ids");

    /* sort_particles(), positions.c:405-443 */
    size_t *idx = Malloc((size_t)n * sizeof *idx);
    Qsort_Index(Omp.NThreads, idx, haloID, n, sizeof *haloID, &compare_int);
#ifdef TOYGPU_SHIM_IO
    /* the file order: record idx[k] of the device's order goes to position k */
    check(tg_set_output_order(Ctx, idx), "tg_set_output_order");
#endif
    for (size_t i = 0; i < (size_t)n; i++) {
        if (idx[i] == i)
            continue;
        size_t dest = i;
        struct ParticleData Ptmp = P[dest];
        struct GasParticleData SphPtmp = SphP[dest];
        size_t src = idx[i];
        for (;;) {
            memcpy(&P[dest], &P[src], sizeof *P);
            memcpy(&SphP[dest], &SphP[src], sizeof *SphP);
            idx[dest] = dest;
            dest = src;
            src = idx[dest];
            if (src == i)
                break;
        }
        memcpy(&P[dest], &Ptmp, sizeof *P);
        memcpy(&SphP[dest], &SphPtmp, sizeof *SphP);
        idx[dest] = dest;
    }
    Free(idx); Free(haloID);

    /* positions.c:289-323 */
    Sub.Ntotal = Sub.Npart[1];
    Sub.Npart[0] = 0;
    for (int i = 0; i < nh; i++) {
        Halo[i].Npart[0] = npart[i];
        Halo[i].Ntotal = Halo[i].Npart[0] + Halo[i].Npart[1];
        if (i >= Sub.First) {
            Sub.Ntotal += npart[i];
            Sub.Npart[0] += npart[i];
        }
    }
    int iGas = Halo[0].Npart[0];
    for (int i = 1; i < nh; i++) {
        Halo[i].Gas = &(P[iGas]);
        Halo[i].SphP = &(SphP[iGas]);
        iGas += Halo[i].Npart[0];
    }
    printf("Particle Distribution after Relaxation :\n"
           "   Main     %8lld   %8lld   %8lld  \n",
           Halo[0].Ntotal, Halo[0].Npart[0], Halo[0].Npart[1]);
    if (Param.Mass_Ratio)
        printf("   Bullet   %8lld   %8lld   %8lld  \n",
               Halo[1].Ntotal, Halo[1].Npart[0], Halo[1].Npart[1]);
#ifdef SUBSTRUCTURE
    printf("   Subhalos %8d   %8d   %8d \n", Sub.Ntotal, Sub.Npart[0], Sub.Npart[1]);
#endif
    Free(npart); Free(rs_gas); Free(stripped);
}
#endif

#ifdef TOYGPU_SHIM_IO
#if !defined(TOYGPU_SHIM_MAGNETIC_FIELD) || !defined(TOYGPU_SHIM_REASSIGN)
#error "TOYGPU_SHIM_IO needs the device to hold the final Bfld and the file order: build with -DTOYGPU_SHIM_MAGNETIC_FIELD -DTOYGPU_SHIM_REASSIGN"
#endif
/* io.c:85-133.  With this defined the shim also replaces add_block(): the write
 * buffers of the blocks the path owns -- the gas range of POS, and RHO, HSML, BFLD, RHOM -- come
 * straight from the device's SoA arrays in the file order Reassign_particles_to_halos() left
 * (tg_set_output_order above), without the per-particle fill_write_buffer() loop over the AoS
 * records.  Nothing after the path changes those fields (Make_temperatures writes U,
 * Make_velocities / Apply_kinematics write Vel).  The other particle types of POS and the
 * blocks VEL, ID, U are filled from the driver's records as before; header, block framing and
 * everything else are io.c's own (Write_output, write_header, set_block_info, my_fwrite).
 * io.o is linked with its add_block weakened (objcopy -W, see oracle/Makefile). */
#include "io.h"

void add_block(FILE *fp, enum iofields iblock)
{
    set_block_info(iblock);

    printf("   Block %d (%s)\n", iblock, Block.Name);

    const size_t vals = Block.Val_per_element;
    const size_t nData = Block.Ntot * vals * Block.Bytes_per_element;
    void *write_buffer = Malloc(nData);

    int dev = -1;
    switch (iblock) {
    case IO_POS: dev = TG_BLOCK_POS; break;
    case IO_RHO: dev = TG_BLOCK_RHO; break;
    case IO_HSML: dev = TG_BLOCK_HSML; break;
    case IO_BFLD: dev = TG_BLOCK_BFLD; break;
    case IO_RHOMODEL: dev = TG_BLOCK_RHOMODEL; break;
    default: break;
    }

    /* tests: prove that the device blocks do not read the driver's records */
    if (iblock == IO_POS && getenv("TOYSHIM_POISON_RECORDS"))
        for (int i = 0; i < Param.Npart[0]; i++) {
            P[i].Pos[0] = P[i].Pos[1] = P[i].Pos[2] = NAN;
            SphP[i].Rho = SphP[i].Hsml = SphP[i].Rho_Model = NAN;
            SphP[i].Bfld[0] = SphP[i].Bfld[1] = SphP[i].Bfld[2] = NAN;
        }

    size_t first = 0;                        /* first particle that comes from the records */
    if (dev >= 0 && Ctx != NULL && Block.Npart[0] > 0) {
        check(tg_fill_block(Ctx, dev, write_buffer), "tg_fill_block");
        first = Block.Npart[0];
    }
    /* io.c:100-114: the six type ranges are consecutive */
    for (size_t ipart = first, ibuf = first * vals; ipart < (size_t)Block.Ntot; ipart++, ibuf += vals)
        fill_write_buffer(iblock, write_buffer, ipart, ibuf);

    int blocksize = sizeof(int) + 4 * sizeof(char);          /* F90 records, io.c:116-129 */
    my_fwrite(&blocksize, sizeof(int), 1, fp);
    my_fwrite(&Block.Label, sizeof(char), 4, fp);
    int nextblock = nData + 2 * sizeof(int);
    my_fwrite(&nextblock, sizeof(int), 1, fp);
    my_fwrite(&blocksize, sizeof(int), 1, fp);

    blocksize = nData;
    my_fwrite(&blocksize, sizeof(int), 1, fp);
    my_fwrite(write_buffer, blocksize, 1, fp);
    my_fwrite(&blocksize, sizeof(int), 1, fp);

    free(write_buffer);
}
#endif

/* wvt_relax.c:227-256; host-side, for callers outside the path (proto.h:46). */
float Global_density_model(const int ipart)
{
    const double boxhalf = Param.Boxsize * 0.5;
    const double x = P[ipart].Pos[0], y = P[ipart].Pos[1], z = P[ipart].Pos[2];
    double rho = 0;

    for (int i = 0; i < Param.Nhalos; i++) {
        if (Halo[i].Mass[0] == 0)
            continue;
        const double dx = x - Halo[i].D_CoM[0] - boxhalf;
        const double dy = y - Halo[i].D_CoM[1] - boxhalf;
        const double dz = z - Halo[i].D_CoM[2] - boxhalf;
        const double rho_i = Gas_density_profile(sqrt(dx * dx + dy * dy + dz * dz), Halo[i].Rho0,
                                Halo[i].Beta, Halo[i].Rcore, Halo[i].Rcut, Halo[i].Have_Cuspy);
        rho = fmax(rho_i, rho);
    }
    return rho;
}
