/*
 * fcapp.h -- C ABI of libfcapp_cuda.so: the B200 (sm_100a) implementation of
 * freeCappuccino's pressure-correction path.
 *
 * Every entry point is `extern "C"`, takes plain pointers / sizes and returns
 * an int status (FC_OK = 0).  This is exactly what a Fortran ISO_C_BINDING
 * interface block binds (see fortran/fcapp_shim.f90 and INTEGRATION.md): the
 * reference keeps its subroutine signatures (`call calcp`, `call dpcg(fi,ifi)`,
 * `call laplacian(mu,phi)`, `call grad(phi,dPhidxi)`, ...) and their bodies
 * become calls into this library.
 *
 * Conventions (the reference's own): INTEGER(4) / REAL(8);
 * index arrays are 1-BASED as Fortran keeps them (`owner`, `neighbour`,
 * `ioffset`, `ja`, `diag`, ...); gradient arrays are dPhidxi(3,numCells), i.e.
 * xyz interleaved per cell; host arrays stay owned by the caller, the library
 * owns device mirrors.  There is no CPU fallback anywhere behind this header.
 *
 * Citations are relative to the reference tree (nikola-m/freeCappuccino).
 */
#ifndef FCAPP_H
#define FCAPP_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fc_context fc_context; /* opaque; one per rank / GPU */

enum {
  FC_OK = 0,
  FC_ERR_ARG = 1,         /* bad argument / call order                    */
  FC_ERR_CUDA = 2,        /* CUDA runtime failure (see fc_last_error)     */
  FC_ERR_NCCL = 3,        /* NCCL failure                                 */
  FC_ERR_UNSUPPORTED = 4, /* feature of the reference not on the GPU path */
  FC_ERR_NODEVICE = 5     /* no CUDA device: the library never falls back */
};

/* Krylov solvers of the path: src/dpcg.f90, src/iccg.f90, src/bicgstab.f90 */
enum { FC_DPCG = 0, FC_ICCG = 1, FC_BICGSTAB = 2 };

/* Device-resident fields (`module variables`, `module sparse_matrix`).
 * Sizes: *_T = numTotal, *_C = numCells, *_P = numCells+npro, G = 3*numCells. */
enum {
  FC_U = 0, FC_V, FC_W, FC_P, FC_PP, FC_DEN,          /* numTotal           */
  FC_FLMASS,                                          /* numInnerFaces      */
  FC_APU, FC_APV, FC_APW,                             /* numCells + npro    */
  FC_DUDXI, FC_DVDXI, FC_DWDXI, FC_DPDXI,             /* (3,numCells)       */
  FC_A,                                               /* nnz                */
  FC_SU, FC_RES,                                      /* numCells           */
  FC_FMI, FC_FMO,                                     /* ninl, nout         */
  FC_APR, FC_FMPRO,                                   /* npro               */
  FC_SCRATCH_T,                                       /* numTotal (user vec)*/
  FC_USER0, FC_USER1, FC_USER2, FC_USER3,             /* numTotal, caller's */
  /* momentum predictor (fc_calcuvw); allocated on first use: */
  FC_VIS, FC_UO, FC_VO, FC_WO, FC_UOO, FC_VOO, FC_WOO, FC_T, /* numTotal   */
  FC_SV, FC_SW, FC_SPU, FC_SPV, FC_SP,                /* numCells           */
  FC_NUM_FIELDS
};

/* `module geometry` (src/mesh_geometry_and_topology.f90:13-98; the parallel
 * twin adds the processor-boundary block, src-parallel/...:497-515,637-660).
 * "FacesStart" values are 0-based offsets exactly like the reference:
 * face = start + i, i = 1..count.  Field slots follow the cells in the order
 * [numCells | npro halo | inlet | outlet | symmetry | wall | prOutlet].     */
typedef struct {
  int numCells, numInnerFaces, numFaces, numTotal;
  int npro;
  int ninl, nout, nsym, nwal, npru, noc;
  int iProcFacesStart, iInletFacesStart, iOutletFacesStart, iSymmetryFacesStart,
      iWallFacesStart, iPressOutletFacesStart, iOCFacesStart;
  const int *owner;      /* [numFaces]      */
  const int *neighbour;  /* [numInnerFaces] */
  const double *xc, *yc, *zc, *vol;             /* [numCells + npro]  */
  const double *arx, *ary, *arz, *xf, *yf, *zf; /* [numFaces]         */
  const double *facint;                         /* [numInnerFaces]    */
  const double *fpro;                           /* [npro] or NULL     */
  /* processor connectivity (`process` file; my_mpi_module): */
  int numConnections;
  const int *neighbProcNo;     /* [numConnections] ranks, 0-based          */
  const int *neighbProcOffset; /* [numConnections+1], 1-based into 1..npro */
  int gloCells;                /* global cell count (ppref = mean in MPI)  */
} fc_mesh_desc;

/* ---- life cycle ------------------------------------------------------- */
int fc_create(int device, fc_context **out);
int fc_destroy(fc_context *ctx);
const char *fc_last_error(const fc_context *ctx);
int fc_version(void);

/* ---- multi-GPU plumbing: replaces MPI_COMM_WORLD of src-parallel -------
 * One rank per GPU.  Rank 0 calls fc_comm_unique_id and broadcasts the 128
 * bytes with whatever the host program has (MPI_Bcast in the Fortran MPI
 * build, torch.distributed in the Python harness).                        */
int fc_comm_unique_id(char id128[128]);
int fc_comm_init(fc_context *ctx, int rank, int nranks, const char id128[128]);
/* Optional peer-to-peer mode for the GPUs of one box (after fc_create_csr): every
 * rank publishes a blob (CUDA IPC handle of its communication arena + its
 * connection table), the host all-gathers them in rank order (MPI_Allgather /
 * torch.distributed) and hands the table back.  Afterwards the Krylov loop's halo
 * exchange and scalar reductions run as direct NVLink stores between the kernels
 * instead of NCCL calls.  Without it everything goes through NCCL.            */
#define FC_P2P_BLOB_BYTES 512
int fc_comm_p2p_blob(fc_context *ctx, char *blob /* FC_P2P_BLOB_BYTES */);
int fc_comm_p2p_open(fc_context *ctx, const char *blobs /* nranks * FC_P2P_BLOB_BYTES */, int nranks);

/* ---- mesh + CSR pattern ----------------------------------------------- */
/* Copies the geometry arrays to the device and builds the cell-to-face map. */
int fc_set_mesh(fc_context *ctx, const fc_mesh_desc *mesh);

/* create_CSR_matrix_from_mesh_data (src/sparse_matrix.f90:42-172): builds the
 * pattern on the device and, for every non-NULL pointer, returns the 1-based
 * arrays bit-identical to the reference's.  Sizes: ioffset[numCells+1],
 * ja[nnz], diag[numCells], icell_jcell/jcell_icell[numInnerFaces].          */
int fc_create_csr(fc_context *ctx, int *ioffset, int *ja, int *diag, int *icell_jcell,
                  int *jcell_icell);

/* ---- field transfer ---------------------------------------------------- */
int fc_field_size(const fc_context *ctx, int field, size_t *n);
int fc_upload(fc_context *ctx, int field, const double *host, size_t n);
int fc_download(fc_context *ctx, int field, double *host, size_t n);
int fc_fill(fc_context *ctx, int field, double value);
int fc_copy(fc_context *ctx, int src_field, int dst_field); /* device-to-device, min of the two sizes */
int fc_synchronize(fc_context *ctx);

/* ---- operators (device-resident fields) -------------------------------- */
/* y = A x over the resident CSR; x, y are field ids of size >= numCells.
 * One launch of the SpMV kernel of the Krylov loop (dpcg.f90:105-110).      */
int fc_spmv(fc_context *ctx, int x_field, int y_field);

/* grad(phi,dPhidxi) with the Gauss option (gradients.f90:95-151 ->
 * grad_gauss.f90:1-124): nigrad fixed-point passes.                         */
int fc_grad_gauss(fc_context *ctx, int phi_field, int grad_field, int nigrad);
/* grad(phi,dPhidxi,'gauss_corrected') (gradients.f90:197-259 ->
 * grad_gauss_corrected.f90): one pass seeded by the current grad_field.
 * `zero_seed` != 0 reproduces the option wrapper, which zeroes dPhidxi first
 * (gradients.f90:222).                                                      */
int fc_grad_gauss_corrected(fc_context *ctx, int phi_field, int grad_field, int zero_seed);
/* The gradient scheme of the `grad` dispatcher (gradients.f90:95-151: the lstsq /
 * lstsq_qr / lstsq_dm / gauss flags and the `limiter` string of the input file).
 * After this call fc_grad, fc_calcp, fc_calcuvw and fc_piso compute their gradients
 * with it; the geometric matrices (create_lsq_gradients_matrix, :65-90) are built
 * here.  Default: Gauss, no limiter.  SURVEY.md 8(f) rank 3; one rank.
 * lstsq_qr follows grad_lsq_qr.f90 and is defined for cells with exactly six
 * neighbours (FC_ERR_UNSUPPORTED otherwise).                                    */
enum { FC_GRAD_GAUSS = 0, FC_GRAD_LSTSQ = 1, FC_GRAD_LSTSQ_QR = 2, FC_GRAD_LSTSQ_DM = 3 };
enum { FC_LIMIT_NONE = 0, FC_LIMIT_BARTH_JESPERSEN = 1, FC_LIMIT_VENKATAKRISHNAN = 2, FC_LIMIT_MVENKATAKRISHNAN = 3 };
int fc_set_gradient(fc_context *ctx, int method, int limiter, double small);
/* grad(phi,dPhidxi) with the configured scheme + limiter.                      */
int fc_grad(fc_context *ctx, int phi_field, int grad_field, int nigrad);
/* bpres(p,istage) (bpres.f90:37-150), gradient taken from FC_DPDXI.         */
int fc_bpres(fc_context *ctx, int p_field, int istage);

/* laplacian(mu,phi) (fvm_laplacian.f90:1-163): fills FC_A and updates FC_SU
 * (wall Dirichlet part).  mu_field has >= numCells (+npro) entries.          */
int fc_laplacian(fc_context *ctx, int mu_field, int phi_field);

typedef struct {
  double sor;    /* sor(ifi): relative L1 tolerance                         */
  int nsw;       /* nsw(ifi): iteration cap                                 */
  double small;  /* `small` of module parameters: (double)1e-20f            */
  double tol;    /* early-return threshold on res0, (double)1e-13f; <0 off  */
  int parallel;  /* 1 = src-parallel arithmetic (+small in preconditioners) */
} fc_solver_opts;

typedef struct {
  double res0, resl; /* initial / final L1 residual (resor(ifi) = res0)     */
  int iters;
} fc_solver_report;

/* dpcg / iccg / bicgstab (fi, ifi): solves A fi = su with the resident FC_A,
 * FC_SU; fi is a numTotal field id (only 1..numCells is updated); FC_RES
 * holds the final residual afterwards.                                      */
int fc_solve(fc_context *ctx, int solver, int fi_field, const fc_solver_opts *o,
             fc_solver_report *rep);

/* The same under the reference's own subroutine names: dpcg(fi,ifi) (src/dpcg.f90:3),
 * iccg(fi,ifi) (src/iccg.f90:3), bicgstab(fi,ifi) (src/bicgstab.f90:1); sor(ifi), nsw(ifi)
 * travel in `o`.                                                                   */
int fc_dpcg(fc_context *ctx, int fi_field, const fc_solver_opts *o, fc_solver_report *rep);
int fc_iccg(fc_context *ctx, int fi_field, const fc_solver_opts *o, fc_solver_report *rep);
int fc_bicgstab(fc_context *ctx, int fi_field, const fc_solver_opts *o, fc_solver_report *rep);

/* Host-buffer form, the drop-in body of `subroutine dpcg(fi,ifi)` when the
 * matrix lives in the Fortran module arrays: uploads a(nnz), su(numCells),
 * fi(numTotal), solves, downloads fi(1:numCells) and res(numCells).  All
 * copies are inside the call.                                               */
int fc_solve_host(fc_context *ctx, int solver, const double *a, const double *su, double *fi,
                  double *res, const fc_solver_opts *o, fc_solver_report *rep);

/* Explicit-array form shaped like the reference's only explicit interfaces:
 * LIS solve_csr(numCells,nnz,ioffset,ja,aval,su,phi)
 * (LIS_linear_solver_library.f95:106-119) and the test copies
 * X(a,ja,ioffset,diag,nnz,numCells,numTotal,su,fi)
 * (tests/test_sparse_solvers.f90:8,233,398).  Needs no mesh: the context
 * adopts the given 1-based pattern.  `hist` (nsw doubles or NULL) receives
 * resl of every iteration, the numbers tests/output.txt prints.            */
int fc_solve_csr(fc_context *ctx, int solver, int numCells, int nnz, const int *ioffset,
                 const int *ja, const int *diag, const double *a, const double *su, double *fi,
                 const fc_solver_opts *o, fc_solver_report *rep, double *hist);

typedef struct {
  int npcor, nigrad, nipgrad; /* parameters: npcor, nigrad, nipgrad(=2)      */
  int pRefCell;               /* 1-based                                      */
  double urf_p;               /* urf(ip)                                      */
  int solver;                 /* FC_DPCG / FC_ICCG (shipped) / FC_BICGSTAB    */
  int const_mflux;            /* .true. skips adjustMassFlow                  */
  double flomas;
  int lsq_flag;               /* lstsq_qr.or.lstsq_dm: extra gauss_corrected  */
  int flux_variant;           /* 0 facefluxmass, 1 facefluxmass2, 2 _piso     */
  fc_solver_opts sol;
} fc_calcp_opts;

typedef struct {
  fc_solver_report rep[8];    /* one per pressure corrector                   */
  double sumLocalContErr, globalContErr; /* continuityErrors.h               */
} fc_calcp_report;

/* The assembly half of calcp (calcp-multiple_correction_SIMPLE.f90:34-107):
 * gradients of U,V,W, face fluxes, FC_A / FC_SU / FC_FLMASS.                */
int fc_calcp_assemble(fc_context *ctx, const fc_calcp_opts *o);
/* `call calcp`: assemble + npcor x (solve, bpres/grad, flux / velocity /
 * pressure correction) + continuity report, all on device-resident fields. */
int fc_calcp(fc_context *ctx, const fc_calcp_opts *o, fc_calcp_report *rep);
/* The post-solve half of pressure corrector `ipcorr` (1..npcor) for a host that solves the system itself between
 * fc_calcp_assemble and this call (calcp :132-223): FC_PP holds the solved correction; boundary pressure and
 * gradient of pp, flux / velocity / pressure correction, boundary velocities and -- unless ipcorr = npcor -- the
 * non-orthogonal corrector source in FC_SU for the next solve.  After the last corrector the multi-rank halo of
 * u, v, w, p is refreshed and, when `rep` is not NULL, its continuity errors are filled (continuityErrors.h).
 * fc_calcp = fc_calcp_assemble, then per corrector: FC_PP = 0, solve, fc_calcp_correct.                        */
int fc_calcp_correct(fc_context *ctx, const fc_calcp_opts *o, int ipcorr, fc_calcp_report *rep);

/* Host-buffer form of `call calcp`: uploads u,v,w,p (numTotal), apu,apv,apw
 * (numCells+npro), runs fc_calcp, downloads u,v,w,p,pp (numTotal) and flmass
 * (numInnerFaces).  den, fmi are uploaded separately (they do not change per
 * SIMPLE iteration).                                                        */
int fc_calcp_host(fc_context *ctx, const fc_calcp_opts *o, double *u, double *v, double *w,
                  double *p, double *pp, const double *apu, const double *apv, const double *apw,
                  double *flmass, fc_calcp_report *rep);

/* ---- momentum predictor: `call calcuvw` (src/calcuvw.f90:3-557) -----------
 * The step immediately before calcp.  Device-resident
 * inputs: FC_U/V/W, FC_P, FC_DEN, FC_VIS (effective viscosity `vis`), FC_FLMASS,
 * FC_FMI, FC_FMO and, for bdf / cn, FC_UO..FC_WOO; buoyancy reads FC_T.
 * Outputs: FC_U/V/W (solved), FC_APU/APV/APW = 1/(a(diag)+small), the boundary
 * slots of FC_P and FC_DPDXI (calcPressDiv, fieldManipulation.f90:82-87),
 * FC_DUDXI/DVDXI/DWDXI, FC_SV/SW/SPU/SPV/SP; FC_A / FC_SU hold the W system
 * afterwards, exactly like the module arrays of the reference.
 * Laminar form (lturb = .false.), no O-C cuts.  One rank: serial `src` semantics.
 * Several ranks (after fc_comm_init): src-parallel/calcuvw.f90 -- processor faces, the
 * running-subtraction diagonal, exchange of u, v, w, apu at the end; FC_VIS must
 * arrive with a current halo (fc_exchange); the processor faces' mass fluxes are
 * FC_FMPRO as fc_calcp leaves them.                                                  */
typedef struct {
  int nigrad, nipgrad;  /* parameters: nigrad, nipgrad (= 2)                        */
  int scheme;           /* convective scheme (read_input.f90:97-133 -> face_value,
                           interpolation.f90:36-57): 0 central, 1 cds-corrected,
                           2 central-f, 3 linear-f, 4 muscl-f, 5 flux limiter       */
  int limiter;          /* scheme 5: 0 smart, 1 avl-smart, 2 muscl, 3 umist, 4 koren,
                           5 charm, 6 ospre, 7 linear (psi = 1)                     */
  double gds;           /* gds(iu): deferred-correction blending                    */
  double urf[3];        /* urf(iu), urf(iv), urf(iw)                                */
  double sor[3];        /* sor(iu..iw)                                              */
  int nsw[3];           /* nsw(iu..iw)                                              */
  int bdf; double btime, timestep; int cn;
  int const_mflux; double gradPcmf;
  int lbuoy, boussinesq; double beta, tref, densit, gravx, gravy, gravz;
  double viscos;        /* molecular viscosity (wall faces, calcuvw.f90:320)        */
  fc_solver_opts sol;   /* small, tol, parallel (sor / nsw come from the arrays)    */
} fc_calcuvw_opts;

typedef struct {
  fc_solver_report rep[3]; /* bicgstab(u,iu), (v,iv), (w,iw)                        */
} fc_calcuvw_report;

/* calcuvw.f90:48-389: gradients, calcPressDiv, sources, face fluxes -> FC_SU/SV/SW,
 * FC_SPU/SPV/SP and the off-diagonals of FC_A.                                     */
int fc_calcuvw_assemble(fc_context *ctx, const fc_calcuvw_opts *o);
/* One component (0 u, 1 v, 2 w): diagonal, under-relaxation, ap*, bicgstab.        */
int fc_calcuvw_component(fc_context *ctx, const fc_calcuvw_opts *o, int comp, fc_solver_report *rep);
/* `call calcuvw`.                                                                  */
int fc_calcuvw(fc_context *ctx, const fc_calcuvw_opts *o, fc_calcuvw_report *rep);
/* Host-buffer form: uploads u,v,w,p,vis (numTotal) and flmass (numInnerFaces), runs
 * fc_calcuvw, downloads u,v,w,p (numTotal) and apu,apv,apw (numCells).  den, fmi,
 * fmo and the old time levels are uploaded separately with fc_upload.              */
int fc_calcuvw_host(fc_context *ctx, const fc_calcuvw_opts *o, double *u, double *v, double *w, double *p,
                    const double *vis, const double *flmass, double *apu, double *apv, double *apw,
                    fc_calcuvw_report *rep);

/* ---- PISO / PIMPLE pressure equation (src/PISO_multiple_correction.f90:2,
 * src/PIMPLE_multiple_correction.f90:2, src/get_rAU_x_UEqnH.f90:2).
 * Call after fc_calcuvw: FC_A must still hold the momentum matrix (it is backed up as
 * `h = a`), FC_APU/APV/APW the reciprocal diagonals.  Works on the pressure itself:
 * FC_PP is the unknown (not reset between correctors), FC_P receives it (PISO: p = pp;
 * PIMPLE: p += urf_p (pp - p)).  Serial semantics, one rank, no O-C cuts.          */
typedef struct {
  int ncorr, npcor, nigrad, nipgrad, pRefCell; /* pRefCell 1-based                   */
  int pimple;           /* 0 PISO_multiple_correction, 1 PIMPLE_multiple_correction   */
  double urf_p;         /* PIMPLE: urf(ip)                                            */
  int const_mflux; double flomas;
  int bdf; double btime, timestep; int cn;      /* sources of get_rAU_x_UEqnH          */
  int lbuoy, boussinesq; double beta, tref, densit, gravx, gravy, gravz;
  fc_solver_opts sol;   /* sor(ip), nsw(ip), small, tol                               */
} fc_piso_opts;

typedef struct {
  fc_solver_report rep[16]; /* iccg(pp,ip) reports in call order (the first 16)      */
  int nsolves;
  double sumLocalContErr, globalContErr; /* last continuityErrors.h report          */
} fc_piso_report;

int fc_piso(fc_context *ctx, const fc_piso_opts *o, fc_piso_report *rep);

/* ---- src-parallel communication (exchange.f90, global_sum_mpi.f90) ------ */
int fc_exchange(fc_context *ctx, int field);          /* halo of a numTotal / numPCells field */
int fc_global_sum(fc_context *ctx, double *value);    /* in-place sum over ranks (host scalar) */

/* ---- measurement helpers (bench.py): device time of the last fc_solve /
 * fc_calcp phases in milliseconds, measured with CUDA events on the library's
 * own stream.                                                               */
typedef struct {
  double solve_ms;     /* Krylov loop of the last solve                       */
  double assemble_ms;  /* gradients + face loop + row gather of last calcp    */
  double correct_ms;   /* post-solve corrections of last calcp                */
  double spmv_ms;      /* mean duration of the SpMV(+dot) launches sampled    */
                       /* inside the last solve (fc_set_spmv_sampling)        */
  int spmv_samples;    /* how many launches that mean is over                 */
  int sweep_tiles;     /* tiles of the tiled sweep schedule in use (FC_TUNE_SWEEP_TILED), */
                       /* 0 = the level schedule                              */
  long long launches;  /* kernels launched by the library since creation      */
  /* persistent DPCG kernel of the last solve (zero when the multi-kernel path ran):
   * device time of the whole kernel and of its three phases summed over the
   * iterations, each phase measured on the GPU's global timer from the grid
   * barrier that starts it to the arrival of the last CTA at the barrier that
   * ends it; persist_ms - the three = barriers, reductions, halo waits.          */
  double persist_ms, persist_pupdate_ms, persist_spmv_ms, persist_update_ms;
  int persist_iters;   /* iterations the phase sums cover                      */
  int persist_grid;    /* CTAs of the persistent kernel                        */
  double persist_mail_ms; /* of the remainder: waiting for the other ranks' partial sums */
  double uvw_assemble_ms; /* last fc_calcuvw: gradients + face + row kernels            */
  double uvw_solve_ms;    /* last fc_calcuvw: the three BiCGStab solves                 */
  int persist_index_bytes; /* bytes per non-zero the persistent kernel read for the column index:
                              1 (one-byte codes, FC_TUNE_JA_CODED) or 4 (`ja`); 0 when it did not run */
  int column_offsets;      /* distinct column offsets ja(k) - row of the pattern, 0 when more than 256  */
} fc_timings;
/* Kernel selection, for measurements and A/B tests (defaults in brackets).     */
enum {
  FC_TUNE_SPMV_KERNEL = 0,     /* stand-alone SpMV launches: 0 CSR-stream kernel,
                                  1 TMA-staged pipeline, [2] chosen by size        */
  FC_TUNE_DPCG_PERSISTENT = 1, /* 0 one launch per vector op, [1] whole DPCG loop
                                  as one persistent cooperative kernel            */
  FC_TUNE_CTAS_PER_SM = 2,     /* persistent kernel: CTAs per SM, [0] = all that fit */
  FC_TUNE_PIPE_GEOMETRY = 3,   /* TMA pipeline (threads, non-zeros staged, stages):
                                  0 256/2304/3, [1] 256/2304/2, 2 256/2048/2, 3 128/1024/2,
                                  4 256/1824/3 with one-byte codes (hexahedra only).  Pipelines with three stages
                                  gather x one chunk ahead into registers; measured slower than two stages
                                  (profiles/r02_spmv_variants.txt), so they stay options                      */
  FC_TUNE_SWEEP_P2P = 4,       /* triangular sweeps (iccg, bicgstab): [0] one counter per level,
                                  1 point-to-point flags between 128-row blocks (experimental:
                                  same row sums, bit-identical results)                      */
  FC_TUNE_FUSED_GRAD = 6,      /* grad(U), grad(V), grad(W) of calcuvw / calcp: [0] three Gauss passes, 1 one kernel
                                  per pass for the three fields (experimental; each gradient bit-identical) */
  FC_TUNE_SWEEP_CHECK = 8,     /* debugging: 1 repeats every tiled sweep with the level schedule and fails the call
                                  (FC_ERR_CUDA, first differing row in fc_last_error) if a single bit differs     */
  FC_TUNE_DPCG_FUSED = 10,     /* persistent DPCG kernel, "fused p" scheme: the product gathers p = q + bet*pold
                                  instead of a separate p-update phase (one phase and one grid barrier less per
                                  iteration, bit-identical iterates): [0] never, 1 always, 2 on partitioned meshes.
                                  Measured slower than the three-phase kernel at every size (the second gather per
                                  non-zero costs what the phase saved), so it stays an option                   */
  FC_TUNE_FACE_OCC = 11,       /* face kernel of calcp: CTAs of 256 threads per SM the register allocation must
                                  allow, 2, [3] or 4 (it is bound by the latency of its ~55 gathers per face;
                                  calcp assembly at 216^3: 2.72 / 2.45 / 2.84 ms)                               */
  FC_TUNE_L2_KEEP = 9,         /* persistent DPCG kernel: pk, zk, res, a_ii marked L2 evict_last (the matrix
                                  stream is evict_first): 0 never, 1 always, [2] when the four vectors fit  */
  FC_TUNE_JA_CODED = 13,       /* persistent DPCG kernel: the column indices travel as one-byte codes,
                                  ja[k] = row + offset[code[k]], when the pattern has at most 256 distinct column
                                  offsets (7 on a structured hexahedral block): 9 instead of 12 bytes per non-zero
                                  and product, same columns, bit-identical results: 0 off, 1 on, [2] from 2 M rows on one rank
                                  (below that the product is bound by the latency of a 256-row chunk, not its bytes) */
  FC_TUNE_DPCG_EAGER = 15,     /* persistent DPCG kernel on small partitions: fi += alf*pk runs behind the beta
                                  reduction (between a CTA's arrival at the barrier and its release) instead of inside
                                  the p-update, and the x/r update hands q = res/a_ii to the p-update (24 instead of
                                  48 bytes per row there, no division); bit-identical iterates: 0 off, 1 on,
                                  [2] up to 3.2 M rows per rank (measured: -4.6 % at 1.26 M, +1.5 % at 10 M)     */
  FC_TUNE_X_PREFETCH = 14,     /* persistent DPCG kernel: while a 256-row chunk is computed, the far x gathers of the
                                  next chunk are prefetched: 0 off, 1 into L1, 2 into L2                       */
  FC_TUNE_MAT_KEEP = 12,       /* persistent DPCG kernel: percent (0..100) of the matrix chunks whose bulk copies are
                                  marked L2 evict_last, spread evenly over every CTA's rows, so that this share of
                                  `a` / `ja` stays in the L2 from one product to the next and only the rest streams
                                  from HBM; [-1] chosen from the L2 size and the footprint of vectors and matrix  */
  FC_TUNE_TILE_CTAS = 7,       /* tiled sweeps: CTAs per SM the kernel's registers allow, [2] or 3           */
  FC_TUNE_SWEEP_TILED = 5      /* triangular sweeps: 0 one hand-over per dependency level,
                                  1 two-level schedule -- spatial tiles of <= 512 cells walked
                                  inside one CTA, hand-overs only between tile levels, [2] the
                                  same with point-to-point flags between tiles, 3 no flags at all:
                                  a row polls the VALUE of an out-of-tile dependency, 4 the flags of
                                  mode 2 with the tile staged in shared memory by 256 threads and
                                  walked branch-free by 64 (experimental;
                                  needs a mesh whose numbering is monotone across the tiles, else
                                  the level schedule stays; same row sums, bit-identical results) */
};
int fc_set_tuning(fc_context *ctx, int key, int value);
/* One line on the schedule the triangular sweeps of the last iccg / bicgstab solve used: row levels of the level
 * schedule and, with FC_TUNE_SWEEP_TILED, the tiling (tiles, tile levels, local levels) or why the mesh got none.
 * The string belongs to the context and is valid until the next call.                                          */
const char *fc_sweep_schedule_info(fc_context *ctx);
/* Bracket up to `max_samples` SpMV launches of every following solve with CUDA
 * events on the library stream (0 switches it off).                          */
int fc_set_spmv_sampling(fc_context *ctx, int max_samples);
int fc_get_timings(const fc_context *ctx, fc_timings *t);
/* Times `reps` back-to-back launches of the SpMV kernel (y = A x) with CUDA
 * events on the library stream; returns the mean in ms.                     */
int fc_time_spmv(fc_context *ctx, int x_field, int y_field, int reps, double *mean_ms);
void *fc_stream(fc_context *ctx); /* cudaStream_t of the context */

#ifdef __cplusplus
}
#endif
#endif /* FCAPP_H */
