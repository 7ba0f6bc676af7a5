/*
 * fc_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, sequential sums, no FMA contraction) of the
 * pressure-correction hot path of nikola-m/freeCappuccino.  It is the checker
 * the CUDA path is compared against; only tests/, __graft_entry__.smoke() and
 * the cpu_baseline / --impl reference legs of bench.py may load it.  The
 * product library (freecappuccino_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED for the three Krylov solvers by the reference's own
 * golden file tests/output.txt (every printed digit, see
 * tests/test_oracle_golden.py); the assembly / gradient routines have no
 * stored outputs in the reference -- they are pinned analytically (Poisson
 * second-order convergence, poisson.f90) and by exact-arithmetic identities.
 *
 * All index arrays are 1-based INTEGER(4) exactly as the Fortran code keeps
 * them; all reals are REAL(8).  Citations are relative to /root/reference.
 */
#ifndef FC_ORACLE_H
#define FC_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Mesh description = the arrays of `module geometry`
 * (src/mesh_geometry_and_topology.f90:13-98; src-parallel twin adds npro). */
typedef struct {
  int numCells, numInnerFaces, numFaces, numTotal;
  int npro;  /* processor-boundary faces (src-parallel only, else 0) */
  int ninl, nout, nsym, nwal, npru, noc;
  /* 0-based "start" offsets exactly like the reference: face = start + i, i=1.. */
  int iProcFacesStart, iInletFacesStart, iOutletFacesStart, iSymmetryFacesStart,
      iWallFacesStart, iPressOutletFacesStart, iOCFacesStart;
  const int *owner;      /* [numFaces]      */
  const int *neighbour;  /* [numInnerFaces] */
  const double *xc, *yc, *zc, *vol;             /* [numCells (+npro)] */
  const double *arx, *ary, *arz, *xf, *yf, *zf; /* [numFaces]         */
  const double *facint;                         /* [numInnerFaces]    */
  const double *fpro;                           /* [npro]             */
  const int *ijl, *ijr, *ijlFace;               /* [noc]              */
  const double *foc;                            /* [noc]              */
} fco_mesh;

/* CSR container = `module sparse_matrix` (src/sparse_matrix.f90:8-22). */
typedef struct {
  int n, nnz;
  const int *ioffset, *ja, *diag;       /* 1-based */
  const int *icell_jcell, *jcell_icell; /* [numInnerFaces], may be NULL for solvers */
} fco_csr;

typedef struct {
  double sor;    /* resmax = sor(ifi)                       */
  int nsw;       /* max sweeps = nsw(ifi)                   */
  double small;  /* (double)1e-20f in src, (double)1e-30f in tests/ */
  double tol;    /* early-return threshold, (double)1e-13f; <0 disables (tests/) */
  int parallel;  /* 1 = src-parallel arithmetic (+small in preconditioners)   */
} fco_solver_opts;

typedef struct {
  double res0, resl;
  int iters;
} fco_report;

int fco_create_csr(int numCells, int numInnerFaces, const int *owner, const int *neighbour,
                   int *ioffset, int *ja, int *diag, int *icell_jcell, int *jcell_icell);

void fco_spmv(const fco_csr *m, const double *a, const double *x, double *y);

/* O-C strips (al, ar, ijl, ijr, noc) and processor strips (apr, npro, owner of
 * proc faces `pown`, halo values live in fi[iProcStart + i]) may be empty.   */
typedef struct {
  int noc;  const int *ijl, *ijr;  const double *al, *ar;
  int npro; const int *pown;       const double *apr;  int iProcStart;
} fco_strips;

int fco_dpcg(const fco_csr *m, const double *a, const double *su, double *fi, double *res,
             const fco_strips *s, const fco_solver_opts *o, fco_report *rep, double *hist);
int fco_iccg(const fco_csr *m, const double *a, const double *su, double *fi, double *res,
             const fco_strips *s, const fco_solver_opts *o, fco_report *rep, double *hist);
int fco_bicgstab(const fco_csr *m, const double *a, const double *su, double *fi, double *res,
                 const fco_strips *s, const fco_solver_opts *o, fco_report *rep, double *hist);

void fco_laplacian(const fco_mesh *g, const fco_csr *m, const double *mu, const double *phi,
                   double *a, double *su, double *al, double *ar);

void fco_grad_gauss(const fco_mesh *g, const double *u, int nigrad, double *dudxi /* (3,numCells) */);
void fco_grad_gauss_corrected(const fco_mesh *g, const double *u, double *dudxi);
void fco_bpres(const fco_mesh *g, double *p, const double *dPdxi, int istage);

/* Fields of `module variables` + `sparse_matrix` touched by calcp. */
typedef struct {
  double *u, *v, *w, *p, *pp;           /* [numTotal] */
  const double *den;                    /* [numTotal] */
  double *flmass;                       /* [numInnerFaces] */
  double *fmi, *fmo, *fmoc;             /* [ninl],[nout],[noc] */
  double *dUdxi, *dVdxi, *dWdxi, *dPdxi;/* (3,numCells) */
  const double *apu, *apv, *apw;        /* [numCells] */
  double *a, *su, *res, *al, *ar;       /* CSR values, rhs, residual, O-C coefs */
} fco_fields;

typedef struct {
  int npcor, nigrad, nipgrad;   /* parameters :114-117 (nipgrad = 2)         */
  int pRefCell;                 /* 1-based                                    */
  double urf_p;                 /* urf(ip)                                    */
  int solver;                   /* 0 dpcg, 1 iccg (shipped), 2 bicgstab       */
  int const_mflux;              /* .true. skips adjustMassFlow                */
  double flomas;
  int lsq_flag;                 /* lstsq_qr.or.lstsq_dm: extra gauss_corrected*/
  int flux_variant;             /* 0 facefluxmass, 1 facefluxmass2, 2 _piso   */
  fco_solver_opts sol;
} fco_calcp_opts;

typedef struct {
  fco_report rep[8];            /* one per pressure corrector                 */
  double sumLocalContErr, globalContErr;
} fco_calcp_report;

void fco_calcp_assemble(const fco_mesh *g, const fco_csr *m, fco_fields *f, const fco_calcp_opts *o);
int  fco_calcp(const fco_mesh *g, const fco_csr *m, fco_fields *f, const fco_calcp_opts *o,
               fco_calcp_report *rep);

void fco_facefluxmass(const fco_mesh *g, const fco_fields *f, int variant, int ijp, int ijn,
                      double xf, double yf, double zf, double arx, double ary, double arz,
                      double lambda, double *cap, double *can, double *fluxmass);
void fco_fluxmc(const fco_mesh *g, const fco_fields *f, int ijp, int ijn,
                double xf, double yf, double zf, double arx, double ary, double arz,
                double lambda, double *fmcor);

/* ---- momentum predictor calcuvw (fc_oracle_uvw.c) ---- */
typedef struct {
  const double *vis;                          /* [numTotal] effective viscosity                    */
  const double *uo, *vo, *wo, *uoo, *voo, *woo; /* [numTotal] previous time levels (bdf / cn)      */
  const double *t;                            /* [numTotal] temperature (buoyancy) or NULL         */
  double *sv, *sw, *spu, *spv, *sp;           /* [numCells] module sparse_matrix                   */
  double *apu, *apv, *apw;                    /* [numCells] written: 1/(a(diag)+small)             */
} fco_uvw;

typedef struct {
  int nigrad, nipgrad;
  int scheme;        /* 0 central, 1 cds-corrected, 2 central-f, 3 linear-f, 4 muscl-f, 5 flux limiter   */
  int limiter;       /* scheme 5: 0 smart, 1 avl-smart, 2 muscl, 3 umist, 4 koren, 5 charm, 6 ospre, 7 linear */
  double gds;        /* gds(iu)                                                                      */
  double urf[3];     /* urf(iu), urf(iv), urf(iw)                                                    */
  double sor[3];     /* sor(iu..iw)                                                                  */
  int nsw[3];        /* nsw(iu..iw)                                                                  */
  int bdf; double btime, timestep; int cn;
  int const_mflux; double gradPcmf;
  int lbuoy, boussinesq; double beta, tref, densit, gravx, gravy, gravz;
  double viscos;
  fco_solver_opts sol; /* small, tol, parallel (sor / nsw taken from the arrays above)               */
} fco_uvw_opts;

typedef struct { fco_report rep[3]; } fco_uvw_report;

void fco_facefluxuvw(const fco_mesh *g, const fco_fields *f, const fco_uvw *x, const fco_uvw_opts *o, int ijp,
                     int ijn, double xf, double yf, double zf, double arx, double ary, double arz, double flomass,
                     double lambda, double gam, double *cap, double *can, double *sup, double *svp, double *swp);
int fco_calcuvw_assemble(const fco_mesh *g, const fco_csr *m, fco_fields *f, fco_uvw *x, const fco_uvw_opts *o);
int fco_calcuvw_component(const fco_mesh *g, const fco_csr *m, fco_fields *f, fco_uvw *x, const fco_uvw_opts *o,
                          int comp, fco_report *rep);
int fco_calcuvw(const fco_mesh *g, const fco_csr *m, fco_fields *f, fco_uvw *x, const fco_uvw_opts *o,
                fco_uvw_report *rep);

/* ---- PISO / PIMPLE pressure equation (fc_oracle_piso.c) ---- */
typedef struct {
  int ncorr, npcor, nigrad, nipgrad, pRefCell;
  int pimple;        /* 0 PISO_multiple_correction, 1 PIMPLE_multiple_correction */
  double urf_p;      /* PIMPLE: urf(ip) */
  int const_mflux; double flomas;
  int bdf; double btime, timestep; int cn;
  int lbuoy, boussinesq; double beta, tref, densit, gravx, gravy, gravz;
  fco_solver_opts sol; /* sor(ip), nsw(ip) */
} fco_piso_opts;

typedef struct {
  fco_report rep[16];  /* iccg reports in call order (ncorr x npcor, the first 16) */
  int nsolves;
  double sumLocalContErr, globalContErr; /* last continuityErrors.h report */
} fco_piso_report;

void fco_get_rAU_x_UEqnH(const fco_mesh *g, const fco_csr *m, fco_fields *f, fco_uvw *x, const fco_piso_opts *o,
                         const double *h);
/* h: scratch of nnz doubles, receives the copy of the momentum matrix (module hcoef) */
int fco_piso(const fco_mesh *g, const fco_csr *m, fco_fields *f, fco_uvw *x, const fco_piso_opts *o, double *h,
             fco_piso_report *rep);

/* ---- least-squares gradients + slope limiters behind `grad` (fc_oracle_grad.c) ---- */
typedef struct {
  int method;      /* 0 gauss, 1 lstsq, 2 lstsq_qr, 3 lstsq_dm  (gradients.f90:107-128)                       */
  int limiter;     /* 0 no-limit, 1 Barth-Jespersen, 2 Venkatakrishnan, 3 mVenkatakrishnan (:133-148)         */
  const double *dmat;   /* (9,numCells) from fco_lsq_matrix, methods 1 and 3                                  */
  const double *dmatqr; /* (3,6,numCells) from fco_lsq_qr_matrix, method 2                                    */
  double small;
} fco_gradient_cfg;
void fco_lsq_matrix(const fco_mesh *g, int weighted, double *dmat);
void fco_grad_lsq(const fco_mesh *g, int weighted, const double *dmat, const double *fi, double *dFidxi);
int  fco_lsq_qr_matrix(const fco_mesh *g, double *D);
void fco_grad_lsq_qr(const fco_mesh *g, const double *D, const double *fi, double *dFidxi);
void fco_slope_limiter(const fco_mesh *g, const fco_csr *m, int which, const double *phi, double *dPhidxi, double small);
void fco_slope_limiter_par(const fco_mesh *g, const fco_csr *m, int which, const double *phi, double *dPhidxi, double small,
                           double glomin, double glomax);
/* process-wide gradient configuration used by fco_grad and, through it, by calcp / calcuvw / piso (NULL = gauss) */
void fco_set_gradient(const fco_gradient_cfg *c);
void fco_grad(const fco_mesh *g, const fco_csr *m, const double *phi, int nigrad, double *dPhidxi);
void fco_limit_configured(const fco_mesh *g, const fco_csr *m, const double *phi, double *dPhidxi);

/* ---- src-parallel semantics: R ranks in lock step inside one process (fc_oracle_par.c) ---- */
typedef struct {
  fco_mesh g;
  fco_csr m;
  fco_fields f;                 /* gradients (3,numCells+npro), apu.. numCells+npro, u.. numTotal */
  double *apr, *fmpro;          /* [npro] */
  int numConnections;
  const int *neighbProcNo;      /* 0-based ranks */
  const int *neighbProcOffset;  /* 1-based, numConnections+1 entries */
} fco_rank;

void fco_par_exchange(fco_rank *R, int nr, double **phi, int stride);
void fco_par_grad_gauss(fco_rank *R, int nr, double **phi, int nigrad, double **grad);
void fco_par_grad_gauss_corrected(fco_rank *R, int nr, double **phi, double **grad);
void fco_par_laplacian(fco_rank *R, int nr, double **mu, double **phi);
/* run the ranks of every lock-step phase of fco_par_solve / fco_par_exchange on up to n host threads
 * (bit-identical results); fco_par_openmp() = 0 when the library was built without OpenMP */
void fco_par_set_threads(int n);
int fco_par_openmp(void);
int fco_par_solve(fco_rank *R, int nr, int solver, double **fi, const fco_solver_opts *o, fco_report *rep,
                  double *hist);
void fco_par_calcp_assemble(fco_rank *R, int nr, const fco_calcp_opts *o);
int fco_par_calcp(fco_rank *R, int nr, const fco_calcp_opts *o, fco_calcp_report *rep);
/* src-parallel/calcuvw.f90 in lock step (fc_oracle_par_uvw.c); X = one fco_uvw per rank, apu.. of numCells+npro */
void fco_par_calcuvw_assemble(fco_rank *R, int nr, fco_uvw *X, const fco_uvw_opts *o);
int fco_par_calcuvw_component(fco_rank *R, int nr, fco_uvw *X, const fco_uvw_opts *o, int comp, fco_report *rep);
int fco_par_calcuvw(fco_rank *R, int nr, fco_uvw *X, const fco_uvw_opts *o, fco_uvw_report *rep);

#ifdef __cplusplus
}
#endif
#endif
