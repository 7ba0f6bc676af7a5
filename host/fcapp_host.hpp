// Host-side mirror of the reference's Fortran interface for the pressure-correction path.
//
// The reference is Fortran and this image has no Fortran compiler, so the code a maintainer
// would keep in Fortran (fortran/fcapp_shim.f90) is mirrored here in C++ with the SAME names,
// argument meaning and side effects: module-global arrays (`module geometry`,
// `sparse_matrix`, `parameters`, `variables`, `title_mod`) and free subroutines
// `create_CSR_matrix_from_mesh_data`, `laplacian(mu,phi)`, `dpcg(fi,ifi)`, `iccg(fi,ifi)`,
// `bicgstab(fi,ifi)`, `grad(phi,dPhidxi)`, `calcp()`, `calcuvw()`.  Index VALUES stay 1-based as in Fortran;
// only the C++ container subscripts are 0-based.  Everything computes through libfcapp_cuda
// (include/fcapp.h); there is no host arithmetic on the path.
#pragma once
#include <string>
#include <vector>

#include "fcapp.h"

namespace fcapp {

using dp = double;

namespace geometry {  // src/mesh_geometry_and_topology.f90:13-98
extern int numCells, numInnerFaces, numFaces, numBoundaryFaces, numTotal, nnz;
extern int ninl, nout, nsym, nwal, npru, noc;
extern int iInletFacesStart, iOutletFacesStart, iSymmetryFacesStart, iWallFacesStart, iPressOutletFacesStart,
    iOCFacesStart;
extern std::vector<int> owner, neighbour;
extern std::vector<dp> xc, yc, zc, vol, arx, ary, arz, xf, yf, zf, facint;
}  // namespace geometry

namespace sparse_matrix {  // src/sparse_matrix.f90:8-22
extern std::vector<int> ioffset, ja, diag, icell_jcell_csr_value_index, jcell_icell_csr_value_index;
extern std::vector<dp> a, su, sv, sw, spu, spv, sp, res, apu, apv, apw;
}  // namespace sparse_matrix

namespace parameters {  // src/modules_allocatable.f90:10-137
constexpr int nphi = 10;
enum { iu = 1, iv, iw, ip, ite, ied, ien, ivis, ivart, icon };  // variable identifiers (1-based like the reference)
extern dp small;                       // 1e-20 as a default-real literal
extern dp sor[nphi + 1], urf[nphi + 1], resor[nphi + 1];
extern int nsw[nphi + 1];
extern int npcor, nigrad, nipgrad, pRefCell, ncorr;
extern bool const_mflux, ltest, lstsq_qr, lstsq_dm;
extern dp flomas;
// calcuvw (src/calcuvw.f90) reads these as well
extern dp densit, viscos, gds[nphi + 1];
extern bool bdf, cn, lturb, lbuoy, boussinesq, lcal[nphi + 1];
extern dp btime, timestep, gradPcmf, beta, tref, gravx, gravy, gravz;
extern std::string convective_scheme;  // as in the `input` file: muscl-f, linear-f, central, smart, ... (read_input.f90:97-133)
}  // namespace parameters

namespace variables {  // src/modules_allocatable.f90:149-200
extern std::vector<dp> u, v, w, p, pp, den, flmass, fmi, fmo;
extern std::vector<dp> vis, uo, vo, wo, uoo, voo, woo, t;
extern std::vector<dp> dUdxi, dVdxi, dWdxi, dPdxi;  // (3,numCells): xyz interleaved
extern dp sumLocalContErr, globalContErr, cumulativeContErr;
}  // namespace variables

namespace title_mod {  // chvarSolver, src/modules_allocatable.f90:208
extern const char *chvarSolver[parameters::nphi + 1];
}

// mesh_geometry for the synthetic boxes of the benchmark configs (poisson.f90 reads a polyMesh
// instead; the array layout it leaves behind is the same)
void mesh_geometry_box(int nx, int ny, int nz, dp lx, dp ly, dp lz, const char *kinds[6]);
// mesh_geometry (src/mesh_geometry_and_topology.f90:310-1081, polyMesh branch): reads points / faces / owner /
// neighbour / boundary of an OpenFOAM polyMesh directory and computes `module geometry` (fcapp_mesh.cpp)
void mesh_geometry(const std::string &polymesh_dir);
// restart files in the reference's unformatted sequential format (fcapp_restart.cpp): write_restart_files.f90:16-62,
// readfiles.f90:12-52
void write_restart_files(const std::string &path, int itime, dp time);
void readfiles(const std::string &path, int *itime, dp *time);

void fcapp_init(int device);  // after mesh_geometry: hand `module geometry` to the GPU
void fcapp_finalize();
void allocate_arrays();       // allocate.f90: fields of `module variables`

void create_CSR_matrix_from_mesh_data();                 // sparse_matrix.f90:42
void laplacian(const dp *mu, const dp *phi);             // fvm_laplacian.f90:1
void grad(const dp *phi, dp *dPhidxi);                   // gradients.f90:95 (Gauss)
void dpcg(dp *fi, int ifi);                              // dpcg.f90:3
void iccg(dp *fi, int ifi);                              // iccg.f90:3
void bicgstab(dp *fi, int ifi);                          // bicgstab.f90:1
void calcp();                                            // calcp-multiple_correction_SIMPLE.f90:3
void calcuvw();                                          // calcuvw.f90:3 (laminar, serial)
void PISO_multiple_correction();                         // PISO_multiple_correction.f90:2
void PIMPLE_multiple_correction();                       // PIMPLE_multiple_correction.f90:2

}  // namespace fcapp
