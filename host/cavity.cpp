// Mirror of the reference's main program for the shipped lid-driven cavity (src/main.f90:118-170 with
// examples/cavity/input): SIMPLE outer iterations  call calcuvw ; call calcp  until
// max(resor(iu), resor(iv), resor(iw), resor(ip)) < sormax or maxit, on an n x n x 1 box whose y+ wall moves
// with U = (1,0,0) (examples/cavity/0/U); z faces are symmetry planes.  Prints the solver report lines in the
// reference's format and, last, "iterations source umax".
//   usage: cavity <n | polyMesh dir> [maxit] [sormax]
// With a directory the mesh is read by mesh_geometry (fcapp_mesh.cpp), e.g. the polyMesh of
// examples/cavity/cavity-setup.tar.gz; with a number the n x n x 1 box is generated.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "fcapp_host.hpp"

using namespace fcapp;

int main(int argc, char **argv) {
  const bool from_files = argc > 1 && std::atoi(argv[1]) <= 0;
  const int n = argc > 1 && !from_files ? std::atoi(argv[1]) : 20;
  const int maxit = argc > 2 ? std::atoi(argv[2]) : 1000;
  const double sormax = argc > 3 ? std::atof(argv[3]) : 1e-6;
  const char *kinds[6] = {"wall", "wall", "wall", "wall", "symmetry", "symmetry"};
  if (from_files) mesh_geometry(argv[1]);
  else mesh_geometry_box(n, n, 1, 0.1, 0.1, 0.01, kinds);   // the shipped mesh is 0.1 x 0.1 x 0.01, 20 x 20 x 1
  using namespace geometry;
  using namespace parameters;
  using namespace variables;
  // examples/cavity/input
  densit = 1.0; viscos = 0.01;
  for (int i = 1; i <= nphi; ++i) { gds[i] = 1.0; urf[i] = 0.7; sor[i] = 1e-2; nsw[i] = 5; resor[i] = 0.0; lcal[i] = false; }
  gds[ip] = 0.0; urf[ip] = 0.3;
  nsw[iu] = nsw[iv] = nsw[iw] = 20; nsw[ip] = 100;
  bdf = true; btime = 0.0; timestep = 1e20; cn = false;
  convective_scheme = "muscl-f";
  npcor = 1; nigrad = 1; pRefCell = 1; const_mflux = false; flomas = 0.0;
  fcapp_init(0);
  allocate_arrays();
  create_CSR_matrix_from_mesh_data();
  // lid: wall faces whose normal points in +y
  const int iWallStart = numCells + ninl + nout + nsym;
  for (int i = 0; i < nwal; ++i)
    if (ary[iWallFacesStart + i] > 0.0) u[iWallStart + i] = 1.0;
  int iter = 0;
  double source = 0.0;
  for (iter = 1; iter <= maxit; ++iter) {
    calcuvw();
    calcp();
    source = std::max(std::max(resor[iu], resor[iv]), std::max(resor[iw], resor[ip]));   // main.f90:157
    if (source < sormax) break;
  }
  double umax = 0.0;
  for (int i = 0; i < numCells; ++i) umax = std::fmax(umax, std::fabs(u[i]));
  std::printf(" \n%d %11.4E %11.4E\n", std::min(iter, maxit), source, umax);
  fcapp_finalize();
  return 0;
}
