// Mirror of the reference's `poisson` driver (src/poisson.f90): -laplacian(p) = 8 pi^2 sin(2 pi x) sin(2 pi y)
// on the unit square, p = 0 on the x/y walls, solved with iccg to sor(ip) = 1e-16 within nsw(ip) = 1000
// sweeps; prints the solver report line and "h, L_inf error" (poisson.f90:104).
//   usage: poisson <n> [nz]       (n x n x nz cells; mesh_geometry is replaced by the box generator)
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "fcapp_host.hpp"

using namespace fcapp;

int main(int argc, char **argv) {
  const int n = argc > 1 ? std::atoi(argv[1]) : 40;
  const int nz = argc > 2 ? std::atoi(argv[2]) : 1;
  const double pi = 4.0 * std::atan(1.0);
  const char *kinds[6] = {"wall", "wall", "wall", "wall", "symmetry", "symmetry"};
  mesh_geometry_box(n, n, nz, 1.0, 1.0, (double)nz / n, kinds);
  fcapp_init(0);
  allocate_arrays();
  create_CSR_matrix_from_mesh_data();

  using namespace geometry;
  using namespace sparse_matrix;
  using namespace variables;
  for (int i = 0; i < numCells; ++i) su[i] = 8 * pi * pi * std::sin(2 * pi * xc[i]) * std::sin(2 * pi * yc[i]) * vol[i];
  for (int i = 0; i < numTotal; ++i) p[i] = 0.0;
  for (int i = 0; i < numCells; ++i) sv[i] = -1.0;
  laplacian(sv.data(), p.data());

  parameters::sor[parameters::ip] = (double)1e-16f;
  parameters::nsw[parameters::ip] = 1000;
  std::printf(" \n");
  iccg(p.data(), parameters::ip);
  std::printf(" \n");

  const double lh = std::fabs(xc[owner[0] - 1] - xc[neighbour[0] - 1]);
  double linf = 0.0;
  for (int i = 0; i < numCells; ++i)
    linf = std::fmax(linf, std::fabs(p[i] - std::sin(2 * pi * xc[i]) * std::sin(2 * pi * yc[i])));
  std::printf(" \n%11.4E%11.4E\n", lh, linf);
  fcapp_finalize();
  return 0;
}
