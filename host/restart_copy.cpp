// Reads a restart file of the reference's format into the host mirror's module arrays and writes it out again
// (fcapp::readfiles / fcapp::write_restart_files).  No GPU needed; used by tests/test_restart_io.py.
//   usage: restart_copy <numCells> <numInnerFaces> <numTotal> <in> <out>
#include <cstdio>
#include <cstdlib>
#include <exception>

#include "fcapp_host.hpp"

using namespace fcapp;

int main(int argc, char **argv) {
  if (argc < 6) { std::fprintf(stderr, "usage: restart_copy <numCells> <numInnerFaces> <numTotal> <in> <out>\n"); return 2; }
  geometry::numCells = std::atoi(argv[1]);
  geometry::numInnerFaces = std::atoi(argv[2]);
  geometry::numTotal = std::atoi(argv[3]);
  geometry::numFaces = geometry::numInnerFaces + (geometry::numTotal - geometry::numCells);
  geometry::ninl = geometry::nout = 0;
  allocate_arrays();
  try {
    int itime = 0;
    double time = 0.0;
    readfiles(argv[4], &itime, &time);
    write_restart_files(argv[5], itime, time);
    std::printf("itime %d time %.17g\n", itime, time);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
