// Reads a polyMesh directory with fcapp::mesh_geometry and writes `module geometry` to a binary file:
// int32 header {numCells, numInnerFaces, numFaces, ninl, nout, nsym, nwal, npru, 5 x FacesStart}, owner,
// neighbour (int32), then xc yc zc vol [numCells], arx ary arz xf yf zf [numFaces], facint [numInnerFaces]
// (float64).  No GPU needed: used by tests/test_polymesh_reader.py and as a mesh check for users.
//   usage: meshdump <polyMesh dir> [out.bin]
#include <cstdio>
#include <exception>

#include "fcapp_host.hpp"

using namespace fcapp;

int main(int argc, char **argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: meshdump <polyMesh dir> [out.bin]\n"); return 2; }
  try {
    mesh_geometry(argv[1]);
  } catch (const std::exception &e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  using namespace geometry;
  std::printf("cells %d innerFaces %d faces %d inlet %d@%d outlet %d@%d symmetry %d@%d wall %d@%d prOutlet %d@%d\n", numCells,
              numInnerFaces, numFaces, ninl, iInletFacesStart, nout, iOutletFacesStart, nsym, iSymmetryFacesStart, nwal,
              iWallFacesStart, npru, iPressOutletFacesStart);
  if (argc > 2) {
    std::FILE *fp = std::fopen(argv[2], "wb");
    if (!fp) { std::fprintf(stderr, "cannot write %s\n", argv[2]); return 1; }
    const int hdr[13] = {numCells, numInnerFaces, numFaces, ninl, nout, nsym, nwal, npru, iInletFacesStart,
                         iOutletFacesStart, iSymmetryFacesStart, iWallFacesStart, iPressOutletFacesStart};
    std::fwrite(hdr, sizeof(int), 13, fp);
    std::fwrite(owner.data(), sizeof(int), owner.size(), fp);
    std::fwrite(neighbour.data(), sizeof(int), neighbour.size(), fp);
    for (auto *v : {&xc, &yc, &zc, &vol, &arx, &ary, &arz, &xf, &yf, &zf, &facint}) std::fwrite(v->data(), sizeof(dp), v->size(), fp);
    std::fclose(fp);
  }
  return 0;
}
