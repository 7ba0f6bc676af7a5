// mesh_geometry for OpenFOAM polyMesh directories: the host-side reader + geometry of the reference
// (src/mesh_geometry_and_topology.f90:310-1081, polyMesh branch) mirrored in C++.
// Reads `points`, `faces`, `owner`, `neighbour` (OpenFOAM ASCII) and the reference's simplified `boundary`
// table ("#type nFaces startFace" header, then one row per patch, :414-470) and fills `module geometry`:
// face area vectors by fan triangulation about node 1 (:868-895), face centres as node averages (:930-940),
// cell volumes by the pyramid sums (:906-910) and centroids (:975-990) in the reference's face-major
// accumulation order, interpolation factors from the intersection of the P-N line with the plane of the
// first three face nodes (find_intersection_point :201-275).  Host preprocessing only: nothing here is on
// the GPU path; the arrays are handed to the device by fcapp_init exactly like the synthetic boxes.
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "fcapp_host.hpp"

namespace fcapp {
namespace {

std::string slurp(const std::string &path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("mesh_geometry: cannot open " + path);
  std::stringstream ss;
  ss << in.rdbuf();
  return ss.str();
}

// strip /* */ and // comments and the FoamFile { ... } header
std::string foam_body(std::string t) {
  std::string o;
  o.reserve(t.size());
  for (size_t i = 0; i < t.size();) {
    if (t.compare(i, 2, "/*") == 0) {
      size_t e = t.find("*/", i + 2);
      i = e == std::string::npos ? t.size() : e + 2;
    } else if (t.compare(i, 2, "//") == 0) {
      size_t e = t.find('\n', i);
      i = e == std::string::npos ? t.size() : e;
    } else {
      o.push_back(t[i++]);
    }
  }
  size_t f = o.find("FoamFile");
  if (f != std::string::npos) {
    size_t e = o.find('}', f);
    if (e != std::string::npos) o = o.substr(e + 1);
  }
  return o;
}

struct cursor {
  const std::string &s;
  size_t i = 0;
  void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
  bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { ++i; return true; } return false; }
  void need(char c, const char *what) {
    if (!eat(c)) throw std::runtime_error(std::string("mesh_geometry: expected '") + c + "' in " + what);
  }
  long integer(const char *what) {
    ws();
    char *e = nullptr;
    long v = std::strtol(s.c_str() + i, &e, 10);
    if (e == s.c_str() + i) throw std::runtime_error(std::string("mesh_geometry: expected an integer in ") + what);
    i = (size_t)(e - s.c_str());
    return v;
  }
  double real(const char *what) {
    ws();
    char *e = nullptr;
    double v = std::strtod(s.c_str() + i, &e);
    if (e == s.c_str() + i) throw std::runtime_error(std::string("mesh_geometry: expected a number in ") + what);
    i = (size_t)(e - s.c_str());
    return v;
  }
};

std::vector<int> read_labels(const std::string &path) {
  const std::string b = foam_body(slurp(path));
  cursor c{b};
  const long n = c.integer("label list");
  c.need('(', "label list");
  std::vector<int> v((size_t)n);
  for (long k = 0; k < n; ++k) v[(size_t)k] = (int)c.integer("label list");
  c.need(')', "label list");
  return v;
}

// find_intersection_point (:201-275), expression produced "by MATLAB symbolic tool", kept term by term
void find_intersection_point(dp x1, dp y1, dp z1, dp x2, dp y2, dp z2, dp x3, dp y3, dp z3, dp x4, dp y4, dp z4, dp x5,
                             dp y5, dp z5, dp &xjp, dp &yjp, dp &zjp) {
  const dp tiny = (dp)1e-30f;   // `tiny = 1e-30`, a default-real literal (:57)
  const dp t = -(x2 * (y3 * z4 - y4 * z3) - x1 * (y3 * z4 - y4 * z3) - x3 * (y2 * z4 - y4 * z2) + x1 * (y2 * z4 - y4 * z2) +
                 x3 * (y1 * z4 - y4 * z1) - x2 * (y1 * z4 - y4 * z1) + x4 * (y2 * z3 - y3 * z2) - x1 * (y2 * z3 - y3 * z2) -
                 x4 * (y1 * z3 - y3 * z1) + x2 * (y1 * z3 - y3 * z1) + x4 * (y1 * z2 - y2 * z1) - x3 * (y1 * z2 - y2 * z1)) /
               (x2 * (y3 * (z5 - z4) - (y5 - y4) * z3) - x1 * (y3 * (z5 - z4) - (y5 - y4) * z3) -
                x3 * (y2 * (z5 - z4) - (y5 - y4) * z2) + x1 * (y2 * (z5 - z4) - (y5 - y4) * z2) +
                x3 * (y1 * (z5 - z4) - (y5 - y4) * z1) - x2 * (y1 * (z5 - z4) - (y5 - y4) * z1) +
                (x5 - x4) * (y2 * z3 - y3 * z2) - (x5 - x4) * (y1 * z3 - y3 * z1) + (x5 - x4) * (y1 * z2 - y2 * z1) + tiny);
  xjp = x4 + (x5 - x4) * t;
  yjp = y4 + (y5 - y4) * t;
  zjp = z4 + (z5 - z4) * t;
}

inline dp cell_volume_part(dp ax, dp ay, dp az, dp nx, dp ny, dp nz) { return 1.0 / 6.0 * (ax * nx + ay * ny + az * nz); }
inline dp centroid_part(dp ax, dp bx, dp cx, dp nx, dp vol) {   // :181
  return 1.0 / (2 * vol) * 1.0 / 24.0 * nx * ((ax + bx) * (ax + bx) + (bx + cx) * (bx + cx) + (cx + ax) * (cx + ax));
}

}  // namespace

void mesh_geometry(const std::string &dir) {
  using namespace geometry;
  // ---- boundary table (:395-470) ----
  ninl = nout = nsym = nwal = npru = noc = 0;
  iInletFacesStart = iOutletFacesStart = iSymmetryFacesStart = iWallFacesStart = iPressOutletFacesStart = iOCFacesStart = 0;
  {
    std::ifstream in(dir + "/boundary");
    if (!in) throw std::runtime_error("mesh_geometry: cannot open " + dir + "/boundary");
    std::string line;
    while (std::getline(in, line)) {
      std::istringstream ls(line);
      std::string kind;
      int nf = 0, st = 0;
      if (!(ls >> kind) || kind[0] == '#') continue;
      if (!(ls >> nf >> st)) throw std::runtime_error("mesh_geometry: bad row in boundary: " + line);
      auto add = [&](int &cnt, int &start) { if (cnt == 0) start = st; cnt += nf; };
      if (kind == "inlet") add(ninl, iInletFacesStart);
      else if (kind == "outlet") add(nout, iOutletFacesStart);
      else if (kind == "symmetry") add(nsym, iSymmetryFacesStart);
      else if (kind == "wall" || kind == "wallIsoth" || kind == "wallAdiab" || kind == "wallQFlux") add(nwal, iWallFacesStart);
      else if (kind == "prOutlet") add(npru, iPressOutletFacesStart);
      else if (kind == "domain" || kind == "cyclic")
        throw std::runtime_error("mesh_geometry: O-C cuts / cyclic boundaries are outside the GPU path");
      else throw std::runtime_error("Non-existing boundary type in polymesh/boundary file: " + kind);
    }
  }
  // ---- points, owner, neighbour ----
  std::vector<dp> x, y, z;
  {
    const std::string b = foam_body(slurp(dir + "/points"));
    cursor c{b};
    const long np = c.integer("points");
    c.need('(', "points");
    x.resize((size_t)np); y.resize((size_t)np); z.resize((size_t)np);
    for (long k = 0; k < np; ++k) {
      c.need('(', "points");
      x[(size_t)k] = c.real("points"); y[(size_t)k] = c.real("points"); z[(size_t)k] = c.real("points");
      c.need(')', "points");
    }
  }
  owner = read_labels(dir + "/owner");
  neighbour = read_labels(dir + "/neighbour");
  for (auto &v : owner) v += 1;       // OpenFOAM is 0-based, the reference adds 1 (:560-575)
  for (auto &v : neighbour) v += 1;
  numFaces = (int)owner.size();
  numInnerFaces = (int)neighbour.size();
  numBoundaryFaces = numFaces - numInnerFaces;
  numCells = 0;
  for (int v : owner) numCells = v > numCells ? v : numCells;
  numTotal = numCells + numBoundaryFaces;
  nnz = numCells + 2 * numInnerFaces;   // :580
  if (ninl + nout + nsym + nwal + npru != numBoundaryFaces)
    throw std::runtime_error("mesh_geometry: the boundary table does not cover the boundary faces");
  // ---- faces ----
  std::vector<int> fptr(1, 0), fnode;
  {
    const std::string b = foam_body(slurp(dir + "/faces"));
    cursor c{b};
    const long nf = c.integer("faces");
    if (nf != numFaces) throw std::runtime_error("mesh_geometry: faces and owner disagree on the number of faces");
    c.need('(', "faces");
    for (long k = 0; k < nf; ++k) {
      const long nn = c.integer("faces");
      c.need('(', "faces");
      for (long q = 0; q < nn; ++q) fnode.push_back((int)c.integer("faces"));
      c.need(')', "faces");
      fptr.push_back((int)fnode.size());
    }
  }
  const dp half = 0.5, one_third = 1.0 / 3.0;
  arx.assign(numFaces, 0.0); ary.assign(numFaces, 0.0); arz.assign(numFaces, 0.0);
  xf.assign(numFaces, 0.0); yf.assign(numFaces, 0.0); zf.assign(numFaces, 0.0);
  vol.assign(numCells, 0.0); xc.assign(numCells, 0.0); yc.assign(numCells, 0.0); zc.assign(numCells, 0.0);
  // ---- areas, face centres, volumes (:859-941) ----
  for (int f = 0; f < numFaces; ++f) {
    const int *nd = &fnode[fptr[f]];
    const int nn = fptr[f + 1] - fptr[f], inp = owner[f] - 1;
    for (int i = 0; i + 2 < nn; ++i) {
      const dp px = x[nd[i + 1]] - x[nd[0]], py = y[nd[i + 1]] - y[nd[0]], pz = z[nd[i + 1]] - z[nd[0]];
      const dp qx = x[nd[i + 2]] - x[nd[0]], qy = y[nd[i + 2]] - y[nd[0]], qz = z[nd[i + 2]] - z[nd[0]];
      const dp nx = py * qz - pz * qy, ny = pz * qx - px * qz, nz = px * qy - py * qx;
      arx[f] = arx[f] + half * nx; ary[f] = ary[f] + half * ny; arz[f] = arz[f] + half * nz;
      const dp cx = one_third * (x[nd[i + 2]] + x[nd[i + 1]] + x[nd[0]]);
      const dp cy = one_third * (y[nd[i + 2]] + y[nd[i + 1]] + y[nd[0]]);
      const dp cz = one_third * (z[nd[i + 2]] + z[nd[i + 1]] + z[nd[0]]);
      vol[inp] = vol[inp] + cell_volume_part(cx, cy, cz, nx, ny, nz);
      if (f < numInnerFaces) {
        const int inn = neighbour[f] - 1;
        vol[inn] = vol[inn] + cell_volume_part(cx, cy, cz, -nx, -ny, -nz);
      }
    }
    for (int q = 0; q < nn; ++q) { xf[f] = xf[f] + x[nd[q]]; yf[f] = yf[f] + y[nd[q]]; zf[f] = zf[f] + z[nd[q]]; }
    xf[f] = xf[f] / (dp)nn; yf[f] = yf[f] / (dp)nn; zf[f] = zf[f] / (dp)nn;
  }
  // ---- cell centres (:960-992) ----
  for (int f = 0; f < numFaces; ++f) {
    const int *nd = &fnode[fptr[f]];
    const int nn = fptr[f + 1] - fptr[f], inp = owner[f] - 1;
    for (int i = 0; i + 2 < nn; ++i) {
      const dp px = x[nd[i + 1]] - x[nd[0]], py = y[nd[i + 1]] - y[nd[0]], pz = z[nd[i + 1]] - z[nd[0]];
      const dp qx = x[nd[i + 2]] - x[nd[0]], qy = y[nd[i + 2]] - y[nd[0]], qz = z[nd[i + 2]] - z[nd[0]];
      const dp nx = py * qz - pz * qy, ny = pz * qx - px * qz, nz = px * qy - py * qx;
      xc[inp] = xc[inp] + centroid_part(x[nd[0]], x[nd[i + 1]], x[nd[i + 2]], nx, vol[inp]);
      yc[inp] = yc[inp] + centroid_part(y[nd[0]], y[nd[i + 1]], y[nd[i + 2]], ny, vol[inp]);
      zc[inp] = zc[inp] + centroid_part(z[nd[0]], z[nd[i + 1]], z[nd[i + 2]], nz, vol[inp]);
      if (f < numInnerFaces) {
        const int inn = neighbour[f] - 1;
        xc[inn] = xc[inn] + centroid_part(x[nd[0]], x[nd[i + 2]], x[nd[i + 1]], -nx, vol[inn]);
        yc[inn] = yc[inn] + centroid_part(y[nd[0]], y[nd[i + 2]], y[nd[i + 1]], -ny, vol[inn]);
        zc[inn] = zc[inn] + centroid_part(z[nd[0]], z[nd[i + 2]], z[nd[i + 1]], -nz, vol[inn]);
      }
    }
  }
  // ---- interpolation factors (:1012-1062) ----
  facint.assign(numInnerFaces, 0.0);
  for (int f = 0; f < numInnerFaces; ++f) {
    const int *nd = &fnode[fptr[f]];
    const int inp = owner[f] - 1, inn = neighbour[f] - 1;
    dp xpn = xc[inn] - xc[inp], ypn = yc[inn] - yc[inp], zpn = zc[inn] - zc[inp];
    const dp dpn = std::sqrt(xpn * xpn + ypn * ypn + zpn * zpn);
    dp xjp, yjp, zjp;
    find_intersection_point(x[nd[0]], y[nd[0]], z[nd[0]], x[nd[1]], y[nd[1]], z[nd[1]], x[nd[2]], y[nd[2]], z[nd[2]],
                            xc[inp], yc[inp], zc[inp], xc[inn], yc[inn], zc[inn], xjp, yjp, zjp);
    xpn = xjp - xc[inp]; ypn = yjp - yc[inp]; zpn = zjp - zc[inp];
    const dp djn = std::sqrt(xpn * xpn + ypn * ypn + zpn * zpn);
    facint[f] = djn / dpn;
  }
}

}  // namespace fcapp
