// write_restart_files / readfiles of the reference (src/write_restart_files.f90:16-62, src/readfiles.f90:12-50) for the
// host mirror: Fortran unformatted sequential records (gfortran: 4-byte byte count before and after each payload) in
// the order (itime,time) [gradpcmf] flmass u v w p te ed t vis uu vv ww uv uw vw uo vo wo teo edo.  The mirror is
// laminar and has no energy equation: te, ed, teo, edo are written as numTotal zeros (the reference allocates them
// unconditionally, allocate.f90:75-80), t and the Reynolds stresses as empty records (unallocated there too).
// SURVEY.md 8(f) rank 4.  Host I/O only.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <cstdio>
#include <stdexcept>

#include "fcapp_host.hpp"

namespace fcapp {
namespace {

void put_record(std::FILE *fp, const void *p, size_t bytes) {
  if (bytes > 2147483639u) throw std::runtime_error("restart record longer than 2 GiB (sub-records not implemented)");
  const int32_t n = (int32_t)bytes;
  std::fwrite(&n, 4, 1, fp);
  if (bytes) std::fwrite(p, 1, bytes, fp);
  std::fwrite(&n, 4, 1, fp);
}

std::vector<unsigned char> get_record(std::FILE *fp) {
  int32_t n = 0, m = 0;
  if (std::fread(&n, 4, 1, fp) != 1 || n < 0) throw std::runtime_error("restart file: bad record header");
  std::vector<unsigned char> b((size_t)n);
  if (n && std::fread(b.data(), 1, (size_t)n, fp) != (size_t)n) throw std::runtime_error("restart file: short record");
  if (std::fread(&m, 4, 1, fp) != 1 || m != n) throw std::runtime_error("restart file: record markers disagree");
  return b;
}

void into(std::vector<dp> &v, const std::vector<unsigned char> &b, const char *name) {
  if (b.empty()) return;   // not allocated in the run that wrote the file
  if (b.size() != v.size() * sizeof(dp)) throw std::runtime_error(std::string("restart file: size of ") + name);
  std::memcpy(v.data(), b.data(), b.size());
}

}  // namespace

void write_restart_files(const std::string &path, int itime, dp time) {
  using namespace variables;
  std::FILE *fp = std::fopen(path.c_str(), "wb");
  if (!fp) throw std::runtime_error("cannot write " + path);
  unsigned char head[12];
  const int32_t it = itime;
  std::memcpy(head, &it, 4);
  std::memcpy(head + 4, &time, 8);
  put_record(fp, head, 12);
  if (parameters::const_mflux) put_record(fp, &parameters::gradPcmf, sizeof(dp));
  const std::vector<dp> zeros_t(geometry::numTotal, 0.0), none;
  std::vector<dp> fl(flmass.begin(), flmass.begin() + geometry::numInnerFaces);   // flmass(numInnerFaces), allocate.f90:209
  const std::vector<dp> *order[] = {&fl, &u, &v, &w, &p, &zeros_t, &zeros_t, &none, &vis, &none, &none, &none, &none, &none,
                                    &none, &uo, &vo, &wo, &zeros_t, &zeros_t};
  for (const auto *a : order) put_record(fp, a->data(), a->size() * sizeof(dp));
  std::fclose(fp);
}

void readfiles(const std::string &path, int *itime, dp *time) {
  using namespace variables;
  std::FILE *fp = std::fopen(path.c_str(), "rb");
  if (!fp) throw std::runtime_error("cannot open " + path);
  try {
    const auto head = get_record(fp);
    if (head.size() != 12) throw std::runtime_error("restart file: first record is not (itime,time)");
    int32_t it;
    std::memcpy(&it, head.data(), 4);
    if (itime) *itime = it;
    if (time) std::memcpy(time, head.data() + 4, 8);
    if (parameters::const_mflux) {
      const auto b = get_record(fp);
      if (b.size() == sizeof(dp)) std::memcpy(&parameters::gradPcmf, b.data(), sizeof(dp));
    }
    std::vector<dp> fl(geometry::numInnerFaces), skip;
    into(fl, get_record(fp), "flmass");
    std::copy(fl.begin(), fl.end(), flmass.begin());
    into(u, get_record(fp), "u"); into(v, get_record(fp), "v"); into(w, get_record(fp), "w"); into(p, get_record(fp), "p");
    get_record(fp); get_record(fp);                      // te, ed
    into(t, get_record(fp), "t");
    into(vis, get_record(fp), "vis");
    for (int k = 0; k < 6; ++k) get_record(fp);           // uu vv ww uv uw vw
    into(uo, get_record(fp), "uo"); into(vo, get_record(fp), "vo"); into(wo, get_record(fp), "wo");
    get_record(fp); get_record(fp);                      // teo, edo
  } catch (...) {
    std::fclose(fp);
    throw;
  }
  std::fclose(fp);
  pp = p;   // readfiles.f90:52
}

}  // namespace fcapp
