// LAMMPS text dump of reconstructed (iSED) frames - host code, no CUDA (N2 of SURVEY.md 8f).
//
// Byte-for-byte the format the reference writes (reference: src/psa/io/writer.py:139-228) and its GUI parses
// back (src/psa/gui/psa_gui.py:1396-1455): per frame `ITEM: TIMESTEP`, atom count, orthogonal (`pp pp pp`) or
// triclinic (`xy xz yz pp pp pp`) box bounds with 8 decimals, then `id type x y z` with 6 decimals.  The reference
// issues one Python `write` per atom per frame (6.4 M lines per point on the 64 000-atom config); here frames are
// formatted by a pool of threads into per-frame buffers and written in order.
//
// `%.6f` of a float32 is produced exactly (correctly rounded, ties to even on the exact binary value - what both
// CPython's and glibc's formatters do) with integer arithmetic: x = m 2^e, so x 10^6 = (m 10^6) / 2^-e is a shift.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace psa {
namespace {

inline char* put_uint(char* p, uint64_t v) {
  char tmp[24];
  int n = 0;
  do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (n) *p++ = tmp[--n];
  return p;
}

// appends "%.6f" of x; returns the new end
inline char* put_fixed6(char* p, float x) {
  uint32_t bits;
  memcpy(&bits, &x, 4);
  const uint32_t expo = (bits >> 23) & 0xFF, frac = bits & 0x7FFFFF;
  if (expo == 0xFF) {                                  // CPython prints nan / inf / -inf
    if (frac) { memcpy(p, "nan", 3); return p + 3; }
    if (bits >> 31) *p++ = '-';
    memcpy(p, "inf", 3);
    return p + 3;
  }
  if (bits >> 31) *p++ = '-';
  const uint64_t m = expo ? (uint64_t)(frac | 0x800000) : frac;
  const int e = (expo ? (int)expo : 1) - 150;          // x = m * 2^e
  uint64_t q;                                          // round(|x| * 10^6)
  if (e >= 0) {
    if (e > 16) {                                      // |x| >= 2^40: rare, leave it to the C library (same rounding)
      return p + snprintf(p, 64, "%.6f", fabs((double)x));
    }
    q = (m << e) * 1000000ull;
  } else {
    const int s = -e;
    const uint64_t n = m * 1000000ull;                 // < 2^44
    if (s >= 64) {
      q = 0;
    } else {
      q = n >> s;
      const uint64_t rem = n & ((1ull << s) - 1), half = 1ull << (s - 1);
      if (rem > half || (rem == half && (q & 1))) ++q;
    }
  }
  p = put_uint(p, q / 1000000ull);
  *p++ = '.';
  uint32_t f = (uint32_t)(q % 1000000ull);
  for (int i = 5; i >= 0; --i) { p[i] = (char)('0' + f % 10); f /= 10; }
  return p + 6;
}

void format_frame(std::string& buf, int64_t i_fr, const std::string& box_txt, const float* xyz, const int32_t* types,
                  int64_t n_at) {
  char head[96];
  const int hn = snprintf(head, sizeof(head), "ITEM: TIMESTEP\n%lld\nITEM: NUMBER OF ATOMS\n%lld\n", (long long)i_fr,
                          (long long)n_at);
  static const char atoms_txt[] = "ITEM: ATOMS id type x y z\n";
  buf.resize((size_t)hn + box_txt.size() + sizeof(atoms_txt) - 1 + (size_t)n_at * 96);
  char* p = &buf[0];
  memcpy(p, head, (size_t)hn); p += hn;
  memcpy(p, box_txt.data(), box_txt.size()); p += box_txt.size();
  memcpy(p, atoms_txt, sizeof(atoms_txt) - 1); p += sizeof(atoms_txt) - 1;
  for (int64_t a = 0; a < n_at; ++a) {
    p = put_uint(p, (uint64_t)(a + 1));
    *p++ = ' ';
    int64_t t = types[a];
    if (t < 0) { *p++ = '-'; t = -t; }
    p = put_uint(p, (uint64_t)t);
    for (int c = 0; c < 3; ++c) {
      *p++ = ' ';
      p = put_fixed6(p, xyz[a * 3 + c]);
    }
    *p++ = '\n';
  }
  buf.resize((size_t)(p - &buf[0]));
}

}  // namespace
}  // namespace psa

using namespace psa;

extern "C" int psa_write_dump(const char* path, const float* frames_host, const int32_t* types_host, int64_t n_frames,
                              int64_t n_atoms, const float* box9_host, int n_threads) {
  PSA_REQUIRE(path && box9_host && n_frames >= 0 && n_atoms >= 0, "psa_write_dump: bad arguments");
  PSA_REQUIRE(n_frames * n_atoms == 0 || (frames_host && types_host), "psa_write_dump: null pointer");
  // box bounds in float32 arithmetic, like the reference's NumPy scalars (writer.py:181-203)
  const float* b = box9_host;
  const float xhi = b[0], yhi = b[4], zhi = b[8];
  float xy = b[1], xz = b[2], yz = b[5];
  auto close0 = [](float v) { return fabs((double)v) <= 1e-8; };          // np.isclose(v, 0.0)
  const bool tri = !(close0(xy) && close0(xz) && close0(yz));
  char line[256];
  std::string box_txt;
  if (tri) {
    const float sum = xy + xz;
    const float lo_x = 0.0f + fminf(fminf(0.0f, xy), fminf(xz, sum)), hi_x = xhi + fmaxf(fmaxf(0.0f, xy), fmaxf(xz, sum));
    const float lo_y = 0.0f + fminf(0.0f, yz), hi_y = yhi + fmaxf(0.0f, yz);
    box_txt = "ITEM: BOX BOUNDS xy xz yz pp pp pp\n";
    snprintf(line, sizeof(line), "%.8f %.8f %.8f\n", (double)lo_x, (double)hi_x, (double)xy); box_txt += line;
    snprintf(line, sizeof(line), "%.8f %.8f %.8f\n", (double)lo_y, (double)hi_y, (double)xz); box_txt += line;
    snprintf(line, sizeof(line), "%.8f %.8f %.8f\n", 0.0, (double)zhi, (double)yz); box_txt += line;
  } else {
    box_txt = "ITEM: BOX BOUNDS pp pp pp\n";
    snprintf(line, sizeof(line), "%.8f %.8f\n", 0.0, (double)xhi); box_txt += line;
    snprintf(line, sizeof(line), "%.8f %.8f\n", 0.0, (double)yhi); box_txt += line;
    snprintf(line, sizeof(line), "%.8f %.8f\n", 0.0, (double)zhi); box_txt += line;
  }
  FILE* fh = fopen(path, "wb");
  if (!fh) {
    set_error("psa_write_dump: cannot open %s for writing", path);
    return PSA_ERR_BAD_ARG;
  }
  int workers = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (workers < 1) workers = 1;
  if (workers > 64) workers = 64;
  if ((int64_t)workers > n_frames) workers = (int)(n_frames > 0 ? n_frames : 1);
  std::vector<std::string> bufs((size_t)workers);
  bool ok = true;
  for (int64_t f0 = 0; f0 < n_frames && ok; f0 += workers) {
    const int n_now = (int)((n_frames - f0) < workers ? (n_frames - f0) : workers);
    std::vector<std::thread> pool;
    for (int w = 1; w < n_now; ++w)
      pool.emplace_back(format_frame, std::ref(bufs[(size_t)w]), f0 + w, std::cref(box_txt),
                        frames_host + (f0 + w) * n_atoms * 3, types_host, n_atoms);
    format_frame(bufs[0], f0, box_txt, frames_host + f0 * n_atoms * 3, types_host, n_atoms);
    for (auto& t : pool) t.join();
    for (int w = 0; w < n_now && ok; ++w)
      ok = fwrite(bufs[(size_t)w].data(), 1, bufs[(size_t)w].size(), fh) == bufs[(size_t)w].size();
  }
  ok = (fclose(fh) == 0) && ok;
  if (!ok) {
    set_error("psa_write_dump: write to %s failed", path);
    return PSA_ERR_CUDA;
  }
  return PSA_OK;
}
