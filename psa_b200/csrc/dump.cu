// LAMMPS text dump of reconstructed (iSED) frames - host code, no CUDA (N2 of SURVEY.md 8f).
//
// Byte-for-byte the format the reference writes (reference: src/psa/io/writer.py:139-228) and its GUI parses
// back (src/psa/gui/psa_gui.py:1396-1455): per frame `ITEM: TIMESTEP`, atom count, orthogonal (`pp pp pp`) or
// triclinic (`xy xz yz pp pp pp`) box bounds with 8 decimals, then `id type x y z` with 6 decimals.  The reference
// issues one Python `write` per atom per frame (6.4 M lines per point on the 64 000-atom config); here a pool of threads
// formats whole frames into per-thread buffers and writes each at its own file offset (pwrite).
//
// `%.6f` of a float32 is produced exactly (correctly rounded, ties to even on the exact binary value - what both
// CPython's and glibc's formatters do) with integer arithmetic: x = m 2^e, so x 10^6 = (m 10^6) / 2^-e is a shift.
#include <errno.h>
#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
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

// formats one frame into buf (only ever grown: no zero fill per frame) and returns the number of bytes
size_t format_frame(std::vector<char>& buf, int64_t i_fr, const std::string& box_txt, const float* xyz, const int32_t* types,
                    int64_t n_at) {
  char head[96];
  const int hn = snprintf(head, sizeof(head), "ITEM: TIMESTEP\n%lld\nITEM: NUMBER OF ATOMS\n%lld\n", (long long)i_fr,
                          (long long)n_at);
  static const char atoms_txt[] = "ITEM: ATOMS id type x y z\n";
  const size_t need = (size_t)hn + box_txt.size() + sizeof(atoms_txt) - 1 + (size_t)n_at * 96;
  if (buf.size() < need) buf.resize(need);
  char* p = buf.data();
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
  return (size_t)(p - buf.data());
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
  const int fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) {
    set_error("psa_write_dump: cannot open %s for writing", path);
    return PSA_ERR_BAD_ARG;
  }
  int workers = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
  if (workers < 1) workers = 1;
  if (workers > 64) workers = 64;
  if ((int64_t)workers > n_frames) workers = (int)(n_frames > 0 ? n_frames : 1);
  // Frames are handed out by a counter; a worker formats its frame into its own buffer, learns the frame's file
  // offset from its predecessor (offset[f + 1] = offset[f] + size[f] is published as soon as frame f is FORMATTED, so
  // the chain never waits for a write), and writes the buffer at that offset itself: formatting and writing of
  // different frames overlap, and the file is the same byte for byte.
  std::vector<std::atomic<int64_t>> offset((size_t)n_frames + 1);
  for (auto& o : offset) o.store(-1, std::memory_order_relaxed);
  offset[0].store(0, std::memory_order_release);
  std::atomic<int64_t> next{0};
  std::atomic<bool> ok{true};
  auto work = [&]() {
    std::vector<char> buf;
    for (;;) {
      const int64_t f = next.fetch_add(1, std::memory_order_relaxed);
      if (f >= n_frames) break;
      size_t left = format_frame(buf, f, box_txt, frames_host + f * n_atoms * 3, types_host, n_atoms);
      int64_t at;
      while ((at = offset[(size_t)f].load(std::memory_order_acquire)) < 0) std::this_thread::yield();
      offset[(size_t)f + 1].store(at + (int64_t)left, std::memory_order_release);
      const char* src = buf.data();
      while (left > 0 && ok.load(std::memory_order_relaxed)) {
        const ssize_t n = pwrite(fd, src, left, (off_t)at);
        if (n < 0) {
          if (errno == EINTR) continue;
          ok.store(false, std::memory_order_relaxed);
          break;
        }
        src += n; left -= (size_t)n; at += n;
      }
    }
  };
  std::vector<std::thread> pool;
  for (int w = 1; w < workers; ++w) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
  const bool closed = close(fd) == 0;
  const bool all_ok = ok.load() && closed;
  if (!all_ok) {
    set_error("psa_write_dump: write to %s failed", path);
    return PSA_ERR_CUDA;
  }
  return PSA_OK;
}
