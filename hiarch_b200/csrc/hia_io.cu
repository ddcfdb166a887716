// f2: the wire format between the pipeline's processes -- 3-column text (row, col, value).
// HOST code only (compiled by nvcc like the rest of the library, no device code).
//
//   hia_mtx_scan / hia_mtx_parse   replace  pd.read_table(sps_file, names=['row','col','value']) + dropna
//                                           (read_sps_file, publish/new_hic_class.py:2580-2581)
//   hia_mtx_write                  replaces np.savetxt(out_file, vstack([rows, cols, values]).T,
//                                           fmt="%d\t%d\t%.2f")   (SpsHiCMtx.to_output, :2290-2297)
//
// At 12 M entries the pandas parse is 2.1 s and savetxt 25 s of a pipeline whose kernels now take
// milliseconds. Reader: the file is mmap'ed and cut at line boundaries into one slice per thread;
// the lines of every slice are counted (memchr) and each slice is parsed straight into the caller's
// arrays at its line offset (gaps left by dropped lines are closed afterwards). Writer: one pass
// computes the exact byte length of every thread's share, then the threads format straight into the
// mmap'ed output file. 12.3 M entries, 8 threads: read 0.10 s, write 0.13 s (tmpfs).
//
// Parity rules:
//   * blank lines are skipped (pandas skip_blank_lines); a line with fewer than three fields or
//     with an empty / 'nan' field is dropped (dropna); anything else that is not a number -> error.
//   * a decimal literal whose digits fit 2^53 with <= 22 fractional digits is mantissa / 10^d in one
//     IEEE division (Clinger's exact case: correctly rounded, and bit-identical to what pandas'
//     default parser produces for up to 15 significant digits, i.e. for every "%.2f" file);
//     everything else (exponents, long literals) goes through strtod (correctly rounded; pandas'
//     xstrtod is up to 1 ulp off there, e.g. 0.30000000000000004 -> 0.3).
//   * "%.2f" output is the correctly rounded decimal of the binary value (what CPython's % operator
//     and glibc printf both produce): the fast path rounds v * 100 only when it is provably far
//     from a tie, otherwise it defers to snprintf. Negative zero and negative values that round
//     to zero print "-0.00" like Python.
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cmath>
#include <string>
#include <thread>
#include <vector>

#include "hia_common.cuh"

namespace hia {
namespace {

struct Mapped {
  const char* p = nullptr;
  size_t n = 0;
  int fd = -1;
  bool open(const char* path) {
    fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) { ::close(fd); fd = -1; return false; }
    n = (size_t)st.st_size;
    if (n == 0) { p = nullptr; return true; }
    void* m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) { ::close(fd); fd = -1; return false; }
    madvise(m, n, MADV_SEQUENTIAL);
    p = static_cast<const char*>(m);
    return true;
  }
  ~Mapped() {
    if (p) munmap(const_cast<char*>(p), n);
    if (fd >= 0) ::close(fd);
  }
};

const double kPow10[23] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// Parse one field [s, e). Returns 0 = number, 1 = missing (empty / nan / NA -> the line is dropped),
// -1 = malformed.
int parse_field(const char* s, const char* e, double* out) {
  while (s < e && (*s == ' ')) ++s;
  while (e > s && (e[-1] == ' ' || e[-1] == '\r')) --e;
  if (s == e) return 1;
  const char* q = s;
  bool neg = false;
  if (*q == '-' || *q == '+') { neg = (*q == '-'); ++q; }
  unsigned long long mant = 0;
  int digits = 0, frac = 0;
  bool any = false, overflow = false;
  while (q < e && *q >= '0' && *q <= '9') {
    any = true;
    if (digits < 19) { mant = mant * 10 + (unsigned)(*q - '0'); if (mant || digits) ++digits; } else overflow = true;
    ++q;
  }
  if (q < e && *q == '.') {
    ++q;
    while (q < e && *q >= '0' && *q <= '9') {
      any = true;
      if (digits < 19) { mant = mant * 10 + (unsigned)(*q - '0'); if (mant || digits) ++digits; ++frac; } else overflow = true;
      ++q;
    }
  }
  if (any && q == e && !overflow && frac <= 22 && mant < (1ull << 53)) {
    const double v = (double)mant / kPow10[frac];                 // exact operands, one rounding
    *out = neg ? -v : v;
    return 0;
  }
  // the general case: exponents, inf, long literals
  const size_t len = (size_t)(e - s);
  if (len == 3 && (strncasecmp(s, "nan", 3) == 0)) return 1;
  if ((len == 2 && strncmp(s, "NA", 2) == 0) || (len == 4 && strncasecmp(s, "null", 4) == 0) ||
      (len == 3 && strncmp(s, "N/A", 3) == 0) || (len == 4 && (strncmp(s, "-nan", 4) == 0 || strncmp(s, "-NaN", 4) == 0)))
    return 1;
  char buf[64];
  if (len >= sizeof(buf)) return -1;
  memcpy(buf, s, len);
  buf[len] = 0;
  char* endp = nullptr;
  errno = 0;
  const double v = strtod(buf, &endp);
  if (endp != buf + len) return -1;
  if (v != v) return 1;
  *out = v;
  return 0;
}

// One line [s, e) (without the '\n'). Returns 1 = entry, 0 = skipped (blank / dropna), -1 = malformed.
int parse_line(const char* s, const char* e, long long* r, long long* c, double* v) {
  const char* t = e;
  while (t > s && (t[-1] == '\r' || t[-1] == ' ' || t[-1] == '\t')) --t;
  const char* b = s;
  while (b < t && (*b == ' ')) ++b;
  if (b == t) return 0;                                  // blank line
  const char* f[3][2];
  int nf = 0;
  const char* cur = s;
  for (const char* p = s; p <= e && nf < 3; ++p) {
    if (p == e || *p == '\t') { f[nf][0] = cur; f[nf][1] = p; ++nf; cur = p + 1; }
  }
  if (nf < 3) return 0;                                  // missing trailing fields -> NaN -> dropna
  if (f[2][1] != e) {                                    // a 4th field: pandas would shift the columns
    for (const char* p = f[2][1]; p < e; ++p)
      if (*p != '\t' && *p != ' ' && *p != '\r') return -1;
  }
  double d[3];
  for (int i = 0; i < 3; ++i) {
    const int st = parse_field(f[i][0], f[i][1], &d[i]);
    if (st == 1) return 0;
    if (st < 0) return -1;
  }
  if (!(fabs(d[0]) < 9.0e18) || !(fabs(d[1]) < 9.0e18)) return -1;
  *r = (long long)d[0];                                  // astype(int64): truncation
  *c = (long long)d[1];
  *v = d[2];
  return 1;
}

// The common line -- "<digits>\t<digits>\t[-]<digits>[.<digits>]" with nothing else on it -- without
// the field splitting and trimming of parse_line. Anything it does not recognise returns -2 and the
// caller falls back to parse_line, which gives the same numbers for the lines accepted here:
// <= 15 digit characters in the value keep the mantissa below 2^53 (the exact-division rule).
inline int parse_line_fast(const char* p, const char* e, long long* r, long long* c, double* v) {
  unsigned long long a = 0;
  const char* q = p;
  while (p < e && (unsigned)(*p - '0') < 10u) a = a * 10 + (unsigned)(*p++ - '0');
  if (p == q || p - q > 18 || p >= e || *p != '\t') return -2;
  ++p;
  unsigned long long b = 0;
  q = p;
  while (p < e && (unsigned)(*p - '0') < 10u) b = b * 10 + (unsigned)(*p++ - '0');
  if (p == q || p - q > 18 || p >= e || *p != '\t') return -2;
  ++p;
  bool neg = false;
  if (p < e && *p == '-') { neg = true; ++p; }
  unsigned long long mant = 0;
  q = p;
  while (p < e && (unsigned)(*p - '0') < 10u) mant = mant * 10 + (unsigned)(*p++ - '0');
  int nd = (int)(p - q), frac = 0;
  if (p < e && *p == '.') {
    ++p;
    q = p;
    while (p < e && (unsigned)(*p - '0') < 10u) mant = mant * 10 + (unsigned)(*p++ - '0');
    frac = (int)(p - q);
    nd += frac;
  }
  if (p != e || nd == 0 || nd > 15) return -2;
  const double x = frac ? (double)mant / kPow10[frac] : (double)mant;
  *r = (long long)a;
  *c = (long long)b;
  *v = neg ? -x : x;
  return 1;
}

inline int parse_line_any(const char* b, const char* e, long long* r, long long* c, double* v) {
  const int st = parse_line_fast(b, e, r, c, v);
  return st != -2 ? st : parse_line(b, e, r, c, v);
}

struct Slice { size_t b, e; long long count; int bad; long long max_r, max_c; int any_neg; };

// first line start at or after byte i
size_t line_start(const Mapped& m, size_t i) {
  if (i == 0) return 0;
  if (i >= m.n) return m.n;
  const void* nl = memchr(m.p + i - 1, '\n', m.n - (i - 1));
  return nl ? (size_t)(static_cast<const char*>(nl) - m.p) + 1 : m.n;
}

// one slice of whole lines per thread (at least 64 KB each)
std::vector<Slice> cut(const Mapped& m, int threads) {
  std::vector<Slice> s;
  if (m.n == 0) return s;
  const size_t T = std::max<size_t>(1, std::min<size_t>((size_t)threads, m.n / (1 << 16) + 1));
  size_t b = 0;
  for (size_t t = 0; t < T; ++t) {
    const size_t e = (t == T - 1) ? m.n : std::max(b, line_start(m, m.n / T * (t + 1)));
    if (e > b) s.push_back({b, e, 0, 0, -1, -1, 0});
    b = e;
  }
  return s;
}

template <typename F>
void for_lines(const Mapped& m, const Slice& sl, F&& fn) {
  const char* p = m.p + sl.b;
  const char* end = m.p + sl.e;
  while (p < end) {
    const char* nl = static_cast<const char*>(memchr(p, '\n', (size_t)(end - p)));
    const char* le = nl ? nl : end;
    fn(p, le);
    p = nl ? nl + 1 : end;
  }
}

int default_threads(int threads) {
  if (threads > 0) return threads;
  const unsigned hc = std::thread::hardware_concurrency();
  return hc ? (int)std::min(hc, 64u) : 8;
}

// ---- "%.2f" ------------------------------------------------------------------------------------
const char kDigits2[201] =
    "0001020304050607080910111213141516171819202122232425262728293031323334353637383940414243444546474849"
    "5051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";

inline int ndigits32(unsigned v) {
  return v < 10u ? 1 : v < 100u ? 2 : v < 1000u ? 3 : v < 10000u ? 4 : v < 100000u ? 5 : v < 1000000u ? 6
       : v < 10000000u ? 7 : v < 100000000u ? 8 : v < 1000000000u ? 9 : 10;
}
// digits written backwards straight into place, two per division (no staging copy: a
// variable-length memcpy per number cost more than the divisions)
inline char* put_u32(char* p, unsigned v) {
  char* const e = p + ndigits32(v);
  char* t = e;
  while (v >= 100u) {
    const unsigned d = v % 100u;
    v /= 100u;
    t -= 2;
    memcpy(t, kDigits2 + 2 * d, 2);
  }
  if (v >= 10u) memcpy(t - 2, kDigits2 + 2 * v, 2);
  else t[-1] = (char)('0' + v);
  return e;
}
inline char* put_uint(char* p, unsigned long long v) {
  if (!(v >> 32)) return put_u32(p, (unsigned)v);
  const unsigned long long hi = v / 100000000ull;
  unsigned lo = (unsigned)(v % 100000000ull);          // exactly 8 digits, zero padded
  p = put_uint(p, hi);
  for (int i = 3; i >= 0; --i) { memcpy(p + 2 * i, kDigits2 + 2 * (lo % 100u), 2); lo /= 100u; }
  return p + 8;
}
inline char* put_int(char* p, long long v) {
  if (v < 0) { *p++ = '-'; return put_uint(p, (unsigned long long)(-(v + 1)) + 1ull); }
  return put_uint(p, (unsigned long long)v);
}
// correctly rounded "%.2f" of v (see the header comment). f2_fast: v * 100 is provably not near a
// tie -> the rounded hundredths q; otherwise the caller defers to snprintf.
inline bool f2_fast(double v, unsigned long long* q) {
  const double a = fabs(v);
  if (!(a < 1.0e11)) return false;                       // also NaN / inf
  const double x = a * 100.0;                            // |error| <= 0.5 ulp(x) <= 1e-3 at 1e13
  const unsigned long long fi = (unsigned long long)x;   // x >= 0: truncation = floor (no libm call)
  const double fr = x - (double)fi;
  if (!(fabs(fr - 0.5) > 4.0e-3)) return false;          // too close to a tie to decide in binary
  *q = fi + (fr > 0.5 ? 1ull : 0ull);
  return true;
}
inline int ndigits64(unsigned long long v) {
  int n = 0;
  while (v >> 32) { v /= 100000000ull; n += 8; }
  return n + ndigits32((unsigned)v);
}
inline char* put_f2(char* p, double v) {
  unsigned long long q;
  if (f2_fast(v, &q)) {
    if (std::signbit(v)) *p++ = '-';
    p = put_uint(p, q / 100);
    *p++ = '.';
    memcpy(p, kDigits2 + 2 * (unsigned)(q % 100), 2);
    return p + 2;
  }
  if (v != v) { memcpy(p, "nan", 3); return p + 3; }
  if (std::isinf(v)) { if (v < 0) *p++ = '-'; memcpy(p, "inf", 3); return p + 3; }
  return p + snprintf(p, 400, "%.2f", v);
}
inline size_t len_f2(double v) {
  unsigned long long q;
  if (f2_fast(v, &q)) return (size_t)(std::signbit(v) ? 1 : 0) + (size_t)ndigits64(q / 100) + 3;
  if (v != v) return 3;
  if (std::isinf(v)) return v < 0 ? 4 : 3;
  return (size_t)snprintf(nullptr, 0, "%.2f", v);
}
inline size_t len_int(long long v) {
  return v < 0 ? 1 + (size_t)ndigits64((unsigned long long)(-(v + 1)) + 1ull) : (size_t)ndigits64((unsigned long long)v);
}
// "%d" % float of Python / numpy.savetxt: truncation toward zero
inline char* put_trunc(char* p, double v) {
  if (!(fabs(v) < 9.0e18)) return p + snprintf(p, 400, "%.0f", trunc(v));
  return put_int(p, (long long)v);
}
inline size_t len_trunc(double v) {
  if (!(fabs(v) < 9.0e18)) return (size_t)snprintf(nullptr, 0, "%.0f", trunc(v));
  return len_int((long long)v);
}

}  // namespace
}  // namespace hia

// Pass 1: number of entries the file holds after blank-line skipping and dropna, the largest row and
// column index (shape when no window file is given) and whether any value is negative.
// Returns the entry count (>= 0) or a negative hia_status (-1: cannot open / malformed line).
HIA_API int64_t hia_mtx_scan(const char* path, int32_t threads, int64_t* max_row, int64_t* max_col,
                             int32_t* has_neg) {
  using namespace hia;
  if (!path) return HIA_ERR_BAD_ARG;
  Mapped m;
  if (!m.open(path)) return HIA_ERR_BAD_ARG;
  std::vector<Slice> sl = cut(m, default_threads(threads));
  std::vector<std::thread> pool;
  for (size_t i = 0; i < sl.size(); ++i)
    pool.emplace_back([&, i] {
      Slice& s = sl[i];
      for_lines(m, s, [&](const char* b, const char* e) {
        long long r, c;
        double v;
        const int st = parse_line_any(b, e, &r, &c, &v);
        if (st < 0) s.bad = 1;
        if (st == 1) {
          ++s.count;
          if (r > s.max_r) s.max_r = r;
          if (c > s.max_c) s.max_c = c;
          if (v < 0) s.any_neg = 1;
        }
      });
    });
  for (auto& t : pool) t.join();
  long long total = 0, mr = -1, mc = -1;
  int neg = 0;
  for (const Slice& s : sl) {
    if (s.bad) return HIA_ERR_BAD_ARG;
    total += s.count;
    mr = std::max(mr, s.max_r);
    mc = std::max(mc, s.max_c);
    neg |= s.any_neg;
  }
  if (max_row) *max_row = mr;
  if (max_col) *max_col = mc;
  if (has_neg) *has_neg = neg;
  return total;
}

// Upper bound on the entries of a file: its number of lines (a memchr pass, no parsing).
HIA_API int64_t hia_mtx_count_lines(const char* path, int32_t threads) {
  using namespace hia;
  if (!path) return HIA_ERR_BAD_ARG;
  Mapped m;
  if (!m.open(path)) return HIA_ERR_BAD_ARG;
  std::vector<Slice> sl = cut(m, default_threads(threads));
  std::vector<std::thread> pool;
  for (size_t i = 0; i < sl.size(); ++i)
    pool.emplace_back([&, i] { for_lines(m, sl[i], [&](const char*, const char*) { ++sl[i].count; }); });
  for (auto& t : pool) t.join();
  long long total = 0;
  for (const Slice& s : sl) total += s.count;
  return total;
}

// Parse into caller-owned HOST arrays of `capacity` entries (>= the entries of the file), file order
// kept. With capacity >= the LINES of the file (hia_mtx_count_lines) the slices parse directly into
// the arrays; otherwise through per-thread buffers. rows / cols: int64 (like the DataFrame columns); vals: fp64; any of
// rows32 / cols32 / vals32 may be given in addition (the int32 / fp32 copies the device kernels take;
// pass pinned memory to upload them without another pass). *n_out = entries written.
HIA_API int hia_mtx_parse(const char* path, int32_t threads, int64_t capacity, int64_t* rows, int64_t* cols,
                          double* vals, int32_t* rows32, int32_t* cols32, float* vals32, int64_t* n_out) {
  using namespace hia;
  if (!path || capacity < 0 || !n_out) return HIA_ERR_BAD_ARG;
  Mapped m;
  if (!m.open(path)) return HIA_ERR_BAD_ARG;
  std::vector<Slice> sl = cut(m, default_threads(threads));
  const size_t S = sl.size();
  std::vector<std::thread> pool;
  auto run = [&](auto&& fn) {
    for (size_t i = 0; i < S; ++i) pool.emplace_back(fn, i);
    for (auto& t : pool) t.join();
    pool.clear();
  };
  auto store = [&](long long k, long long r, long long c, double v) {
    if (rows) rows[k] = r;
    if (cols) cols[k] = c;
    if (vals) vals[k] = v;
    if (rows32) rows32[k] = (int32_t)r;
    if (cols32) cols32[k] = (int32_t)c;
    if (vals32) vals32[k] = (float)v;
  };
  // lines per slice (a memchr pass): an upper bound on its entries, hence a safe write offset
  std::vector<long long> lines(S, 0), off(S + 1, 0), got(S, 0);
  run([&](size_t i) { long long n = 0; for_lines(m, sl[i], [&](const char*, const char*) { ++n; }); lines[i] = n; });
  for (size_t i = 0; i < S; ++i) off[i + 1] = off[i] + lines[i];
  if (off[S] <= capacity) {
    // DIRECT: every slice parses straight into the caller's arrays at its line offset (no staging
    // copy, half the freshly touched memory); slices that dropped lines are closed up afterwards.
    run([&](size_t i) {
      long long k = off[i];
      for_lines(m, sl[i], [&](const char* b, const char* e) {
        long long r, c;
        double v;
        const int st = parse_line_any(b, e, &r, &c, &v);
        if (st < 0) sl[i].bad = 1;
        if (st == 1) store(k++, r, c, v);
      });
      got[i] = k - off[i];
    });
    long long w = 0;
    for (size_t i = 0; i < S; ++i) {
      if (sl[i].bad) return HIA_ERR_BAD_ARG;
      if (w != off[i] && got[i] > 0) {
        const long long src = off[i], cnt = got[i];
        if (rows) memmove(rows + w, rows + src, (size_t)cnt * sizeof(int64_t));
        if (cols) memmove(cols + w, cols + src, (size_t)cnt * sizeof(int64_t));
        if (vals) memmove(vals + w, vals + src, (size_t)cnt * sizeof(double));
        if (rows32) memmove(rows32 + w, rows32 + src, (size_t)cnt * sizeof(int32_t));
        if (cols32) memmove(cols32 + w, cols32 + src, (size_t)cnt * sizeof(int32_t));
        if (vals32) memmove(vals32 + w, vals32 + src, (size_t)cnt * sizeof(float));
      }
      w += got[i];
    }
    *n_out = w;
    return HIA_OK;
  }
  // STAGED (capacity below the line count, e.g. the exact entry count of hia_mtx_scan): parse into
  // per-thread buffers, then copy to the final offsets.
  struct Entry { long long r, c; double v; };
  std::vector<std::vector<Entry>> parsed(S);
  run([&](size_t i) {
    std::vector<Entry>& out = parsed[i];
    out.reserve((size_t)lines[i]);
    for_lines(m, sl[i], [&](const char* b, const char* e) {
      Entry en;
      const int st = parse_line_any(b, e, &en.r, &en.c, &en.v);
      if (st < 0) sl[i].bad = 1;
      if (st == 1) out.push_back(en);
    });
  });
  for (size_t i = 0; i < S; ++i) {
    if (sl[i].bad) return HIA_ERR_BAD_ARG;
    off[i + 1] = off[i] + (long long)parsed[i].size();
  }
  if (off[S] > capacity) return HIA_ERR_WORKSPACE;
  run([&](size_t i) {
    long long k = off[i];
    for (const Entry& en : parsed[i]) store(k++, en.r, en.c, en.v);
  });
  *n_out = off[S];
  return HIA_OK;
}

// np.savetxt(path, [rows, cols, vals].T, fmt="%d\t%d\t%.2f")  (value_format 0)  or  "%d\t%d\t%d"
// (value_format 1: out_int=True, the value truncated toward zero like Python's '%d' % float).
HIA_API int hia_mtx_write(const char* path, int64_t n, const int64_t* rows, const int64_t* cols, const double* vals,
                          int value_format, int32_t threads) {
  using namespace hia;
  if (!path || n < 0 || (n > 0 && (!rows || !cols || !vals)) || value_format < 0 || value_format > 1)
    return HIA_ERR_BAD_ARG;
  const int fd = ::open(path, O_RDWR | O_CREAT | O_TRUNC, 0644);
  if (fd < 0) return HIA_ERR_BAD_ARG;
  if (n == 0) return ::close(fd) == 0 ? HIA_OK : HIA_ERR_BAD_ARG;
  const int T = (int)std::max<int64_t>(1, std::min<int64_t>(default_threads(threads), n / 65536 + 1));
  std::vector<std::thread> pool;
  auto run = [&](auto&& fn) {
    for (int t = 0; t < T; ++t) pool.emplace_back(fn, t);
    for (auto& th : pool) th.join();
    pool.clear();
  };
  // pass 1: exact byte length of every thread's share -> its offset in the file
  std::vector<long long> off(T + 1, 0);
  run([&](int t) {
    const int64_t b = n * t / T, e = n * (t + 1) / T;
    size_t len = 0;
    for (int64_t i = b; i < e; ++i)
      len += len_int(rows[i]) + len_int(cols[i]) + (value_format == 0 ? len_f2(vals[i]) : len_trunc(vals[i])) + 3;
    off[t + 1] = (long long)len;
  });
  for (int t = 0; t < T; ++t) off[t + 1] += off[t];
  const size_t total = (size_t)off[T];
  // pass 2: format straight into the mapped file (no staging buffers; the page-cache pages of
  // different threads fault in concurrently). File systems without shared writable mappings or
  // without ftruncate (pipes, /dev/null) get a staged write instead.
  char* map = nullptr;
  if (ftruncate(fd, (off_t)total) == 0) {
    void* m = mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    if (m != MAP_FAILED) map = static_cast<char*>(m);
  }
  std::vector<char> staged;
  char* base = map;
  if (!base) {
    staged.resize(total + 512);                           // +512: snprintf's terminating NUL at the very end
    base = staged.data();
  }
  std::vector<int> bad(T, 0);
  run([&](int t) {
    const int64_t b = n * t / T, e = n * (t + 1) / T;
    char* p = base + off[t];
    char* const end = base + off[t + 1];
    char tail[512];
    for (int64_t i = b; i < e; ++i) {
      // the last entries of a share are formatted on the stack: snprintf writes a NUL past the text
      char* const q0 = (end - p < 480) ? tail : p;
      char* q = put_int(q0, rows[i]);
      *q++ = '\t';
      q = put_int(q, cols[i]);
      *q++ = '\t';
      q = value_format == 0 ? put_f2(q, vals[i]) : put_trunc(q, vals[i]);
      *q++ = '\n';
      const size_t len = (size_t)(q - q0);
      if (len > (size_t)(end - p)) { bad[t] = 1; return; }   // would disagree with pass 1: never
      if (q0 == tail) memcpy(p, tail, len);
      p += len;
    }
    if (p != end) bad[t] = 1;
  });
  int rc = HIA_OK;
  for (int t = 0; t < T; ++t) if (bad[t]) rc = HIA_ERR_BAD_ARG;
  if (map) {
    if (munmap(map, total) != 0) rc = HIA_ERR_BAD_ARG;
  } else if (rc == HIA_OK) {
    size_t done = 0;
    while (done < total) {
      const ssize_t w = ::write(fd, base + done, total - done);
      if (w <= 0) { rc = HIA_ERR_BAD_ARG; break; }
      done += (size_t)w;
    }
  }
  if (::close(fd) != 0) rc = HIA_ERR_BAD_ARG;
  return rc;
}
