// draco_sharp_b200/synth/drc_synth.cpp
//
// Synthetic Draco v2.2 bitstream generator (bench + test TOOLING, not the decode product and
// not the oracle).  Writes valid point-cloud `.drc` buffers of the BASELINE.json shapes:
// sequential attribute encoding, quantized positions (delta prediction + wrap transform),
// optional octahedral normals (delta + canonicalized octahedron transform) and uint8 RGB
// colours (delta + wrap), symbols coded with the Raw or Tagged rANS scheme.
//
// It is an independent encoder written from the bitstream layout.  The
// reference's own C# encoder cannot be the generator: it emits corrupt probability tables
// (src/Draco/IO/Entropy/RAnsSymbolEncoder.cs:132-135) and cannot write point clouds
// (src/Draco/IO/DracoEncoder.cs:71-74).  Format anchors in the reference:
//   container      src/Draco/IO/DracoDecoder.cs:44-64, ConnectivityDecoder.cs:16-44,
//                  Attributes/AttributesDecoder.cs:19-63, SequentialAttributeDecodersController.cs:16-27
//   symbols        Entropy/SymbolEncoding.cs:8-192 (scheme rule, tagged layout, raw layout)
//   rANS table     Entropy/RAnsSymbolEncoder.cs:15-164 (probability normalisation + table bytes)
//   rANS payload   Entropy/RAnsEncoder.cs:15-30, AnsEncoder.cs (write_end size tag)
//   wrap           Attributes/PredictionSchemes/PredictionSchemeWrapEncodingTransform.cs:45-96
//   oct transform  Attributes/PredictionSchemes/PredictionSchemeNormalOctahedronCanonicalizedEncodingTransform.cs:62-90
//   oct toolbox    Attributes/OctahedronToolBox.cs:28-119
//
// Determinism: all randomness comes from splitmix64/xoshiro256** seeded by spec.seed; the delta
// distribution is a two-sided geometric built with integer arithmetic only (no libm), so the same
// seed gives the same bytes on every machine.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {

typedef struct synth_spec {
  uint64_t seed;
  uint32_t n_points;
  int32_t pos_bits;     // quantization bits of the position attribute (0 = no positions)
  int32_t rho_num;      // two-sided geometric ratio rho = rho_num / rho_den (24/25 ~ Laplace b=24.5)
  int32_t rho_den;
  int32_t scheme;       // -1 = upstream selection rule, 0 = force Tagged, 1 = force Raw
  int32_t normal_bits;  // 0 = no normals, else octahedral quantization bits (10)
  int32_t colors;       // 0 = none, 1 = uint8 x 3 RGB
  int32_t color_step;   // max |step| of the colour random walk (3)
  int32_t reserved;
} synth_spec;

typedef struct synth_truth {
  // optional outputs (may be NULL): the generator's source integers, in entry order
  int32_t *pos_q;    // [n*3]
  int32_t *nrm_st;   // [n*2]
  uint8_t *rgb;      // [n*3]
  int32_t scheme[3]; // scheme chosen per attribute (0 tagged / 1 raw), -1 if attribute absent
  uint64_t sums[3];  // position-weighted 32-bit word checksum of the expected OUTPUT bytes per attribute
} synth_truth;
}

namespace {

struct Rng {
  uint64_t s[4];
  static uint64_t splitmix(uint64_t &x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed) {
    for (auto &v : s) v = splitmix(seed);
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  uint32_t u32() { return (uint32_t)(next() >> 32); }
};

// two-sided geometric sampler P(d = k) ~ rho^|k|, integer-only construction, 32-bit inverse CDF
struct DeltaSampler {
  std::vector<uint32_t> cdf;  // cdf[j] = upper bound (exclusive, scaled 2^32) of outcome j
  std::vector<int32_t> val;   // outcome j -> delta
  std::vector<uint16_t> fast; // top-16-bit bucket -> first outcome index whose cdf exceeds bucket start
  DeltaSampler(int num, int den) {
    // weights w_k = rho^k in 2^62 fixed point, two-sided: k = 0, +1, -1, +2, -2, ...
    std::vector<unsigned __int128> w;
    unsigned __int128 cur = (unsigned __int128)1 << 62, total = 0;
    std::vector<int32_t> ks;
    for (int k = 0; k < 4096 && cur > 0; ++k) {
      if (k == 0) { w.push_back(cur); ks.push_back(0); total += cur; }
      else { w.push_back(cur); ks.push_back(k); w.push_back(cur); ks.push_back(-k); total += 2 * cur; }
      cur = cur * (unsigned)num / (unsigned)den;
      if (cur < ((unsigned __int128)1 << 20)) break;
    }
    unsigned __int128 acc = 0;
    for (size_t j = 0; j < w.size(); ++j) {
      acc += w[j];
      unsigned __int128 b = (acc << 32) / total;
      uint32_t ub = b >= ((unsigned __int128)1 << 32) ? 0xFFFFFFFFu : (uint32_t)b;
      if (j + 1 == w.size()) ub = 0xFFFFFFFFu;
      if (!cdf.empty() && ub <= cdf.back()) continue;  // zero-width outcome at 2^-32 resolution
      cdf.push_back(ub);
      val.push_back(ks[j]);
    }
    fast.resize(65536);
    size_t j = 0;
    for (uint32_t b = 0; b < 65536; ++b) {
      uint32_t start = b << 16;
      while (j + 1 < cdf.size() && cdf[j] <= start) ++j;
      fast[b] = (uint16_t)j;
    }
  }
  int32_t draw(Rng &r) const {
    uint32_t u = r.u32();
    size_t j = fast[u >> 16];
    while (j + 1 < cdf.size() && cdf[j] <= u) ++j;
    return val[j];
  }
};

struct Out {
  std::vector<uint8_t> b;
  void u8(uint8_t v) { b.push_back(v); }
  void bytes(const void *p, size_t n) { const uint8_t *q = (const uint8_t *)p; b.insert(b.end(), q, q + n); }
  void u16(uint16_t v) { bytes(&v, 2); }
  void i32(int32_t v) { bytes(&v, 4); }
  void f32(float v) { bytes(&v, 4); }
  void varint(uint64_t v) {
    do {
      uint8_t c = v & 0x7F;
      v >>= 7;
      if (v) c |= 0x80;
      b.push_back(c);
    } while (v);
  }
};

int msb(uint32_t v) { int m = -1; while (v) { ++m; v >>= 1; } return m; }
int rans_precision(int mbl) { int p = 3 * mbl / 2; return p < 12 ? 12 : (p > 20 ? 20 : p); }

// Probability normalisation: RAnsSymbolEncoder.cs:15-110 (upstream algorithm; ties broken by index)
bool build_probs(const std::vector<uint64_t> &freq, int prec_bits, std::vector<uint32_t> &prob, uint32_t &num_symbols) {
  const uint32_t precision = 1u << prec_bits;
  uint64_t total = 0;
  int max_valid = 0;
  for (size_t i = 0; i < freq.size(); ++i) { total += freq[i]; if (freq[i]) max_valid = (int)i; }
  num_symbols = (uint32_t)max_valid + 1;
  prob.assign(num_symbols, 0);
  if (total == 0) return false;
  const double total_d = (double)total, prec_d = (double)precision;
  int64_t total_prob = 0;
  for (uint32_t i = 0; i < num_symbols; ++i) {
    double p = (double)freq[i] / total_d;
    uint32_t rp = (uint32_t)(p * prec_d + 0.5f);
    if (rp == 0 && freq[i] > 0) rp = 1;
    prob[i] = rp;
    total_prob += rp;
  }
  if (total_prob != precision) {
    std::vector<int> sorted(num_symbols);
    for (uint32_t i = 0; i < num_symbols; ++i) sorted[i] = (int)i;
    std::stable_sort(sorted.begin(), sorted.end(), [&](int a, int b) { return prob[a] < prob[b]; });
    if (total_prob < precision) {
      prob[sorted.back()] += (uint32_t)(precision - total_prob);
    } else {
      int64_t error = total_prob - precision;
      while (error > 0) {
        double rel = prec_d / (double)total_prob;
        for (int j = (int)num_symbols - 1; j >= 0; --j) {
          int sid = sorted[j];
          if (prob[sid] <= 1) { if (j == (int)num_symbols - 1) return false; break; }
          int32_t np = (int32_t)std::floor(rel * prob[sid]);
          int32_t fix = (int32_t)prob[sid] - np;
          if (fix == 0) fix = 1;
          if (fix >= (int32_t)prob[sid]) fix = (int32_t)prob[sid] - 1;
          if (fix > error) fix = (int32_t)error;
          prob[sid] -= (uint32_t)fix;
          total_prob -= fix;
          error -= fix;
          if (total_prob == precision) break;
        }
      }
    }
  }
  return true;
}

// RANS_TABLE bytes: RAnsSymbolEncoder.cs:121-164 with the upstream thresholds (2^6, 2^14, 2^22)
void write_table(Out &o, const std::vector<uint32_t> &prob, uint32_t num_symbols) {
  o.varint(num_symbols);
  for (uint32_t i = 0; i < num_symbols; ++i) {
    uint32_t p = prob[i];
    if (p == 0) {
      uint32_t offset = 0;
      for (; offset < 63; ++offset) {
        if (i + offset + 1 >= num_symbols) break;
        if (prob[i + offset + 1] > 0) break;
      }
      o.u8((uint8_t)((offset << 2) | 3));
      i += offset;
    } else {
      int extra = 0;
      if (p >= (1u << 6)) { ++extra; if (p >= (1u << 14)) ++extra; }
      o.u8((uint8_t)((p << 2) | (uint32_t)extra));
      for (int b = 0; b < extra; ++b) o.u8((uint8_t)(p >> (8 * (b + 1) - 2)));
    }
  }
}

// rANS payload: symbols are pushed in REVERSE so the decoder pops them forward
// (SymbolEncoding.cs:177-183, RAnsEncoder.cs:22-30, upstream ans.h write_end size tag)
struct RansEnc {
  int prec_bits; uint32_t precision, l_base, state;
  std::vector<uint8_t> buf;
  explicit RansEnc(int pb) : prec_bits(pb), precision(1u << pb), l_base(4u << pb), state(4u << pb) {}
  inline void put(uint32_t prob, uint32_t cum) {
    const uint64_t lim = (uint64_t)(l_base / precision) * 256u * prob;
    while (state >= lim) { buf.push_back((uint8_t)(state & 0xFF)); state >>= 8; }
    state = (state / prob) * precision + state % prob + cum;
  }
  void finish() {
    uint32_t s = state - l_base;
    if (s < (1u << 6)) { buf.push_back((uint8_t)s); }
    else if (s < (1u << 14)) { uint32_t v = (1u << 14) + s; buf.push_back(v & 0xFF); buf.push_back(v >> 8); }
    else if (s < (1u << 22)) { uint32_t v = (2u << 22) + s; buf.push_back(v & 0xFF); buf.push_back((v >> 8) & 0xFF); buf.push_back(v >> 16); }
    else { uint32_t v = (3u << 30) + s; buf.push_back(v & 0xFF); buf.push_back((v >> 8) & 0xFF); buf.push_back((v >> 16) & 0xFF); buf.push_back(v >> 24); }
  }
};

struct BitWriter {  // LSB-first within bytes (EncoderBuffer.cs:172-186)
  std::vector<uint8_t> b; uint64_t acc = 0; int nacc = 0;
  inline void put(uint32_t v, int n) {
    if (n == 0) return;
    acc |= (uint64_t)(n == 32 ? v : (v & ((1u << n) - 1u))) << nacc;
    nacc += n;
    while (nacc >= 8) { b.push_back((uint8_t)acc); acc >>= 8; nacc -= 8; }
  }
  void finish() { if (nacc > 0) { b.push_back((uint8_t)acc); acc = 0; nacc = 0; } }
};

double shannon_bits(const std::vector<uint64_t> &freq, uint64_t n, int &unique) {
  double bits = 0; unique = 0;
  for (uint64_t f : freq) if (f) { ++unique; bits += (double)f * std::log2((double)f / (double)n); }
  return -bits;
}
int64_t approx_table_bits(int max_value, int num_unique) {  // RAnsSymbolCoding.cs:35-41
  int64_t zero_bits = 8 * ((int64_t)num_unique + (max_value - num_unique) / 64);
  return 8 * (int64_t)num_unique + zero_bits;
}

// SYMBOLS(n*nc, nc): SymbolEncoding.cs:8-192.  Returns the scheme used.
int encode_symbols(Out &o, const std::vector<uint32_t> &sym, int nc, int force_scheme) {
  if (sym.empty()) return -1;
  const size_t nv = sym.size(), n = nv / nc;
  std::vector<uint8_t> bitlen(n);
  uint32_t max_value = 0;
  for (size_t i = 0; i < n; ++i) {
    uint32_t m = 0;
    for (int c = 0; c < nc; ++c) m = std::max(m, sym[i * nc + c]);
    bitlen[i] = (uint8_t)((m > 0 ? msb(m) : 0) + 1);
    max_value = std::max(max_value, m);
  }
  std::vector<uint64_t> freq((size_t)max_value + 1, 0);
  for (uint32_t s : sym) ++freq[s];
  std::vector<uint64_t> tagfreq(33, 0);
  uint64_t total_bitlen = 0;
  for (uint8_t b : bitlen) { ++tagfreq[b]; total_bitlen += b; }
  int scheme = force_scheme;
  int num_unique = 0;
  double raw_data_bits = shannon_bits(freq, nv, num_unique);
  if (scheme < 0) {  // upstream rule (SymbolEncoding.cs:12-30 mirrors it with the comparison typos)
    int tag_unique = 0;
    int64_t tag_bits = (int64_t)shannon_bits(tagfreq, n, tag_unique);
    int64_t tagged_total = tag_bits + approx_table_bits(tag_unique, tag_unique) + (int64_t)total_bitlen * nc;
    int64_t raw_total = approx_table_bits((int)max_value, num_unique) + (int64_t)raw_data_bits;
    int max_value_bit_length = msb(std::max(1u, max_value)) + 1;
    scheme = (tagged_total < raw_total || max_value_bit_length > 18) ? 0 : 1;
  }
  o.u8((uint8_t)scheme);
  if (scheme == 1) {
    int unique_bit_length = (num_unique > 0 ? msb((uint32_t)num_unique) : 0) + 1;  // compression level 7: no adjustment
    unique_bit_length = std::min(std::max(1, unique_bit_length), 18);
    o.u8((uint8_t)unique_bit_length);
    const int pb = rans_precision(unique_bit_length);
    std::vector<uint32_t> prob; uint32_t ns = 0;
    build_probs(freq, pb, prob, ns);
    write_table(o, prob, ns);
    std::vector<uint32_t> cum(ns);
    uint32_t c = 0;
    for (uint32_t i = 0; i < ns; ++i) { cum[i] = c; c += prob[i]; }
    RansEnc enc(pb);
    enc.buf.reserve((size_t)(raw_data_bits / 8 * 1.1) + 64);
    for (size_t i = nv; i-- > 0;) enc.put(prob[sym[i]], cum[sym[i]]);
    enc.finish();
    o.varint(enc.buf.size());
    o.bytes(enc.buf.data(), enc.buf.size());
  } else {
    const int pb = rans_precision(5);
    std::vector<uint32_t> prob; uint32_t ns = 0;
    build_probs(tagfreq, pb, prob, ns);
    write_table(o, prob, ns);
    std::vector<uint32_t> cum(ns);
    uint32_t c = 0;
    for (uint32_t i = 0; i < ns; ++i) { cum[i] = c; c += prob[i]; }
    RansEnc enc(pb);
    for (size_t i = n; i-- > 0;) enc.put(prob[bitlen[i]], cum[bitlen[i]]);
    enc.finish();
    o.varint(enc.buf.size());
    o.bytes(enc.buf.data(), enc.buf.size());
    BitWriter bw;
    bw.b.reserve((size_t)(total_bitlen * nc / 8) + 16);
    for (size_t i = 0; i < n; ++i)
      for (int cc = 0; cc < nc; ++cc) bw.put(sym[i * nc + cc], bitlen[i]);
    bw.finish();
    o.bytes(bw.b.data(), bw.b.size());
  }
  return scheme;
}

inline uint32_t zigzag_enc(int32_t v) { return v >= 0 ? ((uint32_t)v << 1) : ((((uint32_t)(-(v + 1))) << 1) | 1u); }

// delta prediction + wrap transform, encoder side (PredictionSchemeWrapEncodingTransform.cs:45-88)
void delta_wrap_symbols(const int32_t *q, size_t n, int nc, int32_t &mn, int32_t &mx, std::vector<uint32_t> &sym) {
  sym.resize(n * nc);
  if (n == 0) { mn = 0; mx = 0; return; }
  mn = mx = q[0];
  for (size_t i = 1; i < n * nc; ++i) { mn = std::min(mn, q[i]); mx = std::max(mx, q[i]); }
  const int32_t max_diff = 1 + mx - mn;
  int32_t max_corr = max_diff / 2, min_corr = -max_corr;
  if ((max_diff & 1) == 0) max_corr -= 1;
  for (size_t i = 0; i < n * nc; ++i) {
    int32_t pred = i < (size_t)nc ? 0 : q[i - nc];
    pred = pred > mx ? mx : (pred < mn ? mn : pred);
    int32_t corr = q[i] - pred;
    if (corr < min_corr) corr += max_diff; else if (corr > max_corr) corr -= max_diff;
    sym[i] = zigzag_enc(corr);
  }
}

// ---- octahedral tool box (OctahedronToolBox.cs) ----
struct OctBox {
  int32_t bits, max_q, max_value, center;
  explicit OctBox(int b) : bits(b), max_q((1 << b) - 1), max_value(max_q - 1), center(max_value / 2) {}
  bool in_diamond(int32_t s, int32_t t) const { return (uint32_t)std::abs(s) + (uint32_t)std::abs(t) <= (uint32_t)center; }
  void invert_diamond(int32_t &s, int32_t &t) const {
    int32_t ss, st;
    if (s >= 0 && t >= 0) { ss = 1; st = 1; } else if (s <= 0 && t <= 0) { ss = -1; st = -1; } else { ss = s > 0 ? 1 : -1; st = t > 0 ? 1 : -1; }
    int32_t cs = ss * center, ct = st * center, us = s + s - cs, ut = t + t - ct, tmp = us;
    if (ss * st >= 0) { us = -ut; ut = -tmp; } else { us = ut; ut = tmp; }
    us += cs; ut += ct; s = us / 2; t = ut / 2;
  }
  int32_t make_positive(int32_t x) const { return x < 0 ? x + max_q : x; }
  void canonicalize(int32_t &s, int32_t &t) const {  // :28-54
    if ((s == 0 && t == 0) || (s == 0 && t == max_value) || (s == max_value && t == 0)) { s = max_value; t = max_value; }
    else if (s == 0 && t > center) t = center - (t - center);
    else if (s == max_value && t < center) t = center + (center - t);
    else if (t == max_value && s < center) s = center + (center - s);
    else if (t == 0 && s > center) s = center - (s - center);
  }
  void from_unit(const double v[3], int32_t &s, int32_t &t) const {  // :79-119 + :61-77
    double abs_sum = std::fabs(v[0]) + std::fabs(v[1]) + std::fabs(v[2]);
    double sv[3];
    if (abs_sum > 1e-6) { double sc = 1.0 / abs_sum; sv[0] = v[0] * sc; sv[1] = v[1] * sc; sv[2] = v[2] * sc; }
    else { sv[0] = 1; sv[1] = 0; sv[2] = 0; }
    int32_t iv[3];
    iv[0] = (int32_t)std::floor(sv[0] * center + 0.5);
    iv[1] = (int32_t)std::floor(sv[1] * center + 0.5);
    iv[2] = center - std::abs(iv[0]) - std::abs(iv[1]);
    if (iv[2] < 0) { if (iv[1] > 0) iv[1] += iv[2]; else iv[1] -= iv[2]; iv[2] = 0; }
    if (sv[2] < 0) iv[2] *= -1;
    if (iv[0] >= 0) { s = iv[1] + center; t = iv[2] + center; }
    else {
      s = iv[1] < 0 ? std::abs(iv[2]) : max_value - std::abs(iv[2]);
      t = iv[2] < 0 ? std::abs(iv[1]) : max_value - std::abs(iv[1]);
    }
    canonicalize(s, t);
  }
};
inline int rotation_count(int32_t x, int32_t y) {
  if (x == 0) return y == 0 ? 0 : (y > 0 ? 3 : 1);
  if (x > 0) return y >= 0 ? 2 : 1;
  return y <= 0 ? 0 : 3;
}
inline void rotate(int32_t &a, int32_t &b, int rot) {
  int32_t x = a, y = b;
  if (rot == 1) { a = y; b = -x; } else if (rot == 2) { a = -x; b = -y; } else if (rot == 3) { a = -y; b = x; }
}
// delta prediction + canonicalized octahedron transform, encoder side (upstream semantics)
void delta_oct_symbols(const int32_t *st, size_t n, const OctBox &box, std::vector<uint32_t> &sym) {
  sym.resize(n * 2);
  for (size_t i = 0; i < n; ++i) {
    int32_t o0 = st[2 * i] - box.center, o1 = st[2 * i + 1] - box.center;
    int32_t p0 = (i ? st[2 * i - 2] : 0) - box.center, p1 = (i ? st[2 * i - 1] : 0) - box.center;
    if (!box.in_diamond(p0, p1)) { box.invert_diamond(o0, o1); box.invert_diamond(p0, p1); }
    bool bottom_left = (p0 == 0 && p1 == 0) || (p0 < 0 && p1 <= 0);
    if (!bottom_left) { int rot = rotation_count(p0, p1); rotate(o0, o1, rot); rotate(p0, p1, rot); }
    sym[2 * i] = (uint32_t)box.make_positive(o0 - p0);
    sym[2 * i + 1] = (uint32_t)box.make_positive(o1 - p1);
  }
}

inline uint64_t word_checksum(const void *p, size_t nbytes) {
  const uint8_t *b = (const uint8_t *)p;
  uint64_t sum = 0, i = 0;
  size_t nw = nbytes / 4;
  for (; i < nw; ++i) { uint32_t w; memcpy(&w, b + 4 * i, 4); sum += (i + 1) * (uint64_t)w; }
  if (nbytes & 3) { uint32_t w = 0; memcpy(&w, b + 4 * nw, nbytes & 3); sum += (nw + 1) * (uint64_t)w; }
  return sum;
}

void gen_cloud(const synth_spec &sp, const DeltaSampler &ds, std::vector<uint8_t> &out, synth_truth *truth) {
  Rng rng(sp.seed);
  const size_t n = sp.n_points;
  Out o;
  o.bytes("DRACO", 5);
  o.u8(2); o.u8(2);  // version 2.2
  o.u8(0);           // POINT_CLOUD
  o.u8(0);           // sequential encoding
  o.u16(0);          // flags
  o.i32((int32_t)n);
  o.u8(1);           // one attributes decoder
  const bool has_pos = sp.pos_bits > 0, has_nrm = sp.normal_bits > 0, has_rgb = sp.colors != 0;
  const int n_attr = (int)has_pos + (int)has_nrm + (int)has_rgb;
  o.varint((uint64_t)n_attr);
  uint32_t uid = 0;
  if (has_pos) { o.u8(0); o.u8(9); o.u8(3); o.u8(0); o.varint(uid++); }  // POSITION float32 x3
  if (has_nrm) { o.u8(1); o.u8(9); o.u8(3); o.u8(0); o.varint(uid++); }  // NORMAL   float32 x3
  if (has_rgb) { o.u8(2); o.u8(2); o.u8(3); o.u8(1); o.varint(uid++); }  // COLOR    uint8   x3 normalized
  if (has_pos) o.u8(2);  // SEQUENTIAL_ATTRIBUTE_ENCODER_QUANTIZATION
  if (has_nrm) o.u8(3);  // SEQUENTIAL_ATTRIBUTE_ENCODER_NORMALS
  if (has_rgb) o.u8(1);  // SEQUENTIAL_ATTRIBUTE_ENCODER_INTEGER
  if (truth) { truth->scheme[0] = truth->scheme[1] = truth->scheme[2] = -1; truth->sums[0] = truth->sums[1] = truth->sums[2] = 0; }
  std::vector<uint32_t> sym;
  // ---- positions ----
  float pos_min[3] = {-1.0f, -1.0f, -1.0f};
  float pos_range = 2.0f;
  if (has_pos) {
    const int32_t maxq = (1 << sp.pos_bits) - 1;
    std::vector<int32_t> q(n * 3);
    for (size_t i = 0; i < n; ++i)
      for (int c = 0; c < 3; ++c) {
        int32_t v;
        if (i == 0) v = (int32_t)(rng.u32() % (uint32_t)(maxq + 1));
        else {
          v = q[(i - 1) * 3 + c] + ds.draw(rng);
          while (v < 0 || v > maxq) { if (v < 0) v = -v; if (v > maxq) v = 2 * maxq - v; }  // reflect
        }
        q[i * 3 + c] = v;
      }
    int32_t mn, mx;
    delta_wrap_symbols(q.data(), n, 3, mn, mx, sym);
    o.u8(0);  // PREDICTION_DIFFERENCE
    o.u8(1);  // PREDICTION_TRANSFORM_WRAP
    o.u8(1);  // compressed
    int sch = encode_symbols(o, sym, 3, sp.scheme);
    o.i32(mn); o.i32(mx);
    if (truth) {
      truth->scheme[0] = sch;
      if (truth->pos_q) memcpy(truth->pos_q, q.data(), q.size() * 4);
      const float delta = pos_range / (float)maxq;
      std::vector<float> f(n * 3);
      for (size_t i = 0; i < n * 3; ++i) { volatile float p = (float)q[i] * delta; f[i] = p + pos_min[i % 3]; }
      truth->sums[0] = word_checksum(f.data(), f.size() * 4);
    }
  }
  // ---- normals ----
  if (has_nrm) {
    OctBox box(sp.normal_bits);
    std::vector<int32_t> st(n * 2);
    // smooth random field: a direction doing a small random walk on the sphere (integer steps / 2^14)
    double v[3] = {0.3, 0.5, 0.8};
    for (size_t i = 0; i < n; ++i) {
      for (int c = 0; c < 3; ++c) v[c] += (double)((int32_t)(rng.u32() % 2049u) - 1024) / 16384.0;
      double nn = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      if (nn < 1e-3) { v[0] = 1; v[1] = 0; v[2] = 0; nn = 1; }
      for (int c = 0; c < 3; ++c) v[c] /= nn;
      box.from_unit(v, st[2 * i], st[2 * i + 1]);
    }
    delta_oct_symbols(st.data(), n, box, sym);
    o.u8(0);  // PREDICTION_DIFFERENCE
    o.u8(3);  // PREDICTION_TRANSFORM_NORMAL_OCTAHEDRON_CANONICALIZED
    o.u8(1);
    int sch = encode_symbols(o, sym, 2, sp.scheme);
    o.i32(box.max_q); o.i32(box.center);
    if (truth) {
      truth->scheme[1] = sch;
      if (truth->nrm_st) memcpy(truth->nrm_st, st.data(), st.size() * 4);
    }
  }
  // ---- colours ----
  if (has_rgb) {
    std::vector<int32_t> q(n * 3);
    const uint32_t span = 2u * (uint32_t)sp.color_step + 1u;
    for (size_t i = 0; i < n; ++i)
      for (int c = 0; c < 3; ++c) {
        int32_t v = i == 0 ? (int32_t)(rng.u32() & 255u) : q[(i - 1) * 3 + c] + (int32_t)(rng.u32() % span) - sp.color_step;
        q[i * 3 + c] = v < 0 ? 0 : (v > 255 ? 255 : v);
      }
    int32_t mn, mx;
    delta_wrap_symbols(q.data(), n, 3, mn, mx, sym);
    o.u8(0); o.u8(1); o.u8(1);
    int sch = encode_symbols(o, sym, 3, sp.scheme);
    o.i32(mn); o.i32(mx);
    if (truth) {
      truth->scheme[2] = sch;
      std::vector<uint8_t> b(n * 3);
      for (size_t i = 0; i < n * 3; ++i) b[i] = (uint8_t)q[i];
      if (truth->rgb) memcpy(truth->rgb, b.data(), b.size());
      truth->sums[2] = word_checksum(b.data(), b.size());
    }
  }
  // ---- XFORM_PARAMS, in attribute order ----
  if (has_pos) { o.f32(pos_min[0]); o.f32(pos_min[1]); o.f32(pos_min[2]); o.f32(pos_range); o.u8((uint8_t)sp.pos_bits); }
  if (has_nrm) o.u8((uint8_t)sp.normal_bits);
  out.swap(o.b);
}

}  // namespace

extern "C" {

// One cloud.  Returns the byte size, or -(needed) if cap is too small.
int64_t synth_cloud(const synth_spec *sp, uint8_t *out, uint64_t cap, synth_truth *truth) {
  DeltaSampler ds(sp->rho_num, sp->rho_den);
  std::vector<uint8_t> b;
  gen_cloud(*sp, ds, b, truth);
  if (b.size() > cap) return -(int64_t)b.size();
  memcpy(out, b.data(), b.size());
  return (int64_t)b.size();
}

// Batch: cloud k uses seed base->seed + k.  Clouds are packed back to back in `arena`, each
// starting on a 16-byte boundary.  sums[3*k + a] receives the expected-output checksum of
// attribute slot a (0 positions, 1 normals (0: not computed), 2 colours).  Returns total bytes
// used, or -(needed) if the arena is too small (nothing written in that case).
int64_t synth_batch(const synth_spec *base, uint32_t n_bufs, int n_threads, uint8_t *arena, uint64_t cap,
                    uint64_t *offs, uint64_t *lens, uint64_t *sums, int32_t *schemes) {
  if (n_threads < 1) n_threads = 1;
  if ((uint32_t)n_threads > n_bufs) n_threads = (int)(n_bufs ? n_bufs : 1);
  DeltaSampler ds(base->rho_num, base->rho_den);
  std::vector<std::vector<std::vector<uint8_t>>> chunks((size_t)n_threads);
  std::vector<std::thread> th;
  auto work = [&](int t) {
    uint32_t k0 = (uint32_t)((uint64_t)n_bufs * t / n_threads), k1 = (uint32_t)((uint64_t)n_bufs * (t + 1) / n_threads);
    chunks[t].resize(k1 - k0);
    for (uint32_t k = k0; k < k1; ++k) {
      synth_spec sp = *base;
      sp.seed = base->seed + k;
      synth_truth tr;
      memset(&tr, 0, sizeof tr);
      gen_cloud(sp, ds, chunks[t][k - k0], &tr);
      if (sums) for (int a = 0; a < 3; ++a) sums[3 * (size_t)k + a] = tr.sums[a];
      if (schemes) for (int a = 0; a < 3; ++a) schemes[3 * (size_t)k + a] = tr.scheme[a];
    }
  };
  for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
  for (auto &x : th) x.join();
  uint64_t total = 0;
  for (auto &c : chunks) for (auto &b : c) total = ((total + 15) & ~15ull) + b.size();
  if (total > cap) return -(int64_t)total;
  uint64_t pos = 0; uint32_t k = 0;
  for (auto &c : chunks)
    for (auto &b : c) {
      pos = (pos + 15) & ~15ull;
      memcpy(arena + pos, b.data(), b.size());
      offs[k] = pos; lens[k] = b.size(); ++k;
      pos += b.size();
      std::vector<uint8_t>().swap(b);
    }
  return (int64_t)pos;
}

// ---------------------------------------------------------------------------------------------
// BASELINE configs[3]: a triangulated w x h grid surface (2 (w-1)(h-1) faces).  The Edgebreaker connectivity itself
// is the host's business (an opaque blob in the buffer); what the attribute path consumes is the corner table and the
// depth-first traversal maps, built here the way CornerTable / DepthFirstTraverser define them
// (IO/Mesh/CornerTable.cs:59-83, IO/Mesh/Traverser/DepthFirstTraverser.cs:9-99,
// MeshAttributeIndicesEncodingObserver.cs:14-21).
// ---------------------------------------------------------------------------------------------
static inline uint32_t cnext(uint32_t c) { return c % 3u == 2u ? c - 2u : c + 1u; }
static inline uint32_t cprev(uint32_t c) { return c % 3u == 0u ? c + 2u : c - 1u; }

// opposite / corner_to_vertex: [3F]; data_to_corner: [V]; vertex_to_data: [V].  Returns the number of entries (== V).
int64_t synth_grid_topology(uint32_t w, uint32_t h, uint32_t *opposite, uint32_t *c2v, uint32_t *d2c, int32_t *v2d) {
  if (w < 2 || h < 2) return -1;
  const uint32_t V = w * h, F = 2u * (w - 1) * (h - 1), INV = 0xFFFFFFFFu;
  uint32_t f = 0;
  for (uint32_t j = 0; j + 1 < h; ++j)
    for (uint32_t i = 0; i + 1 < w; ++i) {
      const uint32_t v00 = j * w + i, v10 = v00 + 1, v01 = v00 + w, v11 = v01 + 1;
      c2v[3 * f] = v00; c2v[3 * f + 1] = v10; c2v[3 * f + 2] = v11; ++f;
      c2v[3 * f] = v00; c2v[3 * f + 1] = v11; c2v[3 * f + 2] = v01; ++f;
    }
  // opposite corners: the two corners facing the same undirected edge
  std::vector<std::pair<uint64_t, uint32_t>> edges(3ull * F);
  for (uint32_t c = 0; c < 3 * F; ++c) {
    uint32_t a = c2v[cnext(c)], b = c2v[cprev(c)];
    if (a > b) std::swap(a, b);
    edges[c] = {((uint64_t)a << 32) | b, c};
    opposite[c] = INV;
  }
  std::sort(edges.begin(), edges.end());
  for (size_t k = 0; k + 1 < edges.size(); ++k)
    if (edges[k].first == edges[k + 1].first) {
      opposite[edges[k].second] = edges[k + 1].second;
      opposite[edges[k + 1].second] = edges[k].second;
      ++k;
    }
  auto on_boundary = [&](uint32_t v) { const uint32_t i = v % w, j = v / w; return i == 0 || j == 0 || i == w - 1 || j == h - 1; };
  auto right = [&](uint32_t c) { return opposite[cnext(c)]; };
  auto left = [&](uint32_t c) { return opposite[cprev(c)]; };
  std::vector<uint8_t> fvis(F, 0), vvis(V, 0);
  std::vector<uint32_t> stk;
  uint32_t n = 0;
  for (uint32_t v = 0; v < V; ++v) v2d[v] = -1;
  auto visit = [&](uint32_t v, uint32_t c) { vvis[v] = 1; v2d[v] = (int32_t)n; d2c[n++] = c; };
  for (uint32_t f0 = 0; f0 < F; ++f0) {
    if (fvis[f0]) continue;
    uint32_t corner = 3 * f0;
    stk.assign(1, corner);
    const uint32_t nv = c2v[cnext(corner)], pv = c2v[cprev(corner)];
    if (!vvis[nv]) visit(nv, cnext(corner));
    if (!vvis[pv]) visit(pv, cprev(corner));
    while (!stk.empty()) {
      corner = stk.back();
      if (corner == INV || fvis[corner / 3]) { stk.pop_back(); continue; }
      for (;;) {
        fvis[corner / 3] = 1;
        const uint32_t v = c2v[corner];
        if (!vvis[v]) {
          const bool b = on_boundary(v);
          visit(v, corner);
          if (!b) { corner = right(corner); continue; }
        }
        const uint32_t rc = right(corner), lc = left(corner);
        const bool rvis = rc == INV || fvis[rc / 3], lvis = lc == INV || fvis[lc / 3];
        if (rvis) {
          if (lvis) { stk.pop_back(); break; }
          corner = lc;
        } else {
          if (lvis) corner = rc;
          else { stk.back() = lc; stk.push_back(rc); break; }
        }
      }
    }
  }
  return (int64_t)n;
}

// One mesh over that topology: jittered grid in x / y, a smooth height field in z, `pos_bits`-bit quantization,
// parallelogram prediction (MeshPredictionSchemeParallelogramEncoder semantics = the decoder's
// MeshPredictionSchemeParallelogramDecoder.cs:29-89 read backwards) + wrap transform.  `out` receives the whole .drc
// buffer; *attr_off the ATTRIBUTES section offset, *sum the checksum of the expected output floats (entry order),
// pos_q (nullable, [V*3]) the quantized positions in entry order.  Returns bytes, or -(needed).
int64_t synth_grid_mesh(uint32_t w, uint32_t h, uint64_t seed, int32_t pos_bits, int32_t scheme, const uint32_t *opposite,
                        const uint32_t *c2v, const uint32_t *d2c, const int32_t *v2d, uint8_t *out, uint64_t cap,
                        uint64_t *attr_off, uint64_t *sum, int32_t *pos_q, int32_t *scheme_used) {
  const uint32_t V = w * h, INV = 0xFFFFFFFFu;
  const int32_t maxq = (1 << pos_bits) - 1;
  Rng rng(seed);
  const double fx = 0.004 + 0.004 * (double)(rng.u32() % 1000u) / 1000.0, fy = 0.003 + 0.005 * (double)(rng.u32() % 1000u) / 1000.0;
  const double ph = (double)(rng.u32() % 6283u) / 1000.0;
  std::vector<int32_t> q((size_t)V * 3);  // entry order
  for (uint32_t p = 0; p < V; ++p) {
    const uint32_t v = c2v[d2c[p]], i = v % w, j = v / w;
    const double gx = (double)i * (double)(maxq - 8) / (double)(w - 1) + 4.0, gy = (double)j * (double)(maxq - 8) / (double)(h - 1) + 4.0;
    const uint32_t r = rng.u32();
    int32_t x = (int32_t)std::lround(gx) + (int32_t)(r % 5u) - 2, y = (int32_t)std::lround(gy) + (int32_t)((r >> 8) % 5u) - 2;
    int32_t z = (int32_t)std::lround(0.5 * maxq + 0.45 * maxq * std::sin(fx * i + ph) * std::cos(fy * j)) + (int32_t)((r >> 16) % 3u) - 1;
    q[3ull * p] = x < 0 ? 0 : (x > maxq ? maxq : x);
    q[3ull * p + 1] = y < 0 ? 0 : (y > maxq ? maxq : y);
    q[3ull * p + 2] = z < 0 ? 0 : (z > maxq ? maxq : z);
  }
  int32_t mn = q[0], mx = q[0];
  for (size_t k = 1; k < q.size(); ++k) { mn = std::min(mn, q[k]); mx = std::max(mx, q[k]); }
  const int32_t max_diff = 1 + mx - mn;
  int32_t max_corr = max_diff / 2, min_corr = -max_corr;
  if ((max_diff & 1) == 0) max_corr -= 1;
  std::vector<uint32_t> sym((size_t)V * 3);
  for (uint32_t p = 0; p < V; ++p) {
    int32_t pred[3] = {0, 0, 0};
    if (p > 0) {
      bool para = false;
      const uint32_t oc = opposite[d2c[p]];
      if (oc != INV) {
        const int32_t a = v2d[c2v[oc]], b = v2d[c2v[cnext(oc)]], c = v2d[c2v[cprev(oc)]];
        if (a >= 0 && b >= 0 && c >= 0 && a < (int32_t)p && b < (int32_t)p && c < (int32_t)p) {
          para = true;
          for (int k = 0; k < 3; ++k) pred[k] = q[3ull * b + k] + q[3ull * c + k] - q[3ull * a + k];
        }
      }
      if (!para) for (int k = 0; k < 3; ++k) pred[k] = q[3ull * (p - 1) + k];
    }
    for (int k = 0; k < 3; ++k) {
      const int32_t pr = pred[k] > mx ? mx : (pred[k] < mn ? mn : pred[k]);
      int32_t corr = q[3ull * p + k] - pr;
      if (corr < min_corr) corr += max_diff; else if (corr > max_corr) corr -= max_diff;
      sym[3ull * p + k] = zigzag_enc(corr);
    }
  }
  Out o;
  o.bytes("DRACO", 5);
  o.u8(2); o.u8(2);
  o.u8(1);   // TRIANGULAR_MESH
  o.u8(1);   // MESH_EDGEBREAKER_ENCODING
  o.u16(0);
  o.u8(0);   // standard traversal; the connectivity payload proper is decoded by the host and not part of this path
  for (int k = 0; k < 15; ++k) o.u8(0xEB);
  const uint64_t aoff = o.b.size();
  o.u8(1);                        // one attributes decoder
  o.u8(0xFF); o.u8(0); o.u8(0);   // att_data_id -1 (position), MESH_VERTEX_ATTRIBUTE, depth-first traversal
  o.varint(1);
  o.u8(0); o.u8(9); o.u8(3); o.u8(0); o.varint(0);
  o.u8(2);                        // SEQUENTIAL_ATTRIBUTE_ENCODER_QUANTIZATION
  o.u8(1);                        // MESH_PREDICTION_PARALLELOGRAM
  o.u8(1);                        // PREDICTION_TRANSFORM_WRAP
  o.u8(1);
  const int sch = encode_symbols(o, sym, 3, scheme);
  o.i32(mn); o.i32(mx);
  const float pos_min[3] = {-1.0f, -1.0f, -1.0f}, range = 2.0f;
  o.f32(pos_min[0]); o.f32(pos_min[1]); o.f32(pos_min[2]); o.f32(range); o.u8((uint8_t)pos_bits);
  if (scheme_used) *scheme_used = sch;
  if (attr_off) *attr_off = aoff;
  if (sum) {
    const float delta = range / (float)maxq;
    std::vector<float> fl((size_t)V * 3);
    for (size_t k = 0; k < fl.size(); ++k) { volatile float t = (float)q[k] * delta; fl[k] = t + pos_min[k % 3]; }
    *sum = word_checksum(fl.data(), fl.size() * 4);
  }
  if (pos_q) memcpy(pos_q, q.data(), q.size() * 4);
  if (o.b.size() > cap) return -(int64_t)o.b.size();
  memcpy(out, o.b.data(), o.b.size());
  return (int64_t)o.b.size();
}

uint64_t synth_word_checksum(const void *p, uint64_t nbytes) { return word_checksum(p, (size_t)nbytes); }
}
