// draco_sharp_b200/csrc/dcb_device.cuh -- device helpers shared by the sm_100a kernels.
//
// Reference semantics restated here (src/Draco/IO/...):
//   zig-zag            BitUtilities.cs:72-81
//   wrap transform     Attributes/PredictionSchemes/PredictionSchemeWrapDecodingTransform.cs:46-67,
//                      PredictionSchemeWrapTransform.cs:67-100
//   octahedron xform   PredictionSchemeNormalOctahedron{,Canonicalized}DecodingTransform.cs,
//                      PredictionSchemeNormalOctahedronCanonicalizedTransform.cs:43-89, OctahedronToolBox.cs:144-212
//   dequantise         Attributes/AttributeQuantizationTransform.cs:179-199, Core/Dequantizer.cs:14-23
//   oct -> unit vector Attributes/AttributeOctahedronTransform.cs:82-102, OctahedronToolBox.cs:139-142,220-239
//   narrowing store    Attributes/SequentialIntegerAttributeDecoder.cs:142-160
// Float results are bit-exact with the C# evaluation order: every binary32 operation is rounded
// separately (explicit *_rn intrinsics, never contracted to FMA).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dcb_internal.h"

namespace dcb {

__device__ __forceinline__ int32_t zigzag_dec(uint32_t v) {
  // (v & 1) ? -(v >> 1) - 1 : v >> 1
  return (int32_t)((v >> 1) ^ (0u - (v & 1u)));
}

__device__ __forceinline__ int32_t wrap_original(int32_t pred, int32_t corr, int32_t mn, int32_t mx, int32_t max_diff) {
  pred = pred > mx ? mx : (pred < mn ? mn : pred);
  int32_t o = (int32_t)((uint32_t)pred + (uint32_t)corr);
  if (o > mx)
    o = (int32_t)((uint32_t)o - (uint32_t)max_diff);
  else if (o < mn)
    o = (int32_t)((uint32_t)o + (uint32_t)max_diff);
  return o;
}

struct OctBox {
  int32_t max_q, max_value, center;
  __device__ __forceinline__ void set(int bits) {
    max_q = (int32_t)((1u << bits) - 1u);
    max_value = max_q - 1;
    center = max_value / 2;
  }
};
__device__ __forceinline__ int32_t neg32(int32_t v) { return (int32_t)(0u - (uint32_t)v); }
__device__ __forceinline__ int32_t abs32(int32_t v) { return v < 0 ? neg32(v) : v; }
__device__ __forceinline__ bool oct_in_diamond(const OctBox &t, int32_t s, int32_t u) {
  return (uint32_t)abs32(s) + (uint32_t)abs32(u) <= (uint32_t)t.center;
}
__device__ __forceinline__ void oct_invert_diamond(const OctBox &t, int32_t &s, int32_t &u) {
  int32_t ss, su;
  if (s >= 0 && u >= 0) { ss = 1; su = 1; }
  else if (s <= 0 && u <= 0) { ss = -1; su = -1; }
  else { ss = s > 0 ? 1 : -1; su = u > 0 ? 1 : -1; }
  const int32_t cs = ss * t.center, cu = su * t.center;
  int32_t us = (int32_t)((uint32_t)s + (uint32_t)s - (uint32_t)cs);
  int32_t uu = (int32_t)((uint32_t)u + (uint32_t)u - (uint32_t)cu);
  const int32_t tmp = us;
  if (ss * su >= 0) { us = neg32(uu); uu = neg32(tmp); }
  else { us = uu; uu = tmp; }
  us = (int32_t)((uint32_t)us + (uint32_t)cs);
  uu = (int32_t)((uint32_t)uu + (uint32_t)cu);
  s = us / 2;  // truncating, as C#
  u = uu / 2;
}
__device__ __forceinline__ int32_t oct_mod_max(const OctBox &t, int32_t x) {
  if (x > t.center) return (int32_t)((uint32_t)x - (uint32_t)t.max_q);
  return x < -t.center ? (int32_t)((uint32_t)x + (uint32_t)t.max_q) : x;
}
__device__ __forceinline__ void oct_rotate(int32_t &a, int32_t &b, int rot) {
  const int32_t x = a, y = b;
  if (rot == 1) { a = y; b = neg32(x); }
  else if (rot == 2) { a = neg32(x); b = neg32(y); }
  else if (rot == 3) { a = neg32(y); b = x; }
}
// pred (p0io,p1io) + correction (c0,c1) -> original, in place
__device__ __forceinline__ void oct_original(const OctBox &t, bool canonical, int32_t &p0io, int32_t &p1io,
                                             int32_t c0, int32_t c1) {
  int32_t p0 = (int32_t)((uint32_t)p0io - (uint32_t)t.center);
  int32_t p1 = (int32_t)((uint32_t)p1io - (uint32_t)t.center);
  const bool in_diamond = oct_in_diamond(t, p0, p1);
  if (!in_diamond) oct_invert_diamond(t, p0, p1);
  bool bottom_left = true;
  int rot = 0;
  if (canonical) {
    bottom_left = (p0 == 0 && p1 == 0) ? true : (p0 < 0 && p1 <= 0);
    if (p0 == 0) rot = p1 == 0 ? 0 : (p1 > 0 ? 3 : 1);
    else if (p0 > 0) rot = p1 >= 0 ? 2 : 1;
    else rot = p1 <= 0 ? 0 : 3;
    if (!bottom_left) oct_rotate(p0, p1, rot);
  }
  int32_t o0 = oct_mod_max(t, (int32_t)((uint32_t)p0 + (uint32_t)c0));
  int32_t o1 = oct_mod_max(t, (int32_t)((uint32_t)p1 + (uint32_t)c1));
  if (canonical && !bottom_left) oct_rotate(o0, o1, (4 - rot) & 3);
  if (!in_diamond) oct_invert_diamond(t, o0, o1);
  p0io = (int32_t)((uint32_t)o0 + (uint32_t)t.center);
  p1io = (int32_t)((uint32_t)o1 + (uint32_t)t.center);
}

// oct (s,t) -> unit vector; arithmetic types as the C# (binary32 coords and norm, binary64 1/sqrt and products)
__device__ __forceinline__ void oct_to_unit(int32_t s, int32_t u, float scale, float &ox, float &oy, float &oz) {
  float y = __fsub_rn(__fmul_rn(__int2float_rn(s), scale), 1.0f);
  float z = __fsub_rn(__fmul_rn(__int2float_rn(u), scale), 1.0f);
  const float x = __fsub_rn(__fsub_rn(1.0f, fabsf(y)), fabsf(z));
  const float x_offset = (-x < 0.0f) ? 0.0f : -x;
  y = __fadd_rn(y, (y < 0.0f) ? x_offset : -x_offset);
  z = __fadd_rn(z, (z < 0.0f) ? x_offset : -x_offset);
  const float ns = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
  if ((double)ns < 1E-6) {
    ox = 0.0f; oy = 0.0f; oz = 0.0f;
  } else {
    const double d = __ddiv_rn(1.0, __dsqrt_rn((double)ns));
    ox = __double2float_rn(__dmul_rn((double)x, d));
    oy = __double2float_rn(__dmul_rn((double)y, d));
    oz = __double2float_rn(__dmul_rn((double)z, d));
  }
}

// Per-stream constants of the reconstruction + store stages, loaded once per lane.
struct PostParams {
  int32_t recon, store, mn, mx, max_diff, dsize;
  bool zig;
  float qmin[4];
  float delta, oct_scale;
  OctBox box;
  __device__ void load(const StreamDesc &d) {
    recon = d.recon;
    store = d.store;
    zig = d.zigzag != 0;
    mn = d.xf_a;
    mx = d.xf_b;
    max_diff = (int32_t)(1u + (uint32_t)mx - (uint32_t)mn);
    box.set(2);
    delta = 0.0f;
    oct_scale = 0.0f;
    dsize = 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) qmin[c] = 0.0f;
    if (recon == RECON_DELTA_OCT || recon == RECON_DELTA_OCT_CANON)
      box.set(32 - __clz(d.xf_a));  // bits = msb(max_q) + 1 (PredictionSchemeNormalOctahedronTransform.cs:44-53)
    if (store == STORE_DEQUANT) {
      const int32_t maxq = (int32_t)((1u << d.q_bits) - 1u);
      delta = __fdiv_rn(d.q_range, __int2float_rn(maxq));  // Dequantizer.cs:17
#pragma unroll
      for (int c = 0; c < 4; ++c) qmin[c] = d.q_min[c];
    } else if (store == STORE_OCT_UNIT) {
      const int32_t max_value = (int32_t)((1u << d.q_bits) - 2u);
      oct_scale = __fdiv_rn(2.0f, __int2float_rn(max_value));  // OctahedronToolBox.cs:19
    } else {
      dsize = dcb_dtype_len(d.data_type);
    }
  }
  // words of output per entry group of 4 entries
  __device__ __forceinline__ float dequant(int32_t q, int c) const {
    return __fadd_rn(__fmul_rn(__int2float_rn(q), delta), qmin[c]);
  }
};

// Store ONE entry (NCP portable ints in v) at entry index e of the attribute output.
template <int NCP>
__device__ __forceinline__ void store_entry(const PostParams &pp, uint8_t *optr, uint64_t e, const int32_t *v) {
  if (pp.store == STORE_DEQUANT) {
    float *o = reinterpret_cast<float *>(optr) + e * NCP;
#pragma unroll
    for (int c = 0; c < NCP; ++c) o[c] = pp.dequant(v[c], c);
  } else if (pp.store == STORE_OCT_UNIT) {
    if (NCP == 2) {
      float ox, oy, oz;
      oct_to_unit(v[0], v[NCP - 1], pp.oct_scale, ox, oy, oz);
      float *o = reinterpret_cast<float *>(optr) + e * 3;
      o[0] = ox; o[1] = oy; o[2] = oz;
    }
  } else {
    if (pp.dsize == 1) {
      uint8_t *o = optr + e * NCP;
#pragma unroll
      for (int c = 0; c < NCP; ++c) o[c] = (uint8_t)v[c];
    } else if (pp.dsize == 2) {
      uint16_t *o = reinterpret_cast<uint16_t *>(optr) + e * NCP;
#pragma unroll
      for (int c = 0; c < NCP; ++c) o[c] = (uint16_t)v[c];
    } else {
      int32_t *o = reinterpret_cast<int32_t *>(optr) + e * NCP;
#pragma unroll
      for (int c = 0; c < NCP; ++c) o[c] = v[c];
    }
  }
}

// Store a GROUP of 4 consecutive entries starting at entry index e4 (a multiple of 4): the group
// occupies 4 * bytes_per_entry bytes starting on a 16-byte boundary whenever that size is a multiple
// of 16 (attribute outputs start on 128-byte boundaries), so it is written as 128-bit stores.
template <int NCP>
__device__ __forceinline__ void store_group4(const PostParams &pp, uint8_t *optr, uint64_t e4, const int32_t (*v)[NCP]) {
  if (pp.store == STORE_DEQUANT) {
    float f[4 * NCP];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < NCP; ++c) f[j * NCP + c] = pp.dequant(v[j][c], c);
    float4 *o = reinterpret_cast<float4 *>(reinterpret_cast<float *>(optr) + e4 * NCP);
#pragma unroll
    for (int k = 0; k < NCP; ++k) o[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
  } else if (pp.store == STORE_OCT_UNIT) {
    if (NCP == 2) {
      float f[12];
#pragma unroll
      for (int j = 0; j < 4; ++j) oct_to_unit(v[j][0], v[j][NCP - 1], pp.oct_scale, f[3 * j], f[3 * j + 1], f[3 * j + 2]);
      float4 *o = reinterpret_cast<float4 *>(reinterpret_cast<float *>(optr) + e4 * 3);
#pragma unroll
      for (int k = 0; k < 3; ++k) o[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
    }
  } else if (pp.dsize == 4) {
    int4 *o = reinterpret_cast<int4 *>(reinterpret_cast<int32_t *>(optr) + e4 * NCP);
    const int32_t *f = &v[0][0];
#pragma unroll
    for (int k = 0; k < NCP; ++k) o[k] = make_int4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
  } else if (pp.dsize == 1) {
    // 4 entries x NCP bytes = NCP 32-bit words
    uint32_t *o = reinterpret_cast<uint32_t *>(optr + e4 * NCP);
    const int32_t *f = &v[0][0];
#pragma unroll
    for (int k = 0; k < NCP; ++k)
      o[k] = ((uint32_t)f[4 * k] & 0xFFu) | (((uint32_t)f[4 * k + 1] & 0xFFu) << 8) |
             (((uint32_t)f[4 * k + 2] & 0xFFu) << 16) | (((uint32_t)f[4 * k + 3] & 0xFFu) << 24);
  } else {
    // 4 entries x NCP x 2 bytes = 2 * NCP 32-bit words
    uint32_t *o = reinterpret_cast<uint32_t *>(optr + e4 * NCP * 2);
    const int32_t *f = &v[0][0];
#pragma unroll
    for (int k = 0; k < 2 * NCP; ++k) o[k] = ((uint32_t)f[2 * k] & 0xFFFFu) | (((uint32_t)f[2 * k + 1] & 0xFFFFu) << 16);
  }
}

// ---------------------------------------------------------------------------------------------
// compressed-byte window: bytes are consumed from the END of the payload towards its start
// (RAnsDecoder.cs:60 `Buffer[--BufferOffset]`).  `win` holds the next bytes MSB-first; it is
// refilled from a 16-byte chunk held in registers while the following chunk is already in flight.
// ---------------------------------------------------------------------------------------------
struct ByteWin {
  uint64_t win;
  int nwin;
  uint32_t c0, c1, c2, c3;  // current chunk; c3 = highest addresses = consumed first
  int ncur;
  uint4 nxt;
  const uint8_t *arena;
  int64_t next_addr;

  __device__ __forceinline__ uint4 load16(int64_t a) const {
    a = a < 0 ? 0 : a;
    return __ldg(reinterpret_cast<const uint4 *>(arena + a));
  }
  __device__ __forceinline__ void refill() {
    if (nwin <= 4) {
      const uint32_t w = c3;
      c3 = c2; c2 = c1; c1 = c0;
      --ncur;
      win |= (uint64_t)w << (32 - 8 * nwin);
      nwin += 4;
      if (ncur == 0) {
        c0 = nxt.x; c1 = nxt.y; c2 = nxt.z; c3 = nxt.w;
        ncur = 4;
        nxt = load16(next_addr);
        next_addr -= 16;
      }
    }
  }
  // end = arena offset one past the last unread byte
  __device__ __forceinline__ void init(const uint8_t *arena_, uint64_t end) {
    arena = arena_;
    const int64_t a0 = end == 0 ? 0 : (int64_t)((end - 1) & ~15ull);
    const uint4 c = load16(a0);
    c0 = c.x; c1 = c.y; c2 = c.z; c3 = c.w;
    ncur = 4;
    nxt = load16(a0 - 16);
    next_addr = a0 - 32;
    win = 0;
    nwin = 0;
    int drop = (int)(a0 + 16 - (int64_t)end);
    while (drop >= 4) {
      c3 = c2; c2 = c1; c1 = c0;
      --ncur;
      drop -= 4;
    }
    refill();
    if (drop) { win <<= 8 * drop; nwin -= drop; }
    refill();
  }
};

// ---------------------------------------------------------------------------------------------
// per-lane probability table:
//   lut[0..nb)   index of the symbol owning slot (b << lut_shift)
//   cum[0..ne]   cumulative probability of table entry i (cum[ne] = 2^prec)
//   sym[0..ne)   symbol id of entry i (compact tables only; dense tables index by symbol id)
// T = uint16_t when 2^prec and the ids fit, else uint32_t.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct LaneTable {
  T *lut, *cum, *sym;
};

template <typename T>
__device__ __forceinline__ void carve_table(uint8_t *base, uint32_t slot_bytes, uint32_t prec_bits, uint32_t lut_shift,
                                            bool compact, LaneTable<T> &t, uint32_t &cap_entries) {
  const uint32_t nb = (1u << prec_bits) >> lut_shift;
  const uint32_t words = slot_bytes / sizeof(T);
  cap_entries = words > nb + 1u ? (compact ? (words - nb - 1u) / 2u : (words - nb - 1u)) : 0u;
  t.lut = reinterpret_cast<T *>(base);
  t.cum = t.lut + nb;
  t.sym = t.cum + cap_entries + 1u;
}

// Parse RANS_TABLE at d.table_off into the lane's table (RAnsSymbolDecoder.cs:12-51, RAnsDecoder.cs:69-88).
template <typename T>
__device__ int build_table(const uint8_t *arena, const StreamDesc &d, LaneTable<T> t, uint32_t lut_shift, bool compact,
                           uint32_t cap_entries) {
  uint64_t pos = d.table_off;
  const uint64_t end = d.buf_end;
  for (int i = 0; i < 10; ++i) {  // skip the num_symbols varint (value parsed by the indexer)
    if (pos >= end) return DCB_ERR_EOF;
    if (!(arena[pos++] & 0x80)) break;
  }
  const uint32_t ns = d.num_symbols;
  const uint32_t prec = 1u << d.prec_bits;
  if (!compact && ns > cap_entries) return DCB_ERR_TABLE;
  uint64_t c = 0;
  uint32_t ne = 0;
  bool overflow = false;
  for (uint32_t i = 0; i < ns; ++i) {
    if (pos >= end) return DCB_ERR_EOF;
    const uint32_t pd = arena[pos++];
    const uint32_t token = pd & 3u;
    if (token == 3u) {
      const uint32_t off = pd >> 2;
      if (i + off >= ns) return DCB_ERR_TABLE;
      if (!compact)
        for (uint32_t j = 0; j <= off; ++j) t.cum[i + j] = (T)(c > prec ? prec : c);
      i += off;
    } else {
      uint32_t prob = pd >> 2;
      for (uint32_t b = 0; b < token; ++b) {
        if (pos >= end) return DCB_ERR_EOF;
        prob |= (uint32_t)arena[pos++] << (8 * (b + 1) - 2);
      }
      if (compact) {
        if (prob) {
          if (c + prob > prec || ne >= cap_entries) overflow = true;
          if (!overflow) {
            t.cum[ne] = (T)c;
            t.sym[ne] = (T)i;
            ++ne;
          }
        }
      } else {
        if (c + prob > prec) overflow = true;
        t.cum[i] = (T)(c > prec ? prec : c);
      }
      c += prob;
    }
  }
  if (overflow || c != prec) return DCB_ERR_TABLE;
  if (!compact) ne = ns;
  t.cum[ne] = (T)prec;
  const uint32_t nb = prec >> lut_shift;
  uint32_t i = 0;
  for (uint32_t b = 0; b < nb; ++b) {
    const uint32_t slot = b << lut_shift;
    while ((uint32_t)t.cum[i + 1] <= slot) ++i;
    t.lut[b] = (T)i;
  }
  return DCB_OK;
}

// rANS state of one lane
struct RansState {
  uint32_t x;     // state
  uint32_t off;   // unread payload bytes (BufferOffset)
  uint32_t L, L8, L16, mask, prec_bits;
};

// RAnsDecoder.ReadInit (RAnsDecoder.cs:20-54)
__device__ __forceinline__ int rans_init(const uint8_t *arena, const StreamDesc &d, RansState &s) {
  const uint64_t n = d.payload_len;
  if (n < 1) return DCB_ERR_RANS_INIT;
  const uint8_t *p = arena + d.payload_off;
  const uint32_t tag = (uint32_t)p[n - 1] >> 6;
  if (n < tag + 1) return DCB_ERR_RANS_INIT;
  uint32_t v = 0;
  for (uint32_t i = 0; i <= tag; ++i) v |= (uint32_t)p[n - 1 - tag + i] << (8 * i);
  v &= (tag == 0) ? 0x3Fu : (tag == 1) ? 0x3FFFu : (tag == 2) ? 0x3FFFFFu : 0x3FFFFFFFu;
  s.prec_bits = d.prec_bits;
  s.L = 4u << d.prec_bits;
  s.L8 = s.L >> 8;
  s.L16 = s.L >> 16;
  s.mask = (1u << d.prec_bits) - 1u;
  s.x = v + s.L;
  s.off = (uint32_t)(n - (tag + 1));
  if (s.x >= s.L * 256u) return DCB_ERR_RANS_INIT;
  return DCB_OK;
}

// One RAnsDecoder.Read() (RAnsDecoder.cs:56-67,90-99): renormalise, then decode.  Returns the table
// entry index.
template <typename T>
__device__ __forceinline__ uint32_t rans_step(RansState &s, ByteWin &w, const LaneTable<T> &t, uint32_t lut_shift) {
  uint32_t x = s.x;
  // while (state < L && off > 0) state = state * 256 + buf[--off];   L is a multiple of 256, so the
  // number of iterations depends on x alone: one per threshold L, L/256, L/65536 that x is below.
  uint32_t n = (x < s.L ? 1u : 0u) + (x < s.L8 ? 1u : 0u) + (x < s.L16 ? 1u : 0u);
  n = min(n, s.off);
  const uint32_t sh = 8u * n;
  x = __funnelshift_l((uint32_t)(w.win >> 32), x, sh);
  w.win <<= sh;
  w.nwin -= (int)n;
  s.off -= n;
  w.refill();
  const uint32_t r = x & s.mask;
  const uint32_t q = x >> s.prec_bits;
  uint32_t i = t.lut[r >> lut_shift];
  uint32_t ca = t.cum[i];
  uint32_t cb = t.cum[i + 1];
  while (r >= cb) {
    ++i;
    ca = cb;
    cb = t.cum[i + 1];
  }
  s.x = q * (cb - ca) + (r - ca);
  return i;
}

}  // namespace dcb
