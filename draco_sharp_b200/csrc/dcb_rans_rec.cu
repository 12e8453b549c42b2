// draco_sharp_b200/csrc/dcb_rans_rec.cu -- lane-per-stream rANS kernels with BUCKET-RECORD tables: ONE dependent
// shared-memory access per symbol at full residency.
//
// The chain of a rANS stream is  x -> renormalise -> slot r = x mod 2^prec -> table entry (freq, cum) -> x' = q * freq +
// r - cum  (RAnsDecoder.cs:56-67, 90-99).  Round 1 resolved the entry with two dependent shared-memory accesses (a
// rank byte + block base, then three cum values: ~2 x 35 cycles and a dozen ALU levels between them: 157 cycles per
// symbol even for a warp that carries nothing but the chain, ncu profiles/r2_pc_c2_*).  A direct slot table needs one
// access but 2^prec x 4 bytes per stream -- 32 KB where a full machine has 3.3 KB per stream.
//
// Bucket records get the one access into ~2.7 KB for a 478-entry table at 13 bits.  Draco symbols are zig-zag coded
// corrections, so entry widths fall off with the symbol id; the slot axis is cut twice (RecShape, dcb_internal.h):
//   [0, ta)        one 8-byte RECORD per 2^ka slots   (every such bucket meets at most two table entries)
//   [ta, tb)       one record per 16 slots            (likewise)
//   [tb, tc)       one record per 8 slots             (likewise; usually a few buckets between the two neighbours)
//   [tc, 2^prec)   one BYTE per slot: (freq - 1) << 4 | (r - cum)   (every entry reaching in is <= 16 slots wide)
// A record { c1 = cum[i + 1], f0 = freq[i], f1 = freq[i + 1], rank i } (i = the entry owning the bucket's first slot)
// decides the step without a search:  second = r >= c1;  x' = second ? q * f1 + (r - c1) : (q + 1) * f0 + (r - c1).
// Each lane issues exactly one of the two loads (predicated), then ~4 ALU levels close the chain.  Tables without that
// shape (a wide entry behind narrow ones: unsigned octahedral corrections, some tag alphabets) keep the two-level
// kernels of dcb_kernels.cu / dcb_rans_pc.cu; the host planner knows per stream (StreamDesc::rec_need, computed by
// the container walk with the same RecShape code the device build runs) and per group.
//
// As in dcb_rans_pc.cu the work is split over warp PAIRS: the chain warp runs the chain and pushes one 16-bit code per
// symbol (entry rank, or 0x8000 | slot for the byte region) into a lane-interleaved shared-memory queue; the consumer
// warp turns codes into symbols (byte region: rank = popcount over a bitmap of entry starts), then value map / zig-zag
// (BitUtilities.cs:72-81), PredictionSchemeDeltaDecoder + wrap (PredictionSchemeWrapDecodingTransform.cs:46-67),
// dequantisation (Dequantizer.cs:14-23) or the narrowing store (SequentialIntegerAttributeDecoder.cs:142-160).
// Stream ends are exact as before: the queue carries the warp-uniform main loop, the chain warp finishes every stream
// with the careful per-entry loop (`off > 0` checked per byte).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <algorithm>

#include "dcb_device.cuh"
#include "dcb_internal.h"
#include "dcb_kernels.h"

using namespace dcb;

namespace {

constexpr uint32_t kStages = DCB_REC_STAGES;      // queue depth in groups (power of two)
constexpr uint32_t kRowBytes = DCB_PC_ROW_BYTES;  // one queue row: 32 lanes x 2 bytes
constexpr uint32_t kUnset = 0xFFFFFFFFu;

// what the two warps of a pair tell each other outside the queue, per lane
struct RecHand {
  uint32_t dprefix;   // chain -> consumer: entries below it are their own symbol ids
  int32_t active;     // chain -> consumer: the lane decodes a stream
  uint32_t t_a, t_b, t_c;  // chain -> consumer: region cuts of the lane's table (layout = dcb_rec_layout)
  uint32_t rank_c0;   // chain -> consumer: first entry of the byte region
  int32_t status;     // tags: consumer -> chain
  uint32_t e;         // tags: consumer -> chain: first tag the careful tail has to decode
  int32_t prev[4];    // consumer -> chain: running values of the delta decoder after the last queued group
  uint64_t bits;      // tags: consumer -> chain: bits consumed in the bit area so far
};
static_assert(sizeof(RecHand) == DCB_REC_HAND_BYTES, "RecHand size is part of the shared-memory plan");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory");
}
__device__ __forceinline__ uint32_t mbar_test(uint32_t a, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(a), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "RC_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra RC_DONE;\n"
      "bra RC_WAIT;\n"
      "RC_DONE:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;\n" ::"r"(a), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;\n" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16m(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u32m(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
  return v;
}

// control block of a pair: full[kStages] | empty[kStages] | setup | handoff | flag[kStages]
constexpr uint32_t kNumBarriers = 2u * kStages + 2u;
struct PairCtl {
  uint32_t base;
  __device__ __forceinline__ uint32_t full(uint32_t s) const { return base + 8u * s; }
  __device__ __forceinline__ uint32_t empty(uint32_t s) const { return base + 8u * (kStages + s); }
  __device__ __forceinline__ uint32_t setup() const { return base + 16u * kStages; }
  __device__ __forceinline__ uint32_t handoff() const { return base + 16u * kStages + 8u; }
  __device__ __forceinline__ uint32_t flag(uint32_t s) const { return base + 8u * kNumBarriers + 4u * s; }
};
static_assert(8u * kNumBarriers + 4u * kStages <= DCB_REC_CTL_BYTES, "control block");

// RANS_TABLE reader (RAnsSymbolDecoder.cs:12-51): yields the symbols with a non-zero probability, in order
struct TabReader {
  const uint8_t *a;
  uint64_t pos, bend;
  uint32_t i, ns;
  int err;
  __device__ void open(const uint8_t *arena, const StreamDesc &d) {
    a = arena;
    pos = d.table_off;
    bend = d.buf_end;
    i = 0;
    ns = d.num_symbols;
    err = DCB_OK;
    for (int k = 0; k < 10; ++k) {  // skip the num_symbols varint (value parsed by the indexer)
      if (pos >= bend) {
        err = DCB_ERR_EOF;
        return;
      }
      if (!(a[pos++] & 0x80)) break;
    }
  }
  __device__ bool next(uint32_t &sym, uint32_t &prob) {
    while (!err && i < ns) {
      if (pos >= bend) {
        err = DCB_ERR_EOF;
        return false;
      }
      const uint32_t pd = a[pos++];
      const uint32_t token = pd & 3u;
      if (token == 3u) {
        const uint32_t off = pd >> 2;
        if (i + off >= ns) {
          err = DCB_ERR_TABLE;
          return false;
        }
        i += off + 1u;
        continue;
      }
      uint32_t p = pd >> 2;
      for (uint32_t b = 0; b < token; ++b) {
        if (pos >= bend) {
          err = DCB_ERR_EOF;
          return false;
        }
        p |= (uint32_t)a[pos++] << (8 * (b + 1) - 2);
      }
      const uint32_t s = i++;
      if (p) {
        sym = s;
        prob = p;
        return true;
      }
    }
    return false;
  }
};

// geometry of a launch (uniform over the grid)
struct RecGeom {
  uint32_t ka;          // log2(slots per record) of the wide region
  uint32_t area_bytes;  // per lane: records + byte region + bitmap + value map (multiple of 8)
  uint32_t prec_bits;
  uint32_t cap_exc;     // value slots for entries beyond the dense prefix
  uint32_t zig;
  uint32_t slice_bytes, ring_off, q_off, ctl_off, hand_off;  // a pair's slice of shared memory
};

// ---------------------------------------------------------------------------------------------
// chain side
// ---------------------------------------------------------------------------------------------
struct RecChain {
  // state chain
  uint32_t x, p1;
  // constants
  uint32_t L, L8, mask, prec_bits;
  uint32_t base_a, base_bp, base_b2p, base_cp, t_a, t_b, t_c, ka;
  uint32_t ring;
  // byte supply (see RansLane in dcb_device.cuh: 128-byte ring, direct-mapped by arena address, fed by cp.async)
  const uint8_t *gnext;  // global address of the lowest loaded 16-byte chunk
  uint32_t lo32;         // its arena offset, low 32 bits
  uint32_t end32;        // low 32 bits of the arena offset one past the first unread byte at init
  uint32_t p1_init, off_init;
  uint32_t pk_hi, pk_lo, pk_sh;

  __device__ __forceinline__ uint32_t consumed() const { return p1_init - p1; }
  __device__ __forceinline__ uint32_t bytes_left() const { return off_init - consumed(); }

  template <int MAX_CHUNKS>
  __device__ __forceinline__ void top_up() {
    // keep the ring filled down to (read position rounded up to 16) - 128: distances are far below 2^31, so the low
    // 32 bits of the offsets order them
    const uint32_t ptr = p1 + 1u;
    const uint32_t floor_lo = ((ptr + 15u) & ~15u) - DCB_RING_BYTES;
#pragma unroll
    for (int k = 0; k < MAX_CHUNKS; ++k) {
      if ((int32_t)(lo32 - 16u - floor_lo) >= 0) {
        lo32 -= 16u;
        gnext -= 16;
        cp_async16(ring | (lo32 & (DCB_RING_BYTES - 1u)), gnext);
      }
    }
    cp_async_commit();
  }
  __device__ __forceinline__ void prefetch() {
    pk_hi = lds_u32(ring | (p1 & (DCB_RING_BYTES - 4u)));
    pk_lo = lds_u32(ring | ((p1 + DCB_RING_BYTES - 4u) & (DCB_RING_BYTES - 4u)));
    pk_sh = (p1 & 3u) * 8u + 8u;
  }
  __device__ __forceinline__ uint32_t peek() const { return __funnelshift_rc(pk_lo, pk_hi, pk_sh); }

  // RAnsDecoder.ReadInit (RAnsDecoder.cs:20-54)
  __device__ int init_state(const uint8_t *arena, const StreamDesc &d) {
    const uint64_t n = d.payload_len;
    if (n < 1) return DCB_ERR_RANS_INIT;
    const uint8_t *p = arena + d.payload_off;
    const uint32_t tag = (uint32_t)p[n - 1] >> 6;
    if (n < tag + 1) return DCB_ERR_RANS_INIT;
    uint32_t v = 0;
    for (uint32_t i = 0; i <= tag; ++i) v |= (uint32_t)p[n - 1 - tag + i] << (8 * i);
    v &= (tag == 0) ? 0x3Fu : (tag == 1) ? 0x3FFFu : (tag == 2) ? 0x3FFFFFu : 0x3FFFFFFFu;
    prec_bits = d.prec_bits;
    L = 4u << d.prec_bits;
    L8 = L >> 8;
    mask = (1u << d.prec_bits) - 1u;
    x = v + L;
    if (x >= L * 256u) return DCB_ERR_RANS_INIT;
    off_init = (uint32_t)(n - (tag + 1));
    const uint64_t end = d.payload_off + off_init;
    end32 = (uint32_t)end;
    p1_init = end32 - 1u;
    p1 = p1_init;
    const uint64_t lo = (end + 15ull) & ~15ull;
    lo32 = (uint32_t)lo;
    gnext = arena + lo;
    return DCB_OK;
  }
  __device__ __forceinline__ void init_ring(uint32_t ring_addr) {
    ring = ring_addr;
    top_up<DCB_RING_BYTES / 16>();
    cp_async_wait<0>();
    prefetch();
  }

  // One RAnsDecoder.Read() (RAnsDecoder.cs:56-67, 90-99).  Returns the code the consumer resolves: the rank of the
  // table entry, or 0x8000 | slot in the byte region.  CAREFUL: renormalisation bounded by the bytes left.
  template <bool CAREFUL>
  __device__ __forceinline__ uint32_t step() {
    const uint32_t v = peek();
    // precision <= 15: x >= 4 > L / 65536, at most two bytes.  Both shifted candidates are formed next to the
    // compares; the selects are the only level between them and the slot.
    const bool lt = x < L, lt8 = x < L8;
    uint32_t xr;
    if (CAREFUL) {
      uint32_t nb = (lt ? 1u : 0u) + (lt8 ? 1u : 0u);
      nb = min(nb, bytes_left());
      xr = __funnelshift_l(v, x, 8u * nb);
      p1 -= nb;
    } else {
      const uint32_t x8 = __funnelshift_l(v, x, 8u), x16 = __funnelshift_l(v, x, 16u);
      xr = lt ? x8 : x;
      xr = lt8 ? x16 : xr;
      p1 -= lt ? 1u : 0u;
      p1 -= lt8 ? 1u : 0u;
    }
    prefetch();
    const uint32_t r = xr & mask;
    const uint32_t q = xr >> prec_bits;
    const bool in_a = r < t_a, in_b = r < t_b, in_rec = r < t_c;
    uint32_t k = in_b ? 4u : 3u, base = in_b ? base_bp : base_b2p;
    k = in_a ? ka : k;
    base = in_a ? base_a : base;
    const uint32_t a_rec = base + ((r >> k) << 3);
    const uint32_t a_byte = base_cp + r;
    // each lane issues exactly one of the two loads (predicated: no branch, and fewer lanes per access)
    uint32_t lo = 0, hi = 0, u = 0;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.u32 p, %5, 0;\n"
        "@p ld.shared.v2.u32 {%0, %1}, [%3];\n"
        "@!p ld.shared.u8 %2, [%4];\n"
        "}\n"
        : "+r"(lo), "+r"(hi), "+r"(u)
        : "r"(a_rec), "r"(a_byte), "r"((uint32_t)in_rec));
    const uint32_t c1 = lo & 0xFFFFu, f0 = lo >> 16, f1 = hi & 0xFFFFu, rank = hi >> 16;
    const bool second = r >= c1;
    const uint32_t d = r - c1;
    const uint32_t x_rec = second ? q * f1 + d : (q + 1u) * f0 + d;
    const uint32_t x_byte = q * ((u >> 4) + 1u) + (u & 15u);
    x = in_rec ? x_rec : x_byte;
    return in_rec ? rank + (second ? 1u : 0u) : (0x8000u | r);
  }
};

// Parse the stream's RANS_TABLE twice (sizes, then contents) and lay the lane's table area out.  `area` is the lane's
// own area_bytes of shared memory.  Fills the hand-over record.
__device__ int rec_build(const uint8_t *arena, const StreamDesc &d, const RecGeom &g, uint8_t *area, uint32_t area_addr,
                         RecChain &rc, RecHand &h) {
  const uint32_t prec = 1u << d.prec_bits;
  TabReader rd;
  // ---- pass 1: shape, entry counts ----
  rd.open(arena, d);
  uint32_t c = 0, ne = 0, dprefix = kUnset;
  bool overflow = false;
  uint32_t sym, prob;
  RecShape shape;
  shape.begin();
  while (rd.next(sym, prob)) {
    if (c + prob > prec) {
      overflow = true;
      break;
    }
    if (dprefix == kUnset && sym != ne) dprefix = ne;
    shape.entry(c, prob);
    c += prob;
    ++ne;
  }
  if (rd.err) return rd.err;
  if (overflow || c != prec) return DCB_ERR_TABLE;  // RAnsDecoder.cs:80,87
  if (dprefix == kUnset) dprefix = ne;
  uint32_t ta, tb, tc, rank_c0 = kUnset;
  const bool shape_ok = shape.cuts(prec, g.ka, ta, tb, tc);
  const RecLayout lay = dcb_rec_layout(ta, tb, tc, prec, g.ka);
  const uint32_t n_exc = ne - dprefix;
  // the planner sized the area from the same numbers (StreamDesc::rec_need, n_active - dense_prefix)
  if (!shape_ok || n_exc > g.cap_exc || lay.bytes + 2u * n_exc > g.area_bytes) return DCB_ERR_STATE;
  uint2 *rec_a = reinterpret_cast<uint2 *>(area);
  uint2 *rec_b = reinterpret_cast<uint2 *>(area + lay.off_b);
  uint2 *rec_b2 = reinterpret_cast<uint2 *>(area + lay.off_b2);
  uint8_t *byt = area + lay.off_c;
  uint32_t *bm = reinterpret_cast<uint32_t *>(area + lay.off_bm);
  uint16_t *cnt = reinterpret_cast<uint16_t *>(area + lay.off_cnt);
  uint16_t *val = reinterpret_cast<uint16_t *>(area + lay.bytes);
  const uint32_t words = (lay.n_c + 31u) >> 5;
  for (uint32_t w = 0; w < words; ++w) bm[w] = 0u;
  // ---- pass 2: records, bytes, bitmap, value map ----
  rd.open(arena, d);
  uint32_t sym_n = 0, prob_n = 0;
  rd.next(sym, prob);
  uint32_t c0 = 0, c1 = prob;
  bool more = rd.next(sym_n, prob_n);
  uint32_t c2 = more ? c1 + prob_n : prec;
  uint32_t ja = 0, jb = ta >> 4, jb2 = tb >> 3;
  const uint32_t jb_end = jb + lay.n_b, jb2_end = jb2 + lay.n_b2;
  for (uint32_t i = 0; i < ne; ++i) {
    const uint2 rec = make_uint2(c1 | ((c1 - c0) << 16), (c2 - c1) | (i << 16));
    while (ja < lay.n_a && (ja << g.ka) < c1) rec_a[ja++] = rec;
    while (jb < jb_end && (jb << 4) < c1) {
      rec_b[jb - (ta >> 4)] = rec;
      ++jb;
    }
    while (jb2 < jb2_end && (jb2 << 3) < c1) {
      rec_b2[jb2 - (tb >> 3)] = rec;
      ++jb2;
    }
    if (c1 > tc) {  // reaches into the byte region (tc is an entry boundary: the entry starts at or behind it)
      if (rank_c0 == kUnset) rank_c0 = i;
      const uint32_t f = c1 - c0, first = max(c0, tc);
      for (uint32_t s = first; s < c1; ++s) byt[s - tc] = (uint8_t)(((f - 1u) << 4) | (s - c0));
      bm[(first - tc) >> 5] |= 1u << ((first - tc) & 31u);
    }
    if (i >= dprefix) val[i - dprefix] = (uint16_t)(g.zig ? (uint32_t)zigzag_dec(sym) : sym);
    c0 = c1;
    c1 = c2;
    sym = sym_n;
    if (more) more = rd.next(sym_n, prob_n);
    c2 = more ? c1 + prob_n : prec;
  }
  uint32_t run = 0;
  for (uint32_t w = 0; w < words; ++w) {
    cnt[w] = (uint16_t)run;
    run += (uint32_t)__popc(bm[w]);
  }
  rc.ka = g.ka;
  rc.t_a = ta;
  rc.t_b = tb;
  rc.t_c = tc;
  rc.base_a = area_addr;
  rc.base_bp = area_addr + lay.off_b - ((ta >> 4) << 3);
  rc.base_b2p = area_addr + lay.off_b2 - ((tb >> 3) << 3);
  rc.base_cp = area_addr + lay.off_c - tc;
  h.dprefix = dprefix;
  h.t_a = ta;
  h.t_b = tb;
  h.t_c = tc;
  h.rank_c0 = rank_c0 == kUnset ? ne : rank_c0;
  return DCB_OK;
}

// ---------------------------------------------------------------------------------------------
// consumer side: code -> table entry rank -> symbol value
// ---------------------------------------------------------------------------------------------
struct RecMap {
  uint32_t bm_addr, cnt_addr, val_addr;  // shared-memory addresses inside the lane's area
  uint32_t t_c, rank_c0, dprefix;
  __device__ __forceinline__ void set(uint32_t area_addr, const RecHand &h, uint32_t prec_bits, uint32_t ka) {
    const RecLayout lay = dcb_rec_layout(h.t_a, h.t_b, h.t_c, 1u << prec_bits, ka);
    bm_addr = area_addr + lay.off_bm;
    cnt_addr = area_addr + lay.off_cnt;
    val_addr = area_addr + lay.bytes;
    t_c = h.t_c;
    rank_c0 = h.rank_c0;
    dprefix = h.dprefix;
  }
  __device__ __forceinline__ uint32_t rank_of(uint32_t code) const {
    uint32_t rank = code;
    if (code & 0x8000u) {  // byte region: entries started up to and including this slot
      const uint32_t j = (code & 0x7FFFu) - t_c;
      const uint32_t w = j >> 5;
      const uint32_t bits = lds_u32m(bm_addr + 4u * w);
      const uint32_t before = lds_u16m(cnt_addr + 2u * w);
      rank = rank_c0 + before + (uint32_t)__popc(bits & (0xFFFFFFFFu >> (31u - (j & 31u)))) - 1u;
    }
    return rank;
  }
  __device__ __forceinline__ int32_t value_of_rank(uint32_t rank, bool zig) const {
    if (rank >= dprefix) {
      const uint32_t v = lds_u16m(val_addr + 2u * (rank - dprefix));
      return zig ? (int32_t)(int16_t)v : (int32_t)v;
    }
    return zig ? zigzag_dec(rank) : (int32_t)rank;
  }
  __device__ __forceinline__ int32_t value(uint32_t code, bool zig) const { return value_of_rank(rank_of(code), zig); }
  __device__ __forceinline__ uint32_t symbol(uint32_t code, bool zig) const {
    const int32_t v = value(code, zig);
    if (!zig) return (uint32_t)v;
    return v >= 0 ? ((uint32_t)v << 1) : ((((uint32_t)(-(v + 1))) << 1) | 1u);
  }
};

// Which pair a warp belongs to and on which side.  Packed layout: warps 0..pairs-1 are chain warps, the next `pairs`
// their consumers (a pair shares a sub-partition when pairs == 4).  Split layout (bit 31 of `pairs`, at most 3 pairs,
// 4 * pairs warps): chain warps 0..pairs-1 have a sub-partition each to themselves (warp id mod 4), ALL consumers sit
// on sub-partition 3 (warps 3, 7, 11), the warps in between leave at once.  The ALU and FMA pipes of a sub-partition
// take one warp instruction every two cycles each, so a consumer next to its chain warp takes issue slots from it.
__device__ __forceinline__ void warp_role(uint32_t warp, uint32_t &pairs, uint32_t &pair, uint32_t &role) {
  const bool split = (pairs >> 31) != 0u, chain_low = ((pairs >> 30) & 1u) != 0u;
  pairs &= 0x3FFFFFFFu;
  if (!split) {
    // the sub-partition's arbiter prefers the warp with the higher id: the chain warps take the upper half, so that
    // the consumer only gets the issue slots the chain leaves
    pair = warp % pairs;
    role = chain_low ? warp / pairs : 1u - warp / pairs;
  } else if (warp < pairs) {
    pair = warp;
    role = 0;
  } else if ((warp & 3u) == 3u) {
    pair = warp >> 2;
    role = 1;
  } else {
    pair = 0;
    role = 2;
  }
}

// The chain warp's main loop.  NSYM symbols per group and lane; returns the number of groups queued.
template <int NSYM>
__device__ __forceinline__ uint32_t produce(RecChain &rc, bool active, uint32_t g_min, uint32_t q_addr, const PairCtl &ctl,
                                            uint32_t lane) {
  constexpr uint32_t kGroupBytes = (uint32_t)NSYM * 2u;  // at most two bytes per symbol (precision <= 15)
  uint32_t g = 0;
  bool go = g_min != kUnset && g_min > 0u && __all_sync(0xffffffffu, !active || rc.bytes_left() >= kGroupBytes);
  uint32_t slot_free = 1u;  // the first kStages groups find their slots untouched
  while (go) {
    const uint32_t s = g & (kStages - 1u), par = (g / kStages) & 1u;
    if (!slot_free) mbar_wait(ctl.empty(s), par ^ 1u);
    {
      const uint32_t g1 = g + 1u;
      slot_free = mbar_test(ctl.empty(g1 & (kStages - 1u)), ((g1 / kStages) & 1u) ^ 1u);
    }
    const uint32_t qs = q_addr + s * ((uint32_t)NSYM * kRowBytes) + lane * 2u;
    if (active) {
#pragma unroll
      for (int j = 0; j < NSYM; ++j) sts_u16(qs + (uint32_t)j * kRowBytes, rc.template step<false>());
    }
    const bool next_go = (g + 1u < g_min) && __all_sync(0xffffffffu, !active || rc.bytes_left() >= kGroupBytes);
    if (active) {
      rc.template top_up<(kGroupBytes + 15) / 16 + 1>();
      cp_async_wait<1>();
    }
    __syncwarp();
    if (lane == 0) {
      sts_u32(ctl.flag(s), 0u);
      mbar_arrive(ctl.full(s));
    }
    ++g;
    go = next_go;
  }
  {  // end marker: the consumer hands its running values back and the chain warp finishes every stream itself
    const uint32_t s = g & (kStages - 1u), par = (g / kStages) & 1u;
    if (!slot_free) mbar_wait(ctl.empty(s), par ^ 1u);
    __syncwarp();
    if (lane == 0) {
      sts_u32(ctl.flag(s), 1u);
      mbar_arrive(ctl.full(s));
    }
  }
  return g;
}

template <int NCP, int MODE>
__device__ __forceinline__ void redirect_post(PostParams &pp, uint32_t &dump, uint8_t *&optr, uint8_t *aux, const StreamDesc &d) {
  if (MODE == 3 || MODE == 4 || (MODE == 0 && (pp.recon == RECON_PARA_WRAP || pp.store == STORE_OCT_UNIT))) {
    optr = aux + d.aux_off;  // int32 scratch: corrections for the parallelogram kernel / for oct_chain_kernel
    if (NCP == 2 && pp.store == STORE_OCT_UNIT) {
      pp.recon = RECON_NONE;
      dump &= ~(uint32_t)DCB_DUMP_QINTS;
    }
    pp.store = STORE_NARROW;
    pp.dsize = 4;
  }
}

// reconstruction of one entry from its corrections (same rules as decode_entry in dcb_device.cuh)
template <int NCP>
__device__ __forceinline__ void recon_entry(const PostParams &pp, int recon, int32_t *prev, int32_t *v) {
  if (recon == RECON_DELTA_WRAP) {
#pragma unroll
    for (int c = 0; c < NCP; ++c) {
      prev[c] = wrap_original(prev[c], v[c], pp.mn, pp.mx, pp.max_diff);
      v[c] = prev[c];
    }
  } else if (recon == RECON_DELTA_OCT || recon == RECON_DELTA_OCT_CANON) {
    if (NCP == 2) {
      oct_original(pp.box, recon == RECON_DELTA_OCT_CANON, prev[0], prev[NCP - 1], v[0], v[NCP - 1]);
      v[0] = prev[0];
      v[NCP - 1] = prev[NCP - 1];
    }
  }
}

// set up a lane of the chain warp: tables, state, ring; fills the hand-over record.  Returns true when the lane decodes.
__device__ __forceinline__ bool chain_setup(const uint8_t *arena, StreamDesc *dp, bool have, const RecGeom &geom, uint8_t *slice,
                                            uint32_t slice_addr, uint32_t lane, RecChain &rc, RecHand *hand, bool tags) {
  bool active = false;
  if (have) {
    const StreamDesc &d = *dp;
    int status = DCB_OK;
    if (d.n_entries > 0) {
      uint8_t *area = slice + (size_t)lane * geom.area_bytes;
      status = rec_build(arena, d, geom, area, slice_addr + lane * geom.area_bytes, rc, *hand);
      if (status == DCB_OK) status = rc.init_state(arena, d);
      if (status == DCB_OK) {
        rc.init_ring(slice_addr + geom.ring_off + lane * DCB_RING_BYTES);
        active = true;
      }
    }
    if (status != DCB_OK) {
      dp->status = status;
      if (tags) dp->bits_total = 0;
    }
    hand->active = active ? 1 : 0;
  }
  return active;
}

// ---------------------------------------------------------------------------------------------
// Raw scheme (SymbolDecoding.cs:52-67) fused with inverse prediction + transform + store
// ---------------------------------------------------------------------------------------------
template <int NCP, bool DUMP, int MODE>
__global__ void __launch_bounds__(384) rans_raw_rec_kernel(const uint8_t *__restrict__ arena, StreamDesc *streams,
                                                           const uint32_t *__restrict__ order, uint32_t n_streams,
                                                           uint32_t lanes, uint32_t pairs, RecGeom geom,
                                                           uint8_t *__restrict__ out, uint8_t *__restrict__ dbg,
                                                           uint8_t *__restrict__ aux, uint32_t dump) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int kSym = 4 * NCP;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t pair, role;
  warp_role(warp, pairs, pair, role);
  uint8_t *slice = smem + (size_t)pair * geom.slice_bytes;
  const uint32_t slice_addr = smem_u32(slice);
  const uint32_t q_addr = slice_addr + geom.q_off;
  const PairCtl ctl{slice_addr + geom.ctl_off};
  if (role == 0 && lane == 0) {
#pragma unroll
    for (uint32_t i = 0; i < kNumBarriers; ++i) mbar_init(ctl.base + 8u * i, 1u);
  }
  __syncthreads();
  if (role == 2u) return;
  const uint32_t slot = (blockIdx.x * pairs + pair) * lanes + lane;
  const bool have = lane < lanes && slot < n_streams;
  StreamDesc *dp = have ? &streams[order[slot]] : nullptr;
  RecHand *hand = reinterpret_cast<RecHand *>(slice + geom.hand_off) + (have ? lane : 0u);
  const bool zig = MODE == 0 ? geom.zig != 0 : MODE != 3;

  if (role == 0) {
    // ================================ chain warp ================================
    RecChain rc;
    const bool active = chain_setup(arena, dp, have, geom, slice, slice_addr, lane, rc, hand, false);
    const uint32_t n_entries = active ? dp->n_entries : 0u;
    uint32_t g_min = active ? (n_entries >> 2) : kUnset;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) g_min = min(g_min, __shfl_xor_sync(0xffffffffu, g_min, o));
    __syncwarp();
    if (lane == 0) mbar_arrive(ctl.setup());
    const uint32_t groups = produce<kSym>(rc, active, g_min, q_addr, ctl, lane);
    mbar_wait(ctl.handoff(), 0u);
    if (!active) return;
    // ---- per-lane tail: exact `off > 0` handling, as RAnsDecoder.Read does it byte by byte ----
    PostParams pp;
    pp.load(*dp);
    uint8_t *optr = out + dp->out_off;
    int32_t *dptr = DUMP ? reinterpret_cast<int32_t *>(dbg + dp->dbg_off) : nullptr;
    redirect_post<NCP, MODE>(pp, dump, optr, aux, *dp);
    const int store = store_of<MODE>(pp), dsize = dsize_of<MODE>(pp), recon = recon_of<MODE>(pp);
    RecMap vm;
    vm.set(slice_addr + lane * geom.area_bytes, *hand, geom.prec_bits, geom.ka);
    int32_t prev[NCP];
#pragma unroll
    for (int c = 0; c < NCP; ++c) prev[c] = hand->prev[c];
    for (uint32_t e = groups * 4u; e < n_entries; ++e) {
      int32_t v[NCP];
#pragma unroll
      for (int c = 0; c < NCP; ++c) {
        const uint32_t code = rc.template step<true>();
        v[c] = vm.value(code, zig);
        if (DUMP && (dump & DCB_DUMP_SYMBOLS)) dptr[(uint64_t)e * NCP + c] = (int32_t)vm.symbol(code, zig);
      }
      recon_entry<NCP>(pp, recon, prev, v);
      if (DUMP && MODE != 3 && (dump & DCB_DUMP_QINTS)) {
#pragma unroll
        for (int c = 0; c < NCP; ++c) dptr[(uint64_t)e * NCP + c] = v[c];
      }
      store_entry<NCP>(pp, store, dsize, optr, e, v);
      rc.template top_up<(2 * NCP + 15) / 16 + 1>();
      cp_async_wait<0>();
    }
  } else {
    // ================================ consumer warp ================================
    PostParams pp;
    uint8_t *optr = nullptr;
    int32_t *dptr = nullptr;
    if (have) {
      pp.load(*dp);
      optr = out + dp->out_off;
      dptr = DUMP ? reinterpret_cast<int32_t *>(dbg + dp->dbg_off) : nullptr;
      redirect_post<NCP, MODE>(pp, dump, optr, aux, *dp);
    }
    const int store = store_of<MODE>(pp), dsize = dsize_of<MODE>(pp), recon = recon_of<MODE>(pp);
    mbar_wait(ctl.setup(), 0u);
    bool active = false;
    RecMap vm{};
    if (have) {
      active = hand->active != 0;
      if (active) vm.set(slice_addr + lane * geom.area_bytes, *hand, geom.prec_bits, geom.ka);
    }
    int32_t prev[NCP];
#pragma unroll
    for (int c = 0; c < NCP; ++c) prev[c] = 0;
    for (uint32_t g = 0;; ++g) {
      const uint32_t s = g & (kStages - 1u);
      mbar_wait(ctl.full(s), (g / kStages) & 1u);
      if (lds_u32m(ctl.flag(s)) != 0u) break;
      const uint32_t qs = q_addr + s * ((uint32_t)kSym * kRowBytes) + lane * 2u;
      uint32_t code[4][NCP];
      if (active) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < NCP; ++c) code[j][c] = lds_u16m(qs + (uint32_t)(j * NCP + c) * kRowBytes);
      }
      // every queued code of this group is in registers: the slot may be refilled
      __syncwarp();
      if (lane == 0) mbar_arrive(ctl.empty(s));
      if (active) {
        int32_t v[4][NCP];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int c = 0; c < NCP; ++c) {
            v[j][c] = vm.value(code[j][c], zig);
            if (DUMP && (dump & DCB_DUMP_SYMBOLS)) dptr[((uint64_t)g * 4 + j) * NCP + c] = (int32_t)vm.symbol(code[j][c], zig);
          }
          recon_entry<NCP>(pp, recon, prev, v[j]);
          if (DUMP && MODE != 3 && (dump & DCB_DUMP_QINTS)) {
#pragma unroll
            for (int c = 0; c < NCP; ++c) dptr[((uint64_t)g * 4 + j) * NCP + c] = v[j][c];
          }
        }
        store_group4<NCP>(pp, store, dsize, optr, (uint64_t)g * 4, v);
      }
    }
    if (have) {
#pragma unroll
      for (int c = 0; c < NCP; ++c) hand->prev[c] = prev[c];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(ctl.handoff());
  }
}

// ---------------------------------------------------------------------------------------------
// Tag stream of a Tagged attribute (SymbolDecoding.cs:30-50): one rANS symbol per point = the bit length of its
// values.  16 tags per group; the consumer writes one byte per point, the running bit offset at every DCB_TAG_CHUNK
// points and validates (tag <= 32, DecoderBuffer.cs:141; bit area inside the buffer).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(384) rans_tag_rec_kernel(const uint8_t *__restrict__ arena, StreamDesc *streams,
                                                           const uint32_t *__restrict__ order, uint32_t n_streams,
                                                           uint32_t lanes, uint32_t pairs, RecGeom geom,
                                                           uint8_t *__restrict__ aux) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int kSym = 16;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t pair, role;
  warp_role(warp, pairs, pair, role);
  uint8_t *slice = smem + (size_t)pair * geom.slice_bytes;
  const uint32_t slice_addr = smem_u32(slice);
  const uint32_t q_addr = slice_addr + geom.q_off;
  const PairCtl ctl{slice_addr + geom.ctl_off};
  if (role == 0 && lane == 0) {
#pragma unroll
    for (uint32_t i = 0; i < kNumBarriers; ++i) mbar_init(ctl.base + 8u * i, 1u);
  }
  __syncthreads();
  if (role == 2u) return;
  const uint32_t slot = (blockIdx.x * pairs + pair) * lanes + lane;
  const bool have = lane < lanes && slot < n_streams;
  StreamDesc *dp = have ? &streams[order[slot]] : nullptr;
  RecHand *hand = reinterpret_cast<RecHand *>(slice + geom.hand_off) + (have ? lane : 0u);

  if (role == 0) {
    RecChain rc;
    const bool active = chain_setup(arena, dp, have, geom, slice, slice_addr, lane, rc, hand, true);
    const uint32_t n_entries = active ? dp->n_entries : 0u;
    uint32_t g_min = active ? (n_entries >> 4) : kUnset;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) g_min = min(g_min, __shfl_xor_sync(0xffffffffu, g_min, o));
    __syncwarp();
    if (lane == 0) mbar_arrive(ctl.setup());
    produce<kSym>(rc, active, g_min, q_addr, ctl, lane);
    mbar_wait(ctl.handoff(), 0u);
    if (!active) return;
    // ---- careful tail (exact `off > 0` handling, per-point checks) ----
    const StreamDesc &d = *dp;
    const uint32_t ncp = d.ncp;
    const uint64_t avail_bits = (d.buf_end - d.bits_off) * 8ull;
    uint8_t *tags = aux + d.tag_off;
    uint64_t *chunk_bits = reinterpret_cast<uint64_t *>(aux + d.tag_off + (((uint64_t)n_entries + 15ull) & ~15ull));
    RecMap vm;
    vm.set(slice_addr + lane * geom.area_bytes, *hand, geom.prec_bits, geom.ka);
    int status = hand->status;
    uint64_t bits = hand->bits;
    for (uint32_t e = hand->e; status == DCB_OK && e < n_entries; ++e) {
      if ((e & (DCB_TAG_CHUNK - 1u)) == 0) chunk_bits[e / DCB_TAG_CHUNK] = bits;
      const uint32_t tag = (uint32_t)vm.value(rc.template step<true>(), false) & 0xFFu;  // (byte) cast, SymbolDecoding.cs:41
      if (tag > 32u) {
        status = DCB_ERR_TAG;
        break;
      }
      bits += (uint64_t)tag * ncp;
      if (bits > avail_bits) {
        status = DCB_ERR_EOF;
        break;
      }
      tags[e] = (uint8_t)tag;
      rc.template top_up<1>();
      cp_async_wait<0>();
    }
    dp->bits_total = bits;
    if (status != DCB_OK) dp->status = status;
  } else {
    uint32_t ncp = 0;
    uint64_t avail_bits = 0;
    uint8_t *tags = nullptr;
    uint64_t *chunk_bits = nullptr;
    if (have) {
      const StreamDesc &d = *dp;
      ncp = d.ncp;
      avail_bits = (d.buf_end - d.bits_off) * 8ull;
      tags = aux + d.tag_off;
      chunk_bits = reinterpret_cast<uint64_t *>(aux + d.tag_off + (((uint64_t)d.n_entries + 15ull) & ~15ull));
    }
    mbar_wait(ctl.setup(), 0u);
    bool active = false;
    RecMap vm{};
    if (have) {
      active = hand->active != 0;
      if (active) vm.set(slice_addr + lane * geom.area_bytes, *hand, geom.prec_bits, geom.ka);
    }
    int status = DCB_OK;
    uint64_t bits = 0;
    uint32_t e = 0;
    for (uint32_t g = 0;; ++g) {
      const uint32_t s = g & (kStages - 1u);
      mbar_wait(ctl.full(s), (g / kStages) & 1u);
      if (lds_u32m(ctl.flag(s)) != 0u) break;
      const uint32_t qs = q_addr + s * ((uint32_t)kSym * kRowBytes) + lane * 2u;
      uint32_t t[16];
      if (active) {
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = lds_u16m(qs + (uint32_t)j * kRowBytes);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(ctl.empty(s));
      if (active && status == DCB_OK) {
#pragma unroll
        for (int j = 0; j < 16; ++j) t[j] = (uint32_t)vm.value(t[j], false) & 0xFFu;  // (byte) cast, SymbolDecoding.cs:41
        if ((e & (DCB_TAG_CHUNK - 1u)) == 0) chunk_bits[e / DCB_TAG_CHUNK] = bits;
        uint32_t tmax = 0, tsum = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          tmax = max(tmax, t[j]);
          tsum += t[j];
        }
        const uint64_t nbits = bits + (uint64_t)tsum * ncp;
        if (tmax > 32u || nbits > avail_bits) {
          // first failing point decides the status, as in the sequential reference loop
          for (int j = 0; j < 16 && status == DCB_OK; ++j) {
            if (t[j] > 32u) status = DCB_ERR_TAG;
            else {
              bits += (uint64_t)t[j] * ncp;
              if (bits > avail_bits) status = DCB_ERR_EOF;
            }
          }
        } else {
          uint4 pk;
          pk.x = t[0] | (t[1] << 8) | (t[2] << 16) | (t[3] << 24);
          pk.y = t[4] | (t[5] << 8) | (t[6] << 16) | (t[7] << 24);
          pk.z = t[8] | (t[9] << 8) | (t[10] << 16) | (t[11] << 24);
          pk.w = t[12] | (t[13] << 8) | (t[14] << 16) | (t[15] << 24);
          *reinterpret_cast<uint4 *>(tags + e) = pk;
          bits = nbits;
          e += 16;
        }
      }
    }
    if (have) {
      hand->status = status;
      hand->bits = bits;
      hand->e = e;
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(ctl.handoff());
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static uint32_t chain_low_bit() {  // DCB_CHAIN_LOW=1: chain warps in the lower half of the CTA (experiments)
  static const uint32_t bit = getenv("DCB_CHAIN_LOW") ? 0x40000000u : 0u;
  return bit;
}
// a pair's slice: lanes x area | rings | queue | control block | hand-over records
static RecGeom rec_geom(const RansLaunch &p, uint32_t syms_per_group) {
  RecGeom g{};
  g.ka = p.rec_ka;
  g.area_bytes = (p.rec_bytes + 7u) & ~7u;
  g.prec_bits = p.prec_bits;
  g.cap_exc = p.cap_exc;
  g.zig = p.zig;
  uint32_t o = p.lanes_per_warp * g.area_bytes;
  o = (o + DCB_RING_BYTES - 1u) & ~(DCB_RING_BYTES - 1u);
  g.ring_off = o;
  o += p.lanes_per_warp * DCB_RING_BYTES;
  g.q_off = o;
  o += DCB_REC_STAGES * syms_per_group * DCB_PC_ROW_BYTES;
  g.ctl_off = o;
  o += DCB_REC_CTL_BYTES;
  g.hand_off = o;
  o += p.lanes_per_warp * DCB_REC_HAND_BYTES;
  g.slice_bytes = (o + 127u) & ~127u;  // slices start on 128-byte boundaries: the rings are direct-mapped by address
  return g;
}

// the last slice is followed by a pad: a record probe for a slot of the byte region is predicated off, but a probe of
// region [ta, tb) for a slot behind tb is not formed at all -- nothing reads past a lane's area; the pad only covers
// the 128-byte alignment of the dynamic shared memory base
uint32_t dcb_rans_rec_smem_bytes(const RansLaunch &p, uint32_t syms_per_group) {
  return rec_geom(p, syms_per_group).slice_bytes * std::max(1u, p.pairs) + 128u;
}

template <int NCP, bool DUMP, int MODE>
static cudaError_t launch_raw_rec_t(const RansLaunch &p, const DevArenas &a, cudaStream_t st) {
  auto k = rans_raw_rec_kernel<NCP, DUMP, MODE>;
  const RecGeom g = rec_geom(p, 4u * NCP);
  const uint32_t pairs = std::max(1u, p.pairs);
  const uint32_t smem_bytes = g.slice_bytes * pairs + 128u;
  if (smem_bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
  }
  cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  const uint32_t per_cta = p.lanes_per_warp * pairs;
  const uint32_t grid = (p.n_streams + per_cta - 1) / per_cta;
  k<<<grid, (p.split ? 128 : 64) * pairs, smem_bytes, st>>>(a.in, p.d_streams, p.d_order, p.n_streams, p.lanes_per_warp, pairs | (p.split ? 0x80000000u : 0u) | chain_low_bit(), g, a.out, a.dbg,
                                          a.aux, p.dump);
  return cudaGetLastError();
}

cudaError_t dcb_launch_rans_raw_rec(const RansLaunch &p, int ncp, const DevArenas &a, cudaStream_t st) {
  if (!p.dump) {
#define DCB_SPEC(M, N) \
  if (p.mode == M && ncp == N) return launch_raw_rec_t<N, false, M>(p, a, st);
    DCB_SPEC(1, 3)
    DCB_SPEC(1, 2)
    DCB_SPEC(2, 3)
    DCB_SPEC(2, 4)
    DCB_SPEC(3, 2)
    DCB_SPEC(4, 3)
    DCB_SPEC(4, 2)
#undef DCB_SPEC
  }
#define DCB_CASE(N) \
  case N:           \
    return p.dump ? launch_raw_rec_t<N, true, 0>(p, a, st) : launch_raw_rec_t<N, false, 0>(p, a, st);
  switch (ncp) {
    DCB_CASE(1)
    DCB_CASE(2)
    DCB_CASE(3)
    DCB_CASE(4)
    default:
      return cudaErrorInvalidValue;
  }
#undef DCB_CASE
}

cudaError_t dcb_launch_rans_tag_rec(const RansLaunch &p, const DevArenas &a, cudaStream_t st) {
  const RecGeom g = rec_geom(p, 16u);
  const uint32_t pairs = std::max(1u, p.pairs);
  const uint32_t smem_bytes = g.slice_bytes * pairs + 128u;
  if (smem_bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rans_tag_rec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
  }
  cudaFuncSetAttribute(rans_tag_rec_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  const uint32_t per_cta = p.lanes_per_warp * pairs;
  const uint32_t grid = (p.n_streams + per_cta - 1) / per_cta;
  rans_tag_rec_kernel<<<grid, (p.split ? 128 : 64) * pairs, smem_bytes, st>>>(a.in, p.d_streams, p.d_order, p.n_streams, p.lanes_per_warp, pairs | (p.split ? 0x80000000u : 0u) | chain_low_bit(),
                                                             g, a.aux);
  return cudaGetLastError();
}
