// draco_sharp_b200/csrc/dcb_par_post.cu -- the point-parallel path behind a Tagged (or uncompressed) symbol source.
//
// Once the tags of a Tagged attribute are known (rans_tag kernels: one byte per point = the bit length of its values,
// SymbolDecoding.cs:37-49), every value is position independent: its bit offset is the prefix sum of tag * nc, the
// field itself is an LSB-first bit string (DecoderBuffer.DecodeLeastSignificantBits32, DecoderBuffer.cs:138-154, B-4).
// Delta + wrap (PredictionSchemeDeltaDecoder.cs:23-37 over PredictionSchemeWrapDecodingTransform.cs:46-67) is a prefix
// sum modulo max_diff when every |correction| < max_diff <= 2^30 (the clamp is then the identity and one +-max_diff
// always re-enters the range); streams that break the condition are flagged `irregular` and re-run by the serial kernel.
//
// Shape of the kernel (round 2; the round-1 kernel was one CTA per 1,024 points with five __syncthreads around a
// synchronous staging loop, 26 % of the HBM peak):
//   * persistent WARPS, no CTA-level synchronisation at all.  A warp claims runs of consecutive 256-point chunks with an
//     atomic ticket (so the predecessor of every claimed run belongs to a warp that is running), 8 points per lane, and
//     carries the running delta value through its run in registers.  The run length is a launch parameter with two
//     useful values (dcb_par_post_plan): a WHOLE STREAM when the batch has enough streams to fill the machine (no
//     look-back at all), else ONE chunk (every chunk publishes its aggregate before it looks back).  Anything in
//     between serialises a stream: a run's later chunks cannot publish before the run's look-back has returned.
//   * the chunk's tags and bit fields arrive by 1-D TMA: cp.async.bulk global -> shared with an mbarrier transaction
//     count, issued by lane 0 one chunk ahead (two stages per warp), so the loads of chunk k+1 are in flight while
//     chunk k is scanned and extracted.  Full chunks leave the same way: the decoded entries are staged in shared
//     memory in their final layout and written by ONE cp.async.bulk shared -> global per chunk (fully coalesced,
//     no store instructions on the warp).
//   * one look-back per run, over 64-bit state words `epoch | state | value` per chunk and component (decoupled
//     look-back: a run publishes its first chunk's aggregate before it looks back; every later chunk publishes its
//     inclusive prefix directly).  The look-back is warp-wide: 32 predecessors per probe.
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>

#include "dcb_device.cuh"
#include "dcb_internal.h"
#include "dcb_kernels.h"

using namespace dcb;

namespace {

constexpr uint32_t kWarps = DCB_PAR_WARPS;       // independent warps per CTA
constexpr uint32_t kPts = DCB_TAG_CHUNK / 32u;   // consecutive points per lane (8)
static_assert(kPts == 8, "a lane owns 8 consecutive points (two tag words)");

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PP_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PP_DONE;\n"
      "bra PP_WAIT;\n"
      "PP_DONE:\n"
      "}\n" ::"r"(a),
      "r"(parity)
      : "memory");
}
// 1-D TMA: global -> shared, completion counted in bytes on an mbarrier (size a multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mbar)
               : "memory");
}
// 1-D TMA: shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ uint32_t mod_add(uint32_t a, uint32_t b, uint32_t md) {  // a, b in [0, md), md <= 2^30
  const uint32_t s = a + b;
  return s >= md ? s - md : s;
}
__device__ __forceinline__ uint32_t warp_sum_mod(uint32_t v, uint32_t md) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = mod_add(v, __shfl_xor_sync(0xffffffffu, v, o), md);
  return v;
}

// everything a warp needs to know about the stream its current chunk belongs to
struct StreamView {
  const uint8_t *tags;              // nullptr: fixed-width fields (uncompressed source)
  const uint64_t *sub_bits;         // bit offset of every chunk (Tagged)
  unsigned long long *state;        // [n_chunks][4] look-back words
  uint64_t src_base;                // arena offset of the bit area / raw values
  uint64_t bits_total;
  uint32_t n, n_chunks, fixed_bits;
  uint32_t slot;                    // position in the launch's order list
  int32_t *irregular;
  StreamDesc *desc;
  bool ok;
};

__device__ __forceinline__ StreamView view_of(StreamDesc *streams, const uint32_t *order, uint32_t slot, uint8_t *aux) {
  StreamView v;
  StreamDesc &d = streams[order[slot]];
  v.desc = &d;
  v.slot = slot;
  v.n = d.n_entries;
  v.n_chunks = (v.n + DCB_TAG_CHUNK - 1u) / DCB_TAG_CHUNK;
  uint8_t *base = aux + d.tag_off;
  const uint64_t tag_bytes = ((uint64_t)v.n + 15ull) & ~15ull;
  const bool tagged = d.scheme == SCHEME_TAGGED;
  v.tags = tagged ? base : nullptr;
  v.sub_bits = reinterpret_cast<const uint64_t *>(base + tag_bytes);
  v.state = reinterpret_cast<unsigned long long *>(base + tag_bytes + 8ull * (v.n_chunks + 1u));
  v.src_base = tagged ? d.bits_off : d.raw_off;
  v.bits_total = d.bits_total;
  v.fixed_bits = 8u * d.raw_num_bytes;
  v.irregular = &d.irregular;
  v.ok = d.status == DCB_OK && v.n > 0;
  return v;
}

// one chunk of work, as the warp's pipeline sees it
struct Job {
  uint32_t run;        // global run index (UINT32_MAX: none)
  uint32_t chunk;      // chunk index inside the stream
  uint32_t k;          // position inside the run
  uint32_t lead;       // bits in front of the first field inside the staged bytes
  uint32_t cnt;        // points in the chunk
};

template <int NCP>
struct Geo {
  static constexpr uint32_t kBitsCap = (DCB_TAG_CHUNK * NCP * 4u + 48u + 15u) & ~15u;  // staged bit fields (32-bit fields worst case)
  static constexpr uint32_t kStage = DCB_TAG_CHUNK + kBitsCap;
  static constexpr uint32_t kOutCap = DCB_TAG_CHUNK * NCP * 4u;
  static constexpr uint32_t kWarpBytes = 2u * kStage + kOutCap + 16u;
};

template <int NCP, bool DUMP>
__global__ void __launch_bounds__(kWarps * 32) par_post2_kernel(const uint8_t *__restrict__ arena, StreamDesc *streams,
                                                                const uint32_t *__restrict__ order,
                                                                const uint32_t *__restrict__ run_prefix, uint32_t n_streams,
                                                                uint32_t total_runs, uint32_t kRun, uint32_t kClaim,
                                                                uint32_t n_rounds, unsigned int *ticket, uint8_t *__restrict__ out, uint8_t *__restrict__ dbg,
                                                                uint8_t *__restrict__ aux, uint32_t dump, uint32_t epoch) {
  extern __shared__ __align__(128) uint8_t smem[];
  typedef Geo<NCP> G;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint8_t *ws = smem + (size_t)warp * G::kWarpBytes;
  const uint32_t ws_addr = smem_addr(ws);
  const uint32_t a_stage[2] = {ws_addr, ws_addr + G::kStage};
  const uint32_t a_out = ws_addr + 2u * G::kStage;
  const uint32_t a_bar[2] = {a_out + G::kOutCap, a_out + G::kOutCap + 8u};
  if (lane == 0) {
    mbar_init(a_bar[0], 1u);
    mbar_init(a_bar[1], 1u);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();
  uint32_t parity[2] = {0u, 0u};
  const unsigned long long ep = (unsigned long long)(epoch & 0x3FFFFFFFu);
  const unsigned long long tag_agg = (ep << 34) | (1ull << 32), tag_pre = (ep << 34) | (2ull << 32);

  // ---- ticketing: kClaim consecutive runs per atomic, the next ticket fetched while the current one is worked on ----
  uint32_t claim_base = 0xFFFFFFFFu, claim_next = 0xFFFFFFFFu;
  auto fetch_ticket = [&]() -> uint32_t {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(ticket, kClaim);
    return __shfl_sync(0xffffffffu, t, 0);
  };
  // chunk-sized runs look back at their predecessors: a claim held in reserve would keep its successors waiting for as
  // long as the claim in front of it takes, so the next ticket is only prefetched when runs are whole streams
  const bool prefetch_ticket = kRun > 1u;
  claim_base = fetch_ticket();
  if (prefetch_ticket) claim_next = fetch_ticket();
  uint32_t claim_pos = 0;  // run inside the claim

  uint32_t slot_hint = 0;
  // run -> item: run_prefix[item] <= run < run_prefix[item + 1]; tickets only grow, so the search gallops forward.
  // Items are streams (stream-major runs) or rounds (n_rounds != 0: chunk r of every stream that has one, see the host).
  const uint32_t n_items = n_rounds ? n_rounds : n_streams;
  auto slot_of_run = [&](uint32_t run) -> uint32_t {
    uint32_t lo = slot_hint;
    if (run < run_prefix[lo + 1]) return lo;
    uint32_t step = 1, hi = lo + 1;
    while (hi < n_items && run >= run_prefix[hi + 1 > n_items ? n_items : hi + 1]) {
      lo = hi;
      hi = hi + step > n_items ? n_items : hi + step;
      step <<= 1;
      if (hi == n_items) break;
    }
    // invariant: run_prefix[lo] <= run; answer in [lo, min(hi, n_items - 1)]
    uint32_t a = lo, b = hi >= n_items ? n_items - 1u : hi;
    while (a < b) {
      const uint32_t mid = (a + b + 1u) >> 1;
      if (run_prefix[mid] <= run) a = mid;
      else b = mid - 1u;
    }
    return a;
  };

  // the job iterator: next chunk of this warp (run == UINT32_MAX when the work is exhausted)
  Job cur, nxt;
  StreamView vcur, vnxt;
  vcur.ok = vnxt.ok = false;
  vcur.slot = vnxt.slot = 0xFFFFFFFFu;
  auto advance = [&](const Job &j, const StreamView &vj, Job &o, StreamView &vo) {
    // continue the run?
    if (j.run != 0xFFFFFFFFu && vj.ok && j.k + 1u < kRun && j.chunk + 1u < vj.n_chunks) {
      o.run = j.run;
      o.k = j.k + 1u;
      o.chunk = j.chunk + 1u;
      vo = vj;
    } else {
      for (;;) {
        if (claim_pos >= kClaim) {
          if (prefetch_ticket) {
            claim_base = claim_next;
            claim_next = fetch_ticket();
          } else {
            claim_base = fetch_ticket();
          }
          claim_pos = 0;
        }
        const uint32_t run = claim_base + claim_pos;
        ++claim_pos;
        if (claim_base >= total_runs || run >= total_runs) {
          o.run = 0xFFFFFFFFu;
          o.cnt = 0;
          return;
        }
        const uint32_t item = slot_of_run(run);
        slot_hint = item;
        const uint32_t slot = n_rounds ? run - run_prefix[item] : item;
        if (slot != vj.slot || !vj.ok) vo = view_of(streams, order, slot, aux);
        else vo = vj;
        if (!vo.ok) continue;  // failed or empty stream: its runs are skipped
        o.run = run;
        o.k = 0;
        o.chunk = n_rounds ? item : (run - run_prefix[item]) * kRun;
        if (o.chunk >= vo.n_chunks) continue;
        break;
      }
    }
    o.cnt = min(DCB_TAG_CHUNK, vo.n - o.chunk * DCB_TAG_CHUNK);
  };

  // issue the TMA loads of a job into stage s (lane 0); every lane learns `lead`
  auto prepare = [&](Job &j, const StreamView &v, uint32_t s) {
    if (j.run == 0xFFFFFFFFu) return;
    uint64_t bit_begin, bit_end;
    if (v.tags) {
      bit_begin = v.sub_bits[j.chunk];
      bit_end = (j.chunk + 1u < v.n_chunks) ? v.sub_bits[j.chunk + 1u] : v.bits_total;
    } else {
      bit_begin = (uint64_t)j.chunk * DCB_TAG_CHUNK * NCP * v.fixed_bits;
      bit_end = bit_begin + (uint64_t)j.cnt * NCP * v.fixed_bits;
    }
    const uint64_t byte_begin = v.src_base + (bit_begin >> 3);
    const uint64_t byte_end = v.src_base + ((bit_end + 7ull) >> 3);
    const uint64_t g_lo = byte_begin & ~15ull;
    uint64_t g_hi = (byte_end + 4ull + 15ull) & ~15ull;  // the funnel of the last field reads one word further
    if (g_hi - g_lo > G::kBitsCap) g_hi = g_lo + G::kBitsCap;  // cannot happen for tags <= 32 (the tag kernel enforces it)
    j.lead = (uint32_t)((byte_begin - g_lo) * 8ull + (bit_begin & 7ull));
    if (lane == 0) {
      const uint32_t nb_bits = (uint32_t)(g_hi - g_lo);
      const uint32_t nb_tags = v.tags ? ((j.cnt + 15u) & ~15u) : 0u;
      mbar_expect_tx(a_bar[s], nb_bits + nb_tags);
      if (nb_tags) bulk_g2s(a_stage[s], v.tags + (size_t)j.chunk * DCB_TAG_CHUNK, nb_tags, a_bar[s]);
      bulk_g2s(a_stage[s] + DCB_TAG_CHUNK, arena + g_lo, nb_bits, a_bar[s]);
    }
  };

  // ---- prologue ----
  cur.run = 0xFFFFFFFFu;
  cur.k = cur.chunk = cur.cnt = cur.lead = 0;
  advance(cur, vcur, nxt, vnxt);
  cur = nxt;
  vcur = vnxt;
  uint32_t s = 0;
  prepare(cur, vcur, s);

  // per-run state of the chain of delta values (uniform over the warp)
  uint32_t carry[NCP];
#pragma unroll
  for (int c = 0; c < NCP; ++c) carry[c] = 0;
  PostParams pp;
  uint32_t pp_slot = 0xFFFFFFFFu;
  uint8_t *obase = nullptr;
  int32_t *dptr = nullptr;
  int recon = RECON_NONE;
  bool zig = false, to_scratch = false, chain_follows = false;
  bool out_pending = false;

  while (cur.run != 0xFFFFFFFFu) {
    // the next chunk's loads go out before this one is touched
    advance(cur, vcur, nxt, vnxt);
    prepare(nxt, vnxt, s ^ 1u);

    const StreamView &v = vcur;
    if (pp_slot != v.slot) {
      const StreamDesc &d = *v.desc;
      pp.load(d);
      pp_slot = v.slot;
      recon = d.recon;
      zig = d.zigzag != 0;
      obase = out + d.out_off;
      dptr = DUMP ? reinterpret_cast<int32_t *>(dbg + d.dbg_off) : nullptr;
      // parallelogram corrections and octahedral corrections go to the stream's int32 scratch: para_chain_kernel /
      // oct_chain_kernel run the recurrences that are not scans
      chain_follows = recon == RECON_PARA_WRAP || recon == RECON_DELTA_OCT || recon == RECON_DELTA_OCT_CANON ||
                      recon == RECON_GEO_OCT || recon == RECON_GEO_OCT_CANON;
      to_scratch = chain_follows || pp.store == STORE_OCT_UNIT;  // oct_unit_kernel reads (s, t) pairs from the scratch
      if (to_scratch) {
        pp.store = STORE_NARROW;
        pp.dsize = 4;
        obase = aux + d.aux_off;
      }
    }
    const uint32_t e0 = cur.chunk * DCB_TAG_CHUNK, cnt = cur.cnt;
    const uint32_t p0 = lane * kPts;
    const uint32_t mine_cnt = p0 < cnt ? min(kPts, cnt - p0) : 0u;

    mbar_wait(a_bar[s], parity[s]);
    parity[s] ^= 1u;
    const uint8_t *st_tags = ws + (size_t)s * G::kStage;
    const uint32_t *sm_bits = reinterpret_cast<const uint32_t *>(ws + (size_t)s * G::kStage + DCB_TAG_CHUNK);

    // ---- bit lengths of this lane's points, exclusive scan over the warp ----
    uint32_t tg[kPts];
    if (v.tags) {
      const uint2 w = *reinterpret_cast<const uint2 *>(st_tags + p0);
#pragma unroll
      for (uint32_t j = 0; j < kPts; ++j) {
        const uint32_t ww = j < 4 ? w.x : w.y;
        tg[j] = j < mine_cnt ? ((ww >> (8 * (j & 3))) & 0xFFu) : 0u;
      }
    } else {
#pragma unroll
      for (uint32_t j = 0; j < kPts; ++j) tg[j] = j < mine_cnt ? v.fixed_bits : 0u;
    }
    uint32_t mine = 0;
#pragma unroll
    for (uint32_t j = 0; j < kPts; ++j) mine += tg[j];
    mine *= NCP;
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    // ---- extract + zig-zag ----
    uint32_t bpos = cur.lead + (incl - mine);
    int32_t val[kPts][NCP];
#pragma unroll
    for (uint32_t j = 0; j < kPts; ++j) {
      const uint32_t t = tg[j];
      const uint32_t msk = t >= 32u ? 0xFFFFFFFFu : ((1u << t) - 1u);
#pragma unroll
      for (int c = 0; c < NCP; ++c) {
        const uint32_t w = bpos >> 5;
        const uint32_t sym = __funnelshift_r(sm_bits[w], sm_bits[w + 1], bpos) & msk;
        bpos += t;
        if (DUMP && (dump & DCB_DUMP_SYMBOLS) && j < mine_cnt) dptr[(size_t)(e0 + p0 + j) * NCP + c] = (int32_t)sym;
        val[j][c] = zig ? zigzag_dec(sym) : (int32_t)sym;
      }
    }

    if (recon == RECON_DELTA_WRAP) {
      const uint32_t md = (uint32_t)pp.max_diff;
      const bool md_ok = md >= 1u && md <= (1u << 30);
      bool bad = !md_ok;
      uint32_t tsum[NCP], inc[NCP], agg[NCP];
#pragma unroll
      for (int c = 0; c < NCP; ++c) {
        tsum[c] = 0;
#pragma unroll
        for (uint32_t j = 0; j < kPts; ++j) {
          const int32_t x = val[j][c];
          const uint32_t ax = x < 0 ? 0u - (uint32_t)x : (uint32_t)x;
          bad |= (j < mine_cnt) & (ax >= md);
          const uint32_t res = x < 0 ? (uint32_t)x + md : (uint32_t)x;  // residue in [0, md) when not bad
          val[j][c] = (int32_t)res;
          tsum[c] = (j < mine_cnt && !bad) ? mod_add(tsum[c], res, md) : tsum[c];
        }
        inc[c] = tsum[c];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, inc[c], o);
          if (lane >= (uint32_t)o) inc[c] = mod_add(inc[c], t, md);
        }
        agg[c] = __shfl_sync(0xffffffffu, inc[c], 31);
      }
      if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(reinterpret_cast<unsigned int *>(v.irregular), 1u);
      // ---- prefix of the chunks in front of this one ----
      if (cur.k == 0) {
        if (cur.chunk == 0) {
#pragma unroll
          for (int c = 0; c < NCP; ++c) carry[c] = 0;
        } else {
          // decoupled look-back: the aggregate goes out first, so that successors never wait for this run's own wait
          if (lane < NCP) {
            uint32_t a = agg[0];
#pragma unroll
            for (int c = 1; c < NCP; ++c) a = lane == (uint32_t)c ? agg[c] : a;
            *reinterpret_cast<volatile unsigned long long *>(&v.state[(size_t)cur.chunk * 4 + lane]) = tag_agg | a;
          }
#pragma unroll
          for (int c = 0; c < NCP; ++c) {
            int64_t base = (int64_t)cur.chunk - 1;
            uint32_t acc = 0;
            for (;;) {
              const int64_t idx = base - (int64_t)lane;
              unsigned long long wv = tag_pre;  // in front of chunk 0: prefix 0
              if (idx >= 0) wv = *reinterpret_cast<volatile const unsigned long long *>(&v.state[(size_t)idx * 4 + c]);
              const bool valid = (wv >> 34) == ep && ((wv >> 32) & 3ull) != 0ull;
              const bool pre = valid && ((wv >> 32) & 3ull) == 2ull;
              const uint32_t P = __ballot_sync(0xffffffffu, pre), V = __ballot_sync(0xffffffffu, valid);
              if (P) {
                const uint32_t fp = (uint32_t)__ffs((int)P) - 1u;
                const uint32_t need = fp >= 31u ? 0xFFFFFFFFu : ((2u << fp) - 1u);
                if ((V & need) != need) continue;  // a predecessor in between has not published yet
                acc = mod_add(acc, warp_sum_mod(lane <= fp ? (uint32_t)wv : 0u, md), md);
                break;
              }
              if (V != 0xFFFFFFFFu) continue;
              acc = mod_add(acc, warp_sum_mod((uint32_t)wv, md), md);
              base -= 32;
            }
            carry[c] = acc;
          }
        }
      }
      // this chunk's inclusive prefix: what every later chunk of the stream needs
      uint32_t after[NCP];
#pragma unroll
      for (int c = 0; c < NCP; ++c) after[c] = mod_add(carry[c], agg[c], md);
      if (lane < NCP) {
        uint32_t a = after[0];
#pragma unroll
        for (int c = 1; c < NCP; ++c) a = lane == (uint32_t)c ? after[c] : a;
        *reinterpret_cast<volatile unsigned long long *>(&v.state[(size_t)cur.chunk * 4 + lane]) = tag_pre | a;
      }
      // first element of the stream: prediction = clamp(0) (PredictionSchemeDeltaDecoder.cs:30 + ClampPredictedValue)
      const int32_t p0v = 0 > pp.mx ? pp.mx : (0 < pp.mn ? pp.mn : 0);
#pragma unroll
      for (int c = 0; c < NCP; ++c) {
        uint32_t before = mod_add(inc[c], tsum[c] == 0u ? 0u : md - tsum[c], md);  // lanes in front of this one
        before = mod_add(before, carry[c], md);
        uint32_t acc = mod_add(before, (uint32_t)(p0v - pp.mn), md);
#pragma unroll
        for (uint32_t j = 0; j < kPts; ++j) {
          acc = mod_add(acc, (uint32_t)val[j][c], md);
          val[j][c] = (int32_t)((uint32_t)pp.mn + acc);
        }
        carry[c] = after[c];
      }
    }
    // ---- store ----
    if (DUMP && (dump & DCB_DUMP_QINTS) && !chain_follows) {
#pragma unroll
      for (uint32_t j = 0; j < kPts; ++j)
        if (j < mine_cnt)
#pragma unroll
          for (int c = 0; c < NCP; ++c) dptr[(size_t)(e0 + p0 + j) * NCP + c] = val[j][c];
    }
    const uint32_t esz = pp.store == STORE_DEQUANT ? 4u * NCP : (uint32_t)pp.dsize * NCP;  // bytes per entry
    const uint32_t chunk_bytes = cnt * esz;
    if (cnt == DCB_TAG_CHUNK || (chunk_bytes & 15u) == 0u && (cnt & 7u) == 0u) {
      // full lanes only: stage the chunk in its final layout, one bulk store
      if (out_pending) {
        if (lane == 0) bulk_wait_read0();  // the previous chunk's bulk store has read the staging buffer
        __syncwarp();
      }
      if (mine_cnt) {
        const uint32_t a_mine = a_out + p0 * esz;
        if (pp.store == STORE_DEQUANT || pp.dsize == 4) {
#pragma unroll
          for (uint32_t q = 0; q < 2u * NCP; ++q) {
            uint32_t wd[4];
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
              const uint32_t idx = q * 4u + k, j = idx / NCP, c = idx % NCP;
              wd[k] = pp.store == STORE_DEQUANT ? __float_as_uint(pp.dequant(val[j][c], (int)c)) : (uint32_t)val[j][c];
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(a_mine + 16u * q), "r"(wd[0]), "r"(wd[1]), "r"(wd[2]),
                         "r"(wd[3])
                         : "memory");
          }
        } else if (pp.dsize == 1) {
#pragma unroll
          for (uint32_t q = 0; q < 2u * NCP; ++q) {  // 8 * NCP bytes = 2 * NCP words
            uint32_t wd = 0;
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
              const uint32_t idx = q * 4u + k, j = idx / NCP, c = idx % NCP;
              wd |= ((uint32_t)val[j][c] & 0xFFu) << (8u * k);
            }
            asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(a_mine + 4u * q), "r"(wd) : "memory");
          }
        } else {
#pragma unroll
          for (uint32_t q = 0; q < 4u * NCP; ++q) {  // 16 * NCP bytes = 4 * NCP words
            const uint32_t i0 = q * 2u, i1 = q * 2u + 1u;
            const uint32_t wd = ((uint32_t)val[i0 / NCP][i0 % NCP] & 0xFFFFu) | (((uint32_t)val[i1 / NCP][i1 % NCP] & 0xFFFFu) << 16);
            asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(a_mine + 4u * q), "r"(wd) : "memory");
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        bulk_s2g(obase + (size_t)e0 * esz, a_out, chunk_bytes);
        bulk_commit();
      }
      out_pending = true;
    } else if (mine_cnt) {
      // ragged tail of a stream: plain stores
#pragma unroll
      for (uint32_t q = 0; q < kPts / 4; ++q) {
        if (4 * q + 4 <= mine_cnt) {
          store_group4<NCP>(pp, pp.store, pp.dsize, obase, (uint64_t)e0 + p0 + 4 * q, val + 4 * q);
        } else {
#pragma unroll
          for (uint32_t j = 0; j < 4; ++j)
            if (4 * q + j < mine_cnt) store_entry<NCP>(pp, pp.store, pp.dsize, obase, (uint64_t)e0 + p0 + 4 * q + j, val[4 * q + j]);
        }
      }
    }
    __syncwarp();  // every lane is done with stage s: the chunk after next may land there
    cur = nxt;
    vcur = vnxt;
    s ^= 1u;
  }
  if (out_pending && lane == 0) bulk_wait0();  // shared memory must outlive the last bulk store's reads
  __syncwarp();
}

}  // namespace

void dcb_par_post_plan(uint32_t n_streams, uint64_t total_chunks, uint32_t max_chunks, bool any_delta, uint32_t num_sms, int ncp,
                       uint32_t share, uint32_t *run_len, uint32_t *claim) {
  const uint32_t smem = dcb_par_post_smem_bytes(ncp);
  // `share` pipeline slices decode on this device at the same time: each one fills its part of the machine
  const uint64_t warps = std::max<uint64_t>(
      kWarps, (uint64_t)num_sms * std::max<uint32_t>(1u, (227u * 1024u) / (smem + 1024u)) * kWarps / std::max(1u, share));
  // whole-stream runs: the longest stream must be a small part of one warp's share of the batch
  const bool whole = (uint64_t)max_chunks * warps * 2ull <= total_chunks && n_streams >= 2ull * warps;
  if (whole) {
    *run_len = std::max(1u, max_chunks);
    *claim = 1u;
  } else if (any_delta) {
    *run_len = 1u;
    *claim = 4u;
  } else {
    *run_len = 8u;  // no scan: runs only keep a warp on one stream's descriptor for a while
    *claim = 2u;
  }
}

uint32_t dcb_par_post_smem_bytes(int ncp) {
  switch (ncp) {
    case 1: return Geo<1>::kWarpBytes * kWarps;
    case 2: return Geo<2>::kWarpBytes * kWarps;
    case 3: return Geo<3>::kWarpBytes * kWarps;
    default: return Geo<4>::kWarpBytes * kWarps;
  }
}

template <int NCP>
static cudaError_t launch_par_post_n(StreamDesc *d_streams, const uint32_t *d_order, const uint32_t *d_run_prefix, uint32_t n,
                                     uint32_t total_runs, uint32_t run_len, uint32_t claim, uint32_t n_rounds, unsigned int *d_ticket,
                                     uint32_t num_sms, uint32_t dump, uint32_t epoch, const DevArenas &a, cudaStream_t st) {
  const uint32_t smem = Geo<NCP>::kWarpBytes * kWarps;
  auto k0 = par_post2_kernel<NCP, false>;
  auto k1 = par_post2_kernel<NCP, true>;
  cudaError_t e = cudaFuncSetAttribute(k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaFuncSetAttribute(k0, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  // persistent: as many CTAs as fit the machine, never more than there are claims
  const uint32_t per_sm = std::max<uint32_t>(1u, (227u * 1024u) / (smem + 1024u));
  const uint32_t claims = (total_runs + claim - 1u) / claim;
  uint32_t grid = std::min<uint32_t>(num_sms * per_sm, (claims + kWarps - 1u) / kWarps);
  grid = std::max<uint32_t>(grid, 1u);
  e = cudaMemsetAsync(d_ticket, 0, sizeof(unsigned int), st);
  if (e != cudaSuccess) return e;
  if (dump)
    k1<<<grid, kWarps * 32, smem, st>>>(a.in, d_streams, d_order, d_run_prefix, n, total_runs, run_len, claim, n_rounds, d_ticket, a.out, a.dbg, a.aux, dump, epoch);
  else
    k0<<<grid, kWarps * 32, smem, st>>>(a.in, d_streams, d_order, d_run_prefix, n, total_runs, run_len, claim, n_rounds, d_ticket, a.out, a.dbg, a.aux, dump, epoch);
  return cudaGetLastError();
}

cudaError_t dcb_launch_par_post(StreamDesc *d_streams, const uint32_t *d_order, const uint32_t *d_run_prefix, uint32_t n,
                                uint32_t total_runs, uint32_t run_len, uint32_t claim, uint32_t n_rounds, unsigned int *d_ticket,
                                uint32_t num_sms, int ncp, uint32_t dump, uint32_t epoch, const DevArenas &a, cudaStream_t st) {
  if (n == 0 || total_runs == 0) return cudaSuccess;
  run_len = std::max(1u, run_len);
  claim = std::max(1u, claim);
  switch (ncp) {
    case 1: return launch_par_post_n<1>(d_streams, d_order, d_run_prefix, n, total_runs, run_len, claim, n_rounds, d_ticket, num_sms, dump, epoch, a, st);
    case 2: return launch_par_post_n<2>(d_streams, d_order, d_run_prefix, n, total_runs, run_len, claim, n_rounds, d_ticket, num_sms, dump, epoch, a, st);
    case 3: return launch_par_post_n<3>(d_streams, d_order, d_run_prefix, n, total_runs, run_len, claim, n_rounds, d_ticket, num_sms, dump, epoch, a, st);
    case 4: return launch_par_post_n<4>(d_streams, d_order, d_run_prefix, n, total_runs, run_len, claim, n_rounds, d_ticket, num_sms, dump, epoch, a, st);
    default: return cudaErrorInvalidValue;
  }
}
