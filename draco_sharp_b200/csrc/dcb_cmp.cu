// draco_sharp_b200/csrc/dcb_cmp.cu -- MeshPredictionSchemeConstrainedMultiParallelogramDecoder on the GPU.
//
// Reference: D/IO/Attributes/PredictionSchemes/MeshPredictionSchemeConstrainedMultiParallelogramDecoder.cs:32-141 over
// MeshPredictionSchemeParallelogramDecoder.TryComputeParallelogramPrediction (:62-89) and
// PredictionSchemeWrapDecodingTransform.cs:46-67, in the bitstream's semantics where the C# is defective (SURVEY
// Appendix B-17: the per-parallelogram predictions are never stored and the crease flags never decoded).
//
// For entry p the decoder swings left, then right, around the entry's corner and collects up to four parallelograms whose
// three operand entries were decoded before p (:52-80).  Entries with k parallelograms draw k flags from the flag
// sequence of context k - 1 (:89-101); the prediction is the truncated mean of the parallelograms that are not crease
// edges, or entry p - 1 when none is left (:105-114).
//
//   cmp_deps_kernel   the operand entries of every parallelogram of every entry: connectivity only, point-parallel
//   cmp_flags_kernel  the four rABS-coded flag sequences of a stream, one warp per context (RAnsBitDecoder.cs:26-30)
//   cmp_chain_kernel  the recurrence, two warps per stream over blocks of 32 entries.  The helper warp builds the next
//                     block's records while the chain warp decodes the current one: flag positions by ballot, kept
//                     parallelograms, and the SUM of all operands that are final by then (two blocks back in a
//                     shared-memory ring, older ones in the quantized-int scratch), per component; it also stores the
//                     finished block.  Lanes 0..NCP-1 of the chain warp walk the chain, one component each: pre-summed
//                     base + k * value(p - 1) (operands that ARE entry p - 1 never leave the register) + the near
//                     operands out of the ring, division by 1..4 by shift / multiply-high, wrap.
// Product code: nothing here touches oracle/.
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>

#include "dcb_device.cuh"
#include "dcb_internal.h"
#include "dcb_kernels.h"
#include "dcb_rabs.cuh"

using namespace dcb;

namespace {

constexpr uint32_t kInv = 0xFFFFFFFFu;
constexpr uint32_t kBlk = 32;   // entries per block = lanes
constexpr uint32_t kRing = 64;  // history ring: the block being decoded and the one before it

// scratch of a stream behind corrections int32[n ncp] | quantized ints int32[n ncp] (layout_shard sizes it with
// 59 n + 48 bytes): pad to 16 | deps int32[12 n] | count u8[n] (16-byte aligned size) | flags of context c: u8[(c + 1) n]
struct CmpScratch {
  int32_t *deps;
  uint8_t *cnt;
  uint8_t *flags[4];
  __device__ CmpScratch(uint8_t *aux, const StreamDesc &d) {
    const uint64_t n = d.n_entries;
    uint8_t *base = aux + d.aux_off + ((8ull * n * d.ncp + 15ull) & ~15ull);  // aux_off is 16-byte aligned: so are the deps
    deps = reinterpret_cast<int32_t *>(base);
    cnt = base + 48ull * n;
    uint8_t *f = cnt + ((n + 15ull) & ~15ull);
    flags[0] = f;
    flags[1] = f + n;
    flags[2] = f + 3ull * n;
    flags[3] = f + 6ull * n;
  }
};

__device__ __forceinline__ uint32_t c_next(uint32_t c) { return c == kInv ? c : ((c % 3u == 2u) ? c - 2u : c + 1u); }
__device__ __forceinline__ uint32_t c_prev(uint32_t c) { return c == kInv ? c : ((c % 3u == 0u) ? c + 2u : c - 1u); }

__global__ void cmp_deps_kernel(StreamDesc *streams, const uint32_t *__restrict__ order, uint32_t n_streams,
                                const uint8_t *__restrict__ maps, uint8_t *__restrict__ aux) {
  for (uint32_t si = blockIdx.y; si < n_streams; si += gridDim.y) {
    StreamDesc &d = streams[order[si]];
    if (d.status != DCB_OK) continue;
    const uint32_t n = d.n_entries;
    const uint32_t *opp = reinterpret_cast<const uint32_t *>(maps + d.map_off[0]);
    const uint32_t *c2v = reinterpret_cast<const uint32_t *>(maps + d.map_off[1]);
    const uint32_t *d2c = reinterpret_cast<const uint32_t *>(maps + d.map_off[2]);
    const int32_t *v2d = reinterpret_cast<const int32_t *>(maps + d.map_off[3]);
    const uint32_t n_corners = d.n_corners, n_vertices = d.n_vertices;
    CmpScratch sc(aux, d);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
      int32_t e[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) e[i] = -1;
      uint32_t np = 0;
      bool bad = false;
      if (p > 0) {
        const uint32_t start = d2c[p];
        uint32_t corner = start;
        bool first = true;
        uint32_t guard = 0;
        auto opposite = [&](uint32_t c) -> uint32_t {
          if (c == kInv) return c;
          if (c >= n_corners) { bad = true; return kInv; }
          return opp[c];
        };
        while (corner != kInv && !bad) {
          if (++guard > n_corners + 2u) { bad = true; break; }  // not a corner table: the swing never closes
          const uint32_t oc = opposite(corner);                 // ...ParallelogramDecoder.cs:66
          if (oc != kInv) {
            if (oc >= n_corners) { bad = true; break; }
            const uint32_t v_o = c2v[oc], v_n = c2v[c_next(oc)], v_p = c2v[c_prev(oc)];
            if (v_o >= n_vertices || v_n >= n_vertices || v_p >= n_vertices) { bad = true; break; }
            const int32_t a = v2d[v_o], b = v2d[v_n], c = v2d[v_p];
            if (a < (int32_t)p && b < (int32_t)p && c < (int32_t)p) {  // :75
              if (a < 0 || b < 0 || c < 0) { bad = true; break; }
              // (np is at most 3 here; written through a switch to keep e[] in registers)
#pragma unroll
              for (int s = 0; s < 4; ++s)
                if ((uint32_t)s == np) { e[3 * s] = a; e[3 * s + 1] = b; e[3 * s + 2] = c; }
              if (++np == 4u) break;  // Constants.ConstrainedMultiParallelogramMaxNumParallelograms (:64)
            }
          }
          corner = first ? c_next(opposite(c_next(corner))) : c_prev(opposite(c_prev(corner)));  // SwingLeft / SwingRight (:69)
          if (corner == start) break;
          if (corner == kInv && first) {  // :74-78
            first = false;
            corner = c_prev(opposite(c_prev(start)));
          }
        }
      }
      if (bad) {
        d.status = DCB_ERR_MAPS;
        np = 0;
      }
      int4 *dst = reinterpret_cast<int4 *>(sc.deps + 12ull * p);
      dst[0] = make_int4(e[0], e[1], e[2], e[3]);
      dst[1] = make_int4(e[4], e[5], e[6], e[7]);
      dst[2] = make_int4(e[8], e[9], e[10], e[11]);
      sc.cnt[p] = (uint8_t)np;
    }
  }
}

// One warp per (stream, context): rabs_decode_block (dcb_rabs.cuh).
__global__ void __launch_bounds__(128) cmp_flags_kernel(const uint8_t *__restrict__ arena, StreamDesc *streams,
                                                        const uint32_t *__restrict__ order, uint32_t n_streams,
                                                        uint8_t *__restrict__ aux) {
  __shared__ uint8_t s_win[4][kRabsWin];
  __shared__ uint8_t s_bits[4][kRabsWin];
  const uint32_t ctx = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  for (uint32_t si = blockIdx.x; si < n_streams; si += gridDim.x) {
    const StreamDesc &d = streams[order[si]];
    if (d.status != DCB_OK) continue;
    const unsigned long long cap = (unsigned long long)(ctx + 1u) * d.n_entries, got = d.n_crease[ctx];
    const uint32_t want = (uint32_t)(got < cap ? got : cap);  // more flags than that are never consumed
    if (want == 0) continue;
    CmpScratch sc(aux, d);
    rabs_decode_block(arena + d.crease_off[ctx], want, sc.flags[ctx], s_win[ctx], s_bits[ctx], lane);
  }
}

__device__ __forceinline__ int32_t lds32(uint32_t a) {
  int32_t v;
  asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, int32_t v) {
  asm volatile("st.shared.b32 [%0], %1;\n" ::"r"(a), "r"(v) : "memory");
}

// truncating signed division by 1..4 (:112) without a divide or a branch: powers of two by a biased shift (mask and
// shift come with the entry's record), 3 by a multiply-high
__device__ __forceinline__ int32_t div_small(int32_t s, uint32_t mask, uint32_t shift, bool by3) {
  const int32_t sign = s >> 31;
  const int32_t qp = (s + (sign & (int32_t)mask)) >> shift;
  const int32_t q3 = __mulhi(s, 0x55555556) - sign;  // floor(s / 3) + (s < 0)
  int32_t q;
  asm("{ .reg .pred p; setp.ne.s32 p, %3, 0; selp.s32 %0, %1, %2, p; }" : "=r"(q) : "r"(q3), "r"(qp), "r"((int)by3));
  return q;
}

// Two warps per stream, in lock step over blocks of 32 entries (one __syncthreads per block):
//   warp 0  the chain of block b (lanes 0..NCP-1, one component each)
//   warp 1  stores the finished block b - 1, builds the records of block b + 1, fetches the dependencies of block b + 2
// Operands of an entry of block b + 1, by age: entry p - 1 -> coefficient k (the value never leaves the chain lane's
// register); blocks b and b + 1 -> "near" list, read from the 64-entry ring by the chain; block b - 1 (final, still in
// the ring) and older ones (final, in the quantized-int scratch) -> summed per component by the builder.
template <int NCP, bool DUMP>
__global__ void __launch_bounds__(64) cmp_chain_kernel(StreamDesc *streams, const uint32_t *__restrict__ order,
                                                       uint32_t n_streams, uint8_t *__restrict__ out,
                                                       uint8_t *__restrict__ dbg, uint8_t *__restrict__ aux, uint32_t dump) {
  // x: (k & 0xFF) | bias mask << 8 | shift << 12 | by3 << 14 | n_near << 16   y: near operands 0..3
  __shared__ __align__(8) uint2 s_meta[2][kBlk];
  __shared__ __align__(8) uint2 s_more[2][kBlk];  // near operands 4..11 (one byte each: ring index | 0x80 = subtract)
  __shared__ int32_t s_base[2][kBlk * NCP];
  __shared__ int32_t s_cor[2][kBlk * NCP];
  __shared__ int32_t s_ring[kRing * NCP];
  __shared__ int s_status[2];  // written by the helper while it builds in iteration i: word i & 1; read by all at the top of i + 1
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t a_ring = (uint32_t)__cvta_generic_to_shared(s_ring);
  for (uint32_t si = blockIdx.x; si < n_streams; si += gridDim.x) {
    StreamDesc &d = streams[order[si]];
    if (d.status != DCB_OK) continue;  // uniform over the CTA
    const uint32_t n = d.n_entries;
    if (n == 0) continue;
    PostParams pp;
    pp.load(d);
    uint8_t *optr = out + d.out_off;
    int32_t *dptr = DUMP ? reinterpret_cast<int32_t *>(dbg + d.dbg_off) : nullptr;
    const int32_t *corr = reinterpret_cast<const int32_t *>(aux + d.aux_off);
    int32_t *qints = reinterpret_cast<int32_t *>(aux + d.aux_off) + (uint64_t)n * NCP;
    const CmpScratch sc(aux, d);
    const uint32_t n_blocks = (n + kBlk - 1) / kBlk;
    uint32_t fpos[4] = {0, 0, 0, 0};  // flags consumed per context (uniform over the helper warp)
    uint32_t have[4];  // flags decoded per context (cmp_flags_kernel stops at the most an attribute can consume)
#pragma unroll
    for (uint32_t c = 0; c < 4; ++c) {
      const unsigned long long cap = (unsigned long long)(c + 1u) * n, got = d.n_crease[c];
      have[c] = (uint32_t)(got < cap ? got : cap);
    }
    // helper warp: registers holding a block's dependencies between its fetch and its build
    int4 nd0, nd1, nd2;
    uint32_t ncnt = 0;
    int32_t ncor[NCP];
    auto fetch = [&](uint32_t blk) {
      const uint32_t p = blk * kBlk + lane;
      nd0 = nd1 = nd2 = make_int4(-1, -1, -1, -1);
      ncnt = 0;
#pragma unroll
      for (int c = 0; c < NCP; ++c) ncor[c] = 0;
      if (blk < n_blocks && p < n) {
        const int4 *src = reinterpret_cast<const int4 *>(sc.deps + 12ull * p);
        nd0 = src[0];
        nd1 = src[1];
        nd2 = src[2];
        ncnt = sc.cnt[p];
#pragma unroll
        for (int c = 0; c < NCP; ++c) ncor[c] = corr[(uint64_t)p * NCP + c];
      }
    };
    auto build = [&](uint32_t blk, uint32_t iter) {  // helper warp, while the chain warp decodes block blk - 1
      const uint32_t e0 = blk * kBlk, p = e0 + lane, buf = blk & 1u;
      const bool live = p < n;
      const uint32_t np = live ? ncnt : 0u;
      // flag positions: entries with np parallelograms draw np flags from context np - 1, in entry order (:89-92)
      uint32_t my_pos = 0;
#pragma unroll
      for (uint32_t c = 0; c < 4; ++c) {
        const uint32_t m = __ballot_sync(0xffffffffu, np == c + 1u);
        if (np == c + 1u) my_pos = fpos[c] + (c + 1u) * (uint32_t)__popc(m & ((1u << lane) - 1u));
        fpos[c] += (c + 1u) * (uint32_t)__popc(m);
      }
      const int32_t dep[12] = {nd0.x, nd0.y, nd0.z, nd0.w, nd1.x, nd1.y, nd1.z, nd1.w, nd2.x, nd2.y, nd2.z, nd2.w};
      const uint32_t near_limit = e0 >= kBlk ? e0 - kBlk : 0u;        // block blk - 1 is being decoded right now
      const uint32_t ring_limit = e0 >= 2u * kBlk ? e0 - 2u * kBlk : 0u;  // block blk - 2 is final and still in the ring
      const uint32_t ctx = np ? np - 1u : 0u;
      const bool short_of_flags = np && my_pos + np > have[ctx];  // :93
      const uint8_t *f = sc.flags[0] + (uint64_t)n * (ctx * (ctx + 1u) / 2u) + my_pos;
      uint32_t fl[4];
#pragma unroll
      for (uint32_t i = 0; i < 4; ++i) fl[i] = f[(np && !short_of_flags) ? min(i, np - 1u) : 0u];
      if (__any_sync(0xffffffffu, short_of_flags) && lane == 0) s_status[iter & 1u] = DCB_ERR_PRED;
      int32_t k = 0;
      uint32_t n_near = 0;
      uint32_t base[NCP];
#pragma unroll
      for (int c = 0; c < NCP; ++c) base[c] = 0;
      uint32_t w[3] = {0, 0, 0};
      uint32_t keep = 0;  // bit i: parallelogram i is not a crease edge
#pragma unroll
      for (uint32_t i = 0; i < 4; ++i)
        if (i < np && !short_of_flags && fl[i] == 0) keep |= 1u << i;
      const uint32_t used = (uint32_t)__popc(keep);
      // two parallelograms at a time; the second pair only when some entry of the block has more than two
      auto pair = [&](int first) {
        uint32_t kind[6];  // 0 unused | 1 entry p - 1 | 2 near | 3 ring | 4 scratch
        uint32_t far_v[6][NCP], ring_v[6][NCP];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const int slot = first + i / 3;
          const uint32_t e = (uint32_t)dep[3 * first + i];
          const bool on = (keep >> slot) & 1u;
          kind[i] = !on ? 0u : (e + 1u == p) ? 1u : (e >= near_limit) ? 2u : (e >= ring_limit) ? 3u : 4u;
          const uint64_t g = kind[i] == 4u ? (uint64_t)e * NCP : 0ull;
          const uint32_t r = kind[i] == 3u ? (e & (kRing - 1u)) * NCP : 0u;
#pragma unroll
          for (int c = 0; c < NCP; ++c) {
            far_v[i][c] = (uint32_t)__ldcg(qints + g + c);
            ring_v[i][c] = (uint32_t)s_ring[r + c];
          }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const bool neg = (i % 3) == 0;  // next + prev - opp (...ParallelogramDecoder.cs:84)
          const uint32_t kd = kind[i];
          if (kd == 1u) k += neg ? -1 : 1;
          if (kd == 2u) {
            const uint32_t byte = ((uint32_t)dep[3 * first + i] & (kRing - 1u)) | (neg ? 0x80u : 0u);
#pragma unroll
            for (uint32_t sl = 0; sl < 3; ++sl)
              if ((n_near >> 2) == sl) w[sl] |= byte << (8u * (n_near & 3u));
            ++n_near;
          }
#pragma unroll
          for (int c = 0; c < NCP; ++c) {
            const uint32_t v = kd == 4u ? far_v[i][c] : kd == 3u ? ring_v[i][c] : 0u;
            base[c] += neg ? 0u - v : v;
          }
        }
      };
      pair(0);
      if (__any_sync(0xffffffffu, np > 2u)) pair(2);
      uint32_t dv = used;
      if (used == 0) {  // no parallelogram left: entry p - 1 (:105-108); entry 0 is predicted from zero (:43)
        dv = 1;
        k = p > 0 ? 1 : 0;
      }
      const uint32_t mask = dv == 2u ? 1u : dv == 4u ? 3u : 0u, shift = dv == 2u ? 1u : dv == 4u ? 2u : 0u;
      s_meta[buf][lane] = make_uint2(((uint32_t)k & 0xFFu) | (mask << 8) | (shift << 12) | ((dv == 3u ? 1u : 0u) << 14) | (n_near << 16), w[0]);
      s_more[buf][lane] = make_uint2(w[1], w[2]);
#pragma unroll
      for (int c = 0; c < NCP; ++c) {
        s_base[buf][lane * NCP + c] = (int32_t)base[c];
        s_cor[buf][lane * NCP + c] = ncor[c];
      }
    };
    auto output = [&](uint32_t blk) {  // finished block -> quantized-int scratch (later blocks and the tex-coord / normal
      const uint32_t e0 = blk * kBlk, cnt = min(kBlk, n - e0);  // predictors read it), dump, typed output
      if (lane < cnt) {
        const uint32_t p = e0 + lane;
        int32_t v[NCP];
#pragma unroll
        for (int c = 0; c < NCP; ++c) v[c] = s_ring[(p & (kRing - 1u)) * NCP + c];
#pragma unroll
        for (int c = 0; c < NCP; ++c) qints[(uint64_t)p * NCP + c] = v[c];
        if (DUMP && (dump & DCB_DUMP_QINTS)) {
#pragma unroll
          for (int c = 0; c < NCP; ++c) dptr[(size_t)p * NCP + c] = v[c];
        }
        store_entry<NCP>(pp, pp.store, pp.dsize, optr, p, v);
      }
    };

    __syncthreads();  // the previous stream's last block has left shared memory
    if (threadIdx.x < 2) s_status[threadIdx.x] = DCB_OK;
    __syncthreads();
    if (warp == 1) {
      fetch(0);
      build(0, 1u);  // "iteration -1"
      fetch(1);
    }
    __syncthreads();
    int32_t prev = 0;  // chain lanes: value of entry p - 1, component `lane`
    uint32_t done_blocks = 0;
    for (uint32_t blk = 0; blk < n_blocks; ++blk) {
      if (s_status[(blk + 1u) & 1u] != DCB_OK) break;  // uniform: written before the last barrier, not rewritten before the next
      const uint32_t e0 = blk * kBlk, cnt = min(kBlk, n - e0), buf = blk & 1u;
      if (warp == 0) {
        if (lane < NCP) {
          const uint32_t a_lane = a_ring + lane * 4u;
          const uint32_t a_st = a_lane + (e0 & (kRing - 1u)) * (4u * NCP);  // a block never wraps the ring
          const uint32_t a_meta = (uint32_t)__cvta_generic_to_shared(s_meta[buf]);
          const uint32_t a_more = (uint32_t)__cvta_generic_to_shared(s_more[buf]);
          const uint32_t a_base = (uint32_t)__cvta_generic_to_shared(s_base[buf]) + lane * 4u;
          const uint32_t a_cor = (uint32_t)__cvta_generic_to_shared(s_cor[buf]) + lane * 4u;
          // Records are independent of the chain: they are read two entries ahead, into three register sets that take
          // turns (the loop is unrolled by three so that no loaded value has to be moved -- a move waits for its load).
          auto rec_meta = [&](uint32_t j) { return lds64(a_meta + 8u * min(j, cnt - 1u)); };  // harmless re-read past the end
          auto rec_base = [&](uint32_t j) { return lds32(a_base + min(j, cnt - 1u) * (4u * NCP)); };
          auto rec_cor = [&](uint32_t j) { return lds32(a_cor + min(j, cnt - 1u) * (4u * NCP)); };
          auto entry = [&](uint32_t j, const uint2 &m, int32_t bs, int32_t co) {
            uint32_t sum = (uint32_t)bs + (uint32_t)(int32_t)(int8_t)(m.x & 0xFFu) * (uint32_t)prev;
            const uint32_t n_near = (m.x >> 16) & 15u;
            if (n_near) {  // operands decoded in this block or the one before (never entry j - 1: that one is in k)
              uint32_t wv = m.y;
              const uint2 more = n_near > 4u ? lds64(a_more + 8u * j) : make_uint2(0u, 0u);
              for (uint32_t i = 0; i < n_near; ++i) {
                if (i == 4u) wv = more.x;
                if (i == 8u) wv = more.y;
                const uint32_t b = wv & 0xFFu;
                wv >>= 8;
                const uint32_t v = (uint32_t)lds32(a_lane + (b & (kRing - 1u)) * (4u * NCP));
                sum += (b & 0x80u) ? 0u - v : v;
              }
            }
            const int32_t pred = div_small((int32_t)sum, (m.x >> 8) & 15u, (m.x >> 12) & 3u, (m.x >> 14) & 1u);
            prev = wrap_original(pred, co, pp.mn, pp.mx, pp.max_diff);
            sts32(a_st + j * (4u * NCP), prev);
          };
          uint2 m0 = rec_meta(0), m1 = rec_meta(1), m2;
          int32_t b0 = rec_base(0), b1 = rec_base(1), b2;
          int32_t c0 = rec_cor(0), c1 = rec_cor(1), c2;
          for (uint32_t j = 0; j < cnt; j += 3) {
            m2 = rec_meta(j + 2); b2 = rec_base(j + 2); c2 = rec_cor(j + 2);
            entry(j, m0, b0, c0);
            if (j + 1 >= cnt) break;
            m0 = rec_meta(j + 3); b0 = rec_base(j + 3); c0 = rec_cor(j + 3);
            entry(j + 1, m1, b1, c1);
            if (j + 2 >= cnt) break;
            m1 = rec_meta(j + 4); b1 = rec_base(j + 4); c1 = rec_cor(j + 4);
            entry(j + 2, m2, b2, c2);
          }
        }
      } else {
        if (blk > 0) output(blk - 1);
        __syncwarp();
        if (blk + 1 < n_blocks) build(blk + 1, blk);
        fetch(blk + 2);
      }
      done_blocks = blk + 1;
      __syncthreads();
    }
    const int status = s_status[0] != DCB_OK ? s_status[0] : s_status[1];
    if (status == DCB_OK && warp == 1 && done_blocks == n_blocks) output(n_blocks - 1);
    if (status != DCB_OK && threadIdx.x == 0) d.status = status;
  }
}

}  // namespace

cudaError_t dcb_launch_cmp_flags(StreamDesc *d_streams, const uint32_t *d_order, uint32_t n, const DevArenas &a, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  cmp_flags_kernel<<<n, 128, 0, st>>>(a.in, d_streams, d_order, n, a.aux);
  return cudaGetLastError();
}

cudaError_t dcb_launch_cmp_deps(StreamDesc *d_streams, const uint32_t *d_order, uint32_t n, uint32_t max_entries, const DevArenas &a,
                                cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const uint32_t gx = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)max_entries + 127) / 128, 1u << 20));
  cmp_deps_kernel<<<dim3(gx, n > 65535u ? 65535u : n), 128, 0, st>>>(d_streams, d_order, n, a.maps, a.aux);
  return cudaGetLastError();
}

cudaError_t dcb_launch_cmp(StreamDesc *d_streams, const uint32_t *d_order, uint32_t n, int ncp, uint32_t dump, const DevArenas &a,
                           cudaStream_t st) {
  if (n == 0) return cudaSuccess;
#define DCB_CMP_LAUNCH(N)                                                                                  \
  case N:                                                                                                  \
    if (dump)                                                                                              \
      cmp_chain_kernel<N, true><<<n, 64, 0, st>>>(d_streams, d_order, n, a.out, a.dbg, a.aux, dump);       \
    else                                                                                                   \
      cmp_chain_kernel<N, false><<<n, 64, 0, st>>>(d_streams, d_order, n, a.out, a.dbg, a.aux, dump);      \
    break;
  switch (ncp) {
    DCB_CMP_LAUNCH(1)
    DCB_CMP_LAUNCH(2)
    DCB_CMP_LAUNCH(3)
    DCB_CMP_LAUNCH(4)
    default: return cudaErrorInvalidValue;
  }
#undef DCB_CMP_LAUNCH
  return cudaGetLastError();
}
