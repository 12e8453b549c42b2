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
//   cmp_chain_kernel  the recurrence, one warp per stream over blocks of 32 entries.  All lanes build the block's
//                     records: flag positions by ballot, kept parallelograms, and -- because every entry before the
//                     block is final by then (previous block in a shared-memory ring, older ones in the quantized-int
//                     scratch) -- the SUM of all operands from outside the block, per component.  Lanes 0..NCP-1 then
//                     walk the chain, one component each: pre-summed base + k * value(p - 1) (operands that ARE entry
//                     p - 1 never leave the register) + the few operands inside the block, division by 1..4 by
//                     select, wrap.  (First version: 12 operand loads per entry, 164 ms per million entries; this
//                     one: see profiles/.)
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

// truncating signed division by 1..4 (:112): by select, no divide
__device__ __forceinline__ int32_t div_small(int32_t s, uint32_t used) {
  const int32_t d2 = (s + (int32_t)((uint32_t)s >> 31)) >> 1;
  const int32_t d4 = (s + ((s >> 31) & 3)) >> 2;
  const int32_t d3 = s / 3;
  return used == 1u ? s : used == 2u ? d2 : used == 3u ? d3 : d4;
}

template <int NCP, bool DUMP>
__global__ void __launch_bounds__(32) cmp_chain_kernel(StreamDesc *streams, const uint32_t *__restrict__ order,
                                                       uint32_t n_streams, uint8_t *__restrict__ out,
                                                       uint8_t *__restrict__ dbg, uint8_t *__restrict__ aux, uint32_t dump) {
  // per entry of the block: {meta, first four in-block operands}, eight more in-block operands, per component the
  // pre-summed operands from outside the block and the correction
  __shared__ __align__(8) uint2 s_meta[kBlk];   // x: (k & 0xFF) | used << 8 | n_in << 16   y: in-block operands 0..3
  __shared__ __align__(8) uint2 s_more[kBlk];   // in-block operands 4..11 (one byte each: index in block | 0x80 = subtract)
  __shared__ int32_t s_base[kBlk * NCP];
  __shared__ int32_t s_cor[kBlk * NCP];
  __shared__ int32_t s_ring[kRing * NCP];
  const uint32_t lane = threadIdx.x;
  const uint32_t a_ring = (uint32_t)__cvta_generic_to_shared(s_ring);
  const uint32_t a_meta = (uint32_t)__cvta_generic_to_shared(s_meta);
  const uint32_t a_more = (uint32_t)__cvta_generic_to_shared(s_more);
  const uint32_t a_base = (uint32_t)__cvta_generic_to_shared(s_base);
  const uint32_t a_cor = (uint32_t)__cvta_generic_to_shared(s_cor);
  for (uint32_t si = blockIdx.x; si < n_streams; si += gridDim.x) {
    StreamDesc &d = streams[order[si]];
    if (d.status != DCB_OK) continue;
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
    uint32_t fpos[4] = {0, 0, 0, 0};  // flags consumed per context (uniform)
    uint32_t have[4];  // flags decoded per context (cmp_flags_kernel stops at the most an attribute can consume)
#pragma unroll
    for (uint32_t c = 0; c < 4; ++c) {
      const unsigned long long cap = (unsigned long long)(c + 1u) * n, got = d.n_crease[c];
      have[c] = (uint32_t)(got < cap ? got : cap);
    }
    int status = DCB_OK;
    // registers holding the next block's dependencies (issued before the chain of the current block runs)
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
    // Records of block blk from the fetched registers.  Every entry before the block is final by now: the previous
    // block sits in the ring, older ones in the quantized-int scratch -- so every operand from outside the block is
    // summed HERE, by 32 lanes at once, and the serial chain is left with the operands inside the block only.
    auto build = [&](uint32_t blk) {
      const uint32_t e0 = blk * kBlk, p = e0 + lane;
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
      const uint32_t ring_limit = e0 >= kBlk ? e0 - kBlk : 0u;  // entries below it have left the ring
      // Every load of the block is issued before anything waits on one: the crease flags (their position needs the
      // ballots only) and, per operand, one read of the scratch and one of the ring -- at a harmless address when the
      // operand lives elsewhere -- so that a block costs one global-memory latency, not one per operand.
      const uint32_t ctx = np ? np - 1u : 0u;
      const bool short_of_flags = np && my_pos + np > have[ctx];  // :93
      const uint8_t *f = sc.flags[0] + (uint64_t)n * (ctx * (ctx + 1u) / 2u) + my_pos;
      uint32_t fl[4];
#pragma unroll
      for (uint32_t i = 0; i < 4; ++i) fl[i] = f[(np && !short_of_flags) ? min(i, np - 1u) : 0u];
      uint32_t kind[12];  // 0 unused | 1 entry p - 1 | 2 inside the block | 3 ring | 4 scratch
      uint32_t far_v[12][NCP], ring_v[12][NCP];
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const uint32_t e = (uint32_t)dep[i];
        const bool valid = (uint32_t)(i / 3) < np;
        kind[i] = !valid ? 0u : (e + 1u == p) ? 1u : (e >= e0) ? 2u : (e >= ring_limit) ? 3u : 4u;
        const uint64_t g = kind[i] == 4u ? (uint64_t)e * NCP : 0ull;
        const uint32_t r = kind[i] == 3u ? (e & (kRing - 1u)) * NCP : 0u;
#pragma unroll
        for (int c = 0; c < NCP; ++c) {
          far_v[i][c] = (uint32_t)__ldcg(qints + g + c);
          ring_v[i][c] = (uint32_t)s_ring[r + c];
        }
      }
      if (__any_sync(0xffffffffu, short_of_flags)) status = DCB_ERR_PRED;
      uint32_t keep = 0;  // bit i: parallelogram i is not a crease edge
#pragma unroll
      for (uint32_t i = 0; i < 4; ++i)
        if (i < np && !short_of_flags && fl[i] == 0) keep |= 1u << i;
      int32_t k = 0;
      uint32_t used = (uint32_t)__popc(keep), n_in = 0;
      uint32_t base[NCP];
#pragma unroll
      for (int c = 0; c < NCP; ++c) base[c] = 0;
      uint32_t w[3] = {0, 0, 0};
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const bool on = (keep >> (i / 3)) & 1u;
        const bool neg = (i % 3) == 0;  // next + prev - opp (...ParallelogramDecoder.cs:84)
        const uint32_t kd = on ? kind[i] : 0u;
        if (kd == 1u) k += neg ? -1 : 1;  // entry p - 1 rides in the chain lane's register
        if (kd == 2u) {
          const uint32_t byte = ((uint32_t)dep[i] - e0) | (neg ? 0x80u : 0u);
#pragma unroll
          for (uint32_t sl = 0; sl < 3; ++sl)
            if ((n_in >> 2) == sl) w[sl] |= byte << (8u * (n_in & 3u));
          ++n_in;
        }
#pragma unroll
        for (int c = 0; c < NCP; ++c) {
          const uint32_t v = kd == 4u ? far_v[i][c] : kd == 3u ? ring_v[i][c] : 0u;
          base[c] += neg ? 0u - v : v;
        }
      }
      if (used == 0) {  // no parallelogram left: entry p - 1 (:105-108); entry 0 is predicted from zero (:43)
        used = 1;
        k = p > 0 ? 1 : 0;
      }
      s_meta[lane] = make_uint2(((uint32_t)k & 0xFFu) | (used << 8) | (n_in << 16), w[0]);
      s_more[lane] = make_uint2(w[1], w[2]);
#pragma unroll
      for (int c = 0; c < NCP; ++c) {
        s_base[lane * NCP + c] = (int32_t)base[c];
        s_cor[lane * NCP + c] = ncor[c];
      }
    };

    __syncwarp();
    fetch(0);
    build(0);
    __syncwarp();
    int32_t prev = 0;  // chain lanes: value of entry p - 1, component `lane`
    for (uint32_t blk = 0; blk < n_blocks && status == DCB_OK; ++blk) {
      const uint32_t e0 = blk * kBlk, cnt = min(kBlk, n - e0);
      fetch(blk + 1);  // in flight while the chain runs
      if (lane < NCP) {
        const uint32_t a_blk = a_ring + ((e0 & (kRing - 1u)) * NCP + lane) * 4u;  // this block in the ring (a block never wraps it)
        uint2 meta = lds64(a_meta);
        int32_t bs = lds32(a_base + lane * 4u), co = lds32(a_cor + lane * 4u);
        for (uint32_t j = 0; j < cnt; ++j) {
          const uint32_t j1 = j + 1 < cnt ? j + 1 : j;  // record of the next entry: independent of the chain
          const uint2 nmeta = lds64(a_meta + 8u * j1);
          const int32_t nbs = lds32(a_base + (j1 * NCP + lane) * 4u), nco = lds32(a_cor + (j1 * NCP + lane) * 4u);
          uint32_t sum = (uint32_t)bs + (uint32_t)(int32_t)(int8_t)(meta.x & 0xFFu) * (uint32_t)prev;
          uint32_t n_in = (meta.x >> 16) & 15u;
          if (n_in) {  // operands decoded earlier in this block (never entry j - 1: that one is in k)
            uint32_t wv = meta.y;
            const uint2 more = n_in > 4u ? lds64(a_more + 8u * j) : make_uint2(0u, 0u);
            for (uint32_t i = 0; i < n_in; ++i) {
              if (i == 4u) wv = more.x;
              if (i == 8u) wv = more.y;
              const uint32_t b = wv & 0xFFu;
              wv >>= 8;
              const uint32_t v = (uint32_t)lds32(a_blk + (b & 31u) * (4u * NCP));
              sum += (b & 0x80u) ? 0u - v : v;
            }
          }
          const int32_t pred = div_small((int32_t)sum, (meta.x >> 8) & 7u);
          prev = wrap_original(pred, co, pp.mn, pp.mx, pp.max_diff);
          sts32(a_blk + j * (4u * NCP), prev);
          meta = nmeta;
          bs = nbs;
          co = nco;
        }
      }
      __syncwarp();
      // finished block -> quantized-int scratch (later blocks and the tex-coord / normal predictors read it), dump, output
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
      __syncwarp();
      if (blk + 1 < n_blocks) build(blk + 1);
      __syncwarp();
    }
    if (status != DCB_OK && lane == 0) d.status = status;
    __syncwarp();
  }
}

}  // namespace

cudaError_t dcb_launch_cmp_flags(StreamDesc *d_streams, const uint32_t *d_order, uint32_t n, const DevArenas &a, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  cmp_flags_kernel<<<n, 128, 0, st>>>(a.in, d_streams, d_order, n, a.aux);
  return cudaGetLastError();
}

cudaError_t dcb_launch_cmp(StreamDesc *d_streams, const uint32_t *d_order, uint32_t n, int ncp, uint32_t max_entries,
                           uint32_t dump, const DevArenas &a, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const uint32_t gx = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)max_entries + 127) / 128, 1u << 20));
  cmp_deps_kernel<<<dim3(gx, n > 65535u ? 65535u : n), 128, 0, st>>>(d_streams, d_order, n, a.maps, a.aux);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
#define DCB_CMP_LAUNCH(N)                                                                                  \
  case N:                                                                                                  \
    if (dump)                                                                                              \
      cmp_chain_kernel<N, true><<<n, 32, 0, st>>>(d_streams, d_order, n, a.out, a.dbg, a.aux, dump);       \
    else                                                                                                   \
      cmp_chain_kernel<N, false><<<n, 32, 0, st>>>(d_streams, d_order, n, a.out, a.dbg, a.aux, dump);      \
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
