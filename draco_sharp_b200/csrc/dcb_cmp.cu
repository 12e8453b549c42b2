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
//                     records (flag positions by ballot, kept parallelograms, operand ADDRESSES: the 64-entry history
//                     ring in shared memory, or operands gathered from the quantized-int scratch when older); lanes
//                     0..NCP-1 then walk the chain, one component each, branch-free: 12 operand loads (unused slots
//                     point at a zero word), + k * value(p - 1) for operands that ARE entry p - 1 so that value never
//                     leaves its register, division by 1..4 by select, wrap.
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
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
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

struct Operands {
  int32_t v[12];
};

template <int NCP, bool DUMP>
__global__ void __launch_bounds__(32) cmp_chain_kernel(StreamDesc *streams, const uint32_t *__restrict__ order,
                                                       uint32_t n_streams, uint8_t *__restrict__ out,
                                                       uint8_t *__restrict__ dbg, uint8_t *__restrict__ aux, uint32_t dump) {
  __shared__ __align__(16) uint32_t s_addr[kBlk * 12];  // operand addresses of component 0; slot i: opp, next, prev
  __shared__ uint32_t s_meta[kBlk];                     // (k & 0xFF) | used << 8
  __shared__ int32_t s_cor[kBlk * NCP];
  __shared__ int32_t s_far[kBlk * 12 * NCP];            // operands older than the ring
  __shared__ int32_t s_ring[kRing * NCP];
  __shared__ int32_t s_zero[4];
  const uint32_t lane = threadIdx.x;
  const uint32_t a_ring = (uint32_t)__cvta_generic_to_shared(s_ring);
  const uint32_t a_far = (uint32_t)__cvta_generic_to_shared(s_far);
  const uint32_t a_zero = (uint32_t)__cvta_generic_to_shared(s_zero);
  const uint32_t a_addr = (uint32_t)__cvta_generic_to_shared(s_addr);
  const uint32_t a_meta = (uint32_t)__cvta_generic_to_shared(s_meta);
  const uint32_t a_cor = (uint32_t)__cvta_generic_to_shared(s_cor);
  if (lane < 4) s_zero[lane] = 0;
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
    // records of block blk from the fetched registers; every entry before block blk - 1 is in the scratch by now
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
      bool short_of_flags = false;
      uint32_t keep = 0;  // bit i: parallelogram i is not a crease edge
      if (np) {
        if (my_pos + np > have[np - 1u]) {
          short_of_flags = true;  // :93
        } else {
          const uint8_t *f = sc.flags[np - 1u] + my_pos;
          for (uint32_t i = 0; i < np; ++i)
            if (f[i] == 0) keep |= 1u << i;
        }
      }
      if (__any_sync(0xffffffffu, short_of_flags)) status = DCB_ERR_PRED;
      const int32_t dep[12] = {nd0.x, nd0.y, nd0.z, nd0.w, nd1.x, nd1.y, nd1.z, nd1.w, nd2.x, nd2.y, nd2.z, nd2.w};
      const uint32_t far_limit = e0 >= kBlk ? e0 - kBlk : 0u;  // entries below it have left the ring
      int32_t k = 0;
      uint32_t used = 0, slot = 0;
      uint32_t addr[12];
#pragma unroll
      for (int i = 0; i < 12; ++i) addr[i] = a_zero;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (keep & (1u << i)) {
          ++used;
#pragma unroll
          for (int o = 0; o < 3; ++o) {
            const uint32_t e = (uint32_t)dep[3 * i + o];
            uint32_t a;
            if (e + 1u == p) {
              k += (o == 0) ? -1 : 1;
              a = a_zero;
            } else if (e < far_limit) {
              const uint32_t idx = (lane * 12u + slot * 3u + (uint32_t)o) * NCP;
#pragma unroll
              for (int c = 0; c < NCP; ++c) s_far[idx + c] = qints[(uint64_t)e * NCP + c];
              a = a_far + idx * 4u;
            } else {
              a = a_ring + ((e & (kRing - 1u)) * NCP) * 4u;
            }
            // kept parallelograms are packed into the leading slots (written through selects: addr[] stays in registers)
#pragma unroll
            for (int s = 0; s < 4; ++s)
              if ((uint32_t)s == slot) addr[3 * s + o] = a;
          }
          ++slot;
        }
      }
      if (used == 0) {  // no parallelogram left: entry p - 1 (:105-108); entry 0 is predicted from zero (:43)
        used = 1;
        k = p > 0 ? 1 : 0;
      }
      uint4 *ra = reinterpret_cast<uint4 *>(&s_addr[lane * 12u]);
      ra[0] = make_uint4(addr[0], addr[1], addr[2], addr[3]);
      ra[1] = make_uint4(addr[4], addr[5], addr[6], addr[7]);
      ra[2] = make_uint4(addr[8], addr[9], addr[10], addr[11]);
      s_meta[lane] = ((uint32_t)k & 0xFFu) | (used << 8);
#pragma unroll
      for (int c = 0; c < NCP; ++c) s_cor[lane * NCP + c] = ncor[c];
    };
    auto load_ops = [&](uint32_t j, Operands &o) {  // the twelve operands of the block's entry j, this lane's component
      const uint4 r0 = lds128(a_addr + j * 48u), r1 = lds128(a_addr + j * 48u + 16u), r2 = lds128(a_addr + j * 48u + 32u);
      const uint32_t co = 4u * lane;
      o.v[0] = lds32(r0.x + co); o.v[1] = lds32(r0.y + co); o.v[2] = lds32(r0.z + co); o.v[3] = lds32(r0.w + co);
      o.v[4] = lds32(r1.x + co); o.v[5] = lds32(r1.y + co); o.v[6] = lds32(r1.z + co); o.v[7] = lds32(r1.w + co);
      o.v[8] = lds32(r2.x + co); o.v[9] = lds32(r2.y + co); o.v[10] = lds32(r2.z + co); o.v[11] = lds32(r2.w + co);
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
        const uint32_t a_st = a_ring + ((e0 & (kRing - 1u)) * NCP + lane) * 4u;  // a block never wraps the ring
        Operands cur, nxt;
        load_ops(0, cur);
        uint32_t meta = (uint32_t)lds32(a_meta);
        int32_t co = lds32(a_cor + lane * 4u);
        for (uint32_t j = 0; j < cnt; ++j) {
          // operands of entry j + 1: everything up to entry j - 1 is in shared memory, entry j itself rides in k
          const uint32_t j1 = j + 1 < cnt ? j + 1 : j;
          load_ops(j1, nxt);
          const uint32_t nmeta = (uint32_t)lds32(a_meta + 4u * j1);
          const int32_t nco = lds32(a_cor + (j1 * NCP + lane) * 4u);
          uint32_t sum = 0;
#pragma unroll
          for (int i = 0; i < 4; ++i) sum += (uint32_t)cur.v[3 * i + 1] + (uint32_t)cur.v[3 * i + 2] - (uint32_t)cur.v[3 * i];
          const int32_t kk = (int32_t)(int8_t)(meta & 0xFFu);
          sum += (uint32_t)kk * (uint32_t)prev;
          const int32_t pred = div_small((int32_t)sum, meta >> 8);
          prev = wrap_original(pred, co, pp.mn, pp.mx, pp.max_diff);
          sts32(a_st + j * (4u * NCP), prev);
          cur = nxt;
          meta = nmeta;
          co = nco;
        }
      }
      __syncwarp();
      // finished block -> quantized-int scratch (later gathers and the tex-coord predictor read it), dump, typed output
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

cudaError_t dcb_launch_cmp(StreamDesc *d_streams, const uint32_t *d_order, uint32_t n, int ncp, uint32_t max_entries,
                           uint32_t dump, const DevArenas &a, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const uint32_t gx = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)max_entries + 127) / 128, 1u << 20));
  cmp_deps_kernel<<<dim3(gx, n > 65535u ? 65535u : n), 128, 0, st>>>(d_streams, d_order, n, a.maps, a.aux);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  cmp_flags_kernel<<<n, 128, 0, st>>>(a.in, d_streams, d_order, n, a.aux);
  e = cudaGetLastError();
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
