// draco_sharp_b200/csrc/dcb_texcoord.cu -- MeshPredictionSchemeTexCoordsPortableDecoder on the GPU.
//
// Reference: D/IO/Attributes/PredictionSchemes/MeshPredictionSchemeTexCoordsPortableDecoder.cs:49-84 (the loop over
// entries, the orientation flags) over MeshPredictionSchemeTexCoordsPortablePredictor.cs:33-151 (the prediction) and
// PredictionSchemeWrapDecodingTransform.cs:46-67.  The predictor projects the tip of the entry's triangle onto the
// edge (next, prev) in POSITION space -- the quantized positions of the buffer's position attribute, which the
// parallelogram kernels have left in that stream's scratch -- and carries the result over to UV space; one rABS-coded
// flag per projected entry picks the side of the edge.
//
// Everything that depends only on the maps and on the decoded positions is point-parallel (tex_prep_kernel): the two
// operand entries, the 64-bit edge length / dot product / the truncating projection and IntSqrt (MathUtilities.cs:5-25).
// What is left is a serial chain per stream over (u, v) pairs (tex_chain_kernel): two 64-bit multiply-adds, two
// truncating 64-bit divisions and the wrap per entry, one warp per stream -- lane 0 walks the chain out of shared
// memory, the other lanes stage records, operands and flags of the next 32 entries and write the finished block.
// Product code: nothing here touches oracle/.
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>

#include "dcb_device.cuh"
#include "dcb_internal.h"
#include "dcb_kernels.h"

using namespace dcb;

namespace {

// per entry, written by tex_prep_kernel behind the stream's corrections and quantized ints
struct __align__(16) TexRec {
  int32_t nd, pd;     // VertexToData of the next / previous corner's vertex (:58-59); may be >= p or negative
  uint64_t pn2;       // |prev - next|^2 in position space (:75)
  int64_t cdp;        // (prev - next) . (tip - next) (:76)
  uint32_t nrm;       // IntSqrt(|tip - projection|^2 * pn2) (:94)
  uint32_t flag;      // 0 ok | 1 a position lookup of (tip, next, prev) is out of range | 2 the overflow guard :90 fails
};
static_assert(sizeof(TexRec) == 32, "TexRec is 32 bytes (scratch layout in dcb_api.cu)");

__device__ __forceinline__ uint64_t abs64(int64_t v) { return v < 0 ? 0ull - (uint64_t)v : (uint64_t)v; }

// MathUtilities.IntSqrt (D/IO/Core/MathUtilities.cs:5-25), same iteration
__device__ uint64_t int_sqrt(uint64_t number) {
  if (number == 0) return 0;
  uint64_t act = number, root = 1;
  while (act >= 2) {
    root *= 2;
    act /= 4;
  }
  do {
    root = (root + number / root) / 2;
  } while (root * root > number);
  return root;
}

constexpr int64_t kI64Max = 0x7FFFFFFFFFFFFFFFll;

__global__ void tex_prep_kernel(StreamDesc *streams, const uint32_t *__restrict__ order, uint32_t n_streams,
                                const uint8_t *__restrict__ maps, uint8_t *__restrict__ aux) {
  for (uint32_t si = blockIdx.y; si < n_streams; si += gridDim.y) {
    StreamDesc &d = streams[order[si]];
    if (d.status != DCB_OK) continue;
    const uint32_t n = d.n_entries;
    // the parent: the buffer's position attribute, three portable components, decoded by the parallelogram kernels
    // (their quantized ints sit in its scratch).  Anything else is outside this path.
    if (d.parent < 0) {
      if (blockIdx.x == 0 && threadIdx.x == 0) d.status = DCB_ERR_PRED;
      continue;
    }
    const StreamDesc &pa = streams[d.parent];
    if (pa.status != DCB_OK || pa.ncp != 3 || pa.recon != RECON_PARA_WRAP || (pa.pred_method != PRED_PARALLELOGRAM && pa.pred_method != PRED_CONSTRAINED_MULTI) ||
        pa.attr_index >= d.attr_index || !pa.has_maps) {
      if (blockIdx.x == 0 && threadIdx.x == 0) d.status = pa.status != DCB_OK ? pa.status : DCB_ERR_UNSUPPORTED;
      continue;
    }
    const uint32_t *c2v = reinterpret_cast<const uint32_t *>(maps + d.map_off[1]);
    const uint32_t *d2c = reinterpret_cast<const uint32_t *>(maps + d.map_off[2]);
    const int32_t *v2d = reinterpret_cast<const int32_t *>(maps + d.map_off[3]);
    const uint32_t *p_c2v = reinterpret_cast<const uint32_t *>(maps + pa.map_off[1]);
    const int32_t *p_v2d = reinterpret_cast<const int32_t *>(maps + pa.map_off[3]);
    const uint32_t n_corners = d.n_corners, n_vertices = d.n_vertices;
    const uint32_t pn_corners = pa.n_corners, pn_vertices = pa.n_vertices, n_pos = pa.n_entries;
    const int32_t *pos_q = reinterpret_cast<const int32_t *>(aux + pa.aux_off) + 3ull * n_pos;
    TexRec *recs = reinterpret_cast<TexRec *>(aux + d.aux_off + 16ull * n);  // behind corr int32[2n] | qints int32[2n]
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
      TexRec r;
      r.nd = r.pd = 0x7FFFFFFF;
      r.pn2 = 0;
      r.cdp = 0;
      r.nrm = 0;
      r.flag = 0;
      const uint32_t corner = d2c[p];
      if (corner == 0xFFFFFFFFu || corner >= n_corners) {
        d.status = DCB_ERR_MAPS;
      } else {
        const uint32_t nx = (corner % 3u == 2u) ? corner - 2u : corner + 1u;
        const uint32_t pv = (corner % 3u == 0u) ? corner + 2u : corner - 1u;
        const uint32_t v_n = c2v[nx], v_p = c2v[pv];
        if (v_n >= n_vertices || v_p >= n_vertices) {
          d.status = DCB_ERR_MAPS;
        } else {
          r.nd = v2d[v_n];
          r.pd = v2d[v_p];
          if (r.pd < (int32_t)p && r.nd < (int32_t)p) {
            if (r.pd < 0 || r.nd < 0) {
              d.status = DCB_ERR_MAPS;
            } else {
              // GetPositionForEntryId (:34-39): entry -> corner it was first reached at -> position value
              int64_t P[3][3];
              const uint32_t ids[3] = {p, (uint32_t)r.nd, (uint32_t)r.pd};
              bool ok = true;
#pragma unroll
              for (int k = 0; k < 3; ++k) {
                const uint32_t c = d2c[ids[k]];
                uint32_t v = 0xFFFFFFFFu;
                int32_t e = -1;
                if (c < pn_corners) v = p_c2v[c];
                if (v < pn_vertices) e = p_v2d[v];
                if (e < 0 || (uint32_t)e >= n_pos) {
                  ok = false;
#pragma unroll
                  for (int j = 0; j < 3; ++j) P[k][j] = 0;
                } else {
#pragma unroll
                  for (int j = 0; j < 3; ++j) P[k][j] = pos_q[3ull * (uint32_t)e + j];
                }
              }
              if (!ok) {
                r.flag = 1;
              } else {
                int64_t pn[3], cn[3];
                uint64_t pn2 = 0, cdp = 0;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                  pn[j] = P[2][j] - P[1][j];
                  cn[j] = P[0][j] - P[1][j];
                  pn2 += (uint64_t)pn[j] * (uint64_t)pn[j];
                  cdp += (uint64_t)pn[j] * (uint64_t)cn[j];
                }
                r.pn2 = pn2;
                r.cdp = (int64_t)cdp;
                if ((int64_t)pn2 != 0) {
                  uint64_t m = 0;
#pragma unroll
                  for (int j = 0; j < 3; ++j) m = max(m, abs64(pn[j]));
                  if (m == 0 || r.cdp > kI64Max / (int64_t)m) {
                    r.flag = 2;
                  } else {
                    uint64_t cx2 = 0;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                      const int64_t xp = P[1][j] + (int64_t)((uint64_t)r.cdp * (uint64_t)pn[j]) / (int64_t)pn2;  // :91, truncating
                      const int64_t dlt = P[0][j] - xp;
                      cx2 += (uint64_t)dlt * (uint64_t)dlt;
                    }
                    r.nrm = (uint32_t)int_sqrt(cx2 * pn2);
                  }
                }
              }
            }
          }
        }
      }
      recs[p] = r;
    }
  }
}

// rABS bit decoder (AnsDecoder.RAbsRead through RAnsBitDecoder.DecodeNextBit, D/IO/BitCoders/RAnsBitDecoder.cs:26-30),
// run by lane 0 over a 256-byte shared-memory window the warp refills.
struct RabsLane {
  uint32_t state, p;
  int64_t off;  // bytes of the block not consumed yet
};

constexpr uint32_t kTexBlock = 32;
constexpr uint32_t kWin = 256;

template <bool DUMP>
__global__ void __launch_bounds__(32) tex_chain_kernel(const uint8_t *__restrict__ arena, StreamDesc *streams,
                                                       const uint32_t *__restrict__ order, uint32_t n_streams,
                                                       uint8_t *__restrict__ out, uint8_t *__restrict__ dbg,
                                                       uint8_t *__restrict__ aux, uint32_t dump) {
  __shared__ TexRec s_rec[kTexBlock];
  __shared__ int2 s_nuv[kTexBlock], s_puv[kTexBlock], s_fb[kTexBlock], s_out[kTexBlock];
  __shared__ int2 s_cor[kTexBlock];
  __shared__ uint8_t s_flag[kTexBlock];
  __shared__ uint8_t s_win[kWin];
  const uint32_t lane = threadIdx.x;
  for (uint32_t si = blockIdx.x; si < n_streams; si += gridDim.x) {
    StreamDesc &d = streams[order[si]];
    if (d.status != DCB_OK) continue;
    const uint32_t n = d.n_entries;
    if (n == 0) continue;
    PostParams pp;
    pp.load(d);
    uint8_t *optr = out + d.out_off;
    int32_t *dptr = DUMP ? reinterpret_cast<int32_t *>(dbg + d.dbg_off) : nullptr;
    const int2 *corr = reinterpret_cast<const int2 *>(aux + d.aux_off);
    int2 *qints = reinterpret_cast<int2 *>(aux + d.aux_off) + n;
    const TexRec *recs = reinterpret_cast<const TexRec *>(aux + d.aux_off + 16ull * n);
    uint8_t *flags = aux + d.aux_off + 48ull * n;  // u8[n_orient <= 2n + 1]: Orientations, in decoding order
    const uint32_t n_or = d.n_orient;

    // ---- orientation flags: rABS bits, then the running "flip on 0" of DecodePredictionData (:76-82) ----
    {
      const uint8_t *blk = arena + d.orient_off;
      const uint32_t prob_zero = blk[0];
      uint64_t pos = 1, nb = 0;
      for (int i = 0, shift = 0; i < 10; ++i, shift += 7) {  // varint size (validated by the container walk)
        const uint32_t b = blk[pos++];
        nb |= (uint64_t)(b & 0x7Fu) << shift;
        if (!(b & 0x80u)) break;
      }
      const uint8_t *data = blk + pos;
      RabsLane a;
      a.p = (256u - prob_zero) & 0xFFu;
      const uint32_t x = (uint32_t)data[nb - 1] >> 6;
      a.off = (int64_t)nb - 1 - x;
      a.state = 0;
      for (uint32_t i = 0; i <= x; ++i) a.state |= (uint32_t)data[nb - 1 - x + i] << (8 * i);
      a.state &= (x == 0) ? 0x3Fu : (x == 1) ? 0x3FFFu : 0x3FFFFFu;
      a.state += 4096u;
      uint32_t done = 0;
      int last = 1;
      while (done < n_or) {  // uniform over the warp
        // window = the kWin bytes in front of `off`
        const int64_t lo = a.off > (int64_t)kWin ? a.off - (int64_t)kWin : 0;
        for (uint32_t i = lane; i < (uint32_t)(a.off - lo); i += 32) s_win[i] = data[lo + i];
        __syncwarp();
        if (lane == 0) {
          while (done < n_or) {
            if (a.state < 4096u && a.off > 0) {
              if (a.off <= lo) break;  // refill
              a.state = a.state * 256u + s_win[--a.off - lo];
            }
            const uint32_t xx = a.state, quot = xx >> 8, rem = xx & 255u, xn = quot * a.p;
            const bool val = rem < a.p;
            a.state = val ? xn + rem : xx - xn - a.p;
            if (!val) last = !last;
            flags[done++] = (uint8_t)last;
          }
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        a.off = __shfl_sync(0xffffffffu, a.off, 0);
        a.state = __shfl_sync(0xffffffffu, a.state, 0);
        __syncwarp();
      }
    }
    __syncwarp();

    uint32_t left = n_or;  // flags not popped yet (the predictor pops from the back, :126-127)
    int status = DCB_OK;
    for (uint32_t e0 = 0; e0 < n && status == DCB_OK; e0 += kTexBlock) {
      const uint32_t cnt = min(kTexBlock, n - e0);
      // ---- stage: records, corrections, operands decoded in earlier blocks, fall-back operands, flags ----
      if (lane < cnt) {
        const uint32_t p = e0 + lane;
        const TexRec r = recs[p];
        s_rec[lane] = r;
        s_cor[lane] = corr[p];
        int2 nuv = make_int2(0, 0), puv = make_int2(0, 0), fb = make_int2(0, 0);
        if (r.nd >= 0 && (uint32_t)r.nd < e0) nuv = qints[r.nd];
        if (r.pd >= 0 && (uint32_t)r.pd < e0) puv = qints[r.pd];
        // fall-back operand (:135-160) when it is older than this block
        int64_t off = -1;
        if (r.pd < (int32_t)p) off = r.pd;
        if (r.nd < (int32_t)p) off = r.nd;
        else if (p > 0) off = (int64_t)p - 1;
        if (off >= 0 && (uint64_t)off < e0) fb = qints[off];
        s_nuv[lane] = nuv;
        s_puv[lane] = puv;
        s_fb[lane] = fb;
      }
      {
        const uint32_t take = min(left, kTexBlock);  // flags [left - take, left)
        if (lane < take) s_flag[lane] = flags[left - take + lane];
      }
      __syncwarp();
      // ---- chain: lanes 0 and 1, one component each (they meet in shared memory after every entry) ----
      uint32_t used = 0;
      if (lane < 2) {
        const uint32_t take = min(left, kTexBlock);
        const uint32_t pair = 0x3u;
        for (uint32_t j = 0; j < cnt; ++j) {
          const uint32_t p = e0 + j;
          const TexRec r = s_rec[j];
          int64_t pred = 0;
          bool have = false;
          auto value_of = [&](int32_t e, const int2 &staged) -> int2 {
            return (uint32_t)e >= e0 ? s_out[(uint32_t)e - e0] : staged;
          };
          if (r.pd < (int32_t)p && r.nd < (int32_t)p) {
            const int2 nuv = value_of(r.nd, s_nuv[j]), puv = value_of(r.pd, s_puv[j]);
            if (puv.x == nuv.x && puv.y == nuv.y) {  // :66-71
              pred = lane ? puv.y : puv.x;
              have = true;
            } else if (r.flag == 1) {
              status = DCB_ERR_MAPS;
            } else if ((int64_t)r.pn2 != 0) {  // :78
              const uint64_t pn2 = r.pn2;
              const int64_t n0 = nuv.x, n1 = nuv.y;
              const int64_t d0 = (int64_t)puv.x - n0, d1 = (int64_t)puv.y - n1;
              const uint64_t amax = max(abs64(n0), abs64(n1)), bmax = max(abs64(d0), abs64(d1));
              // a > INT64_MAX / b  <=>  a * b > INT64_MAX  (a, b > 0): the guards :85, :87 without a division
              const bool g85 = (int64_t)pn2 < 0 ? (int64_t)amax > kI64Max / (int64_t)pn2  // wrapped edge length: as written
                                                : (__umul64hi(amax, pn2) != 0 || amax * pn2 > (uint64_t)kI64Max);
              const bool g87 = r.cdp > 0 && (__umul64hi((uint64_t)r.cdp, bmax) != 0 || (uint64_t)r.cdp * bmax > (uint64_t)kI64Max);
              if (g85 || g87 || r.flag == 2) {  // :85, :87, :90
                status = DCB_ERR_PRED;
              } else if (used >= take) {  // no flag left (:125)
                status = DCB_ERR_PRED;
              } else {
                const int64_t nc = lane ? n1 : n0, dc = lane ? d1 : d0;
                const int64_t x = (int64_t)((uint64_t)nc * pn2 + (uint64_t)r.cdp * (uint64_t)dc);
                const int64_t cx = lane ? (int64_t)(0ull - (uint64_t)d0 * (uint64_t)r.nrm) : (int64_t)((uint64_t)d1 * (uint64_t)r.nrm);
                const bool o = s_flag[take - 1 - used] != 0;  // Last() + PopBack()
                ++used;
                pred = (o ? (int64_t)((uint64_t)x + (uint64_t)cx) : (int64_t)((uint64_t)x - (uint64_t)cx)) / (int64_t)pn2;  // :128
                have = true;
              }
            }
          }
          if (status != DCB_OK) break;  // both lanes see the same operands: the same verdict
          if (!have) {  // :135-160
            int64_t off = -1;
            bool any = true;
            if (r.pd < (int32_t)p) off = r.pd;
            if (r.nd < (int32_t)p) off = r.nd;
            else if (p > 0) off = (int64_t)p - 1;
            else any = false;
            if (any) {
              if (off < 0) {
                status = DCB_ERR_MAPS;
                break;
              }
              const int2 f = (uint64_t)off >= e0 ? s_out[(uint32_t)off - e0] : s_fb[j];
              pred = lane ? f.y : f.x;
            }
          }
          const int2 co = s_cor[j];
          const int32_t o = wrap_original((int32_t)pred, lane ? co.y : co.x, pp.mn, pp.mx, pp.max_diff);
          reinterpret_cast<int32_t *>(&s_out[j])[lane] = o;
          __syncwarp(pair);
        }
      }
      status = __shfl_sync(0xffffffffu, status, 0);
      used = __shfl_sync(0xffffffffu, used, 0);
      left -= used;
      __syncwarp();
      if (status != DCB_OK) break;
      // ---- finished block: quantized ints (later blocks gather from them), dump, typed output ----
      if (lane < cnt) {
        const uint32_t p = e0 + lane;
        const int2 o = s_out[lane];
        qints[p] = o;
        if (DUMP && (dump & DCB_DUMP_QINTS)) {
          dptr[2ull * p] = o.x;
          dptr[2ull * p + 1] = o.y;
        }
        const int32_t v[2] = {o.x, o.y};
        store_entry<2>(pp, pp.store, pp.dsize, optr, p, v);
      }
      __syncwarp();
    }
    if (status != DCB_OK && lane == 0) d.status = status;
  }
}

}  // namespace

cudaError_t dcb_launch_tex(StreamDesc *d_streams, const uint32_t *d_order, uint32_t n, uint32_t max_entries, uint32_t dump,
                           const DevArenas &a, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const uint32_t gx = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)max_entries + 127) / 128, 1u << 20));
  tex_prep_kernel<<<dim3(gx, n > 65535u ? 65535u : n), 128, 0, st>>>(d_streams, d_order, n, a.maps, a.aux);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const uint32_t grid = n;
  if (dump)
    tex_chain_kernel<true><<<grid, 32, 0, st>>>(a.in, d_streams, d_order, n, a.out, a.dbg, a.aux, dump);
  else
    tex_chain_kernel<false><<<grid, 32, 0, st>>>(a.in, d_streams, d_order, n, a.out, a.dbg, a.aux, dump);
  return cudaGetLastError();
}
