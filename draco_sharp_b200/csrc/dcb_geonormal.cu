// draco_sharp_b200/csrc/dcb_geonormal.cu -- MeshPredictionSchemeGeometricNormalDecoder on the GPU.
//
// Reference: D/IO/Attributes/PredictionSchemes/MeshPredictionSchemeGeometricNormalDecoder.cs:44-82 over
// MeshPredictionSchemeGeometricNormalPredictorArea.cs:11-60 (TriangleArea mode, the only one v2.2 streams use),
// OctahedronToolBox.cs:28-77,121-137 and the octahedron transforms, in the bitstream's semantics where the C# is
// defective (the corner iterator skips its first corner, AbsSum takes no absolute values, the
// predictor returns (n0, n1, n0), CanonicalizeIntegerVector multiplies in 32 bits).
//
// The predicted normal of an entry is the sum of the cross products of the triangles around the entry's vertex in
// POSITION space (the quantized positions the parallelogram kernels left in the parent stream's scratch); it depends on
// no other normal.  Unlike every other prediction scheme of the format the reconstruction is therefore point-parallel:
//   geo_flips_kernel   one rABS-coded flip bit per entry, one warp per stream (the only serial piece)
//   geo_normal_kernel  one thread per entry: swing around the vertex, 64-bit cross products (wrapping sums), scale below
//                      2^29, canonicalise to an L1 norm of the centre value, flip, octahedral coordinates, the
//                      octahedron transform with the entry's correction, unit vector, 12-byte store
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
constexpr long long kI64Max = 0x7FFFFFFFFFFFFFFFll;

__device__ __forceinline__ uint32_t c_next(uint32_t c) { return c == kInv ? c : ((c % 3u == 2u) ? c - 2u : c + 1u); }
__device__ __forceinline__ uint32_t c_prev(uint32_t c) { return c == kInv ? c : ((c % 3u == 0u) ? c + 2u : c - 1u); }

// scratch of a stream: corrections / decoded (s, t) pairs int32[2 n] | flip bits u8[n]
__global__ void __launch_bounds__(32) geo_flips_kernel(const uint8_t *__restrict__ arena, StreamDesc *streams,
                                                       const uint32_t *__restrict__ order, uint32_t n_streams,
                                                       uint8_t *__restrict__ aux) {
  __shared__ uint8_t s_win[kRabsWin];
  __shared__ uint8_t s_bits[kRabsWin];
  for (uint32_t si = blockIdx.x; si < n_streams; si += gridDim.x) {
    const StreamDesc &d = streams[order[si]];
    if (d.status != DCB_OK || d.n_entries == 0) continue;
    rabs_decode_block(arena + d.orient_off, d.n_entries, aux + d.aux_off + 8ull * d.n_entries, s_win, s_bits, threadIdx.x);
  }
}

// Vector<long>.AbsSum (D/IO/Core/Vector.cs:211-226) with the absolute values the C# forgets: saturates at INT64_MAX
__device__ __forceinline__ long long abs_sum3_sat(const long long v[3]) {
  long long r = 0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    if (v[i] == (-kI64Max - 1)) return kI64Max;
    const long long a = v[i] < 0 ? -v[i] : v[i];
    if (r > kI64Max - a) return kI64Max;
    r += a;
  }
  return r;
}

template <bool DUMP>
__global__ void geo_normal_kernel(StreamDesc *streams, const uint32_t *__restrict__ order, uint32_t n_streams,
                                  const uint8_t *__restrict__ maps, uint8_t *__restrict__ out, uint8_t *__restrict__ dbg,
                                  uint8_t *__restrict__ aux, uint32_t dump) {
  for (uint32_t si = blockIdx.y; si < n_streams; si += gridDim.y) {
    StreamDesc &d = streams[order[si]];
    if (d.status != DCB_OK) continue;
    const uint32_t n = d.n_entries;
    if (n == 0) continue;
    // the parent: the buffer's position attribute, three portable components, decoded by a mesh chain kernel (its
    // quantized ints sit in its scratch).  Anything else is outside this path.
    if (d.parent < 0) {
      if (blockIdx.x == 0 && threadIdx.x == 0) d.status = DCB_ERR_PRED;
      continue;
    }
    const StreamDesc &pa = streams[d.parent];
    if (pa.status != DCB_OK || pa.ncp != 3 || pa.recon != RECON_PARA_WRAP || pa.attr_index >= d.attr_index || !pa.has_maps ||
        d.ncp != 2) {
      if (blockIdx.x == 0 && threadIdx.x == 0) d.status = pa.status != DCB_OK ? pa.status : DCB_ERR_UNSUPPORTED;
      continue;
    }
    const uint32_t *opp = reinterpret_cast<const uint32_t *>(maps + d.map_off[0]);
    const uint32_t *d2c = reinterpret_cast<const uint32_t *>(maps + d.map_off[2]);
    const uint32_t *p_c2v = reinterpret_cast<const uint32_t *>(maps + pa.map_off[1]);
    const int32_t *p_v2d = reinterpret_cast<const int32_t *>(maps + pa.map_off[3]);
    const uint32_t n_corners = d.n_corners, pn_corners = pa.n_corners, pn_vertices = pa.n_vertices, n_pos = pa.n_entries;
    const int32_t *pos_q = reinterpret_cast<const int32_t *>(aux + pa.aux_off) + 3ull * n_pos;
    int32_t *st_io = reinterpret_cast<int32_t *>(aux + d.aux_off);  // corrections in, decoded (s, t) out
    const uint8_t *flips = aux + d.aux_off + 8ull * n;
    uint8_t *optr = out + d.out_off;
    int32_t *dptr = DUMP ? reinterpret_cast<int32_t *>(dbg + d.dbg_off) : nullptr;
    PostParams pp;
    pp.load(d);
    OctBox box;
    box.set(32 - __clz(d.xf_a));  // bits = msb(max_q) + 1 (PredictionSchemeNormalOctahedronTransform.cs:44-53)
    const bool canonical = d.recon == RECON_GEO_OCT_CANON;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
      bool bad = false;
      auto pos_of = [&](uint32_t c, long long P[3]) {  // GetPositionForCorner (...GeometricNormalPredictor.cs:27-32)
        uint32_t v = kInv;
        int32_t e = -1;
        if (c < pn_corners) v = p_c2v[c];
        if (v < pn_vertices) e = p_v2d[v];
        if (e < 0 || (uint32_t)e >= n_pos) {
          bad = true;
          P[0] = P[1] = P[2] = 0;
        } else {
#pragma unroll
          for (int k = 0; k < 3; ++k) P[k] = pos_q[3ull * (uint32_t)e + k];
        }
      };
      auto opposite = [&](uint32_t c) -> uint32_t {
        if (c == kInv) return c;
        if (c >= n_corners) { bad = true; return kInv; }
        return opp[c];
      };
      const uint32_t start = d2c[p];
      unsigned long long nrm[3] = {0, 0, 0};
      if (start >= n_corners) {
        bad = true;
      } else {
        long long cent[3];
        pos_of(start, cent);
        uint32_t corner = start, guard = 0;
        bool left = true;
        while (corner != kInv && !bad) {  // VertexCornersIterator (D/IO/Mesh/VertexCornersIterator.cs:19-44), start included
          if (++guard > n_corners + 2u) { bad = true; break; }
          long long nx[3], pv[3], dn[3], dp[3];
          pos_of(c_next(corner), nx);
          pos_of(c_prev(corner), pv);
#pragma unroll
          for (int k = 0; k < 3; ++k) { dn[k] = nx[k] - cent[k]; dp[k] = pv[k] - cent[k]; }
          // CrossProduct, summed as unsigned (...PredictorArea.cs:31-36)
          nrm[0] += (unsigned long long)dn[1] * (unsigned long long)dp[2] - (unsigned long long)dn[2] * (unsigned long long)dp[1];
          nrm[1] += (unsigned long long)dn[2] * (unsigned long long)dp[0] - (unsigned long long)dn[0] * (unsigned long long)dp[2];
          nrm[2] += (unsigned long long)dn[0] * (unsigned long long)dp[1] - (unsigned long long)dn[1] * (unsigned long long)dp[0];
          if (left) {
            corner = c_next(opposite(c_next(corner)));  // SwingLeft
            if (corner == kInv) {
              corner = c_prev(opposite(c_prev(start)));  // open fan: continue to the right of the start
              left = false;
            } else if (corner == start) {
              corner = kInv;
            }
          } else {
            corner = c_prev(opposite(c_prev(corner)));  // SwingRight
          }
        }
      }
      if (bad) {
        d.status = DCB_ERR_MAPS;
        continue;
      }
      long long nv[3] = {(long long)nrm[0], (long long)nrm[1], (long long)nrm[2]};
      const long long upper = 1ll << 29;  // :38
      const long long abs_sum = abs_sum3_sat(nv);
      if (abs_sum > upper) {              // :49-53
        const long long q = abs_sum / upper;
#pragma unroll
        for (int k = 0; k < 3; ++k) nv[k] /= q;
      }
      int32_t v[3] = {(int32_t)nv[0], (int32_t)nv[1], (int32_t)nv[2]};
      {  // OctahedronToolBox.CanonicalizeIntegerVector (:121-137), 64-bit products
        const long long l1 = (long long)abs32(v[0]) + (long long)abs32(v[1]) + (long long)abs32(v[2]);
        if (l1 == 0) {
          v[0] = box.center;
        } else {
          v[0] = (int32_t)(((long long)v[0] * (long long)box.center) / l1);
          v[1] = (int32_t)(((long long)v[1] * (long long)box.center) / l1);
          const int32_t rest = box.center - abs32(v[0]) - abs32(v[1]);
          v[2] = v[2] >= 0 ? rest : -rest;
        }
      }
      if (flips[p]) {  // ...GeometricNormalDecoder.cs:58-61
#pragma unroll
        for (int k = 0; k < 3; ++k) v[k] = neg32(v[k]);
      }
      int32_t s, t;  // OctahedronToolBox.IntegerVectorToQuantizedOctahedralCoords (:61-77)
      if (v[0] >= 0) {
        s = v[1] + box.center;
        t = v[2] + box.center;
      } else {
        s = v[1] < 0 ? abs32(v[2]) : box.max_value - abs32(v[2]);
        t = v[2] < 0 ? abs32(v[1]) : box.max_value - abs32(v[1]);
      }
      {  // CanonicalizeOctahedralCoords (:28-54)
        const int32_t mv = box.max_value, ce = box.center;
        if ((s == 0 && t == 0) || (s == 0 && t == mv) || (s == mv && t == 0)) { s = mv; t = mv; }
        else if (s == 0 && t > ce) t = ce - (t - ce);
        else if (s == mv && t < ce) t = ce + (ce - t);
        else if (t == mv && s < ce) s = ce + (ce - s);
        else if (t == 0 && s > ce) s = ce - (s - ce);
      }
      const int2 co = reinterpret_cast<const int2 *>(st_io)[p];
      oct_original(box, canonical, s, t, co.x, co.y);  // ...GeometricNormalDecoder.cs:65
      reinterpret_cast<int2 *>(st_io)[p] = make_int2(s, t);
      if (DUMP && (dump & DCB_DUMP_QINTS)) {
        dptr[2ull * p] = s;
        dptr[2ull * p + 1] = t;
      }
      const int32_t o[2] = {s, t};
      store_entry<2>(pp, pp.store, pp.dsize, optr, p, o);
    }
  }
}

}  // namespace

cudaError_t dcb_launch_geo_flips(StreamDesc *d_streams, const uint32_t *d_order, uint32_t n, const DevArenas &a, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  geo_flips_kernel<<<n, 32, 0, st>>>(a.in, d_streams, d_order, n, a.aux);
  return cudaGetLastError();
}

cudaError_t dcb_launch_geo_normal(StreamDesc *d_streams, const uint32_t *d_order, uint32_t n, uint32_t max_entries, uint32_t dump,
                                  const DevArenas &a, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const uint32_t gx = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)max_entries + 127) / 128, 1u << 20));
  const dim3 grid(gx, n > 65535u ? 65535u : n);
  if (dump)
    geo_normal_kernel<true><<<grid, 128, 0, st>>>(d_streams, d_order, n, a.maps, a.out, a.dbg, a.aux, dump);
  else
    geo_normal_kernel<false><<<grid, 128, 0, st>>>(d_streams, d_order, n, a.maps, a.out, a.dbg, a.aux, dump);
  return cudaGetLastError();
}
