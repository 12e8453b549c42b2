// Single-warp issue-rate probe (B200): how many cycles does ONE resident warp need per instruction, by pipe mix and by
// number of active lanes?  Answers whether the lane-per-stream rANS kernels (one warp per SM sub-partition, 17 of 32
// lanes active) are bound by pipe cadence, and whether a half-empty warp issues faster.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o issue_rate issue_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP 256
template <int MODE>
__global__ void probe(uint32_t *out, long long *cyc, uint32_t lanes, uint32_t seed) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t a0 = seed + lane, a1 = seed * 3 + lane, a2 = seed * 5 + lane, a3 = seed * 7 + lane;
  uint32_t b0 = a0 ^ 11, b1 = a1 ^ 13, b2 = a2 ^ 17, b3 = a3 ^ 19;
  long long t0 = 0, t1 = 0;
  __shared__ uint32_t chase[256];
  if (MODE == 6) {  // every word points at itself: a dependent LDS chain with a constant address per lane
    const uint32_t self = (uint32_t)__cvta_generic_to_shared(&chase[threadIdx.x]);
    chase[threadIdx.x] = self;
    a0 = self;
    __syncthreads();
  }
  if (lane < lanes) {
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
      for (int r = 0; r < REP / 8; ++r) {
        if (MODE == 0) {  // 8 independent alu-pipe ops (LOP3)
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b0), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a1) : "r"(b1), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(b2), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a3) : "r"(b3), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b0) : "r"(a1), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b1) : "r"(a2), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b2) : "r"(a3), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b3) : "r"(a0), "r"(seed));
        } else if (MODE == 1) {  // alternate alu-pipe (LOP3) and fma-pipe (IMAD)
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b0), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a1) : "r"(b1), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a2) : "r"(b2), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(b3), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b0) : "r"(a1), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b1) : "r"(a2), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b2) : "r"(a3), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b3) : "r"(a0), "r"(seed));
        } else if (MODE == 2) {  // 8 independent fma-pipe ops (IMAD)
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(b0), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a1) : "r"(b1), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a2) : "r"(b2), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a3) : "r"(b3), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b0) : "r"(a1), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b1) : "r"(a2), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b2) : "r"(a3), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b3) : "r"(a0), "r"(seed));
        } else if (MODE == 3) {  // dependent alu chain
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b0), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b1), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b2), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b3), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b0), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b1), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b2), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b3), "r"(seed));
        } else if (MODE == 4) {  // dependent alternating alu -> fma chain (cross-pipe)
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b0), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(b1), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b2), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(b3), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b0), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(b1), "r"(seed));
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a0) : "r"(b2), "r"(seed));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a0) : "r"(b3), "r"(seed));
        } else if (MODE == 5) {  // 8 independent 2-operand alu ops with an immediate (SHF by constant)
          asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(a0) : "r"(b0));
          asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(a1) : "r"(b1));
          asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(a2) : "r"(b2));
          asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(a3) : "r"(b3));
          asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(b0) : "r"(a1));
          asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(b1) : "r"(a2));
          asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(b2) : "r"(a3));
          asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(b3) : "r"(a0));
        } else if (MODE == 6) {  // dependent shared-memory pointer chase (LDS latency), address in a0
          asm volatile("ld.shared.u32 %0, [%0];" : "+r"(a0));
          asm volatile("ld.shared.u32 %0, [%0];" : "+r"(a0));
          asm volatile("ld.shared.u32 %0, [%0];" : "+r"(a0));
          asm volatile("ld.shared.u32 %0, [%0];" : "+r"(a0));
          asm volatile("ld.shared.u32 %0, [%0];" : "+r"(a0));
          asm volatile("ld.shared.u32 %0, [%0];" : "+r"(a0));
          asm volatile("ld.shared.u32 %0, [%0];" : "+r"(a0));
          asm volatile("ld.shared.u32 %0, [%0];" : "+r"(a0));
        } else if (MODE == 7) {  // setp + selp pairs (predicate round trip), dependent
          asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(a0) : "r"(b0), "r"(b1));
          asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(a0) : "r"(b1), "r"(b2));
          asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(a0) : "r"(b2), "r"(b3));
          asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p;}" : "+r"(a0) : "r"(b3), "r"(b0));
        }
      }
    }
    t1 = clock64();
  }
  out[threadIdx.x + blockIdx.x * blockDim.x] = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
  if (lane == 0 && blockIdx.x == 0 && threadIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
static void run(const char *name, int n_instr_per_rep8) {
  uint32_t *out;
  long long *cyc;
  cudaMalloc(&out, 4096 * 4);
  cudaMalloc(&cyc, 8);
  const uint32_t lane_cfg[] = {1, 8, 16, 17, 24, 32};
  for (int warps = 1; warps <= 8; warps *= 2) {  // warps per CTA (= 1..2 per sub-partition at 4, 8)
    for (uint32_t lanes : lane_cfg) {
      long long h = 0;
      for (int k = 0; k < 2; ++k) {
        probe<MODE><<<1, 32 * warps>>>(out, cyc, lanes, MODE == 6 ? 0u : 12345u);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      const double n = 64.0 * (REP / 8) * n_instr_per_rep8;
      printf("%-34s warps/CTA %d lanes %2u : %.2f cycles per warp-instruction\n", name, warps, lanes, (double)h / n);
    }
  }
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("independent LOP3 (alu)", 8);
  run<2>("independent IMAD (fma)", 8);
  run<1>("alternating LOP3/IMAD", 8);
  run<5>("independent SHF imm (alu)", 8);
  run<3>("dependent LOP3 chain", 8);
  run<4>("dependent LOP3->IMAD chain", 8);
  run<7>("dependent SETP+SEL pairs", 8);
  run<6>("dependent LDS chase (addr 0)", 8);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
