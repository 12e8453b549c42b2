// draco_sharp_b200/csrc/dcb_rabs.cuh -- one rABS-coded bit sequence, decoded by a warp.
//
// Reference: RAnsBitDecoder.StartDecoding / DecodeNextBit (D/IO/BitCoders/RAnsBitDecoder.cs:12-30) over
// AnsDecoder.ReadInit / RAbsRead (D/IO/Entropy/AnsDecoder.cs:20-56).  The block (u8 prob_zero | varint size | data) has
// been located and validated by the container walk (walk_rabs_block, dcb_walk.h).  Lane 0 runs the serial state
// recurrence over a 256-byte shared-memory window of the (backwards-read) data that the warp refills, and leaves up to
// 256 decoded bits per round in shared memory; the warp writes them to `dst` (one byte per bit) coalesced.
#pragma once
#include <stdint.h>

namespace dcb {

constexpr uint32_t kRabsWin = 256;

// s_win, s_bits: kRabsWin bytes each, private to the calling warp.  Every lane of the warp must call.
__device__ __forceinline__ void rabs_decode_block(const uint8_t *__restrict__ blk, uint32_t want, uint8_t *__restrict__ dst,
                                                  uint8_t *s_win, uint8_t *s_bits, uint32_t lane) {
  if (want == 0) return;
  const uint32_t prob_zero = blk[0];
  uint64_t pos = 1, nb = 0;
  for (int i = 0, shift = 0; i < 10; ++i, shift += 7) {  // varint size
    const uint32_t b = blk[pos++];
    nb |= (uint64_t)(b & 0x7Fu) << shift;
    if (!(b & 0x80u)) break;
  }
  const uint8_t *data = blk + pos;
  const uint32_t p1 = (256u - prob_zero) & 0xFFu;
  const uint32_t x = (uint32_t)data[nb - 1] >> 6;  // AnsDecoder.ReadInit: the last byte says how many bytes hold the state
  int64_t off = (int64_t)nb - 1 - x;
  uint32_t state = 0;
  for (uint32_t i = 0; i <= x; ++i) state |= (uint32_t)data[nb - 1 - x + i] << (8 * i);
  state &= (x == 0) ? 0x3Fu : (x == 1) ? 0x3FFFu : 0x3FFFFFu;
  state += 4096u;
  uint32_t done = 0;
  while (done < want) {  // uniform over the warp
    const int64_t lo = off > (int64_t)kRabsWin ? off - (int64_t)kRabsWin : 0;
    for (uint32_t i = lane; i < (uint32_t)(off - lo); i += 32) s_win[i] = data[lo + i];
    __syncwarp();
    uint32_t made = 0;
    if (lane == 0) {
      const uint32_t room = min(kRabsWin, want - done);
      while (made < room) {
        if (state < 4096u && off > 0) {  // RAbsRead: one byte of renormalisation at most
          if (off <= lo) break;          // window used up: refill
          state = state * 256u + s_win[--off - lo];
        }
        const uint32_t quot = state >> 8, rem = state & 255u, xn = quot * p1;
        const bool val = rem < p1;
        state = val ? xn + rem : state - xn - p1;
        s_bits[made++] = val ? 1 : 0;
      }
    }
    made = __shfl_sync(0xffffffffu, made, 0);
    off = __shfl_sync(0xffffffffu, off, 0);
    state = __shfl_sync(0xffffffffu, state, 0);
    __syncwarp();
    for (uint32_t i = lane; i < made; i += 32) dst[done + i] = s_bits[i];
    done += made;
    __syncwarp();
  }
}

}  // namespace dcb
