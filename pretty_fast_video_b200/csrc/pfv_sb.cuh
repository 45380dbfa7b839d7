// pfv_sb.cuh — pieces shared by the register-resident sub-block kernels (one thread = one 8x8 sub-block).
//
// The DC-only shortcut used by all of them: with v[1..7] = 0 the 1-D inverse transform (src/dct.rs:241-293)
// returns v[0] in every output (b0 = b1 = c0, every other term is 0 or 0/k), so a sub-block whose 63 AC
// coefficients are zero decodes to clamp((c0 * deq0 + 32768) >> 8) in all 64 pixels — exactly, in wrapping
// arithmetic, what the two full passes produce.
#pragma once
#include "pfv_device.cuh"

namespace pfv {

constexpr int SB_THREADS = 128;             // threads per CTA of the dense sub-block kernels (= SB_MBS_PER_CTA * 4)

// src/dct.rs:44-47 ZIGZAG_TABLE: raster index of scan position s (used only with compile-time indices)
#define PFV_ZIGZAG_INIT { \
     0,  1,  8, 16,  9,  2,  3, 10, 17, 24, 32, 25, 18, 11,  4,  5, \
    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,  6,  7, 14, 21, 28, \
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, \
    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 }

// columns then rows (src/common.rs:315-316), +128 folded into the DC input of each row (see decode_mb_core)
__device__ __forceinline__ void idct8x8_regs(int (&m)[64])
{
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        int v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = m[r * 8 + c];
        idct8(v);
#pragma unroll
        for (int r = 0; r < 8; ++r) m[r * 8 + c] = v[r];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = m[r * 8 + c];
        v[0] += 128 << 8;
        idct8(v);
#pragma unroll
        for (int c = 0; c < 8; ++c) m[r * 8 + c] = v[c] >> 8;
    }
}

// The two int16 halves of a word are sign-extended by IDP.2A (a 16x8-bit dot product with the constant bytes (1, 0)):
// it issues on the multiply pipe, where the transform kernels have slack, instead of PRMT/SHF on the ALU pipe.
__device__ __forceinline__ void unpack_dequant(const uint4 (&raw)[8], const int32_t *deq, int (&m)[64])
{
    constexpr int zz[64] = PFV_ZIGZAG_INIT;
#pragma unroll
    for (int s = 0; s < 64; ++s) {
        const uint4 &q = raw[s >> 3];
        const uint32_t w = ((s >> 1) & 3) == 0 ? q.x : ((s >> 1) & 3) == 1 ? q.y : ((s >> 1) & 3) == 2 ? q.z : q.w;
        const int c = (s & 1) ? __dp2a_hi((int)w, 0x01000000, 0) : __dp2a_lo((int)w, 0x00000001, 0);
        m[zz[s]] = c * deq[s];                               // src/dct.rs:78-83 (tables by scan position)
    }
}

// out = clamp(prev + delta) on four packed pixels; pos4/neg4 = max(delta,0) / min(max(-delta,0),255) in every byte
// (src/common.rs:100-102; delta is in [-256, 254], and prev - 255 already clamps to 0 for every prev)
__device__ __forceinline__ uint32_t add_delta_sat4(uint32_t prev, uint32_t pos4, uint32_t neg4)
{
    return __vsubus4(__vaddus4(prev, pos4), neg4);            // one of pos4/neg4 is zero
}

}  // namespace pfv
