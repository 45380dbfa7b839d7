// pfv_kernels_sb.cu — "register-resident sub-block" kernels (sm_100a).
//
// One THREAD owns one 8x8 sub-block: its 64 coefficients live in registers, both 1-D passes run without any
// exchange, the zig-zag permutation is a compile-time register renaming and the (de)quantiser tables are
// kernel parameters, i.e. constant-bank operands of the multiplies.  A warp covers 8 consecutive macroblocks
// (lane = mb*4 + sub-block), so one 8-byte row store of the warp forms two full 128-byte lines of the plane.
// Compared with the warp-per-macroblock kernels of pfv_kernels.cu this removes every shared-memory round trip
// and amortises the addressing over 8 macroblocks (ncu: 255 -> ~150 warp instructions per macroblock).
//
// Arithmetic is the reference's (see pfv_device.cuh for the file:line map).
#include "pfv_internal.h"
#include "pfv_device.cuh"

namespace pfv {

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

constexpr int SB_WARPS = 4;                 // 4 warps = 32 macroblocks per CTA, never straddling a plane

__global__ void __launch_bounds__(SB_WARPS * 32, 4)
decode_i_sb_kernel(const __grid_constant__ SbParams P, const DecJob *__restrict__ jobs)
{
    constexpr int zz[64] = PFV_ZIGZAG_INIT;
    // plane from the CTA index alone: everything derived from it stays in the uniform datapath
    const uint32_t cta = blockIdx.x;
    const int p = (cta >= P.cta_base[1] ? 1 : 0) + (cta >= P.cta_base[2] ? 1 : 0);
    const PlaneGeom &pl = p == 0 ? P.g.pl[0] : (p == 1 ? P.g.pl[1] : P.g.pl[2]);
    const uint32_t lm = (cta - (p == 0 ? P.cta_base[0] : (p == 1 ? P.cta_base[1] : P.cta_base[2]))) * (SB_WARPS * 8) +
                        (threadIdx.x >> 2);                  // macroblock inside the plane
    if (lm >= pl.bw * pl.bh) return;
    const int sb = threadIdx.x & 3;
    const DecJob job = jobs[blockIdx.y];

    const uint4 *src = reinterpret_cast<const uint4 *>(job.coeff + ((size_t)(pl.mb_base + lm) * 256 + sb * 64));
    uint4 raw[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) raw[k] = __ldcs(src + k);

    const int32_t *deq = p == 0 ? P.deq[0] : (p == 1 ? P.deq[1] : P.deq[2]);
    int m[64];
#pragma unroll
    for (int s = 0; s < 64; ++s) {
        const uint4 &q = raw[s >> 3];
        const uint32_t w = ((s >> 1) & 3) == 0 ? q.x : ((s >> 1) & 3) == 1 ? q.y : ((s >> 1) & 3) == 2 ? q.z : q.w;
        const int c = (s & 1) ? ((int)w >> 16) : (int)(int16_t)(w & 0xffffu);
        m[zz[s]] = c * deq[s];                               // src/dct.rs:78-83 (tables by scan position)
    }
    idct8x8_regs(m);

    uint32_t col;
    const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
    uint8_t *dst = job.dst + pl.off + (size_t)(row * 16u + (uint32_t)(sb >> 1) * 8u) * pl.pw + col * 16u + (uint32_t)(sb & 1) * 8u;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        uint2 o;
        o.x = pack4_sat_u8(m[r * 8 + 0], m[r * 8 + 1], m[r * 8 + 2], m[r * 8 + 3]);
        o.y = pack4_sat_u8(m[r * 8 + 4], m[r * 8 + 5], m[r * 8 + 6], m[r * 8 + 7]);
        __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pl.pw), o);
    }
}

cudaError_t launch_decode_i_sb(const SbParams &P, const DecJob *d_jobs, uint32_t njobs, cudaStream_t s)
{
    dim3 grid(P.cta_total, njobs, 1), block(SB_WARPS * 32, 1, 1);
    decode_i_sb_kernel<<<grid, block, 0, s>>>(P, d_jobs);
    return cudaGetLastError();
}

}  // namespace pfv
