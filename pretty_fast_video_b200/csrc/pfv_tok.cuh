// pfv_tok.cuh — the run-length pass of rle_encode (src/rle.rs:9-39) over one macroblock held by one warp: a lane owns the 8
// consecutive coefficients 8*lane .. 8*lane+7 (the uint4 the encode kernels store).  Shared by the encode kernels (which count
// the entries of a macroblock while its coefficients are still in registers) and the tokenizer (pfv_kernels_tok.cu).
#pragma once
#include <stdint.h>

namespace pfv {
namespace tok {

constexpr unsigned FULL = 0xffffffffu;

// escapes a zero run of `run` costs before its final entry: `while run > 15 { push(15,0,0); run -= 15 }` (src/rle.rs:18-21)
__device__ __forceinline__ int escapes_of(int run)
{
    return (run - 1) / 15;                                           // run = 0 -> 0 (C division truncates)
}

struct LaneCoeffs {
    int      v[8];      // this lane's coefficients: macroblock positions 8*lane .. 8*lane+7
    uint32_t nz;        // bit k set = v[k] != 0
    int      prev;      // position of the last non-zero coefficient before 8*lane (-1: none)
    int      last;      // position of the macroblock's last non-zero coefficient (-1: none), all lanes
};

__device__ __forceinline__ LaneCoeffs unpack_lane(const uint4 raw, uint32_t lane)
{
    LaneCoeffs c;
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    c.nz = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        c.v[k] = (int)(int16_t)(w[k >> 1] >> (16 * (k & 1)));
        c.nz |= (c.v[k] != 0 ? 1u : 0u) << k;
    }
    int incl = c.nz ? (int)(8u * lane) + (31 - __clz((int)c.nz)) : -1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, d);
        if ((int)lane >= d) incl = max(incl, t);
    }
    c.prev = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) c.prev = -1;
    c.last = __shfl_sync(FULL, incl, 31);
    return c;
}

// RLE entries this lane produces: its non-zero coefficients with the escapes in front of them; lane 31 also owns the
// tail of the macroblock (src/rle.rs:31-38)
__device__ __forceinline__ uint32_t lane_count(const LaneCoeffs &c, uint32_t lane)
{
    uint32_t n = 0;
    int p = c.prev;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (c.nz & (1u << k)) {
            const int pos = (int)(8u * lane) + k;
            n += 1u + (uint32_t)escapes_of(pos - p - 1);
            p = pos;
        }
    if (lane == 31) {
        const int run = 255 - c.last;
        if (run > 0) n += 1u + (uint32_t)escapes_of(run);
    }
    return n;
}

// entries of the whole macroblock (all lanes get the sum)
__device__ __forceinline__ uint32_t warp_count(const uint4 raw, uint32_t lane)
{
    const LaneCoeffs c = unpack_lane(raw, lane);
    return __reduce_add_sync(FULL, lane_count(c, lane));
}

}  // namespace tok
}  // namespace pfv
