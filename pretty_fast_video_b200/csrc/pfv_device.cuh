// pfv_device.cuh — device-side building blocks of the macroblock engine (sm_100a).
//
// Work decomposition (all kernels): one warp owns one 16x16 macroblock; lane l owns one 8-vector
// of sub-block (l >> 3): a column (index l & 7) during a column pass, a row during a row pass, so
// every 1-D butterfly runs entirely in registers.  The 8x8 transposes between passes, the zig-zag
// gather/scatter and the coefficient staging go through a small per-warp shared-memory scratch.
//
// Arithmetic follows the reference exactly (paths relative to the reference root):
//   src/dct.rs:176-239 fdct, :241-293 idct — `/` truncates toward zero (C++ int `/` does too)
//   src/dct.rs:75-86   decode: raster[i] = c[s]*SCALE[s]*q[s], s = INV_ZIGZAG[i]  (tables by SCAN position)
//   src/dct.rs:88-99   encode: c[i] = ((m[z]*SCALE[z]) >> 16) / q[z], z = ZIGZAG[i] (tables by RASTER position)
//   src/common.rs:287-325 sub-block drivers: encode = rows then columns, decode = columns then rows
//   src/common.rs:98-104  apply_residuals, :108-123 calc_residuals, :125-139 calc_error (SSD)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pfv_dct.cuh"

namespace pfv {

// ---- per-warp shared scratch -------------------------------------------------------------------
// coef: 4 sub-blocks x 64 int16 in scan order, sub-block stride padded 64 -> 80 halfwords so the
//       four sub-blocks of a gather land on different banks.
// t   : 4 sub-blocks x 8 rows x 8 int32, row stride 12 words and sub-block stride 104 words: the
//       column-side scalar accesses and the row-side 128-bit accesses are both conflict free.
constexpr int COEF_SB_STRIDE = 80;   // halfwords
constexpr int T_ROW_STRIDE   = 12;   // words
constexpr int T_SB_STRIDE    = 104;  // words
struct __align__(16) WarpScratch {
    int16_t coef[4 * COEF_SB_STRIDE];   // 640 B
    int32_t t[4 * T_SB_STRIDE];         // 1664 B
};

// "Lane transposed" constant tables: entry [c*8 + row] belongs to column c.
// src/dct.rs:39-42 transposed: c_izT[c*8+row] = INV_ZIGZAG_TABLE[row*8+c]
__constant__ uint8_t c_izT[64] = {
     0,  2,  3,  9, 10, 20, 21, 35,
     1,  4,  8, 11, 19, 22, 34, 36,
     5,  7, 12, 18, 23, 33, 37, 48,
     6, 13, 17, 24, 32, 38, 47, 49,
    14, 16, 25, 31, 39, 46, 50, 57,
    15, 26, 30, 40, 45, 51, 56, 58,
    27, 29, 41, 44, 52, 55, 59, 62,
    28, 42, 43, 53, 54, 60, 61, 63,
};
// src/dct.rs:4-13 transposed: c_scaleT[c*8+row] = DCT_SCALE_FACTOR[row*8+c] (the table is symmetric)
__constant__ int32_t c_scaleT[64] = {
    32, 37, 34, 26, 32, 26, 34, 37,
    37, 43, 39, 31, 37, 31, 39, 43,
    34, 39, 35, 28, 34, 28, 35, 39,
    26, 31, 28, 22, 26, 22, 28, 31,
    32, 37, 34, 26, 32, 26, 34, 37,
    26, 31, 28, 22, 26, 22, 28, 31,
    34, 39, 35, 28, 34, 28, 35, 39,
    37, 43, 39, 31, 37, 31, 39, 43,
};

// d = (c << 16) | (sat_u8(a) << 8) | sat_u8(b)   (PTX cvt.pack, SASS I2IP)
__device__ __forceinline__ uint32_t pack_sat_u8(int a, int b, uint32_t c)
{
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// four ints -> four saturated bytes, p0 in the lowest byte
__device__ __forceinline__ uint32_t pack4_sat_u8(int p0, int p1, int p2, int p3)
{
    return pack_sat_u8(p1, p0, pack_sat_u8(p3, p2, 0u));
}

// Per-lane constants that do not change between macroblocks of a plane.
struct LaneTables {
    uint32_t gaddr[8];   // byte offset inside WarpScratch::coef of the coefficient feeding (row, c)
    int32_t  deq[8];     // decode multiplier for (row, c)
};

__device__ __forceinline__ void lane_gather_offsets(int lane, uint32_t (&gaddr)[8])
{
    const int sb = lane >> 3, c = lane & 7;
#pragma unroll
    for (int row = 0; row < 8; ++row)
        gaddr[row] = (uint32_t)(sb * COEF_SB_STRIDE + c_izT[c * 8 + row]) * 2u;
}

__device__ __forceinline__ void lane_load8(const int32_t *tbl, int c, int32_t (&out)[8])
{
    const int4 lo = __ldg(reinterpret_cast<const int4 *>(tbl + c * 8));
    const int4 hi = __ldg(reinterpret_cast<const int4 *>(tbl + c * 8 + 4));
    out[0] = lo.x; out[1] = lo.y; out[2] = lo.z; out[3] = lo.w;
    out[4] = hi.x; out[5] = hi.y; out[6] = hi.z; out[7] = hi.w;
}

// Stage one macroblock's 256 coefficients (lane holds scan positions j*8..j*8+7 of sub-block sb).
__device__ __forceinline__ void stage_coeffs(WarpScratch &ws, int lane, const uint4 &craw)
{
    const int sb = lane >> 3, j = lane & 7;
    *reinterpret_cast<uint4 *>(&ws.coef[sb * COEF_SB_STRIDE + j * 8]) = craw;
}

// src/common.rs:313-325 for the four sub-blocks of one macroblock (src/common.rs:238-252).
// Input: coefficients staged in ws.coef (caller has issued __syncwarp()).  Output: for lane (sb, r)
// the 8 values of row r of sub-block sb as ints ALREADY shifted: y[k] = (m >> 8) + 128 before the
// clamp, i.e. the caller only has to saturate.  The +128 is folded into the DC input of the row
// pass (every idct output carries v[0] with coefficient +1 and no division touches it), which is
// exact in wrapping arithmetic.
__device__ __forceinline__ void decode_mb_core(WarpScratch &ws, int lane, const uint32_t (&gaddr)[8],
                                               const int32_t (&deq)[8], int (&y)[8])
{
    const int sb = lane >> 3, c = lane & 7;
    const char *cbase = reinterpret_cast<const char *>(ws.coef);
    int x[8];
#pragma unroll
    for (int row = 0; row < 8; ++row) {
        const int cv = *reinterpret_cast<const int16_t *>(cbase + gaddr[row]);
        x[row] = cv * deq[row];                          // src/dct.rs:79-82
    }
    idct8(x);                                            // columns first, src/common.rs:315
    int32_t *tcol = &ws.t[sb * T_SB_STRIDE + c];
#pragma unroll
    for (int row = 0; row < 8; ++row) tcol[row * T_ROW_STRIDE] = x[row];
    __syncwarp();
    const int4 lo = *reinterpret_cast<const int4 *>(&ws.t[sb * T_SB_STRIDE + c * T_ROW_STRIDE]);
    const int4 hi = *reinterpret_cast<const int4 *>(&ws.t[sb * T_SB_STRIDE + c * T_ROW_STRIDE + 4]);
    y[0] = lo.x + (128 << 8); y[1] = lo.y; y[2] = lo.z; y[3] = lo.w;
    y[4] = hi.x; y[5] = hi.y; y[6] = hi.z; y[7] = hi.w;
    idct8(y);                                            // then rows, src/common.rs:316
#pragma unroll
    for (int k = 0; k < 8; ++k) y[k] >>= 8;              // src/common.rs:321 (arithmetic shift)
    __syncwarp();                                        // scratch may be reused by the caller
}

// Intra output: clamp to u8 and pack (src/common.rs:321).
__device__ __forceinline__ uint2 pack_row_u8(const int (&y)[8])
{
    uint2 o;
    o.x = pack4_sat_u8(y[0], y[1], y[2], y[3]);
    o.y = pack4_sat_u8(y[4], y[5], y[6], y[7]);
    return o;
}

// src/common.rs:98-104: out = clamp(prev + (d - 128) * 2, 0, 255), d = clamp(y, 0, 255).
// d is clamped and packed four at a time by the saturating pack (I2IP); prev + 2 d - 256 is then two byte dot
// products per pixel (IDP.4A with one-hot constants picks the byte and scales it) - no per-byte shifts and masks.
__device__ __forceinline__ uint2 apply_residual_row(const int (&y)[8], uint2 prev)
{
    const uint32_t d_lo = pack4_sat_u8(y[0], y[1], y[2], y[3]);
    const uint32_t d_hi = pack4_sat_u8(y[4], y[5], y[6], y[7]);
    int o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t dw = (k < 4) ? d_lo : d_hi, pw = (k < 4) ? prev.x : prev.y;
        const uint32_t t = __dp4a(dw, 2u << (8 * (k & 3)), (uint32_t)-256);
        o[k] = (int)__dp4a(pw, 1u << (8 * (k & 3)), t);
    }
    uint2 r;
    r.x = pack4_sat_u8(o[0], o[1], o[2], o[3]);
    r.y = pack4_sat_u8(o[4], o[5], o[6], o[7]);
    return r;
}

// Forward path of one macroblock.  x[k] = the 8 level-shifted inputs of row r of sub-block sb
// ((p-128)<<8, src/common.rs:291, or (delta/2)<<8, src/common.rs:304).  Leaves the 256 quantised
// coefficients staged in ws.coef in scan order (exactly the layout decode_mb_core expects) and
// returns this lane's 8 consecutive coefficients (scan positions j*8.. of sub-block sb) as a uint4.
__device__ __forceinline__ uint4 encode_mb_core(WarpScratch &ws, int lane, const uint32_t (&gaddr)[8],
                                                const uint32_t (&encM)[8], const int32_t (&scale)[8],
                                                int (&x)[8])
{
    const int sb = lane >> 3, c = lane & 7;
    fdct8_exact(x);                                           // rows first, src/common.rs:294
    int32_t *trow = &ws.t[sb * T_SB_STRIDE + c * T_ROW_STRIDE];
    *reinterpret_cast<int4 *>(trow)     = make_int4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<int4 *>(trow + 4) = make_int4(x[4], x[5], x[6], x[7]);
    __syncwarp();
    const int32_t *tcol = &ws.t[sb * T_SB_STRIDE + c];
    int v[8];
#pragma unroll
    for (int row = 0; row < 8; ++row) v[row] = tcol[row * T_ROW_STRIDE];
    fdct8_exact(v);                                           // then columns, src/common.rs:295
    char *cbase = reinterpret_cast<char *>(ws.coef);
#pragma unroll
    for (int row = 0; row < 8; ++row) {
        const int n = (v[row] * scale[row]) >> 16;       // src/dct.rs:92, |n| <= 32768
        // n / q truncating (src/dct.rs:95) as sign * floor(|n| * ceil(2^31/q) / 2^31): exact for
        // |n| <= 2^15 and q < 2^16.
        const uint32_t an = (uint32_t)abs(n);
        const int qa = (int)__umulhi(an << 1, encM[row]);
        const int qv = n < 0 ? -qa : qa;
        *reinterpret_cast<int16_t *>(cbase + gaddr[row]) = (int16_t)qv;   // scan position INV_ZIGZAG[row*8+c]
    }
    __syncwarp();
    return *reinterpret_cast<const uint4 *>(&ws.coef[sb * COEF_SB_STRIDE + (lane & 7) * 8]);
}

// ---- small helpers ---------------------------------------------------------------------------------
// n / d for n < 2^24 with r = 1/d precomputed in float (host) — used once per warp to turn a linear
// macroblock index into (row, column).
__device__ __forceinline__ uint32_t div_small(uint32_t n, uint32_t d, float rcp, uint32_t &rem)
{
    uint32_t q = (uint32_t)(__uint2float_rz(n) * rcp);
    int r = (int)(n - q * d);
    if (r < 0) { --q; r += (int)d; }
    else if (r >= (int)d) { ++q; r -= (int)d; }
    rem = (uint32_t)r;
    return q;
}

// 8 bytes at an arbitrary byte address of global memory (three aligned 32-bit loads + funnel shifts).
// May touch up to 3 bytes past the 8 requested: the frame pool is allocated with slack for that.
__device__ __forceinline__ uint2 ldg_u8x8_unaligned(const uint8_t *p)
{
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~uintptr_t(3));
    const uint32_t sh = (uint32_t)(a & 3u) * 8u;
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    uint2 r;
    r.x = __funnelshift_r(w0, w1, sh);
    r.y = __funnelshift_r(w1, w2, sh);
    return r;
}

}  // namespace pfv
