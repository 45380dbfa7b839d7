// pfv_dct.cuh — the exact integer transforms and the quantiser of the reference as plain functions over registers.
//
// Everything here is `__host__ __device__`: the kernels inline it, and tests/hostmath compiles the very same source
// for the CPU and checks it against the oracle without a GPU (the memory side of the kernels is what the -m gpu
// tests are left to prove).  Reference (paths relative to the reference root):
//   src/dct.rs:176-239 fdct, :241-293 idct — `/` truncates toward zero
//   src/dct.rs:88-99   encode: c[i] = ((m[z]*SCALE[z]) >> 16) / q[z], z = ZIGZAG[i] (tables by RASTER position)
//   src/common.rs:287-325 sub-block drivers: encode = rows then columns, decode = columns then rows
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define PFV_HD __host__ __device__ __forceinline__
#else
#define PFV_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define PFV_UNROLL _Pragma("unroll")
#else
#define PFV_UNROLL
#endif

namespace pfv {

// ---- 1-D transforms ------------------------------------------------------------------------------
// src/dct.rs:241-293 / :176-239.  Rust's `/` truncates toward zero: x / 2^k = (x + (x < 0 ? 2^k - 1 : 0)) >> k.
// Written as (x - (2^k - 1) * s) >> k with s = x >> 31 (0 or -1) so that the bias is ONE multiply-add on the FMA
// pipe (IMAD) instead of a LEA.HI on the ALU pipe, and the sign is shared by the two divisions every operand of the
// butterflies takes: per operand 3 ALU + 2 FMA instructions instead of 5 ALU.  ncu had the ALU pipe as the busiest
// unit of every transform kernel (both pipes issue one warp instruction per 2 cycles per sub-partition, and the
// transforms were 2.2 : 1 ALU : FMA).  `+ - *` wrap like release-mode Rust; the bias never overflows (it is only
// added to negative values).
struct Tdiv {
    int x, s;
    PFV_HD explicit Tdiv(int v) : x(v), s(v >> 31) {}
    PFV_HD int d2() const { return (x - s) >> 1; }
    PFV_HD int d4() const { return (x - 3 * s) >> 2; }
    PFV_HD int d16() const { return (x - 15 * s) >> 4; }
};

PFV_HD void idct8(int (&v)[8])
{
    const int c0 = v[0], d4 = v[1], c2 = v[2], d6 = v[3], c1 = v[4], d5 = v[5], c3 = v[6], d7 = v[7];
    const int c4 = d4, c5 = d5 + d6, c7 = d5 - d6, c6 = d7;
    const int b4 = c4 + c5, b5 = c4 - c5, b6 = c6 + c7, b7 = c6 - c7;
    const int b0 = c0 + c1, b1 = c0 - c1;
    const Tdiv t2(c2), t3(c3), t4(b4), t5(b5), t6(b6), t7(b7);
    const int b2 = c2 + t2.d4() + t3.d2();
    const int b3 = t2.d2() - c3 - t3.d4();
    const int a4 = t7.d4() + b4 + t4.d4() - t4.d16();
    const int a7 = t4.d4() - b7 - t7.d4() + t7.d16();
    const int a5 = b5 - b6 + t6.d4() + t6.d16();
    const int a6 = b6 + b5 - t5.d4() - t5.d16();
    const int a0 = b0 + b2, a1 = b1 + b3, a2 = b1 - b3, a3 = b0 - b2;
    v[0] = a0 + a4; v[1] = a1 + a5; v[2] = a2 + a6; v[3] = a3 + a7;
    v[4] = a3 - a7; v[5] = a2 - a6; v[6] = a1 - a5; v[7] = a0 - a4;
}

// The forward transform only ever sees multiples of 256 ((p - 128) << 8, src/common.rs:291; (delta / 2) << 8, :304).  Every
// term of the row pass is such a value divided by at most 16, so its divisions are EXACT and its outputs are multiples
// of 16; the column pass divides those by at most 16 again: exact too.  An exact division needs no rounding toward
// zero - a plain arithmetic shift is the quotient - which takes the sign extraction and the bias (18 of 66 instructions)
// out of each of the 16 one-dimensional passes of a sub-block.  tests/test_hostmath.py compares with the oracle's
// truncating divisions on extreme and random blocks.
struct Xdiv {
    int x;
    PFV_HD explicit Xdiv(int v) : x(v) {}
    PFV_HD int d2() const { return x >> 1; }
    PFV_HD int d4() const { return x >> 2; }
    PFV_HD int d16() const { return x >> 4; }
};

template <typename DIV>
PFV_HD void fdct8_t(int (&v)[8])
{
    const int a0 = v[0] + v[7], a1 = v[1] + v[6], a2 = v[2] + v[5], a3 = v[3] + v[4];
    const int a4 = v[0] - v[7], a5 = v[1] - v[6], a6 = v[2] - v[5], a7 = v[3] - v[4];
    const int b0 = a0 + a3, b1 = a1 + a2, b2 = a0 - a3, b3 = a1 - a2;
    const int c0 = b0 + b1, c1 = b0 - b1;
    const DIV t2(b2), t3(b3), t4(a4), t5(a5), t6(a6), t7(a7);
    const int c2 = b2 + t2.d4() + t3.d2();
    const int c3 = t2.d2() - b3 - t3.d4();
    const int b4 = t7.d4() + a4 + t4.d4() - t4.d16();
    const int b7 = t4.d4() - a7 - t7.d4() + t7.d16();
    const int b5 = a5 + a6 - t6.d4() - t6.d16();
    const int b6 = a6 - a5 + t5.d4() + t5.d16();
    const int c4 = b4 + b5, c5 = b4 - b5, c6 = b6 + b7, c7 = b6 - b7;
    v[0] = c0; v[1] = c4; v[2] = c2; v[3] = c5 - c7;
    v[4] = c1; v[5] = c5 + c7; v[6] = c3; v[7] = c6;
}

// src/dct.rs:176-239 for ANY input (truncating divisions)
PFV_HD void fdct8(int (&v)[8]) { fdct8_t<Tdiv>(v); }
// the same for inputs that are multiples of 256 (row pass) or of 16 (column pass): see Xdiv
PFV_HD void fdct8_exact(int (&v)[8]) { fdct8_t<Xdiv>(v); }

// src/dct.rs:44-47 ZIGZAG_TABLE: raster index of scan position s (used only with compile-time indices)
#define PFV_ZIGZAG_INIT { \
     0,  1,  8, 16,  9,  2,  3, 10, 17, 24, 32, 25, 18, 11,  4,  5, \
    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,  6,  7, 14, 21, 28, \
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, \
    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63 }

// columns then rows (src/common.rs:315-316), +128 folded into the DC input of each row (see decode_mb_core)
PFV_HD void idct8x8_regs(int (&m)[64])
{
PFV_UNROLL
    for (int c = 0; c < 8; ++c) {
        int v[8];
PFV_UNROLL
        for (int r = 0; r < 8; ++r) v[r] = m[r * 8 + c];
        idct8(v);
PFV_UNROLL
        for (int r = 0; r < 8; ++r) m[r * 8 + c] = v[r];
    }
PFV_UNROLL
    for (int r = 0; r < 8; ++r) {
        int v[8];
PFV_UNROLL
        for (int c = 0; c < 8; ++c) v[c] = m[r * 8 + c];
        v[0] += 128 << 8;
        idct8(v);
PFV_UNROLL
        for (int c = 0; c < 8; ++c) m[r * 8 + c] = v[c] >> 8;
    }
}


// The same with ONE copy of the eight 1-D transforms in the code: `t` holds the sub-block TRANSPOSED (t[c * 8 + r] = the
// coefficient of row r, column c - a renaming where it is filled), both passes run the same loop body over contiguous
// groups of eight and swap the roles of rows and columns on the way out, so the result comes back in raster order in `t`.
// 64 register moves per pass buy half the instruction footprint (the straight-line form is ~15 KB of code; kernels that
// carry a forward transform next to it outgrow the 32 KB instruction cache).  Same values as idct8x8_regs.
PFV_HD void idct8x8_regs_rolled(int (&t)[64])
{
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int pass = 0; pass < 2; ++pass) {
        int o[64];
PFV_UNROLL
        for (int g = 0; g < 8; ++g) {
            int v[8];
PFV_UNROLL
            for (int k = 0; k < 8; ++k) v[k] = t[g * 8 + k];
            idct8(v);
PFV_UNROLL
            for (int k = 0; k < 8; ++k) o[k * 8 + g] = v[k];
        }
        // after the column pass o[r * 8 + c] is raster order: the row pass wants +128 on each row's DC input (see
        // idct8x8_regs); after the row pass the groups were rows, so o is the transposed result
        const int bias = pass == 0 ? (128 << 8) : 0;
PFV_UNROLL
        for (int i = 0; i < 64; ++i) t[i] = o[i] + ((i & 7) == 0 ? bias : 0);
    }
    // t[c * 8 + r] now holds pixel (r, c) << 8: hand it back in raster order
    int o[64];
PFV_UNROLL
    for (int i = 0; i < 64; ++i) o[(i & 7) * 8 + (i >> 3)] = t[i] >> 8;
PFV_UNROLL
    for (int i = 0; i < 64; ++i) t[i] = o[i];
}

// src/dct.rs:4-13 DCT_SCALE_FACTOR by raster position (used only with compile-time indices)
#define PFV_SCALE_INIT { \
    32, 37, 34, 26, 32, 26, 34, 37, \
    37, 43, 39, 31, 37, 31, 39, 43, \
    34, 39, 35, 28, 34, 28, 35, 39, \
    26, 31, 28, 22, 26, 22, 28, 31, \
    32, 37, 34, 26, 32, 26, 34, 37, \
    26, 31, 28, 22, 26, 22, 28, 31, \
    34, 39, 35, 28, 34, 28, 35, 39, \
    37, 43, 39, 31, 37, 31, 39, 43 }

// rows then columns (src/common.rs:294-295 / :307-308); m = multiples of 256 (see Xdiv)
PFV_HD void fdct8x8_regs(int (&m)[64])
{
PFV_UNROLL
    for (int r = 0; r < 8; ++r) {
        int v[8];
PFV_UNROLL
        for (int c = 0; c < 8; ++c) v[c] = m[r * 8 + c];
        fdct8_exact(v);
PFV_UNROLL
        for (int c = 0; c < 8; ++c) m[r * 8 + c] = v[c];
    }
PFV_UNROLL
    for (int c = 0; c < 8; ++c) {
        int v[8];
PFV_UNROLL
        for (int r = 0; r < 8; ++r) v[r] = m[r * 8 + c];
        fdct8_exact(v);
PFV_UNROLL
        for (int r = 0; r < 8; ++r) m[r * 8 + c] = v[r];
    }
}

PFV_HD float bits_as_float(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; __builtin_memcpy(&f, &u, 4); return f;
#endif
}
PFV_HD uint32_t float_as_bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; __builtin_memcpy(&u, &f, 4); return u;
#endif
}

// The same transform in fp32.  With exact divisions (above) the butterfly is LINEAR with dyadic coefficients
// (x + x/4 - x/16 = 19/16 x, x - x/4 - x/16 = 11/16 x, ...): computed on the UNSCALED inputs y = p - 128 (or delta / 2), |y| <= 128,
// every intermediate is a multiple of 2^-8 below 2^14 - 22 significant bits, exactly representable, and an FMA whose exact
// result is representable returns it.  30 FMA-pipe instructions per pass instead of 48 ALU/FMA ones, and the quantiser's
// input 256 * V comes out of ONE FFMA that also does the float -> int conversion (the 1.5 * 2^23 trick: |256 V| < 2^22).
PFV_HD void fdct8_f32(float (&v)[8])
{
    const float a0 = v[0] + v[7], a1 = v[1] + v[6], a2 = v[2] + v[5], a3 = v[3] + v[4];
    const float a4 = v[0] - v[7], a5 = v[1] - v[6], a6 = v[2] - v[5], a7 = v[3] - v[4];
    const float b0 = a0 + a3, b1 = a1 + a2, b2 = a0 - a3, b3 = a1 - a2;
    const float c0 = b0 + b1, c1 = b0 - b1;
    const float c2 = fmaf(b2, 1.25f, b3 * 0.5f);              // b2 + b2/4 + b3/2
    const float c3 = fmaf(b3, -1.25f, b2 * 0.5f);             // b2/2 - b3 - b3/4
    const float b4 = fmaf(a4, 1.1875f, a7 * 0.25f);           // a7/4 + a4 + a4/4 - a4/16
    const float b7 = fmaf(a7, -1.1875f, a4 * 0.25f);          // a4/4 - a7 - a7/4 + a7/16
    const float b5 = fmaf(a6, 0.6875f, a5);                   // a5 + a6 - a6/4 - a6/16
    const float b6 = fmaf(a5, -0.6875f, a6);                  // a6 - a5 + a5/4 + a5/16
    const float c4 = b4 + b5, c5 = b4 - b5, c6 = b6 + b7, c7 = b6 - b7;
    v[0] = c0; v[1] = c4; v[2] = c2; v[3] = c5 - c7;
    v[4] = c1; v[5] = c5 + c7; v[6] = c3; v[7] = c6;
}

// the two halves of fdct8_f32: `a` = the first butterfly layer (sums a0..a3, differences a4..a7)
PFV_HD void fdct8_f32_tail(const float (&a)[8], float (&v)[8])
{
    const float b0 = a[0] + a[3], b1 = a[1] + a[2], b2 = a[0] - a[3], b3 = a[1] - a[2];
    const float c0 = b0 + b1, c1 = b0 - b1;
    const float c2 = fmaf(b2, 1.25f, b3 * 0.5f);
    const float c3 = fmaf(b3, -1.25f, b2 * 0.5f);
    const float b4 = fmaf(a[4], 1.1875f, a[7] * 0.25f);
    const float b7 = fmaf(a[7], -1.1875f, a[4] * 0.25f);
    const float b5 = fmaf(a[6], 0.6875f, a[5]);
    const float b6 = fmaf(a[5], -0.6875f, a[6]);
    const float c4 = b4 + b5, c5 = b4 - b5, c6 = b6 + b7, c7 = b6 - b7;
    v[0] = c0; v[1] = c4; v[2] = c2; v[3] = c5 - c7;
    v[4] = c1; v[5] = c5 + c7; v[6] = c3; v[7] = c6;
}

// One row of eight PIXELS (two little-endian words) through the level shift (p - 128, src/common.rs:291) and fdct8_f32 without
// converting the bytes one by one: 0x4B000000 | p is the float 2^23 + p, so a difference of two such floats is p_i - p_j
// exactly, and a sum with its two level shifts is (f_i - (2^24 + 256)) + f_j - every intermediate an integer below 2^24.
// 12 additions for the first layer instead of 8 conversions + 8.
PFV_HD void fdct8_f32_row_of_bytes(uint32_t lo, uint32_t hi, float (&v)[8])
{
    float f[8];
PFV_UNROLL
    for (int k = 0; k < 8; ++k) {
        const uint32_t w = k < 4 ? lo : hi;
#if defined(__CUDA_ARCH__)
        f[k] = bits_as_float(__byte_perm(w, 0x4B000000u, 0x7440 | (k & 3)));
#else
        f[k] = bits_as_float(0x4B000000u | ((w >> (8 * (k & 3))) & 0xffu));
#endif
    }
    float a[8];
PFV_UNROLL
    for (int k = 0; k < 4; ++k) {
        a[k] = (f[k] - 16777472.0f) + f[7 - k];                  // p_k + p_(7-k) - 256
        a[4 + k] = f[k] - f[7 - k];
    }
    fdct8_f32_tail(a, v);
}

// column pass of the fp32 forward transform over 64 values whose rows are already transformed
PFV_HD void fdct8x8_f32_columns(float (&m)[64])
{
PFV_UNROLL
    for (int c = 0; c < 8; ++c) {
        float v[8];
PFV_UNROLL
        for (int r = 0; r < 8; ++r) v[r] = m[r * 8 + c];
        fdct8_f32(v);
PFV_UNROLL
        for (int r = 0; r < 8; ++r) m[r * 8 + c] = v[r];
    }
}

PFV_HD void fdct8x8_f32(float (&m)[64])
{
PFV_UNROLL
    for (int r = 0; r < 8; ++r) {
        float v[8];
PFV_UNROLL
        for (int c = 0; c < 8; ++c) v[c] = m[r * 8 + c];
        fdct8_f32(v);
PFV_UNROLL
        for (int c = 0; c < 8; ++c) m[r * 8 + c] = v[c];
    }
PFV_UNROLL
    for (int c = 0; c < 8; ++c) {
        float v[8];
PFV_UNROLL
        for (int r = 0; r < 8; ++r) v[r] = m[r * 8 + c];
        fdct8_f32(v);
PFV_UNROLL
        for (int r = 0; r < 8; ++r) m[r * 8 + c] = v[r];
    }
}

// byte k of a word as (p - 128) in fp32 without a conversion instruction: 0x4B000000 | p is the float 2^23 + p
PFV_HD float byte_minus_128_f32(uint32_t w, int k)
{
#if defined(__CUDA_ARCH__)
    const uint32_t bits = __byte_perm(w, 0x4B000000u, 0x7440 | k);       // bytes: [w.k, 00, 00, 4B]
#else
    const uint32_t bits = 0x4B000000u | ((w >> (8 * k)) & 0xffu);
#endif
    return bits_as_float(bits) - 8388736.0f;                              // 2^23 + 128
}

PFV_HD int mulhi_s32(int a, int b)
{
#if defined(__CUDA_ARCH__)
    return __mulhi(a, b);
#else
    return (int)(((long long)a * (long long)b) >> 32);
#endif
}

// The quantiser's divisor as a multiplier: floor(2^30 / q) + 1 (strictly above 2^30 / q, also for powers of two).
inline uint32_t quant_magic(int32_t q) { return q > 0 ? (uint32_t)((1u << 30) / (uint32_t)q) + 1u : 0u; }
constexpr int32_t QUANT_MAX_DIVISOR = 65535;

// src/dct.rs:92-95: n = (v * scale) >> 16 (arithmetic), then n / q truncating toward zero, as
//     mulhi(4 n, M) + (n < 0),   M = quant_magic(q).
// 4 n M / 2^32 = n / q + 4 n eps / 2^32 with 0 < eps <= 1: for n >= 0 the error is below 1/q as long as 4 n q < 2^32, so the
// floor is floor(n / q); for n < 0 the error is strictly negative and smaller than 1/q, so the floor is ceil(n / q) - 1.
// |n| <= 1 100 for every 8x8 of pixels or halved differences (the transform's gain is at most 8.75^2 per pass pair and
// SCALE <= 43), q <= 65 535: exact.  tests/test_hostmath.py checks it exhaustively over that range.
PFV_HD int quant_one(int v, int scale, uint32_t M)
{
    const int a = v * scale;
    const int t = (a >> 14) & ~3;                               // 4 * (a >> 16)
    return mulhi_s32(t, (int)M) + (int)((uint32_t)a >> 31);
}

// The quantiser on the fp32 transform's output V (unscaled: the reference's value is 256 V), entirely on the FMA pipe:
//   n = (256 V * scale) >> 16 = floor(V * scale / 256): ONE fused multiply-add that rounds toward minus infinity onto
//       1.5 * 2^23 (ulp 1 there; the product is exact inside the FMA, so the rounding IS the floor), |n| < 2^12;
//   n / q truncating = trunc(n * R), R = quant_recip_f32(q) a little ABOVE 1/q: n R >= n/q keeps exact multiples, and the
//       excess |n| * 2^-15 / q stays below the 1/q that separates n/q from the next integer.
// FFMA.RM, FADD, FMUL, F2I.TRUNC: 4 instructions per coefficient (the integer form above takes 6 to 8).
// tests/test_hostmath.py checks trunc(n * R) == n / q for every |n| <= 4096 and every q in 1..65535.
inline float quant_recip_f32(int32_t q) { return q > 0 ? (float)((1.0 / (double)q) * (1.0 + 1.0 / 32768.0)) : 0.0f; }

PFV_HD int quant_trunc_f32(float nf, float R) { return (int)(nf * R); }         // cvt.rzi.s32.f32

PFV_HD int quant_one_f32(float V, int scale, float R)
{
#if defined(__CUDA_ARCH__)
    const float nf = __fmaf_rd(V, (float)scale * 0.00390625f, 12582912.0f) - 12582912.0f;
#else
    const float nf = (float)floor((double)V * (double)scale * 0.00390625);        // exact in double
#endif
    return quant_trunc_f32(nf, R);
}

// two quantised coefficients -> one word of the dense layout (low half first)
PFV_HD uint32_t pack_i16x2(int lo, int hi)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410);
#else
    return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16);
#endif
}

// src/common.rs:287-298 / :300-311 after the level shift: x = the 64 inputs in raster order ((p - 128) << 8 or
// (delta / 2) << 8).  Leaves the quantised coefficients in SCAN order, two per word, exactly as the dense layout
// stores them (src/dct.rs:88-99).  encM = quant_magic of the q-table by RASTER position.
// the same on fp32 inputs y = p - 128 (or delta / 2), unscaled (see fdct8_f32); encR = quant_recip_f32 of the q-table by
// RASTER position
PFV_HD void quantise_sb_f32(const float (&y)[64], const float *encR, uint32_t (&w)[32])
{
    constexpr int zz[64] = PFV_ZIGZAG_INIT;
    constexpr int sc[64] = PFV_SCALE_INIT;
PFV_UNROLL
    for (int i = 0; i < 32; ++i) {
        const int z0 = zz[2 * i], z1 = zz[2 * i + 1];
        w[i] = pack_i16x2(quant_one_f32(y[z0], sc[z0], encR[z0]), quant_one_f32(y[z1], sc[z1], encR[z1]));
    }
}

PFV_HD void encode_sb_regs_f32(float (&y)[64], const float *encR, uint32_t (&w)[32])
{
    fdct8x8_f32(y);
    quantise_sb_f32(y, encR, w);
}

// ... on the sub-block's eight rows of PIXELS as they were loaded (rows[r] = bytes 0..3, 4..7 of row r)
PFV_HD void encode_sb_pixels_f32(const uint32_t (&lo)[8], const uint32_t (&hi)[8], const float *encR, uint32_t (&w)[32])
{
    float y[64];
PFV_UNROLL
    for (int r = 0; r < 8; ++r) {
        float v[8];
        fdct8_f32_row_of_bytes(lo[r], hi[r], v);
PFV_UNROLL
        for (int c = 0; c < 8; ++c) y[r * 8 + c] = v[c];
    }
    fdct8x8_f32_columns(y);
    quantise_sb_f32(y, encR, w);
}

PFV_HD void encode_sb_regs(int (&x)[64], const uint32_t *encM, uint32_t (&w)[32])
{
    constexpr int zz[64] = PFV_ZIGZAG_INIT;
    constexpr int sc[64] = PFV_SCALE_INIT;
    fdct8x8_regs(x);
PFV_UNROLL
    for (int i = 0; i < 32; ++i) {
        const int z0 = zz[2 * i], z1 = zz[2 * i + 1];
        w[i] = pack_i16x2(quant_one(x[z0], sc[z0], encM[z0]), quant_one(x[z1], sc[z1], encM[z1]));
    }
}

// ---- run-length bookkeeping of one sub-block (src/rle.rs:9-39) ----------------------------------------------
// What rle_encode makes of a macroblock's 256 coefficients is: one entry per non-zero coefficient, floor((g - 1) / 15)
// escape entries in front of it when g >= 1 zeros precede it, and for the zeros after the last one the same escapes
// plus one closing entry.  A thread that owns 64 of the 256 coefficients can count everything that happens INSIDE its
// sub-block from the 64-bit mask of non-zeros alone: an inner gap of g zeros costs one escape for each of the
// thresholds 16, 31, 46, 61 it reaches.
struct SbRuns {
    int      first, last;   // positions 0..63 of the first / last non-zero coefficient, -1 if none
    uint32_t inner;         // non-zero coefficients + the escapes of the gaps between them
    uint64_t mask;          // bit i = coefficient i is non-zero
};

PFV_HD int clz64_(uint64_t v)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}
PFV_HD int ffs64_(uint64_t v)     // 1-based, 0 if none
{
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)v);
#else
    return v ? __builtin_ctzll(v) + 1 : 0;
#endif
}
PFV_HD int popc64_(uint64_t v)
{
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}

PFV_HD SbRuns sb_runs(const uint32_t (&w)[32])
{
    uint32_t lo = 0, hi = 0;
PFV_UNROLL
    for (int i = 0; i < 32; ++i) {
        const uint32_t b = ((w[i] & 0xffffu) ? 1u : 0u) | ((w[i] >> 16) ? 2u : 0u);
        if (i < 16) lo |= b << (2 * i); else hi |= b << (2 * (i - 16));
    }
    const uint64_t M = ((uint64_t)hi << 32) | lo;
    SbRuns r;
    r.mask = M;
    r.first = ffs64_(M) - 1;
    r.last = 63 - clz64_(M);
    // z16 bit i: the 16 positions i-16 .. i-1 exist and are all zero
    const uint64_t Z = ~M;
    uint64_t a = Z << 1;                 // bit i: position i-1 is zero
    a &= a << 1;                         // i-1, i-2
    a &= a << 2;                         // 4 positions
    a &= a << 4;                         // 8
    const uint64_t z16 = a & (a << 8);
    const uint64_t z31 = z16 & (z16 << 15);
    const uint64_t z46 = z31 & (z16 << 30);
    const uint64_t z61 = z46 & (z16 << 45);
    const uint64_t inner = M & (M - 1);  // every non-zero but the first: its gap lies inside the sub-block
    r.inner = (uint32_t)(popc64_(M) + popc64_(inner & z16) + popc64_(inner & z31) + popc64_(inner & z46) + popc64_(inner & z61));
    return r;
}

// escapes in front of an entry that closes a run of `run` zeros: floor((run - 1) / 15) for 1 <= run <= 256, 0 for run = 0
PFV_HD uint32_t rle_escapes(int run) { return run > 0 ? ((uint32_t)(run - 1) * 4370u) >> 16 : 0u; }

// The macroblock's entry count from its four sub-blocks (sub-block s covers positions 64 s .. 64 s + 63).
PFV_HD uint32_t mb_entry_count(const SbRuns (&sb)[4])
{
    uint32_t n = 0;
    int last = -1;
PFV_UNROLL
    for (int s = 0; s < 4; ++s) {
        n += sb[s].inner;
        if (sb[s].first >= 0) {
            n += rle_escapes(64 * s + sb[s].first - last - 1);
            last = 64 * s + sb[s].last;
        }
    }
    const int run = 255 - last;
    if (run > 0) n += 1u + rle_escapes(run);
    return n;
}

}  // namespace pfv
