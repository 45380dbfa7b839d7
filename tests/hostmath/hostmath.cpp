// hostmath.cpp — the kernels' register-level arithmetic (pretty_fast_video_b200/csrc/pfv_dct.cuh) compiled for the CPU.
//
// Test infrastructure: tests/test_hostmath.py loads the shared library built from this file and compares every function
// with the oracle, so the transforms, the quantiser's reciprocal multiply and the run-length bookkeeping of the CUDA
// kernels are checked here, without a GPU, from the very source the kernels inline.  What is left to the -m gpu tests is
// the memory side (addressing, staging, hand-off between warps).  Built by tests/hostmath/Makefile with g++ -fwrapv
// (the kernels' `+ - *` wrap like release-mode Rust).
#include <stdint.h>
#include <string.h>

#include "../../pretty_fast_video_b200/csrc/pfv_dct.cuh"

using namespace pfv;

static const int kZigzag[64] = PFV_ZIGZAG_INIT;
static const int kScale[64] = PFV_SCALE_INIT;

extern "C" {

// src/common.rs:287-298 (delta == 0: px = 64 bytes) / :300-311 (delta == 1: px = 64 int16 residuals, already clamped)
void pfv_hm_encode_sb(const void *in, int delta, const int32_t q[64], int16_t out[64])
{
    int x[64];
    for (int i = 0; i < 64; i++) {
        if (delta) x[i] = (((const int16_t *)in)[i] / 2) * 256;            // src/common.rs:304
        else       x[i] = ((int)((const uint8_t *)in)[i] - 128) * 256;     // src/common.rs:291
    }
    uint32_t M[64];
    for (int i = 0; i < 64; i++) M[i] = quant_magic(q[i]);
    uint32_t w[32];
    encode_sb_regs(x, M, w);
    memcpy(out, w, 128);
}

// the fp32 formulation of the same (fdct8_f32 + quant_one_f32): must give the same bits.  Intra blocks go the way the
// encode-I kernel takes them: rows of packed pixels through fdct8_f32_row_of_bytes.
void pfv_hm_encode_sb_f32(const void *in, int delta, const int32_t q[64], int16_t out[64])
{
    float R[64];
    for (int i = 0; i < 64; i++) R[i] = quant_recip_f32(q[i]);
    uint32_t w[32];
    if (delta) {
        float y[64];
        for (int i = 0; i < 64; i++) y[i] = (float)(((const int16_t *)in)[i] / 2);
        encode_sb_regs_f32(y, R, w);
    } else {
        uint32_t lo[8], hi[8];
        for (int r = 0; r < 8; r++) {
            memcpy(&lo[r], (const uint8_t *)in + 8 * r, 4);
            memcpy(&hi[r], (const uint8_t *)in + 8 * r + 4, 4);
        }
        encode_sb_pixels_f32(lo, hi, R, w);
    }
    memcpy(out, w, 128);
}

// ... intra blocks through the generic fp32 path (byte_minus_128_f32 + encode_sb_regs_f32)
void pfv_hm_encode_sb_f32_generic(const uint8_t in[64], const int32_t q[64], int16_t out[64])
{
    float R[64], y[64];
    for (int i = 0; i < 64; i++) R[i] = quant_recip_f32(q[i]);
    for (int i = 0; i < 64; i++) y[i] = byte_minus_128_f32((uint32_t)in[i], 0);
    uint32_t w[32];
    encode_sb_regs_f32(y, R, w);
    memcpy(out, w, 128);
}

// trunc(n * quant_recip_f32(q)) against plain C division for every n in [-n_max, n_max] and q in [q_lo, q_hi]: mismatches
long pfv_hm_quant_f32_check(int n_max, int q_lo, int q_hi)
{
    long bad = 0;
    for (int q = q_lo; q <= q_hi; q++) {
        const float R = quant_recip_f32(q);
        for (int n = -n_max; n <= n_max; n++)
            if (quant_trunc_f32((float)n, R) != n / q) bad++;
    }
    return bad;
}

// the floor step of quant_one_f32 (host flavour) against the integer shift: (256 V * scale) >> 16 for 256 V = v
long pfv_hm_quant_f32_floor_check(int v_lo, int v_hi, int v_step)
{
    static const int scales[] = {22, 26, 28, 31, 32, 34, 35, 37, 39, 43};
    long bad = 0;
    for (int si = 0; si < 10; si++)
        for (int v = v_lo; v <= v_hi; v += v_step)
            if (quant_one_f32((float)v * 0.00390625f, scales[si], 1.0f) != ((v * scales[si]) >> 16)) bad++;
    return bad;
}

// src/common.rs:313-325 through idct8x8_regs (the +128 folded into the row pass): out = clamp(m, 0, 255)
void pfv_hm_decode_sb(const int16_t c[64], const int32_t q[64], uint8_t out[64])
{
    int m[64];
    for (int s = 0; s < 64; s++)
        m[kZigzag[s]] = (int)((uint32_t)(int)c[s] * ((uint32_t)kScale[s] * (uint32_t)q[s]));   // src/dct.rs:78-83, wrapping
    idct8x8_regs(m);
    for (int i = 0; i < 64; i++) out[i] = (uint8_t)(m[i] < 0 ? 0 : (m[i] > 255 ? 255 : m[i]));
}

// the same through idct8x8_regs_rolled (one copy of the 1-D transforms, transposed input)
void pfv_hm_decode_sb_rolled(const int16_t c[64], const int32_t q[64], uint8_t out[64])
{
    int t[64];
    for (int s = 0; s < 64; s++) {
        const int z = kZigzag[s];
        t[(z & 7) * 8 + (z >> 3)] = (int)((uint32_t)(int)c[s] * ((uint32_t)kScale[s] * (uint32_t)q[s]));
    }
    idct8x8_regs_rolled(t);
    for (int i = 0; i < 64; i++) out[i] = (uint8_t)(t[i] < 0 ? 0 : (t[i] > 255 ? 255 : t[i]));
}

// quant_one against plain C division over v in [v_lo, v_hi], every scale of the table, q in [q_lo, q_hi]: mismatches
long pfv_hm_quant_check(int v_lo, int v_hi, int v_step, int q_lo, int q_hi)
{
    static const int scales[] = {22, 26, 28, 31, 32, 34, 35, 37, 39, 43};
    long bad = 0;
    for (int q = q_lo; q <= q_hi; q++) {
        const uint32_t M = quant_magic(q);
        for (int si = 0; si < 10; si++)
            for (int v = v_lo; v <= v_hi; v += v_step) {
                const int n = (v * scales[si]) >> 16;
                if (quant_one(v, scales[si], M) != n / q) bad++;
            }
    }
    return bad;
}

long pfv_hm_escapes_check(void)
{
    long bad = 0;
    for (int run = 0; run <= 256; run++)
        if (rle_escapes(run) != (uint32_t)(run > 0 ? (run - 1) / 15 : 0)) bad++;
    return bad;
}

// RLE entries rle_encode makes of one macroblock's 256 coefficients, from the four sub-blocks' masks
uint32_t pfv_hm_mb_entry_count(const int16_t c[256])
{
    SbRuns r[4];
    for (int s = 0; s < 4; s++) {
        uint32_t w[32];
        memcpy(w, c + 64 * s, 128);
        r[s] = sb_runs(w);
    }
    return mb_entry_count(r);
}

}  // extern "C"
