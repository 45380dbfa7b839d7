/*
 * pfv_oracle.c — CPU oracle: plain-C restatement of pfv-rs 0.2.2 (codec 2.1.1).
 *
 * TEST INFRASTRUCTURE ONLY (see pfv_oracle.h).  PARITY UNPINNED by the reference's own
 * tests; cross-checked by oracle/pfv_ref.py and the KATs in tests/test_oracle_kat.py.
 *
 * Every function names the reference lines it follows (paths relative to the
 * reference root).  Build with -fwrapv: the reference is release-mode Rust, whose
 * i32 + - * wrap; C `/` truncates toward zero exactly like Rust's, and gcc's `>>`
 * on negative ints is an arithmetic shift like Rust's.
 */
#include "pfv_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define FP_BITS 8 /* dct.rs:1 */

/* dct.rs:4-13 */
static const int32_t DCT_SCALE_FACTOR[64] = {
    32, 37, 34, 26, 32, 26, 34, 37,
    37, 43, 39, 31, 37, 31, 39, 43,
    34, 39, 35, 28, 34, 28, 35, 39,
    26, 31, 28, 22, 26, 22, 28, 31,
    32, 37, 34, 26, 32, 26, 34, 37,
    26, 31, 28, 22, 26, 22, 28, 31,
    34, 39, 35, 28, 34, 28, 35, 39,
    37, 43, 39, 31, 37, 31, 39, 43,
};

/* dct.rs:16-25 */
static const int32_t Q_TABLE_INTRA[64] = {
     8, 16, 19, 22, 26, 27, 29, 34,
    16, 16, 22, 24, 27, 29, 34, 37,
    19, 22, 26, 27, 29, 34, 34, 38,
    22, 22, 26, 27, 29, 34, 37, 40,
    22, 26, 27, 29, 32, 35, 40, 48,
    26, 27, 29, 32, 35, 40, 48, 58,
    26, 27, 29, 34, 38, 46, 56, 69,
    27, 29, 35, 38, 46, 56, 69, 83,
};

/* dct.rs:28-37 */
static const int32_t Q_TABLE_INTER[64] = {
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16,
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16,
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16,
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 16,
};

/* dct.rs:39-42 */
static const uint8_t INV_ZIGZAG_TABLE[64] = {
     0,  1,  5,  6, 14, 15, 27, 28,  2,  4,  7, 13, 16, 26, 29, 42,
     3,  8, 12, 17, 25, 30, 41, 43,  9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,
    21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63,
};

/* dct.rs:44-47 */
static const uint8_t ZIGZAG_TABLE[64] = {
     0,  1,  8, 16,  9,  2,  3, 10, 17, 24, 32, 25, 18, 11,  4,  5,
    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,  6,  7, 14, 21, 28,
    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
};

const int32_t *pfvo_dct_scale_factor(void) { return DCT_SCALE_FACTOR; }
const int32_t *pfvo_q_table_intra(void) { return Q_TABLE_INTRA; }
const int32_t *pfvo_q_table_inter(void) { return Q_TABLE_INTER; }
const uint8_t *pfvo_zigzag(void) { return ZIGZAG_TABLE; }
const uint8_t *pfvo_inv_zigzag(void) { return INV_ZIGZAG_TABLE; }

/* ------------------------------------------------------------------------------------
 * geometry — frame.rs:28-49 (VideoFrame::new_padded)
 * ---------------------------------------------------------------------------------- */
static uint32_t pad16(uint32_t v) { return v + (16 - (v % 16)) % 16; } /* frame.rs:29-30 */

void pfvo_geometry_for(uint32_t width, uint32_t height, pfvo_geometry *g)
{
    g->width = width;
    g->height = height;
    g->cwidth = width / 2;   /* frame.rs:32 */
    g->cheight = height / 2; /* frame.rs:33 */
    g->pw = pad16(width);
    g->ph = pad16(height);
    g->cpw = pad16(g->cwidth);  /* frame.rs:35 */
    g->cph = pad16(g->cheight); /* frame.rs:36 */
    g->nb_y = (g->pw / 16) * (g->ph / 16);
    g->nb_c = (g->cpw / 16) * (g->cph / 16);
    g->nb = g->nb_y + 2 * g->nb_c; /* dec.rs:255 */
}

size_t pfvo_frame_bytes(const pfvo_geometry *g)
{
    return (size_t)g->pw * g->ph + 2 * (size_t)g->cpw * g->cph;
}

void pfvo_frame_init(const pfvo_geometry *g, uint8_t *frame)
{
    size_t ny = (size_t)g->pw * g->ph, nc = (size_t)g->cpw * g->cph;
    memset(frame, 0, ny);            /* frame.rs:38  plane_y = zeros */
    memset(frame + ny, 128, 2 * nc); /* frame.rs:42-43 */
}

/* ------------------------------------------------------------------------------------
 * q-tables — enc.rs:40-51
 * ---------------------------------------------------------------------------------- */
static int32_t f32_to_i32_trunc(float f) { return (int32_t)f; } /* Rust `as i32`: trunc (values here are small) */

void pfvo_make_qtables(int quality, int32_t out[4][64])
{
    volatile float qscale = (float)quality * 0.25f; /* enc.rs:40 */
    for (int i = 0; i < 64; i++) {
        /* each product rounded to f32 as in Rust; volatile stops any re-association */
        volatile float a, b;
        a = (float)Q_TABLE_INTRA[i] * qscale; b = a * 0.5f;
        out[0][i] = f32_to_i32_trunc(fmaxf(b, 1.0f)); /* enc.rs:50 qtable_intra_l */
        out[1][i] = f32_to_i32_trunc(fmaxf(a, 1.0f)); /* enc.rs:51 qtable_intra_c */
        a = (float)Q_TABLE_INTER[i] * qscale; b = a * 0.5f;
        out[2][i] = f32_to_i32_trunc(fmaxf(b, 1.0f)); /* enc.rs:48 qtable_inter_l */
        out[3][i] = f32_to_i32_trunc(fmaxf(a, 1.0f)); /* enc.rs:49 qtable_inter_c */
    }
}

float pfvo_px_err(int quality) { return (float)quality * 1.5f; } /* enc.rs:41 */

/* ------------------------------------------------------------------------------------
 * 1-D transforms — dct.rs:176-293
 * ---------------------------------------------------------------------------------- */
void pfvo_fdct8(int32_t v[8])
{
    /* dct.rs:178-185 */
    int32_t i0 = v[0], i1 = v[1], i2 = v[2], i3 = v[3], i4 = v[4], i5 = v[5], i6 = v[6], i7 = v[7];
    /* stage 1, dct.rs:188-195 */
    int32_t a0 = i0 + i7, a1 = i1 + i6, a2 = i2 + i5, a3 = i3 + i4;
    int32_t a4 = i0 - i7, a5 = i1 - i6, a6 = i2 - i5, a7 = i3 - i4;
    /* even stage 2, dct.rs:198-201 */
    int32_t b0 = a0 + a3, b1 = a1 + a2, b2 = a0 - a3, b3 = a1 - a2;
    /* even stage 3, dct.rs:204-207 */
    int32_t c0 = b0 + b1;
    int32_t c1 = b0 - b1;
    int32_t c2 = b2 + b2 / 4 + b3 / 2;
    int32_t c3 = b2 / 2 - b3 - b3 / 4;
    /* odd stage 2, dct.rs:211-214 */
    int32_t b4 = a7 / 4 + a4 + a4 / 4 - a4 / 16;
    int32_t b7 = a4 / 4 - a7 - a7 / 4 + a7 / 16;
    int32_t b5 = a5 + a6 - a6 / 4 - a6 / 16;
    int32_t b6 = a6 - a5 + a5 / 4 + a5 / 16;
    /* odd stage 3, dct.rs:217-220 */
    int32_t c4 = b4 + b5, c5 = b4 - b5, c6 = b6 + b7, c7 = b6 - b7;
    /* odd stage 4, dct.rs:223-226 */
    int32_t d4 = c4, d5 = c5 + c7, d6 = c5 - c7, d7 = c6;
    /* permute/output, dct.rs:229-236 */
    v[0] = c0; v[1] = d4; v[2] = c2; v[3] = d6;
    v[4] = c1; v[5] = d5; v[6] = c3; v[7] = d7;
}

void pfvo_idct8(int32_t v[8])
{
    /* dct.rs:243-250 (input permutation) */
    int32_t c0 = v[0], d4 = v[1], c2 = v[2], d6 = v[3], c1 = v[4], d5 = v[5], c3 = v[6], d7 = v[7];
    /* odd stage 4, dct.rs:253-256 */
    int32_t c4 = d4, c5 = d5 + d6, c7 = d5 - d6, c6 = d7;
    /* odd stage 3, dct.rs:259-262 */
    int32_t b4 = c4 + c5, b5 = c4 - c5, b6 = c6 + c7, b7 = c6 - c7;
    /* even stage 3, dct.rs:265-268 */
    int32_t b0 = c0 + c1;
    int32_t b1 = c0 - c1;
    int32_t b2 = c2 + c2 / 4 + c3 / 2;
    int32_t b3 = c2 / 2 - c3 - c3 / 4;
    /* odd stage 2, dct.rs:271-274 */
    int32_t a4 = b7 / 4 + b4 + b4 / 4 - b4 / 16;
    int32_t a7 = b4 / 4 - b7 - b7 / 4 + b7 / 16;
    int32_t a5 = b5 - b6 + b6 / 4 + b6 / 16;
    int32_t a6 = b6 + b5 - b5 / 4 - b5 / 16;
    /* even stage 2, dct.rs:277-280 */
    int32_t a0 = b0 + b2, a1 = b1 + b3, a2 = b1 - b3, a3 = b0 - b2;
    /* stage 1, dct.rs:283-290 */
    v[0] = a0 + a4; v[1] = a1 + a5; v[2] = a2 + a6; v[3] = a3 + a7;
    v[4] = a3 - a7; v[5] = a2 - a6; v[6] = a1 - a5; v[7] = a0 - a4;
}

/* dct.rs:139-145 / 148-154 / 157-163 / 166-172: row and column drivers */
static void rows_apply(int32_t m[64], void (*f)(int32_t *))
{
    for (int r = 0; r < 8; r++) f(&m[r * 8]);
}
static void cols_apply(int32_t m[64], void (*f)(int32_t *))
{
    for (int c = 0; c < 8; c++) {
        int32_t col[8];
        for (int r = 0; r < 8; r++) col[r] = m[c + r * 8]; /* dct.rs:118-128 get_column */
        f(col);
        for (int r = 0; r < 8; r++) m[c + r * 8] = col[r]; /* dct.rs:130-136 set_column */
    }
}

/* dct.rs:88-99 DctMatrix8x8::encode — tables indexed by RASTER position z = ZIGZAG[i] */
void pfvo_quant_encode(const int32_t m[64], const int32_t q[64], int16_t out[64])
{
    for (int i = 0; i < 64; i++) {
        int z = ZIGZAG_TABLE[i];
        int32_t n = (m[z] * DCT_SCALE_FACTOR[z]) >> (FP_BITS * 2); /* dct.rs:92 */
        int32_t d = q[z];                                          /* dct.rs:93 */
        out[i] = (int16_t)(n / d);                                 /* dct.rs:95 (wrapping cast) */
    }
}

/* dct.rs:75-86 DctMatrix8x8::decode — tables indexed by SCAN position idx = INV_ZIGZAG[i] */
void pfvo_quant_decode(const int16_t c[64], const int32_t q[64], int32_t m[64])
{
    for (int i = 0; i < 64; i++) {
        int idx = INV_ZIGZAG_TABLE[i];
        int32_t n = (int32_t)c[idx] * DCT_SCALE_FACTOR[idx]; /* dct.rs:79 */
        int32_t d = q[idx];                                  /* dct.rs:80 */
        m[i] = n * d;                                        /* dct.rs:82 */
    }
}

/* common.rs:287-298 */
void pfvo_encode_subblock(const uint8_t px[64], const int32_t q[64], int16_t out[64])
{
    int32_t m[64];
    for (int i = 0; i < 64; i++) m[i] = ((int32_t)px[i] - 128) * (1 << FP_BITS); /* common.rs:291 */
    rows_apply(m, pfvo_fdct8); /* common.rs:294 */
    cols_apply(m, pfvo_fdct8); /* common.rs:295 */
    pfvo_quant_encode(m, q, out);
}

/* common.rs:300-311 */
void pfvo_encode_subblock_delta(const int16_t d[64], const int32_t q[64], int16_t out[64])
{
    int32_t m[64];
    for (int i = 0; i < 64; i++) m[i] = ((int32_t)d[i] / 2) * (1 << FP_BITS); /* common.rs:304 (trunc /2) */
    rows_apply(m, pfvo_fdct8); /* common.rs:307 */
    cols_apply(m, pfvo_fdct8); /* common.rs:308 */
    pfvo_quant_encode(m, q, out);
}

/* common.rs:313-325 */
void pfvo_decode_subblock(const int16_t c[64], const int32_t q[64], uint8_t out[64])
{
    int32_t m[64];
    pfvo_quant_decode(c, q, m);
    cols_apply(m, pfvo_idct8); /* common.rs:315 */
    rows_apply(m, pfvo_idct8); /* common.rs:316 */
    for (int i = 0; i < 64; i++) {
        int32_t p = (m[i] >> FP_BITS) + 128; /* common.rs:321 */
        out[i] = (uint8_t)(p < 0 ? 0 : (p > 255 ? 255 : p));
    }
}

/* ------------------------------------------------------------------------------------
 * small plane helpers — plane.rs:20-36, common.rs:88-96, 327-349
 * ---------------------------------------------------------------------------------- */
static void get_slice(const uint8_t *plane, int pw, int sx, int sy, int sw, int sh, uint8_t *dst)
{
    for (int r = 0; r < sh; r++) memcpy(dst + r * sw, plane + (size_t)(r + sy) * pw + sx, (size_t)sw);
}

static void blit_block(uint8_t *plane, int pw, const uint8_t blk[256], int dx, int dy) /* common.rs:341-349 */
{
    for (int r = 0; r < 16; r++) memcpy(plane + (size_t)(r + dy) * pw + dx, blk + r * 16, 16);
}

static void blit_subblock(uint8_t mb[256], const uint8_t sb[64], int dx, int dy) /* common.rs:88-96 */
{
    for (int r = 0; r < 8; r++) memcpy(mb + (r + dy) * 16 + dx, sb + r * 8, 8);
}

/* sub-block visiting order (0,0),(8,0),(0,8),(8,8): common.rs:145-149, 239-249 */
static const int SB_X[4] = {0, 8, 0, 8};
static const int SB_Y[4] = {0, 0, 8, 8};

/* common.rs:141-152 */
void pfvo_encode_block(const uint8_t px[256], const int32_t q[64], int16_t out[256])
{
    for (int s = 0; s < 4; s++) {
        uint8_t sb[64];
        get_slice(px, 16, SB_X[s], SB_Y[s], 8, 8, sb);
        pfvo_encode_subblock(sb, q, out + s * 64);
    }
}

/* common.rs:238-252 */
void pfvo_decode_block(const int16_t c[256], const int32_t q[64], uint8_t out[256])
{
    for (int s = 0; s < 4; s++) {
        uint8_t sb[64];
        pfvo_decode_subblock(c + s * 64, q, sb);
        blit_subblock(out, sb, SB_X[s], SB_Y[s]);
    }
}

/* common.rs:125-139: f32 sum of squared differences with early-out */
static float calc_error(const uint8_t a[256], const uint8_t b[256], float ref_lms)
{
    float sum = 0.0f;
    for (int i = 0; i < 256; i++) {
        float diff = (float)a[i] - (float)b[i];
        sum += diff * diff;
        if (sum >= ref_lms) return sum; /* common.rs:133 */
    }
    return sum;
}

/* common.rs:154-204 (recursive, as in the reference) */
float pfvo_block_search(const uint8_t src[256], const uint8_t *ref, int rw, int rh,
                        int cx, int cy, int stepsize, int *dx_out, int *dy_out, uint8_t best[256])
{
    int best_dx = 0, best_dy = 0;
    float best_err = INFINITY;
    uint8_t slice[256];

    /* centre first, common.rs:161-165 */
    get_slice(ref, rw, cx, cy, 16, 16, slice);
    memcpy(best, slice, 256);
    best_err = calc_error(src, slice, best_err);

    for (int my = -1; my < 2; my++) {                  /* common.rs:168 */
        int offsy = cy + my * stepsize;
        if (offsy < 0 || offsy > rh - 16) continue;    /* common.rs:171 */
        for (int mx = -1; mx < 2; mx++) {              /* common.rs:175 */
            if (my == 0 && mx == 0) continue;
            int offsx = cx + mx * stepsize;
            if (offsx < 0 || offsx > rw - 16) continue; /* common.rs:182 */
            get_slice(ref, rw, offsx, offsy, 16, 16, slice);
            float err = calc_error(src, slice, best_err);
            if (err < best_err) {                      /* common.rs:189 strict */
                memcpy(best, slice, 256);
                best_err = err;
                best_dx = mx * stepsize;
                best_dy = my * stepsize;
            }
        }
    }

    if (stepsize > 1) { /* common.rs:198-200 */
        int dx2, dy2;
        float err2 = pfvo_block_search(src, ref, rw, rh, cx + best_dx, cy + best_dy, stepsize / 2, &dx2, &dy2, best);
        *dx_out = best_dx + dx2;
        *dy_out = best_dy + dy2;
        return err2;
    }
    *dx_out = best_dx;
    *dy_out = best_dy;
    return best_err;
}

/* common.rs:108-123 */
static void calc_residuals(const uint8_t from[256], const uint8_t to[256], int16_t out[256])
{
    for (int i = 0; i < 256; i++) {
        int16_t delta = (int16_t)((int16_t)from[i] - (int16_t)to[i]);
        out[i] = delta < -255 ? -255 : (delta > 255 ? 255 : delta);
    }
}

/* common.rs:206-236 */
void pfvo_encode_block_delta(const uint8_t src[256], const uint8_t *ref, int rw, int rh,
                             int bx, int by, const int32_t q[64], float px_err,
                             pfvo_mbhdr *hdr, int16_t out[256])
{
    float min_err = px_err * px_err * 256.0f; /* common.rs:209 */
    int best_dx, best_dy;
    uint8_t prev_block[256];
    float best_err = pfvo_block_search(src, ref, rw, rh, bx, by, 8, &best_dx, &best_dy, prev_block); /* :212 */

    hdr->mx = (int8_t)best_dx;
    hdr->my = (int8_t)best_dy;
    hdr->reserved = 0;
    if (best_err <= min_err) { /* common.rs:221-222 */
        hdr->has_coeff = 0;
        memset(out, 0, 256 * sizeof(int16_t));
        return;
    }
    hdr->has_coeff = 1;
    int16_t delta[256];
    calc_residuals(src, prev_block, delta); /* common.rs:225 */
    for (int s = 0; s < 4; s++) {           /* common.rs:228-232 */
        int16_t sb[64];
        for (int r = 0; r < 8; r++)
            memcpy(sb + r * 8, delta + (r + SB_Y[s]) * 16 + SB_X[s], 8 * sizeof(int16_t));
        pfvo_encode_subblock_delta(sb, q, out + s * 64);
    }
}

/* common.rs:98-104 */
static void apply_residuals(uint8_t blk[256], const uint8_t from[256])
{
    for (int i = 0; i < 256; i++) {
        int16_t d = (int16_t)(((int16_t)blk[i] - 128) * 2);
        int16_t p = (int16_t)from[i];
        int16_t s = (int16_t)(p + d);
        blk[i] = (uint8_t)(s < 0 ? 0 : (s > 255 ? 255 : s));
    }
}

/* common.rs:254-285 */
void pfvo_decode_block_delta(const pfvo_mbhdr *hdr, const int16_t c[256], const uint8_t *ref, int rw, int rh,
                             int bx, int by, const int32_t q[64], uint8_t out[256])
{
    int sx = bx + hdr->mx; /* common.rs:255 */
    int sy = by + hdr->my; /* common.rs:256 */
    (void)rh;
    uint8_t prev_block[256];
    get_slice(ref, rw, sx, sy, 16, 16, prev_block); /* common.rs:261 get_block */
    if (hdr->has_coeff) {
        pfvo_decode_block(c, q, out);         /* common.rs:265-275 */
        apply_residuals(out, prev_block);     /* common.rs:277 */
    } else {
        memcpy(out, prev_block, 256);         /* common.rs:281-283 */
    }
}

/* ------------------------------------------------------------------------------------
 * plane-level loops — common.rs:351-521.  `#pragma omp parallel for` stands where the
 * reference uses rayon par_iter (common.rs:374,411,429,454,481,502).
 * ---------------------------------------------------------------------------------- */
static int clamp_threads(int n)
{
#ifdef _OPENMP
    if (n <= 0) n = omp_get_max_threads();
    return n;
#else
    (void)n;
    return 1;
#endif
}

/* common.rs:352-356: pad with clear colour */
static uint8_t *pad_plane(const uint8_t *src, int w, int h, uint8_t clear_color, int *pw_out, int *ph_out)
{
    int pw = (int)pad16((uint32_t)w), ph = (int)pad16((uint32_t)h);
    uint8_t *img = (uint8_t *)malloc((size_t)pw * ph);
    memset(img, clear_color, (size_t)pw * ph);
    for (int r = 0; r < h; r++) memcpy(img + (size_t)r * pw, src + (size_t)r * w, (size_t)w);
    *pw_out = pw;
    *ph_out = ph;
    return img;
}

void pfvo_encode_plane(const uint8_t *src, int w, int h, const int32_t q[64], uint8_t clear_color,
                       int16_t *coeff_out, int nthreads)
{
    int pw, ph;
    uint8_t *img = pad_plane(src, w, h, clear_color, &pw, &ph);
    int bw = pw / 16, bh = ph / 16, total = bw * bh;
    nthreads = clamp_threads(nthreads);
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int i = 0; i < total; i++) {
        uint8_t blk[256];
        get_slice(img, pw, (i % bw) * 16, (i / bw) * 16, 16, 16, blk); /* common.rs:364-370 */
        pfvo_encode_block(blk, q, coeff_out + (size_t)i * 256);         /* common.rs:374-378 */
    }
    free(img);
}

void pfvo_encode_plane_delta(const uint8_t *src, int w, int h, const uint8_t *ref,
                             const int32_t q[64], float px_err, uint8_t clear_color,
                             pfvo_mbhdr *hdr_out, int16_t *coeff_out, int nthreads)
{
    int pw, ph;
    uint8_t *img = pad_plane(src, w, h, clear_color, &pw, &ph);
    int bw = pw / 16, bh = ph / 16, total = bw * bh;
    nthreads = clamp_threads(nthreads);
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 16)
    for (int i = 0; i < total; i++) {
        uint8_t blk[256];
        int bx = (i % bw) * 16, by = (i / bw) * 16;
        get_slice(img, pw, bx, by, 16, 16, blk);                         /* common.rs:401-407 */
        pfvo_encode_block_delta(blk, ref, pw, ph, bx, by, q, px_err,     /* common.rs:411-413 */
                                &hdr_out[i], coeff_out + (size_t)i * 256);
    }
    free(img);
}

void pfvo_decode_plane_into(const int16_t *coeff, int pw, int ph, const int32_t q[64],
                            uint8_t *target, int nthreads)
{
    int bw = pw / 16, bh = ph / 16, total = bw * bh;
    uint8_t *results = (uint8_t *)malloc((size_t)total * 256);
    nthreads = clamp_threads(nthreads);
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int i = 0; i < total; i++) /* common.rs:481-483 */
        pfvo_decode_block(coeff + (size_t)i * 256, q, results + (size_t)i * 256);
    for (int i = 0; i < total; i++) /* common.rs:490-495 serial blit */
        blit_block(target, pw, results + (size_t)i * 256, (i % bw) * 16, (i / bw) * 16);
    free(results);
}

void pfvo_decode_plane_delta_into(const pfvo_mbhdr *hdr, const int16_t *coeff, int pw, int ph,
                                  const int32_t q[64], uint8_t *refplane, int nthreads)
{
    int bw = pw / 16, bh = ph / 16, total = bw * bh;
    uint8_t *results = (uint8_t *)malloc((size_t)total * 256);
    nthreads = clamp_threads(nthreads);
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int i = 0; i < total; i++) /* common.rs:502-506: every MB reads the OLD plane */
        pfvo_decode_block_delta(&hdr[i], coeff + (size_t)i * 256, refplane, pw, ph,
                                (i % bw) * 16, (i / bw) * 16, q, results + (size_t)i * 256);
    for (int i = 0; i < total; i++) /* common.rs:515-520: then all are written */
        blit_block(refplane, pw, results + (size_t)i * 256, (i % bw) * 16, (i / bw) * 16);
    free(results);
}

/* ------------------------------------------------------------------------------------
 * frame level seam
 * ---------------------------------------------------------------------------------- */
static uint8_t *plane_ptr(const pfvo_geometry *g, uint8_t *frame, int p)
{
    size_t ny = (size_t)g->pw * g->ph, nc = (size_t)g->cpw * g->cph;
    return p == 0 ? frame : (p == 1 ? frame + ny : frame + ny + nc);
}

void pfvo_decode_iframe_coeffs(const pfvo_geometry *g, const int32_t (*qtables)[64], const uint8_t qidx[3],
                               const int16_t *coeff, uint8_t *frame, int nthreads)
{
    /* dec.rs:303-310: Y, U, V consume the sub-block iterator in that order */
    pfvo_decode_plane_into(coeff, (int)g->pw, (int)g->ph, qtables[qidx[0]], plane_ptr(g, frame, 0), nthreads);
    coeff += (size_t)g->nb_y * 256;
    pfvo_decode_plane_into(coeff, (int)g->cpw, (int)g->cph, qtables[qidx[1]], plane_ptr(g, frame, 1), nthreads);
    coeff += (size_t)g->nb_c * 256;
    pfvo_decode_plane_into(coeff, (int)g->cpw, (int)g->cph, qtables[qidx[2]], plane_ptr(g, frame, 2), nthreads);
}

void pfvo_decode_pframe_coeffs(const pfvo_geometry *g, const int32_t (*qtables)[64], const uint8_t qidx[3],
                               const pfvo_mbhdr *hdr, const int16_t *coeff, uint8_t *frame, int nthreads)
{
    /* dec.rs:425-432 */
    pfvo_decode_plane_delta_into(hdr, coeff, (int)g->pw, (int)g->ph, qtables[qidx[0]], plane_ptr(g, frame, 0), nthreads);
    hdr += g->nb_y; coeff += (size_t)g->nb_y * 256;
    pfvo_decode_plane_delta_into(hdr, coeff, (int)g->cpw, (int)g->cph, qtables[qidx[1]], plane_ptr(g, frame, 1), nthreads);
    hdr += g->nb_c; coeff += (size_t)g->nb_c * 256;
    pfvo_decode_plane_delta_into(hdr, coeff, (int)g->cpw, (int)g->cph, qtables[qidx[2]], plane_ptr(g, frame, 2), nthreads);
}

void pfvo_encode_iframe_coeffs(const pfvo_geometry *g, const int32_t (*qtables)[64],
                               const uint8_t *y, const uint8_t *u, const uint8_t *v,
                               int16_t *coeff_out, uint8_t *prev_frame, int nthreads)
{
    /* enc.rs:84-97: encode_plane, decode_plane (closed loop), blit into prev_frame */
    const uint8_t *src[3] = {y, u, v};
    const int qsel[3] = {0, 1, 1};        /* intra_l, intra_c, intra_c */
    const uint8_t clear[3] = {0, 128, 128};
    for (int p = 0; p < 3; p++) {
        int w = p ? (int)g->cwidth : (int)g->width, h = p ? (int)g->cheight : (int)g->height;
        int pw = p ? (int)g->cpw : (int)g->pw, ph = p ? (int)g->cph : (int)g->ph;
        pfvo_encode_plane(src[p], w, h, qtables[qsel[p]], clear[p], coeff_out, nthreads);
        /* decode_plane (common.rs:423-446) writes every MB of a fresh plane, then enc.rs:95-97 blits it
         * over the whole of prev_frame's plane: same as decoding into prev_frame's plane. */
        pfvo_decode_plane_into(coeff_out, pw, ph, qtables[qsel[p]], plane_ptr(g, prev_frame, p), nthreads);
        coeff_out += (size_t)(p ? g->nb_c : g->nb_y) * 256;
    }
}

void pfvo_encode_pframe_coeffs(const pfvo_geometry *g, const int32_t (*qtables)[64], float px_err,
                               const uint8_t *y, const uint8_t *u, const uint8_t *v,
                               pfvo_mbhdr *hdr_out, int16_t *coeff_out, uint8_t *prev_frame, int nthreads)
{
    /* enc.rs:134-147 */
    const uint8_t *src[3] = {y, u, v};
    const int qsel[3] = {2, 3, 3};        /* inter_l, inter_c, inter_c */
    const uint8_t clear[3] = {0, 128, 128};
    for (int p = 0; p < 3; p++) {
        int w = p ? (int)g->cwidth : (int)g->width, h = p ? (int)g->cheight : (int)g->height;
        int pw = p ? (int)g->cpw : (int)g->pw, ph = p ? (int)g->cph : (int)g->ph;
        uint8_t *ref = plane_ptr(g, prev_frame, p);
        pfvo_encode_plane_delta(src[p], w, h, ref, qtables[qsel[p]], px_err, clear[p], hdr_out, coeff_out, nthreads);
        /* decode_plane_delta (common.rs:448-475) then blit over prev_frame == two-phase in-place update */
        pfvo_decode_plane_delta_into(hdr_out, coeff_out, pw, ph, qtables[qsel[p]], ref, nthreads);
        size_t n = p ? g->nb_c : g->nb_y;
        hdr_out += n;
        coeff_out += n * 256;
    }
}

void pfvo_crop_frame(const pfvo_geometry *g, const uint8_t *frame, uint8_t *y, uint8_t *u, uint8_t *v)
{
    /* dec.rs:195-197 */
    const uint8_t *py = plane_ptr(g, (uint8_t *)frame, 0);
    const uint8_t *pu = plane_ptr(g, (uint8_t *)frame, 1);
    const uint8_t *pv = plane_ptr(g, (uint8_t *)frame, 2);
    for (uint32_t r = 0; r < g->height; r++) memcpy(y + (size_t)r * g->width, py + (size_t)r * g->pw, g->width);
    for (uint32_t r = 0; r < g->cheight; r++) {
        memcpy(u + (size_t)r * g->cwidth, pu + (size_t)r * g->cpw, g->cwidth);
        memcpy(v + (size_t)r * g->cwidth, pv + (size_t)r * g->cpw, g->cwidth);
    }
}

/* ------------------------------------------------------------------------------------
 * colour / format helpers next to the path
 * ---------------------------------------------------------------------------------- */
/* Rust `f as u8`: truncation toward zero, saturating, NaN -> 0. */
static uint8_t f32_as_u8(float f)
{
    if (!(f > 0.0f)) return 0;
    if (f >= 255.0f) return 255;
    return (uint8_t)(int)f;
}

/* common.rs:523-536 */
void pfvo_plane_reduce(const uint8_t *src, int w, int h, uint8_t *dst)
{
    const int nw = w / 2, nh = h / 2;
    for (int iy = 0; iy < nh; iy++)
        for (int ix = 0; ix < nw; ix++) dst[ix + iy * nw] = src[ix * 2 + (iy * 2) * w];
}

/* common.rs:538-556 */
void pfvo_plane_double(const uint8_t *src, int w, int h, uint8_t *dst)
{
    const int nw = w * 2;
    for (int iy = 0; iy < h; iy++)
        for (int ix = 0; ix < w; ix++) {
            const uint8_t px = src[ix + iy * w];
            const int d = ix * 2 + (iy * 2) * nw;
            dst[d] = px; dst[d + 1] = px; dst[d + nw] = px; dst[d + nw + 1] = px;
        }
}

/* lib.rs:337-363 load_frame (f32, evaluated left to right, `as u8`) + frame.rs:51-60 from_planes (reduce) */
void pfvo_rgb_to_yuv420(const uint8_t *rgb, int w, int h, uint8_t *y, uint8_t *u, uint8_t *v)
{
    uint8_t *uf = (uint8_t *)malloc((size_t)w * h), *vf = (uint8_t *)malloc((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const float r = (float)rgb[3 * i], g = (float)rgb[3 * i + 1], b = (float)rgb[3 * i + 2];
        const float fy = (0.299f * r) + (0.587f * g) + (0.114f * b);
        const float fu = 128.0f - (0.168736f * r) - (0.331264f * g) + (0.5f * b);
        const float fv = 128.0f + (0.5f * r) - (0.418688f * g) - (0.081312f * b);
        y[i] = f32_as_u8(fy); uf[i] = f32_as_u8(fu); vf[i] = f32_as_u8(fv);
    }
    pfvo_plane_reduce(uf, w, h, u);
    pfvo_plane_reduce(vf, w, h, v);
    free(uf); free(vf);
}

/* lib.rs:365-395 save_frame: double() the chroma planes, f32 conversion, `as u8` */
void pfvo_yuv420_to_rgb(const uint8_t *y, const uint8_t *u, const uint8_t *v, int w, int h, uint8_t *rgb)
{
    uint8_t *uf = (uint8_t *)malloc((size_t)w * h), *vf = (uint8_t *)malloc((size_t)w * h);
    pfvo_plane_double(u, w / 2, h / 2, uf);
    pfvo_plane_double(v, w / 2, h / 2, vf);
    for (size_t i = 0; i < (size_t)w * h; i++) {
        const float fy = (float)y[i], fu = (float)uf[i] - 128.0f, fv = (float)vf[i] - 128.0f;
        const float r = fy + (1.402f * fv);
        const float g = fy - (0.344136f * fu) - (0.714136f * fv);
        const float b = fy + (1.772f * fu);
        rgb[3 * i] = f32_as_u8(r); rgb[3 * i + 1] = f32_as_u8(g); rgb[3 * i + 2] = f32_as_u8(b);
    }
    free(uf); free(vf);
}

/* ------------------------------------------------------------------------------------
 * growable byte buffer + LSB-first bit writer/reader
 * (bitstream-io 1.6.0 LittleEndian semantics as used at the call sites in SURVEY §8c)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    uint8_t *p;
    size_t len, cap;
} bytebuf;

static void bb_reserve(bytebuf *b, size_t extra)
{
    if (b->len + extra <= b->cap) return;
    size_t nc = b->cap ? b->cap * 2 : 4096;
    while (nc < b->len + extra) nc *= 2;
    b->p = (uint8_t *)realloc(b->p, nc);
    b->cap = nc;
}
static void bb_put(bytebuf *b, const void *src, size_t n)
{
    bb_reserve(b, n);
    memcpy(b->p + b->len, src, n);
    b->len += n;
}
static void bb_u8(bytebuf *b, uint8_t v) { bb_put(b, &v, 1); }
static void bb_u16le(bytebuf *b, uint16_t v) { uint8_t t[2] = {(uint8_t)v, (uint8_t)(v >> 8)}; bb_put(b, t, 2); }
static void bb_u32le(bytebuf *b, uint32_t v)
{
    uint8_t t[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)};
    bb_put(b, t, 4);
}

typedef struct {
    bytebuf *out;
    uint32_t acc;  /* pending bits, LSB = oldest */
    uint32_t nacc; /* < 8 */
} bitwriter;

/* BitWriter::<LittleEndian>::write(bits, value): low `bits` bits, least significant first */
static void bw_write(bitwriter *w, uint32_t bits, uint32_t value)
{
    for (uint32_t i = 0; i < bits; i++) {
        w->acc |= ((value >> i) & 1u) << w->nacc;
        if (++w->nacc == 8) {
            bb_u8(w->out, (uint8_t)w->acc);
            w->acc = 0;
            w->nacc = 0;
        }
    }
}
static void bw_write_bit(bitwriter *w, int bit) { bw_write(w, 1, bit ? 1u : 0u); }
/* LittleEndian::write_signed: (bits-1) magnitude bits of the two's-complement value, then the sign bit
 * == the low `bits` bits of the two's-complement value, LSB first. */
static void bw_write_signed(bitwriter *w, uint32_t bits, int32_t value)
{
    bw_write(w, bits - 1, (uint32_t)value & ((1u << (bits - 1)) - 1u));
    bw_write_bit(w, value < 0);
}
static void bw_byte_align(bitwriter *w)
{
    while (w->nacc != 0) bw_write_bit(w, 0);
}

typedef struct {
    const uint8_t *p;
    uint64_t nbits; /* total */
    uint64_t pos;   /* in bits */
} bitreader;

static int br_read(bitreader *r, uint32_t bits, uint32_t *out)
{
    if (r->pos + bits > r->nbits) return -1; /* UnexpectedEof */
    uint32_t v = 0;
    for (uint32_t i = 0; i < bits; i++) {
        uint64_t b = r->pos + i;
        v |= (uint32_t)((r->p[b >> 3] >> (b & 7)) & 1u) << i;
    }
    r->pos += bits;
    *out = v;
    return 0;
}
static int br_read_bit(bitreader *r, int *bit)
{
    uint32_t v;
    if (br_read(r, 1, &v)) return -1;
    *bit = (int)v;
    return 0;
}
/* LittleEndian::read_signed::<i16>(bits) */
static int br_read_signed(bitreader *r, uint32_t bits, int32_t *out)
{
    uint32_t mag;
    int neg;
    if (br_read(r, bits - 1, &mag)) return -1;
    if (br_read_bit(r, &neg)) return -1;
    *out = neg ? (int32_t)mag - (int32_t)(1u << (bits - 1)) : (int32_t)mag;
    return 0;
}

/* ------------------------------------------------------------------------------------
 * rle.rs
 * ---------------------------------------------------------------------------------- */
typedef struct {
    uint8_t num_zeroes;
    uint8_t coeff_size;
    int16_t coeff;
} rle_seq; /* rle.rs:3-7 */

typedef struct {
    rle_seq *p;
    size_t len, cap;
} rle_vec;

static void rv_push(rle_vec *v, uint8_t z, uint8_t s, int16_t c)
{
    if (v->len == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 1024;
        v->p = (rle_seq *)realloc(v->p, v->cap * sizeof(rle_seq));
    }
    v->p[v->len].num_zeroes = z;
    v->p[v->len].coeff_size = s;
    v->p[v->len].coeff = c;
    v->len++;
}

static uint32_t bitlen16(uint16_t c)
{
    uint32_t n = 0;
    while (c) { n++; c >>= 1; }
    return n; /* 16 - leading_zeros */
}

/* rle.rs:9-39 */
static void rle_encode(rle_vec *into, const int16_t *data, size_t n)
{
    uint32_t run = 0;
    for (size_t i = 0; i < n; i++) {
        int16_t val = data[i];
        if (val == 0) {
            run++;
        } else {
            while (run > 15) { rv_push(into, 15, 0, 0); run -= 15; }    /* rle.rs:18-21 */
            uint16_t c = (uint16_t)(val < 0 ? -val : val);              /* rle.rs:23 */
            uint32_t numbits = bitlen16(c) + 1;                          /* rle.rs:24 */
            rv_push(into, (uint8_t)run, (uint8_t)numbits, val);
            run = 0;
        }
    }
    while (run > 15) { rv_push(into, 15, 0, 0); run -= 15; }            /* rle.rs:31-34 */
    if (run > 0) rv_push(into, (uint8_t)run, 0, 0);                     /* rle.rs:36-38 */
}

/* rle.rs:41-47 */
static void update_table(int32_t table[16], const rle_seq *s, size_t n)
{
    for (size_t i = 0; i < n; i++) {
        table[s[i].num_zeroes & 15] += 1;
        table[s[i].coeff_size & 15] += 1;
    }
}

/* ------------------------------------------------------------------------------------
 * huffman.rs
 * ---------------------------------------------------------------------------------- */
typedef struct {
    uint32_t val, len;
    uint8_t symbol;
} hcode; /* huffman.rs:18-22 */

typedef struct hnode {
    uint32_t freq;
    int ch; /* -1 = internal */
    struct hnode *left, *right;
} hnode;

typedef struct {
    hcode codes[16];
    uint8_t table[16];
    hcode dec_table[256];
    hnode *root;
    hnode pool[40];
    int npool;
} htree;

static void assign_codes(const hnode *p, hcode h[16], hcode s) /* huffman.rs:204-217 */
{
    if (p->ch >= 0) {
        s.symbol = (uint8_t)p->ch;
        h[p->ch] = s;
    } else {
        if (p->left) {
            hcode l = s; /* append(false): huffman.rs:30-32 */
            l.len = s.len + 1;
            assign_codes(p->left, h, l);
        }
        if (p->right) {
            hcode r = s; /* append(true) */
            r.val = s.val | (1u << s.len);
            r.len = s.len + 1;
            assign_codes(p->right, h, r);
        }
    }
}

/* huffman.rs:71-119 */
static void htree_from_table(htree *t, const uint8_t table[16])
{
    memset(t, 0, sizeof(*t));
    memcpy(t->table, table, 16);
    hnode *p[40];
    int n = 0;
    for (int ch = 0; ch < 16; ch++) {
        if (table[ch] > 0) {
            hnode *nd = &t->pool[t->npool++];
            nd->freq = table[ch];
            nd->ch = ch;
            nd->left = nd->right = NULL;
            p[n++] = nd;
        }
    }
    /* stable sort, descending by freq (huffman.rs:81; Rust's sort_by is stable) */
    for (int i = 1; i < n; i++) {
        hnode *x = p[i];
        int j = i - 1;
        while (j >= 0 && p[j]->freq < x->freq) { p[j + 1] = p[j]; j--; }
        p[j + 1] = x;
    }
    while (n > 1) {
        hnode *a = p[--n]; /* huffman.rs:84 */
        hnode *b = p[--n]; /* huffman.rs:85 */
        hnode *c = &t->pool[t->npool++];
        c->freq = a->freq + b->freq;
        c->ch = -1;
        c->left = a;
        c->right = b;
        int pos = n; /* huffman.rs:61-69 get_insert_index: first i with c.freq > p[i].freq */
        for (int i = 0; i < n; i++) {
            if (c->freq > p[i]->freq) { pos = i; break; }
        }
        for (int i = n; i > pos; i--) p[i] = p[i - 1];
        p[pos] = c;
        n++;
    }
    if (n == 0) { /* huffman.rs:95-97 HuffmanTree::empty */
        hnode *r = &t->pool[t->npool++];
        r->freq = 0; r->ch = -1; r->left = r->right = NULL;
        t->root = r;
        return;
    }
    t->root = p[0];
    hcode zero = {0, 0, 0};
    assign_codes(t->root, t->codes, zero);
    /* huffman.rs:107-116: 8-bit decode LUT */
    for (uint32_t val = 0; val < 256; val++) {
        for (int i = 0; i < 16; i++) {
            hcode c = t->codes[i];
            if (c.len > 0 && c.len <= 8 && (val & ((1u << c.len) - 1u)) == c.val) {
                t->dec_table[val] = c;
                break;
            }
        }
    }
}

/* rle.rs:49-66 */
static void rle_create_huffman(htree *t, const int32_t table[16])
{
    int32_t max = 0;
    for (int i = 0; i < 16; i++) if (table[i] > max) max = table[i];
    uint8_t w[16];
    for (int i = 0; i < 16; i++) {
        if (table[i] > 0) {
            int32_t v = (table[i] * 255) / max;
            if (v < 1) v = 1;
            w[i] = (uint8_t)v;
        } else {
            w[i] = 0;
        }
    }
    htree_from_table(t, w);
}

/* huffman.rs:125-154.  returns symbol, -1 IO error, -2 decode error */
static int htree_read_slow(const htree *t, bitreader *r)
{
    const hnode *node = t->root;
    for (;;) {
        if (node->ch >= 0) return node->ch;
        int bit;
        if (br_read_bit(r, &bit)) return -1;
        const hnode *nx = bit ? node->right : node->left;
        if (!nx) return -2;
        node = nx;
    }
}

/* huffman.rs:156-197 */
static int htree_read(const htree *t, bitreader *r, uint64_t max_bits)
{
    uint64_t bit_pos = r->pos;
    uint64_t remaining = max_bits - bit_pos;
    uint32_t read_bits = (uint32_t)(remaining < 8 ? remaining : 8);
    uint32_t cur;
    if (br_read(r, read_bits, &cur)) return -1;
    hcode c = t->dec_table[cur & 255];
    if (c.len == 0) {
        r->pos -= read_bits; /* seek back, then slow path */
        return htree_read_slow(t, r);
    }
    r->pos = r->pos - read_bits + c.len; /* rewind = read_bits - len */
    return c.symbol;
}

/* ------------------------------------------------------------------------------------
 * token (de)serialisation shared by I and P packets
 * ---------------------------------------------------------------------------------- */
static int write_tokens(bitwriter *w, const htree *t, const rle_seq *s, size_t n)
{
    for (size_t i = 0; i < n; i++) { /* enc.rs:301-315 / 454-466 */
        if (s[i].coeff_size > 15) return -1; /* reference would panic in update_table */
        hcode z = t->codes[s[i].num_zeroes], b = t->codes[s[i].coeff_size];
        bw_write(w, z.len, z.val);
        bw_write(w, b.len, b.val);
        if (s[i].coeff_size > 0) bw_write_signed(w, s[i].coeff_size, s[i].coeff);
    }
    return 0;
}

/* dec.rs:261-296 (I) and 383-414 (P): fill `n` coefficients */
static int read_tokens(bitreader *r, const htree *t, uint64_t nbits, int16_t *out, size_t n)
{
    size_t out_idx = 0;
    while (out_idx < n) {
        int nz = htree_read(t, r, nbits);
        if (nz < 0) return -1;
        out_idx += (size_t)nz;
        int nb = htree_read(t, r, nbits);
        if (nb < 0) return -1;
        if (nb > 0) {
            int32_t c;
            if (br_read_signed(r, (uint32_t)nb, &c)) return -1;
            if (out_idx >= n) return -1; /* reference: index out of bounds panic */
            out[out_idx] = (int16_t)c;
            out_idx++;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------
 * Encoder — enc.rs
 * ---------------------------------------------------------------------------------- */
static const uint8_t PFV_MAGIC[8] = {'P', 'F', 'V', 'I', 'D', 'E', 'O', 0}; /* common.rs:1 */
#define PFV_VERSION 211u                                                    /* common.rs:2 */

struct pfvo_encoder {
    pfvo_geometry g;
    int framerate, nthreads, finished;
    float px_err;
    int32_t qtables[4][64]; /* intra_l, intra_c, inter_l, inter_c */
    uint8_t *prev_frame;
    int16_t *coeffs;
    pfvo_mbhdr *hdrs;
    bytebuf out;
};

pfvo_encoder *pfvo_encoder_new(int width, int height, int framerate, int quality, int nthreads)
{
    if (quality < 0 || quality > 10) return NULL; /* enc.rs:38 */
    if (width <= 0 || height <= 0 || (width & 1) || (height & 1)) return NULL;
    pfvo_encoder *e = (pfvo_encoder *)calloc(1, sizeof(*e));
    pfvo_geometry_for((uint32_t)width, (uint32_t)height, &e->g);
    e->framerate = framerate;
    e->nthreads = nthreads;
    e->px_err = pfvo_px_err(quality);
    pfvo_make_qtables(quality, e->qtables);
    e->prev_frame = (uint8_t *)malloc(pfvo_frame_bytes(&e->g));
    pfvo_frame_init(&e->g, e->prev_frame); /* enc.rs:46 VideoFrame::new_padded */
    e->coeffs = (int16_t *)calloc((size_t)e->g.nb * 256, sizeof(int16_t));
    e->hdrs = (pfvo_mbhdr *)calloc(e->g.nb, sizeof(pfvo_mbhdr));
    /* write_header, enc.rs:190-219 */
    bb_put(&e->out, PFV_MAGIC, 8);
    bb_u32le(&e->out, PFV_VERSION);
    bb_u16le(&e->out, (uint16_t)width);
    bb_u16le(&e->out, (uint16_t)height);
    bb_u16le(&e->out, (uint16_t)framerate);
    bb_u16le(&e->out, 4);
    for (int t = 0; t < 4; t++)
        for (int i = 0; i < 64; i++) bb_u16le(&e->out, (uint16_t)e->qtables[t][i]);
    return e;
}

/* enc.rs:237-330 */
static int write_iframe_packet(pfvo_encoder *e)
{
    bytebuf pkt = {0};
    bitwriter w = {&pkt, 0, 0};
    rle_vec seq = {0};
    size_t *starts = (size_t *)malloc(((size_t)e->g.nb + 1) * sizeof(size_t));
    int32_t symbol_table[16] = {0};
    for (uint32_t b = 0; b < e->g.nb; b++) { /* Y blocks, then U, then V: enc.rs:246-283 */
        starts[b] = seq.len;
        rle_encode(&seq, e->coeffs + (size_t)b * 256, 256);
        update_table(symbol_table, seq.p + starts[b], seq.len - starts[b]);
    }
    starts[e->g.nb] = seq.len;
    htree tree;
    rle_create_huffman(&tree, symbol_table);
    for (int i = 0; i < 16; i++) bw_write(&w, 8, tree.table[i]); /* enc.rs:290-292 */
    bw_write(&w, 8, 0); bw_write(&w, 8, 1); bw_write(&w, 8, 1);  /* enc.rs:296-298 */
    int rc = write_tokens(&w, &tree, seq.p, seq.len);
    bw_byte_align(&w); /* enc.rs:318 */
    if (rc == 0) {
        bb_u8(&e->out, 1); /* enc.rs:325 */
        bb_u32le(&e->out, (uint32_t)pkt.len);
        bb_put(&e->out, pkt.p, pkt.len);
    }
    free(pkt.p); free(seq.p); free(starts);
    return rc;
}

/* enc.rs:332-481 */
static int write_pframe_packet(pfvo_encoder *e)
{
    bytebuf pkt = {0};
    bitwriter w = {&pkt, 0, 0};
    rle_vec seq = {0};
    int32_t symbol_table[16] = {0};
    for (uint32_t b = 0; b < e->g.nb; b++) { /* enc.rs:341-396: only blocks with coefficients */
        if (!e->hdrs[b].has_coeff) continue;
        size_t start = seq.len;
        rle_encode(&seq, e->coeffs + (size_t)b * 256, 256);
        update_table(symbol_table, seq.p + start, seq.len - start);
    }
    htree tree;
    rle_create_huffman(&tree, symbol_table);
    for (int i = 0; i < 16; i++) bw_write(&w, 8, tree.table[i]); /* enc.rs:403-405 */
    bw_write(&w, 8, 2); bw_write(&w, 8, 3); bw_write(&w, 8, 3);  /* enc.rs:409-411 */
    for (uint32_t b = 0; b < e->g.nb; b++) { /* enc.rs:414-451 block headers */
        const pfvo_mbhdr *h = &e->hdrs[b];
        int has_mvec = h->mx != 0 || h->my != 0;
        bw_write_bit(&w, has_mvec);
        bw_write_bit(&w, h->has_coeff != 0);
        if (has_mvec) {
            bw_write_signed(&w, 7, h->mx);
            bw_write_signed(&w, 7, h->my);
        }
    }
    int rc = write_tokens(&w, &tree, seq.p, seq.len);
    bw_byte_align(&w); /* enc.rs:469 */
    if (rc == 0) {
        bb_u8(&e->out, 2); /* enc.rs:476 */
        bb_u32le(&e->out, (uint32_t)pkt.len);
        bb_put(&e->out, pkt.p, pkt.len);
    }
    free(pkt.p); free(seq.p);
    return rc;
}

int pfvo_encoder_encode_iframe(pfvo_encoder *e, const uint8_t *y, const uint8_t *u, const uint8_t *v)
{
    if (e->finished) return -1; /* enc.rs:80 */
    pfvo_encode_iframe_coeffs(&e->g, (const int32_t (*)[64])e->qtables, y, u, v, e->coeffs, e->prev_frame, e->nthreads);
    memset(e->hdrs, 0, e->g.nb * sizeof(pfvo_mbhdr));
    return write_iframe_packet(e);
}

int pfvo_encoder_encode_pframe(pfvo_encoder *e, const uint8_t *y, const uint8_t *u, const uint8_t *v)
{
    if (e->finished) return -1; /* enc.rs:130 */
    pfvo_encode_pframe_coeffs(&e->g, (const int32_t (*)[64])e->qtables, e->px_err, y, u, v,
                              e->hdrs, e->coeffs, e->prev_frame, e->nthreads);
    return write_pframe_packet(e);
}

int pfvo_encoder_encode_dropframe(pfvo_encoder *e)
{
    if (e->finished) return -1;
    bb_u8(&e->out, 1); /* enc.rs:229-235 */
    bb_u32le(&e->out, 0);
    return 0;
}

int pfvo_encoder_finish(pfvo_encoder *e)
{
    if (e->finished) return -1; /* enc.rs:183 */
    e->finished = 1;
    bb_u8(&e->out, 0); /* enc.rs:221-227 */
    bb_u32le(&e->out, 0);
    return 0;
}

const uint8_t *pfvo_encoder_bytes(const pfvo_encoder *e, size_t *len) { *len = e->out.len; return e->out.p; }
const int16_t *pfvo_encoder_last_coeffs(const pfvo_encoder *e) { return e->coeffs; }
const pfvo_mbhdr *pfvo_encoder_last_headers(const pfvo_encoder *e) { return e->hdrs; }
const uint8_t *pfvo_encoder_prev_frame(const pfvo_encoder *e) { return e->prev_frame; }

void pfvo_encoder_free(pfvo_encoder *e)
{
    if (!e) return;
    free(e->prev_frame); free(e->coeffs); free(e->hdrs); free(e->out.p); free(e);
}

/* ------------------------------------------------------------------------------------
 * Decoder — dec.rs
 * ---------------------------------------------------------------------------------- */
struct pfvo_decoder {
    const uint8_t *data;
    size_t len, pos, reset_pos;
    pfvo_geometry g;
    int framerate, nthreads, eof, nq;
    int32_t (*qtables)[64];
    uint8_t *framebuffer;
    int16_t *coeffs;
    pfvo_mbhdr *hdrs;
    uint8_t qidx[3];
    int last_kind;
};

static int rd_bytes(pfvo_decoder *d, void *dst, size_t n)
{
    if (d->pos + n > d->len) return -1;
    memcpy(dst, d->data + d->pos, n);
    d->pos += n;
    return 0;
}
static int rd_u16(pfvo_decoder *d, uint16_t *v)
{
    uint8_t t[2];
    if (rd_bytes(d, t, 2)) return -1;
    *v = (uint16_t)(t[0] | (t[1] << 8));
    return 0;
}
static int rd_u32(pfvo_decoder *d, uint32_t *v)
{
    uint8_t t[4];
    if (rd_bytes(d, t, 4)) return -1;
    *v = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
    return 0;
}

pfvo_decoder *pfvo_decoder_new(const uint8_t *data, size_t len, int nthreads, int *err)
{
    pfvo_decoder *d = (pfvo_decoder *)calloc(1, sizeof(*d));
    d->data = data; d->len = len; d->nthreads = nthreads;
    *err = 0;
    uint8_t magic[8];
    uint32_t ver;
    uint16_t w, h, fps, nq;
    if (rd_bytes(d, magic, 8)) { *err = 3; goto fail; }                 /* dec.rs:40-46 */
    if (memcmp(magic, PFV_MAGIC, 8) != 0) { *err = 1; goto fail; }      /* dec.rs:48-52 */
    if (rd_u32(d, &ver)) { *err = 3; goto fail; }
    if (ver != PFV_VERSION) { *err = 2; goto fail; }                    /* dec.rs:55-66 */
    if (rd_u16(d, &w) || rd_u16(d, &h) || rd_u16(d, &fps) || rd_u16(d, &nq)) { *err = 3; goto fail; }
    d->nq = nq;
    d->qtables = (int32_t (*)[64])calloc(nq ? nq : 1, sizeof(int32_t[64]));
    for (int t = 0; t < nq; t++)
        for (int i = 0; i < 64; i++) { /* dec.rs:98-111 */
            uint16_t v;
            if (rd_u16(d, &v)) { *err = 3; goto fail; }
            d->qtables[t][i] = v;
        }
    d->reset_pos = d->pos; /* dec.rs:113 */
    d->framerate = fps;
    pfvo_geometry_for(w, h, &d->g);
    d->framebuffer = (uint8_t *)malloc(pfvo_frame_bytes(&d->g));
    pfvo_frame_init(&d->g, d->framebuffer); /* dec.rs:123 */
    d->coeffs = (int16_t *)calloc((size_t)d->g.nb * 256, sizeof(int16_t));
    d->hdrs = (pfvo_mbhdr *)calloc(d->g.nb, sizeof(pfvo_mbhdr));
    return d;
fail:
    pfvo_decoder_free(d);
    return NULL;
}

int pfvo_decoder_width(const pfvo_decoder *d) { return (int)d->g.width; }
int pfvo_decoder_height(const pfvo_decoder *d) { return (int)d->g.height; }
int pfvo_decoder_framerate(const pfvo_decoder *d) { return d->framerate; }

int pfvo_decoder_reset(pfvo_decoder *d)
{
    d->eof = 0; /* dec.rs:148-152: framebuffer is NOT cleared */
    d->pos = d->reset_pos;
    return 0;
}

/* dec.rs:226-326 */
static int decode_iframe(pfvo_decoder *d, const uint8_t *payload, size_t plen)
{
    bitreader r = {payload, (uint64_t)plen * 8, 0};
    uint8_t table[16];
    uint32_t v;
    for (int i = 0; i < 16; i++) { if (br_read(&r, 8, &v)) return -1; table[i] = (uint8_t)v; }
    htree tree;
    htree_from_table(&tree, table);
    for (int i = 0; i < 3; i++) {
        if (br_read(&r, 8, &v)) return -1;
        if ((int)v >= d->nq) return -1; /* reference: index panic */
        d->qidx[i] = (uint8_t)v;
    }
    size_t ncoeff = (size_t)d->g.nb * 256;
    memset(d->coeffs, 0, ncoeff * sizeof(int16_t)); /* dec.rs:258 */
    if (read_tokens(&r, &tree, r.nbits, d->coeffs, ncoeff)) return -1;
    memset(d->hdrs, 0, d->g.nb * sizeof(pfvo_mbhdr));
    pfvo_decode_iframe_coeffs(&d->g, (const int32_t (*)[64])d->qtables, d->qidx, d->coeffs, d->framebuffer, d->nthreads);
    d->last_kind = 1;
    return 0;
}

/* dec.rs:328-448 */
static int decode_pframe(pfvo_decoder *d, const uint8_t *payload, size_t plen)
{
    bitreader r = {payload, (uint64_t)plen * 8, 0};
    uint8_t table[16];
    uint32_t v;
    for (int i = 0; i < 16; i++) { if (br_read(&r, 8, &v)) return -1; table[i] = (uint8_t)v; }
    htree tree;
    htree_from_table(&tree, table);
    for (int i = 0; i < 3; i++) {
        if (br_read(&r, 8, &v)) return -1;
        if ((int)v >= d->nq) return -1;
        d->qidx[i] = (uint8_t)v;
    }
    for (uint32_t b = 0; b < d->g.nb; b++) { /* dec.rs:361-372 */
        int has_mvec, has_coeff;
        pfvo_mbhdr h = {0, 0, 0, 0};
        if (br_read_bit(&r, &has_mvec) || br_read_bit(&r, &has_coeff)) return -1;
        h.has_coeff = (uint8_t)has_coeff;
        if (has_mvec) {
            int32_t mx, my;
            if (br_read_signed(&r, 7, &mx) || br_read_signed(&r, 7, &my)) return -1;
            h.mx = (int8_t)mx;
            h.my = (int8_t)my;
        }
        d->hdrs[b] = h;
    }
    memset(d->coeffs, 0, (size_t)d->g.nb * 256 * sizeof(int16_t)); /* dec.rs:376 */
    for (uint32_t b = 0; b < d->g.nb; b++) {                         /* dec.rs:378-417 */
        if (!d->hdrs[b].has_coeff) continue;
        if (read_tokens(&r, &tree, r.nbits, d->coeffs + (size_t)b * 256, 256)) return -1;
    }
    pfvo_decode_pframe_coeffs(&d->g, (const int32_t (*)[64])d->qtables, d->qidx, d->hdrs, d->coeffs, d->framebuffer, d->nthreads);
    d->last_kind = 2;
    return 0;
}

int pfvo_decoder_advance_frame(pfvo_decoder *d, uint8_t *y, uint8_t *u, uint8_t *v, int *frame_out)
{
    *frame_out = 0;
    if (d->eof) return 0; /* dec.rs:171-173 */
    for (;;) {
        uint8_t type;
        uint32_t plen;
        if (rd_bytes(d, &type, 1) || rd_u32(d, &plen)) return -3; /* dec.rs:179-180 `?` */
        if (type == 0) { /* dec.rs:183-187 */
            d->eof = 1;
            return 0;
        } else if (type == 1) { /* dec.rs:188-202 */
            if (plen > 0) {
                if (d->pos + plen > d->len) return -3;
                const uint8_t *payload = d->data + d->pos;
                d->pos += plen;
                if (decode_iframe(d, payload, plen)) return -3;
                pfvo_crop_frame(&d->g, d->framebuffer, y, u, v);
                *frame_out = 1;
            }
            break;
        } else if (type == 2) { /* dec.rs:203-215 */
            if (d->pos + plen > d->len) return -3;
            const uint8_t *payload = d->data + d->pos;
            d->pos += plen;
            if (decode_pframe(d, payload, plen)) return -3;
            pfvo_crop_frame(&d->g, d->framebuffer, y, u, v);
            *frame_out = 1;
            break;
        } else { /* dec.rs:216-219: skip unknown packet */
            if (d->pos + plen > d->len) return -3;
            d->pos += plen;
        }
    }
    return 1;
}

int pfvo_decoder_last_kind(const pfvo_decoder *d) { return d->last_kind; }
const int16_t *pfvo_decoder_last_coeffs(const pfvo_decoder *d) { return d->coeffs; }
const pfvo_mbhdr *pfvo_decoder_last_headers(const pfvo_decoder *d) { return d->hdrs; }
const uint8_t *pfvo_decoder_last_qidx(const pfvo_decoder *d) { return d->qidx; }
const uint8_t *pfvo_decoder_framebuffer(const pfvo_decoder *d) { return d->framebuffer; }
const int32_t *pfvo_decoder_qtables(const pfvo_decoder *d, int *nq) { *nq = d->nq; return &d->qtables[0][0]; }

void pfvo_decoder_free(pfvo_decoder *d)
{
    if (!d) return;
    free(d->qtables); free(d->framebuffer); free(d->coeffs); free(d->hdrs); free(d);
}

/* rle.rs:9-39 + rle.rs:41-47 as a callable: the RLE sequence of `data` (one call per macroblock, enc.rs:256-262 / :370-378) and,
 * added to table[16], its symbol counts.  Returns the sequence length; writes at most cap entries. */
size_t pfvo_rle_encode(const int16_t *data, size_t n, uint8_t *num_zeroes, uint8_t *coeff_size, int16_t *coeff, size_t cap,
                       int32_t table[16])
{
    rle_vec seq = {0};
    rle_encode(&seq, data, n);
    if (table) update_table(table, seq.p, seq.len);
    for (size_t i = 0; i < seq.len && i < cap; i++) {
        num_zeroes[i] = seq.p[i].num_zeroes;
        coeff_size[i] = seq.p[i].coeff_size;
        coeff[i] = seq.p[i].coeff;
    }
    size_t len = seq.len;
    free(seq.p);
    return len;
}

/* lib.rs:96-158 / 160-239: the reference's only asserting tests, as a callable */
long pfvo_entropy_roundtrip(const int16_t *data, size_t n, int16_t *decoded)
{
    rle_vec seq = {0};
    rle_encode(&seq, data, n);
    int32_t table[16] = {0};
    update_table(table, seq.p, seq.len);
    htree tree;
    rle_create_huffman(&tree, table);
    bytebuf buf = {0};
    bitwriter w = {&buf, 0, 0};
    if (write_tokens(&w, &tree, seq.p, seq.len)) { free(seq.p); free(buf.p); return -3; }
    bw_byte_align(&w);
    bitreader r = {buf.p, (uint64_t)buf.len * 8, 0};
    memset(decoded, 0, n * sizeof(int16_t));
    /* lib.rs:216-234: also checks each token read equals the token written */
    size_t out_idx = 0, run_idx = 0;
    long rc = (long)buf.len;
    while (out_idx < n) {
        int nz = htree_read(&tree, &r, r.nbits);
        int nb = nz < 0 ? -1 : (out_idx += (size_t)nz, htree_read(&tree, &r, r.nbits));
        if (nz < 0 || nb < 0 || run_idx >= seq.len ||
            seq.p[run_idx].num_zeroes != nz || seq.p[run_idx].coeff_size != nb) { rc = -4; break; }
        if (nb > 0) {
            int32_t c;
            if (br_read_signed(&r, (uint32_t)nb, &c) || out_idx >= n) { rc = -5; break; }
            decoded[out_idx++] = (int16_t)c;
        }
        run_idx++;
    }
    if (rc >= 0 && memcmp(decoded, data, n * sizeof(int16_t)) != 0) rc = -6;
    free(seq.p); free(buf.p);
    return rc;
}
