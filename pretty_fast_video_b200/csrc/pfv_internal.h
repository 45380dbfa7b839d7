// pfv_internal.h — structures shared by the host engine (pfv_ctx.cu) and the kernels (pfv_kernels.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pfv_b200.h"

namespace pfv {

// Derived tables, one set per q-table, built on the host by pfv_ctx_create.  "Lane transposed":
// entry [c*8 + row] belongs to column c, so a lane reads its 8 values as two 128-bit loads.
struct QTables {
    int32_t  deqT[64];   // decode: SCALE[s]*q[s] with s = INV_ZIGZAG[row*8+c]   (wrapping product)
    uint32_t encM[64];   // encode: ceil(2^31 / q[row*8+c])
};

// One padded plane of the frame (src/frame.rs:28-49), in macroblock units and bytes.
struct PlaneGeom {
    uint32_t pw, ph;          // padded size in pixels
    uint32_t vw, vh;          // visible size in pixels (tight source planes of the encoder)
    uint32_t bw, bh;          // macroblocks per row / column
    uint32_t mb_base;         // index of the plane's first macroblock in the frame (Y, U, V order)
    uint32_t off;             // byte offset of the plane inside a frame slot
    float    rcp_bw;          // 1.0f / bw
    uint32_t tiles_per_row;   // ceil(bw / 8): encode-P tiles of 8 horizontally adjacent macroblocks
    uint32_t tile_base;       // index of the plane's first tile
    float    rcp_tiles_per_row;   // 1.0f / tiles_per_row
    uint32_t clear4;          // clear colour replicated in 4 bytes: 0 for Y, 0x80808080 for U,V (src/enc.rs:84-90)
};

struct FrameGeom {
    PlaneGeom pl[3];
    uint32_t  nb;             // macroblocks per frame
    uint32_t  total_tiles;    // encode-P tiles per frame
    uint32_t  frame_bytes;
};

// Device-side job records (built on the host per submit, copied with the job's headers).
struct DecJob {
    const int16_t   *coeff;   // nb*256, device
    const pfv_mbhdr *hdr;     // nb, device (P only)
    uint8_t         *dst;     // frame slot
    const uint8_t   *ref;     // frame slot (P only)
    const QTables   *qt[3];   // per plane
    int32_t          ref_slot; // the same slot as an index (TMA coordinate)
};

// sparse coefficient transport (pfv_decode_submit_sparse): one frame's token lists and where to expand them
struct SparseJob {
    const uint32_t  *mb_off;  // nb + 1, device
    const uint32_t  *tok;     // device
    const pfv_mbhdr *hdr;     // P: nb headers (macroblocks without coefficients are not expanded); I: nullptr
    int16_t         *coeff;   // nb*256 dense destination, device
};

// sparse encode transport (pfv_encode_submit_sparse): where one frame's dense coefficients are and where its RLE sequence goes
struct TokJob {
    const int16_t   *coeff;      // nb*256, device (what the encode kernels just wrote)
    const pfv_mbhdr *hdr;        // P: nb headers; I: nullptr
    uint32_t        *mb_off;     // nb + 1, device staging
    uint32_t        *tok;        // nb*256, device staging
    uint32_t        *stats;      // PFV_TOKSTATS_WORDS, device staging
    uint32_t        *out_tok;    // caller's buffers as device-accessible addresses (pinned host or device memory)
    uint32_t        *out_stats;
    uint32_t        *out_mb_off; // may be nullptr
    uint32_t         tok_cap;
    uint32_t         padded;     // key frames: macroblock m's entries sit at tok[m * 256 ..] (tok_emit_sb_kernel), gathered by the store
};

struct EncJob {
    const uint8_t *src[3];    // tight source planes, device
    pfv_mbhdr     *hdr;       // nb, device (P only)
    int16_t       *coeff;     // nb*256, device
    uint8_t       *dst;       // frame slot receiving the reconstruction
    const uint8_t *ref;       // frame slot holding prev_frame (P only)
    int32_t        ref_slot;  // the same slot as an index (TMA coordinate)
    float          min_err;   // px_err*px_err*256 (src/common.rs:209)
    uint32_t      *mb_cnt;    // sparse encode seam: nb entry counts of rle_encode per macroblock (0 = none), else nullptr
};

// Parameters of the register-resident sub-block kernels (pfv_kernels_sb.cu): the three per-plane tables of
// ONE q-index triple travel as kernel parameters so the multiplies take them as constant-bank operands.
constexpr uint32_t SB_MBS_PER_CTA = 32;
struct SbParams {
    FrameGeom g;
    int32_t   deq[3][64];     // per plane, by SCAN position s: SCALE[s]*q[s] (wrapping)
    uint32_t  cta_base[3];    // first CTA of each plane (a CTA = 32 consecutive macroblocks of one plane)
    uint32_t  cta_total;
    uint32_t  tiles_per_warp; // streaming kernels: consecutive 8-macroblock tiles one warp walks
};

// Parameters of the thread-per-sub-block encode kernels (pfv_kernels_enc.cu): the tables of ONE (luma, chroma) pair travel
// as kernel parameters so that the quantiser's multiplies take them as constant-bank operands.
struct EncSbParams {
    FrameGeom g;
    float     encR[2][64];    // luma / chroma: quant_recip_f32(q) by RASTER position (src/dct.rs:93-95)
    int32_t   deq[2][64];     // luma / chroma: SCALE[s]*q[s] by SCAN position (the closed-loop reconstruction)
    uint32_t  cta_base[3];
    uint32_t  cta_total;
    uint32_t  tiles_per_warp;
};

// fused decode-P kernel (pfv_kernels_pf.cu): macroblock rows per window and the TMA box that holds every predictor of them
constexpr int PF_ROWS = 3;
#ifndef PFV_PF_WIN_W
#define PFV_PF_WIN_W 160
#endif
constexpr int PF_WIN_W = PFV_PF_WIN_W;                       // 16 + 8*16 + 15, rounded up to 16 (160) - or 176
constexpr int PF_WIN_H = PF_ROWS * 16 + 30;

// decode-P window items of one frame: a window covers 8 x `rows` macroblocks of one plane
struct McWin {
    uint32_t base[3];         // first item of each plane
    uint32_t tiles_x[3];      // windows per row of windows
    float    rcp_tiles_x[3];
    uint32_t total;
};
inline McWin make_mc_windows(const FrameGeom &g, uint32_t rows)
{
    McWin W;
    uint32_t t = 0;
    for (int p = 0; p < 3; p++) {
        W.base[p] = t;
        W.tiles_x[p] = (g.pl[p].bw + 7u) / 8u;
        W.rcp_tiles_x[p] = 1.0f / (float)W.tiles_x[p];
        t += W.tiles_x[p] * ((g.pl[p].bh + rows - 1) / rows);
    }
    W.total = t;
    return W;
}

// error bits the kernels OR into the context's device error word
enum { ERRBIT_BAD_MV = 1, ERRBIT_TIMEOUT = 2 };

// search window of an encode-P tile: 16 px left margin + 8 macroblocks + 15 px right margin, rounded so the
// row pitch (44 words) spreads the eight rows of a sub-block over distinct banks; 15 + 16 + 15 rows.
constexpr int WIN_W = 176;
constexpr int WIN_H = 46;
constexpr int WIN_BYTES = WIN_W * WIN_H;
// The same window for the warp-per-tile kernel (encode_p2_kernel): 16 + 128 + 15 = 159 -> 160 bytes per row and no more.
// A row pitch of 40 words moves a macroblock's four banks by 8 per row: after the first search levels the eight macroblocks
// of a warp read rows of their own (centres 4, 2, 1 rows apart), and with 44 words (12 banks per row) almost any two of them
// met in a bank - ncu: 2.2 - 2.5 wavefronts per load at the fine levels.
#ifndef PFV_EP2_WIN_W
#define PFV_EP2_WIN_W 160
#endif
constexpr int EP2_WIN_W = PFV_EP2_WIN_W;
constexpr int EP2_WIN_BYTES = EP2_WIN_W * WIN_H;
constexpr int WINP_W = 188;                  // re-pitched window: 47 words per row (odd: rows spread over all banks)
constexpr int WINP_BYTES = WINP_W * WIN_H;

// true the first time it is called with the calling thread's current device (per-device one-time setup such as
// cudaFuncSetAttribute; a race between two host threads only repeats the setup)
inline bool first_use_on_device(bool (&done)[64])
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
}

// records the calling thread's error text (pfv_last_error) and returns `code`
int set_error(int code, const char *fmt, ...) __attribute__((format(printf, 2, 3)));

// kernel launchers (pfv_kernels.cu).  mbs_per_warp-style tuning lives inside.
cudaError_t launch_decode(bool inter, const FrameGeom &g, const DecJob *d_jobs, uint32_t njobs,
                          int *d_err, cudaStream_t s);
cudaError_t launch_decode_i_sb(const SbParams &P, const DecJob *d_jobs, uint32_t njobs, cudaStream_t s);
cudaError_t launch_decode_i_stream(SbParams P, const DecJob *d_jobs, uint32_t njobs, cudaStream_t s);
// frames flagged dense: cp.async-staged tiles, every sub-block through the full transform (pfv_kernels_sb.cu)
cudaError_t launch_decode_i_direct(SbParams P, const DecJob *d_jobs, uint32_t njobs, cudaStream_t s);
// window copy + list-driven residual pass (pfv_kernels_p.cu); d_done != nullptr: the residual kernel clears d_counts itself
cudaError_t launch_decode_p_two_pass4(const SbParams &P, const DecJob *d_jobs, uint32_t njobs, uint32_t *d_lists, uint32_t *d_counts,
                                      int *d_err, const CUtensorMap &tm_luma, const CUtensorMap &tm_chroma, cudaStream_t s,
                                      uint32_t *d_done);
// the fused warp-specialised decode-P kernel (pfv_kernels_pf.cu)
cudaError_t launch_decode_p_fused(const SbParams &P, const DecJob *d_jobs, uint32_t njobs, int *d_err,
                                  const CUtensorMap &tm_luma, const CUtensorMap &tm_chroma, cudaStream_t s);
// completes cta_base / cta_total / tiles_per_warp of the streaming kernels (pfv_kernels_sb.cu)
void sbw_split(SbParams &P, uint32_t njobs, uint32_t warps_per_cta, uint32_t resident_ctas, uint32_t max_tpw, uint32_t start_cost, uint32_t forced_tpw);
cudaError_t launch_expand_tokens(uint32_t nb, const SparseJob *d_jobs, uint32_t njobs, cudaStream_t s);
// the first n_key jobs are key frames (thread-per-sub-block emit into padded slots, then the scan), the others P frames
// (scan of the counts the encode kernel left, then the warp-per-macroblock emit)
cudaError_t launch_tokenize(uint32_t nb, const TokJob *d_jobs, uint32_t n_key, uint32_t njobs, cudaStream_t s);
cudaError_t launch_token_store(uint32_t nb, const TokJob *d_jobs, uint32_t njobs, cudaStream_t s);
cudaError_t launch_rgb_to_yuv420(const uint8_t *d_rgb, uint32_t w, uint32_t h, uint8_t *d_y, uint8_t *d_u, uint8_t *d_v, cudaStream_t s);
cudaError_t launch_yuv420_to_rgb_batch(const uint8_t *d_pool, size_t slot_stride, uint32_t off_u, uint32_t off_v, const uint32_t *slots,
                                       uint32_t n, uint32_t w, uint32_t h, uint32_t pw, uint32_t cpw, uint8_t *d_out, size_t out_stride,
                                       cudaStream_t s);
// count: every job of the launch has EncJob::mb_cnt set (sparse encode seam)
// persistent, chunks of tiles handed out by a device counter; d_work: two zeroed device words (left zeroed)
cudaError_t launch_encode_i_persist(EncSbParams P, const EncJob *d_jobs, uint32_t njobs, uint32_t *d_work, cudaStream_t s);
cudaError_t launch_encode_i(const FrameGeom &g, const EncJob *d_jobs, uint32_t njobs,
                            const QTables *d_qt, bool count, cudaStream_t s);
// variant 0: warp per tile, column-strip search (default); 1: first generation (CTA per tile, warp per macroblock)
// d_work: two zeroed device words (the default kernel's tile counter; it leaves them zeroed)
cudaError_t launch_encode_p(const FrameGeom &g, const EncJob *d_jobs, uint32_t njobs,
                            const QTables *d_qt, const CUtensorMap &tm_luma, const CUtensorMap &tm_chroma,
                            bool count, int variant, uint32_t *d_work, cudaStream_t s);

}  // namespace pfv

// pfv_ctx.cu, for the codec layer (pfv_codec.cpp): allocate the lazily allocated staging buffers now
struct pfv_ctx;
int pfv_ctx_reserve_staging(pfv_ctx *c, bool decode_sparse, bool encode);
// pfv_ctx_wait_submit that polls the event for up to spin_s seconds before it blocks (for the one thread that returns the frame)
int pfv_ctx_wait_submit_polling(pfv_ctx *c, uint64_t id, double spin_s);
// pfv_decode_submit_sparse without the walk over mb_off (the caller is this library's own entropy decoder)
struct pfv_decode_job_sparse;
int pfv_decode_submit_sparse_trusted(pfv_ctx *c, const pfv_decode_job_sparse *jobs, uint32_t njobs);
