// pfv_ctx.cu — host side of the engine: context, frame-slot pool, staging rings, the three-stream
// H2D / compute / D2H pipeline and the C ABI of include/pfv_b200.h.
//
// There is deliberately NO CPU fallback in this file: without a CUDA device every entry point that
// would compute fails with PFV_ERR_NO_DEVICE / PFV_ERR_CUDA.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <chrono>
#include <vector>

#include "pfv_internal.h"
#include "pfv_dct.cuh"
#include "pfv_pool.h"

using namespace pfv;

// ---------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// the same error slot for the host codec layer (pfv_codec.cpp)
int pfv::set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(expr)                                                                                   \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? PFV_ERR_NO_DEVICE \
                                                                                     : PFV_ERR_CUDA,   \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__);   \
    } while (0)

extern "C" const char *pfv_last_error(void) { return g_err; }
extern "C" int pfv_abi_version(void) { return PFV_B200_ABI_VERSION; }

extern "C" int pfv_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(PFV_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

// ---------------------------------------------------------------------------------------------------
// geometry and q-tables (host arithmetic only)
// ---------------------------------------------------------------------------------------------------
static uint32_t pad16(uint32_t v) { return v + (16 - (v % 16)) % 16; }   // src/frame.rs:29-30

extern "C" void pfv_geometry_for(uint32_t width, uint32_t height, pfv_geometry *g)
{
    g->width = width;
    g->height = height;
    g->cwidth = width / 2;                  // src/frame.rs:32
    g->cheight = height / 2;                // src/frame.rs:33
    g->pw = pad16(width);
    g->ph = pad16(height);
    g->cpw = pad16(g->cwidth);              // src/frame.rs:35
    g->cph = pad16(g->cheight);             // src/frame.rs:36
    g->nb_y = (g->pw / 16) * (g->ph / 16);
    g->nb_c = (g->cpw / 16) * (g->cph / 16);
    g->nb = g->nb_y + 2 * g->nb_c;          // src/dec.rs:255
    g->frame_bytes = g->pw * g->ph + 2 * g->cpw * g->cph;
}

// src/dct.rs:16-37
static const int32_t kQIntra[64] = {
     8, 16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37,
    19, 22, 26, 27, 29, 34, 34, 38, 22, 22, 26, 27, 29, 34, 37, 40,
    22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32, 35, 40, 48, 58,
    26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83,
};
static const int32_t kQInterValue = 16;
// src/dct.rs:4-13
static const int32_t kScale[64] = {
    32, 37, 34, 26, 32, 26, 34, 37, 37, 43, 39, 31, 37, 31, 39, 43,
    34, 39, 35, 28, 34, 28, 35, 39, 26, 31, 28, 22, 26, 22, 28, 31,
    32, 37, 34, 26, 32, 26, 34, 37, 26, 31, 28, 22, 26, 22, 28, 31,
    34, 39, 35, 28, 34, 28, 35, 39, 37, 43, 39, 31, 37, 31, 39, 43,
};
// src/dct.rs:39-42
static const uint8_t kInvZigzag[64] = {
     0,  1,  5,  6, 14, 15, 27, 28,  2,  4,  7, 13, 16, 26, 29, 42,
     3,  8, 12, 17, 25, 30, 41, 43,  9, 11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,
    21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63,
};

// src/enc.rs:40-51.  Each product is rounded to f32 like the Rust expression
// `(x as f32 * qscale * 0.5).max(1.0) as i32`.
extern "C" int pfv_make_qtables(int quality, int32_t out[4][64], float *px_err_out)
{
    if (quality < 0 || quality > 10) return fail(PFV_ERR_BAD_ARG, "quality %d outside 0..=10 (src/enc.rs:38)", quality);
    volatile float qscale = (float)quality * 0.25f;
    for (int i = 0; i < 64; i++) {
        volatile float a, b;
        a = (float)kQIntra[i] * qscale;
        b = a * 0.5f;
        out[0][i] = (int32_t)fmaxf(b, 1.0f);   // intra_l, src/enc.rs:50
        out[1][i] = (int32_t)fmaxf(a, 1.0f);   // intra_c, src/enc.rs:51
        a = (float)kQInterValue * qscale;
        b = a * 0.5f;
        out[2][i] = (int32_t)fmaxf(b, 1.0f);   // inter_l, src/enc.rs:48
        out[3][i] = (int32_t)fmaxf(a, 1.0f);   // inter_c, src/enc.rs:49
    }
    if (px_err_out) *px_err_out = (float)quality * 1.5f;   // src/enc.rs:41
    return PFV_OK;
}

extern "C" int pfv_host_alloc(void **out, size_t bytes)
{
    if (!out) return fail(PFV_ERR_BAD_ARG, "pfv_host_alloc: out is NULL");
    *out = nullptr;
    CU_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped));   // reachable by every device (sparse encode seam)
    return PFV_OK;
}

extern "C" void pfv_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
namespace {

// cudaSetDevice costs a driver call; the calling thread is almost always on the context's device already.  A thread that
// has never set a device reports device 0 WITHOUT having the primary context bound (cudaPointerGetAttributes then returns
// no device pointer for pinned host memory): the first call on every thread always binds.
inline cudaError_t ensure_device(int device)
{
    static thread_local bool bound = false;
    int cur = -1;
    if (bound && cudaGetDevice(&cur) == cudaSuccess && cur == device) return cudaSuccess;
    const cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) bound = true;
    return e;
}

constexpr int STAGES = 4;        // staging ring depth: H2D of submit n+1 overlaps the kernels of submit n
constexpr int D2H_RING = 64;     // D2H completion events kept for slot-reuse ordering and pfv_ctx_wait_submit

struct Stage {
    int16_t   *d_coeff = nullptr;    // max_jobs * nb * 256
    pfv_mbhdr *d_hdr = nullptr;      // max_jobs * nb
    uint8_t   *d_src = nullptr;      // max_jobs * src_stride (encode only, lazily allocated)
    void      *d_jobs = nullptr;     // max_jobs * max(sizeof(DecJob), sizeof(EncJob))
    void      *h_jobs = nullptr;     // pinned mirror of d_jobs
    uint32_t  *d_pack = nullptr;     // sparse transport, per job slot: [mb_off (nb+1) | headers (nb) | tokens (nb*256)], lazily
                                     // allocated; a caller that keeps the three arrays adjacent gets ONE copy per frame
    SparseJob *d_sjobs = nullptr;    // max_jobs
    SparseJob *h_sjobs = nullptr;    // pinned mirror
    uint32_t  *d_tokpack = nullptr;  // sparse encode, per job slot: [mb_off (nb+1) | RLE entries (nb*256) | statistics], lazily allocated
    TokJob    *d_tjobs = nullptr;    // max_jobs
    TokJob    *h_tjobs = nullptr;    // pinned mirror
    uint8_t   *d_rgb_src = nullptr;  // encode jobs with PFV_JOB_SRC_RGB: max_jobs * w*h*3 (lazily allocated)
    uint32_t  *h_tok = nullptr;      // host compaction of dense host buffers: max_jobs * nb * 128 tokens, pinned (lazily allocated)
    uint32_t  *h_mboff = nullptr;    // max_jobs * (nb + 1), pinned
    cudaEvent_t ev_h2d = nullptr;    // job table + inputs are on the device
    cudaEvent_t ev_kernel = nullptr; // kernels that read/write the stage's device buffers are done
    cudaEvent_t ev_d2h = nullptr;    // copies out of the stage's device buffers are done (encode)
    bool        d2h_used = false;    // an encode submit has recorded ev_d2h on this stage since a decode submit last waited for it
};

}  // namespace

struct pfv_ctx {
    int device = 0;
    pfv_geometry geo{};
    FrameGeom fg{};
    uint32_t nq = 0, nslots = 0, max_jobs = 0;
    size_t slot_stride = 0;          // frame_bytes rounded up to 256
    size_t src_stride = 0;           // per-job source staging (encode)
    size_t pack_words = 0;           // per-job sparse staging in 32-bit words: (nb+1) + nb + nb*256, rounded to 16 bytes
    uint32_t src_off[3] = {0, 0, 0};
    uint8_t *d_pool = nullptr;
    QTables *d_qt = nullptr;
    int *d_err = nullptr;
    uint32_t *d_work = nullptr;            // 16 zeroed words: work counters of the persistent kernels ([0..1] encode-P, [4..5] encode-I)
    int *h_err = nullptr;            // pinned
    cudaStream_t s_h2d = nullptr, s_compute = nullptr, s_d2h = nullptr;
    cudaStream_t s_d2h2 = nullptr;         // second copy-out stream for batches of large pictures (decode_submit_impl)
    cudaEvent_t ev_d2h_join = nullptr;
    bool d2h_streams2 = true;              // PFV_D2H_STREAMS=1 keeps everything on one
    uint8_t *d_rgb = nullptr;              // pfv_slot_read_rgb: device staging of one packed RGB picture (lazily allocated)
    cudaEvent_t ev_rgb = nullptr;          // the last D2H copy out of d_rgb
    bool trace = false;                    // PFV_TRACE=1: host time of the encode submit path, printed at destroy
    double t_enc_alloc = 0, t_enc_wait = 0, t_enc_copy = 0, t_enc_launch = 0, t_enc_d2h = 0;
    uint64_t n_enc_submits = 0;
    double t_dec_check = 0, t_dec_wait = 0, t_dec_copy = 0, t_dec_launch = 0, t_dec_d2h = 0;   // PFV_TRACE: the same split for decode submits
    uint64_t n_dec_submits = 0, n_dec_jobs = 0;
    int host_compact = 0;                  // PFV_HOST_COMPACT=1: compact dense host coefficients to tokens on a host pool before the
                                           // copy.  Off by default: on the bench box (16 host threads) scanning 6.27 MB per 1080p
                                           // frame cost ~1 ms per frame per thread and halved e2e (7.9 k -> 3.9 k frames/s); hosts
                                           // that can, hand over tokens directly (pfv_decode_submit_sparse: 13 k frames/s)
    Pool *pool = nullptr;                  // host threads for the compaction (created on first use)
    bool pcount_clean = false;             // the last residual kernel left d_pcount[0 .. 4*njobs) zeroed
    bool own_compute = true;
    Stage st[STAGES];
    cudaEvent_t ev_d2h_ring[D2H_RING]{};
    std::vector<uint64_t> slot_last_d2h;   // submit id (1-based) whose D2H last read the slot; 0 = none
    uint64_t submit_id = 0;
    uint64_t failed_ring[D2H_RING]{};      // failed_ring[id % D2H_RING] == id: that submit broke off after its id was issued
    uint64_t launches = 0;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;
    bool have_kernel_time = false;
    bool want_kernel_time = false;         // pfv_ctx_last_kernel_ms has been called once: submits bracket their kernels with events
                                           // from then on (two driver calls per submit that a caller who never asks does not pay)
    std::vector<int32_t> h_deq_scan;       // nq * 64: SCALE[s]*q[s] by scan position (src/dct.rs:78-83)
    std::vector<float> h_enc_recip;        // nq * 64: quant_recip_f32(q[raster]) by raster position (src/dct.rs:93-95)
    bool enc_tables_ok = false;            // tables 0..3 exist and every divisor is in 1..65535 (what an encoder needs)
    uint32_t cta_base[3] = {0, 0, 0}, cta_total = 0;   // sub-block kernels: CTAs of 32 macroblocks per plane
    // One default kernel per path plus independently written second implementations, selectable for the parity tests:
    int decode_i_variant = 0;              // PFV_DECODE_I_VARIANT: 0 "stream" (default: TMA-staged classify/compact/transform), 1 "sb" (plain
                                           // thread-per-sub-block), 2 "warp" (first generation, warp per macroblock)
    int decode_p_variant = 0;              // PFV_DECODE_P_VARIANT: 0 "fused" (default: warp-specialised copy + residual in one kernel),
                                           // 1 "win" (window copy kernel + list-driven residual kernel), 2 "warp" (also used without TMA)
    int encode_i_variant = 0;              // PFV_ENCODE_I_VARIANT: 0 "persist" (default: thread per sub-block, persistent), 2 "warp" (first generation)
    int encode_p_variant = 0;              // PFV_ENCODE_P_VARIANT: 0 "strip" (default: warp per tile, column-strip search), 1 "v1" (warp per macroblock)
    uint32_t *d_plist = nullptr;           // max_jobs * nb: coded macroblocks per (job, plane), filled by mc_copy_kernel
    uint32_t *d_pcount = nullptr;          // max_jobs * 4
    CUtensorMap tm_luma{}, tm_chroma{};          // encode-P search window boxes (first-generation kernel: 176 x 46)
    CUtensorMap tm_ep2_luma{}, tm_ep2_chroma{};  // ... of the warp-per-tile kernel (EP2_WIN_W x 46)
    CUtensorMap tm_win_luma{}, tm_win_chroma{};  // decode-P windows of 8 x 4 macroblocks (176 x 94; the copy + residual pair)
    CUtensorMap tm_pf_luma{}, tm_pf_chroma{};    // decode-P windows of 8 x PF_ROWS macroblocks (the fused kernel)
    bool have_tma = false;
    char tma_err[160] = "";
};

namespace {

int build_tensor_maps(pfv_ctx *c)
{
    typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                      CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                      CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        snprintf(c->tma_err, sizeof(c->tma_err), "cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
        return -1;
    }
    EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(fn);
    const pfv_geometry &g = c->geo;
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    for (int which = 0; which < 4; which++) {
    const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    // which == 0: the search window of an encode-P tile of 8 macroblocks (176 x 46)
    // which == 1: the window of every possible predictor of 8 x 4 macroblocks (176 x 94; decode-P copy kernel)
    // which == 2: the same for 8 x PF_ROWS macroblocks (fused decode-P kernel)
    // which == 3: the search window again, at the warp-per-tile kernel's own pitch
    const cuuint32_t box[4] = {which == 0 ? (cuuint32_t)WIN_W : (which == 3 ? (cuuint32_t)EP2_WIN_W : (which == 2 ? (cuuint32_t)PF_WIN_W : 176u)),
                               which == 0 || which == 3 ? (cuuint32_t)WIN_H : (which == 1 ? 94u : (cuuint32_t)PF_WIN_H), 1, 1};
    CUtensorMap *out_l = which == 0 ? &c->tm_luma : (which == 1 ? &c->tm_win_luma : (which == 2 ? &c->tm_pf_luma : &c->tm_ep2_luma));
    CUtensorMap *out_c = which == 0 ? &c->tm_chroma : (which == 1 ? &c->tm_win_chroma : (which == 2 ? &c->tm_pf_chroma : &c->tm_ep2_chroma));
    {   // luma: (x, y, 1, slot)
        const cuuint64_t dims[4] = {g.pw, g.ph, 1, c->nslots};
        const cuuint64_t strides[3] = {g.pw, c->slot_stride, c->slot_stride};
        CUresult r = encode(out_l, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, c->d_pool, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            snprintf(c->tma_err, sizeof(c->tma_err), "cuTensorMapEncodeTiled(luma) -> CUresult %d", (int)r);
            return -1;
        }
    }
    {   // chroma: (x, y, u|v, slot)
        const cuuint64_t nc = (cuuint64_t)g.cpw * g.cph;
        const cuuint64_t dims[4] = {g.cpw, g.cph, 2, c->nslots};
        const cuuint64_t strides[3] = {g.cpw, nc, c->slot_stride};
        CUresult r = encode(out_c, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, c->d_pool + (size_t)g.pw * g.ph, dims,
                            strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            snprintf(c->tma_err, sizeof(c->tma_err), "cuTensorMapEncodeTiled(chroma) -> CUresult %d", (int)r);
            return -1;
        }
    }
    }
    c->have_tma = true;
    return 0;
}

void fill_frame_geom(pfv_ctx *c)
{
    const pfv_geometry &g = c->geo;
    FrameGeom &f = c->fg;
    const uint32_t pw[3] = {g.pw, g.cpw, g.cpw}, ph[3] = {g.ph, g.cph, g.cph};
    const uint32_t vw[3] = {g.width, g.cwidth, g.cwidth}, vh[3] = {g.height, g.cheight, g.cheight};
    uint32_t mb = 0, tile = 0, off = 0;
    for (int p = 0; p < 3; p++) {
        PlaneGeom &pl = f.pl[p];
        pl.pw = pw[p]; pl.ph = ph[p]; pl.vw = vw[p]; pl.vh = vh[p];
        pl.bw = pw[p] / 16; pl.bh = ph[p] / 16;
        pl.mb_base = mb; pl.off = off;
        pl.rcp_bw = 1.0f / (float)pl.bw;
        pl.tiles_per_row = (pl.bw + 7) / 8;
        pl.rcp_tiles_per_row = 1.0f / (float)pl.tiles_per_row;
        pl.tile_base = tile;
        pl.clear4 = p == 0 ? 0u : 0x80808080u;
        mb += pl.bw * pl.bh;
        tile += pl.tiles_per_row * pl.bh;
        off += pw[p] * ph[p];
    }
    f.nb = mb;
    f.total_tiles = tile;
    f.frame_bytes = off;
}

void build_qtables(const int32_t (*qtables)[64], uint32_t nq, std::vector<QTables> &out)
{
    out.resize(nq);
    for (uint32_t t = 0; t < nq; t++) {
        for (int c = 0; c < 8; c++)
            for (int row = 0; row < 8; row++) {
                const int raster = row * 8 + c;
                const int s = kInvZigzag[raster];
                // src/dct.rs:78-83: tables indexed by the SCAN position s; wrapping i32 product
                const uint32_t prod = (uint32_t)kScale[s] * (uint32_t)qtables[t][s];
                out[t].deqT[c * 8 + row] = (int32_t)prod;
                // src/dct.rs:93-95: divisor indexed by the RASTER position
                const uint32_t q = (uint32_t)qtables[t][raster];
                out[t].encM[c * 8 + row] = q ? (uint32_t)(((1ull << 31) + q - 1) / q) : 0u;
            }
    }
}

uint8_t *slot_ptr(pfv_ctx *c, uint32_t slot) { return c->d_pool + (size_t)slot * c->slot_stride; }

int init_slot(pfv_ctx *c, uint32_t slot, cudaStream_t s)
{
    const size_t ny = (size_t)c->geo.pw * c->geo.ph, nc = (size_t)c->geo.cpw * c->geo.cph;
    CU_TRY(cudaMemsetAsync(slot_ptr(c, slot), 0, ny, s));              // src/frame.rs:38
    CU_TRY(cudaMemsetAsync(slot_ptr(c, slot) + ny, 128, 2 * nc, s));   // src/frame.rs:42-43
    return PFV_OK;
}

int ensure_sparse_staging(pfv_ctx *c)
{
    if (c->st[0].d_pack) return PFV_OK;
    for (int i = 0; i < STAGES; i++) {
        Stage &s = c->st[i];
        CU_TRY(cudaMalloc(&s.d_pack, (size_t)c->max_jobs * c->pack_words * sizeof(uint32_t)));
        CU_TRY(cudaMalloc(&s.d_sjobs, sizeof(SparseJob) * c->max_jobs));
        CU_TRY(cudaHostAlloc(&s.h_sjobs, sizeof(SparseJob) * c->max_jobs, cudaHostAllocDefault));
    }
    return PFV_OK;
}

int ensure_compact_staging(pfv_ctx *c)
{
    if (c->st[0].h_tok) return PFV_OK;
    for (int i = 0; i < STAGES; i++) {
        Stage &s = c->st[i];
        CU_TRY(cudaHostAlloc(&s.h_tok, (size_t)c->max_jobs * c->geo.nb * 128 * sizeof(uint32_t), cudaHostAllocDefault));
        CU_TRY(cudaHostAlloc(&s.h_mboff, (size_t)c->max_jobs * (c->geo.nb + 1) * sizeof(uint32_t), cudaHostAllocDefault));
    }
    if (!c->pool) {
        unsigned n = std::thread::hardware_concurrency();
        if (const char *v = getenv("PFV_HOST_THREADS")) n = (unsigned)atoi(v);
        n = n < 1 ? 1 : (n > 16 ? 16 : n);
        c->pool = new (std::nothrow) Pool(n);
        if (!c->pool) return fail(PFV_ERR_NOMEM, "out of host memory");
    }
    return PFV_OK;
}

// Dense coefficients of one frame -> the token form of pfv_decode_submit_sparse, on the host.  The dense array the
// reference hands to the macroblock loops is >90 % zeros on real streams, so scanning it here (32 bytes per test) and
// sending only the non-zero coefficients over PCIe is several times cheaper than copying it.  Returns the token count,
// or UINT32_MAX when the frame is too dense for `cap` tokens (then the dense copy is used).
uint32_t compact_dense(const int16_t *coeff, const pfv_mbhdr *hdr, uint32_t nb, uint32_t *mb_off, uint32_t *tok, uint32_t cap)
{
    uint32_t n = 0;
    mb_off[0] = 0;
    for (uint32_t m = 0; m < nb; m++) {
        if (!hdr || hdr[m].has_coeff) {                             // skipped P macroblocks are never read (src/dec.rs:381)
            if (n + 256 > cap) return UINT32_MAX;
            const int16_t *c = coeff + (size_t)m * 256;
            for (uint32_t i = 0; i < 256; i += 16) {
                uint64_t w[4];
                memcpy(w, c + i, 32);
                if ((w[0] | w[1] | w[2] | w[3]) == 0) continue;
                for (uint32_t k = 0; k < 16; k++) {
                    const int16_t v = c[i + k];
                    if (v) tok[n++] = ((i + k) << 16) | (uint32_t)(uint16_t)v;
                }
            }
        }
        mb_off[m + 1] = n;
    }
    return n;
}

inline double host_now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int ensure_src_staging(pfv_ctx *c)
{
    if (c->st[0].d_src) return PFV_OK;
    for (int i = 0; i < STAGES; i++) CU_TRY(cudaMalloc(&c->st[i].d_src, c->src_stride * c->max_jobs));
    return PFV_OK;
}

// Make the compute stream wait until earlier D2H reads of the slots in `slots` have finished.
int wait_slot_readers(pfv_ctx *c, const uint32_t *slots, uint32_t n)
{
    uint64_t need = 0;
    for (uint32_t i = 0; i < n; i++)
        if (c->slot_last_d2h[slots[i]] > need) need = c->slot_last_d2h[slots[i]];
    if (need == 0) return PFV_OK;
    // events older than the ring have been re-recorded by a LATER submit on the same stream: waiting on
    // the later one is conservative and correct.
    uint64_t oldest = c->submit_id > D2H_RING ? c->submit_id - D2H_RING + 1 : 1;
    if (need < oldest) need = oldest;
    CU_TRY(cudaStreamWaitEvent(c->s_compute, c->ev_d2h_ring[need % D2H_RING], 0));
    return PFV_OK;
}

}  // namespace

// (internal, pfv_internal.h) The staging buffers of the sparse decode seam / of encode submits are allocated on first use - a dozen
// cudaMalloc / cudaHostAlloc calls, several milliseconds.  The Decoder and Encoder objects know at open what they will submit and
// allocate there, like the reference's Decoder::new / Encoder::new allocate their planes (src/dec.rs:31-52, src/enc.rs:37-73).
int pfv_ctx_reserve_staging(pfv_ctx *c, bool decode_sparse, bool encode)
{
    if (!c) return fail(PFV_ERR_BAD_ARG, "NULL context");
    CU_TRY(ensure_device(c->device));
    if (decode_sparse) { int rc = ensure_sparse_staging(c); if (rc) return rc; }
    if (encode) { int rc = ensure_src_staging(c); if (rc) return rc; }
    return PFV_OK;
}

extern "C" void pfv_ctx_destroy(pfv_ctx *c)
{
    if (!c) return;
    if (c->trace && c->n_enc_submits)
        fprintf(stderr, "[pfv ctx] %llu encode submits, host time per submit: staging alloc %.1f us, wait for the stage %.1f us, "
                        "copies in %.1f us, launches %.1f us, copies out %.1f us\n", (unsigned long long)c->n_enc_submits,
                1e6 * c->t_enc_alloc / c->n_enc_submits, 1e6 * c->t_enc_wait / c->n_enc_submits, 1e6 * c->t_enc_copy / c->n_enc_submits,
                1e6 * c->t_enc_launch / c->n_enc_submits, 1e6 * c->t_enc_d2h / c->n_enc_submits);
    if (c->trace && c->n_dec_submits)
        fprintf(stderr, "[pfv ctx] %llu decode submits (%.2f jobs each), host time per submit: checks + staging %.1f us, wait for the stage %.1f us, "
                        "copies in %.1f us, launches %.1f us, copies out %.1f us\n", (unsigned long long)c->n_dec_submits,
                (double)c->n_dec_jobs / c->n_dec_submits, 1e6 * c->t_dec_check / c->n_dec_submits, 1e6 * c->t_dec_wait / c->n_dec_submits,
                1e6 * c->t_dec_copy / c->n_dec_submits, 1e6 * c->t_dec_launch / c->n_dec_submits, 1e6 * c->t_dec_d2h / c->n_dec_submits);
    cudaSetDevice(c->device);
    if (c->s_h2d) cudaStreamSynchronize(c->s_h2d);
    if (c->s_compute) cudaStreamSynchronize(c->s_compute);
    if (c->s_d2h) cudaStreamSynchronize(c->s_d2h);
    if (c->s_d2h2) cudaStreamSynchronize(c->s_d2h2);
    for (int i = 0; i < STAGES; i++) {
        Stage &s = c->st[i];
        cudaFree(s.d_coeff); cudaFree(s.d_hdr); cudaFree(s.d_src); cudaFree(s.d_jobs);
        cudaFree(s.d_pack); cudaFree(s.d_sjobs); cudaFree(s.d_rgb_src); cudaFree(s.d_tokpack); cudaFree(s.d_tjobs);
        if (s.h_tjobs) cudaFreeHost(s.h_tjobs);
        if (s.h_jobs) cudaFreeHost(s.h_jobs);
        if (s.h_sjobs) cudaFreeHost(s.h_sjobs);
        if (s.h_tok) cudaFreeHost(s.h_tok);
        if (s.h_mboff) cudaFreeHost(s.h_mboff);
        if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
        if (s.ev_kernel) cudaEventDestroy(s.ev_kernel);
        if (s.ev_d2h) cudaEventDestroy(s.ev_d2h);
    }
    for (int i = 0; i < D2H_RING; i++) if (c->ev_d2h_ring[i]) cudaEventDestroy(c->ev_d2h_ring[i]);
    if (c->ev_k0) cudaEventDestroy(c->ev_k0);
    if (c->ev_k1) cudaEventDestroy(c->ev_k1);
    cudaFree(c->d_pool); cudaFree(c->d_qt); cudaFree(c->d_err); cudaFree(c->d_work); cudaFree(c->d_plist); cudaFree(c->d_pcount);
    if (c->h_err) cudaFreeHost(c->h_err);
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    if (c->s_d2h2) cudaStreamDestroy(c->s_d2h2);
    if (c->ev_d2h_join) cudaEventDestroy(c->ev_d2h_join);
    delete c->pool;
    cudaFree(c->d_rgb);
    if (c->ev_rgb) cudaEventDestroy(c->ev_rgb);
    if (c->s_compute && c->own_compute) cudaStreamDestroy(c->s_compute);
    delete c;
}

static int ctx_create_impl(pfv_ctx *c, const int32_t (*qtables)[64], void *ext_stream)
{
    CU_TRY(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, c->device));
    if (prop.major != 10)
        return fail(PFV_ERR_NO_DEVICE, "device %d is sm_%d%d; this engine is built for sm_100a only", c->device,
                    prop.major, prop.minor);

    fill_frame_geom(c);
    c->slot_stride = ((size_t)c->geo.frame_bytes + 255) & ~(size_t)255;
    // tight source planes, each start rounded to 16 bytes so the 8-byte vector path applies whenever w % 8 == 0
    c->src_off[0] = 0;
    c->src_off[1] = (c->geo.width * c->geo.height + 15u) & ~15u;
    c->src_off[2] = (c->src_off[1] + c->geo.cwidth * c->geo.cheight + 15u) & ~15u;
    c->src_stride = ((size_t)c->src_off[2] + (size_t)c->geo.cwidth * c->geo.cheight + 255) & ~(size_t)255;
    c->pack_words = (((size_t)c->geo.nb + 1) + c->geo.nb + (size_t)c->geo.nb * 256 + 3) & ~(size_t)3;

    CU_TRY(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    CU_TRY(cudaStreamCreateWithFlags(&c->s_d2h2, cudaStreamNonBlocking));
    CU_TRY(cudaEventCreateWithFlags(&c->ev_d2h_join, cudaEventDisableTiming));
    if (const char *v = getenv("PFV_D2H_STREAMS")) c->d2h_streams2 = atoi(v) != 1;
    if (const char *v = getenv("PFV_HOST_COMPACT")) c->host_compact = atoi(v) != 0;
    if (const char *v = getenv("PFV_TRACE")) c->trace = atoi(v) != 0;
    if (ext_stream) {
        c->s_compute = (cudaStream_t)ext_stream;
        c->own_compute = false;
    } else {
        CU_TRY(cudaStreamCreateWithFlags(&c->s_compute, cudaStreamNonBlocking));
    }

    // frame pool (+256 bytes of slack: unaligned 8-byte fetches read up to 3 bytes past their end)
    CU_TRY(cudaMalloc(&c->d_pool, c->slot_stride * c->nslots + 256));
    CU_TRY(cudaMemsetAsync(c->d_pool + c->slot_stride * c->nslots, 0, 256, c->s_compute));
    for (uint32_t s = 0; s < c->nslots; s++) {
        int rc = init_slot(c, s, c->s_compute);
        if (rc) return rc;
    }
    c->slot_last_d2h.assign(c->nslots, 0);

    std::vector<QTables> qt;
    build_qtables(qtables, c->nq, qt);
    c->h_deq_scan.resize((size_t)c->nq * 64);
    for (uint32_t t = 0; t < c->nq; t++)
        for (int i = 0; i < 64; i++)
            c->h_deq_scan[t * 64 + i] = (int32_t)((uint32_t)kScale[i] * (uint32_t)qtables[t][i]);
    c->h_enc_recip.resize((size_t)c->nq * 64);
    c->enc_tables_ok = c->nq >= 4;
    for (uint32_t t = 0; t < c->nq; t++)
        for (int i = 0; i < 64; i++) {
            const int32_t q = qtables[t][i];
            if (t < 4 && (q < 1 || q > QUANT_MAX_DIVISOR)) c->enc_tables_ok = false;
            c->h_enc_recip[t * 64 + i] = quant_recip_f32(q);
        }
    {
        uint32_t cta = 0;
        for (int p = 0; p < 3; p++) {
            c->cta_base[p] = cta;
            cta += (c->fg.pl[p].bw * c->fg.pl[p].bh + SB_MBS_PER_CTA - 1) / SB_MBS_PER_CTA;
        }
        c->cta_total = cta;
    }
    if (const char *v = getenv("PFV_DECODE_I_VARIANT")) c->decode_i_variant = strcmp(v, "warp") == 0 ? 2 : (strcmp(v, "sb") == 0 ? 1 : 0);
    if (const char *v = getenv("PFV_DECODE_P_VARIANT")) c->decode_p_variant = strcmp(v, "warp") == 0 ? 2 : (strcmp(v, "win") == 0 ? 1 : 0);
    if (const char *v = getenv("PFV_ENCODE_I_VARIANT")) c->encode_i_variant = strcmp(v, "warp") == 0 ? 2 : 0;
    if (const char *v = getenv("PFV_ENCODE_P_VARIANT")) c->encode_p_variant = strcmp(v, "v1") == 0 ? 1 : 0;
    CU_TRY(cudaMalloc(&c->d_qt, sizeof(QTables) * c->nq));
    CU_TRY(cudaMemcpy(c->d_qt, qt.data(), sizeof(QTables) * c->nq, cudaMemcpyHostToDevice));

    CU_TRY(cudaMalloc(&c->d_plist, (size_t)c->max_jobs * c->geo.nb * sizeof(uint32_t)));
    // list counts | "CTAs done" counter (4 words) | live mode: windows done per frame | chunk tickets per frame
    CU_TRY(cudaMalloc(&c->d_pcount, ((size_t)c->max_jobs * 6 + 4) * sizeof(uint32_t)));
    CU_TRY(cudaMemset(c->d_pcount, 0, ((size_t)c->max_jobs * 6 + 4) * sizeof(uint32_t)));
    CU_TRY(cudaMalloc(&c->d_err, sizeof(int)));
    CU_TRY(cudaMemset(c->d_err, 0, sizeof(int)));
    CU_TRY(cudaMalloc(&c->d_work, 16 * sizeof(uint32_t)));
    CU_TRY(cudaMemset(c->d_work, 0, 16 * sizeof(uint32_t)));
    CU_TRY(cudaHostAlloc(&c->h_err, sizeof(int), cudaHostAllocDefault));
    *c->h_err = 0;

    const size_t job_bytes = sizeof(DecJob) > sizeof(EncJob) ? sizeof(DecJob) : sizeof(EncJob);
    for (int i = 0; i < STAGES; i++) {
        Stage &s = c->st[i];
        CU_TRY(cudaMalloc(&s.d_coeff, (size_t)c->max_jobs * c->geo.nb * 256 * sizeof(int16_t)));
        CU_TRY(cudaMemsetAsync(s.d_coeff, 0, (size_t)c->max_jobs * c->geo.nb * 256 * sizeof(int16_t), c->s_compute));
        CU_TRY(cudaMalloc(&s.d_hdr, (size_t)c->max_jobs * c->geo.nb * sizeof(pfv_mbhdr)));
        CU_TRY(cudaMalloc(&s.d_jobs, job_bytes * c->max_jobs));
        CU_TRY(cudaHostAlloc(&s.h_jobs, job_bytes * c->max_jobs, cudaHostAllocDefault));
        CU_TRY(cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&s.ev_kernel, cudaEventDisableTiming));
        CU_TRY(cudaEventCreateWithFlags(&s.ev_d2h, cudaEventDisableTiming));
    }
    {
        // pfv_ctx_wait_submit is what a host's entropy threads block in, one thread per frame in flight: they sleep in the driver
        // (cudaEventBlockingSync) instead of spinning, which on a 16-thread host took the cores the submitting thread needs.
        // PFV_EVENT_SPIN=1 restores busy-waiting (lowest wake-up latency, one core per waiter).
        const char *env = getenv("PFV_EVENT_SPIN");
        const unsigned flags = cudaEventDisableTiming | ((env && atoi(env) != 0) ? 0u : (unsigned)cudaEventBlockingSync);
        for (int i = 0; i < D2H_RING; i++) CU_TRY(cudaEventCreateWithFlags(&c->ev_d2h_ring[i], flags));
    }
    CU_TRY(cudaEventCreate(&c->ev_k0));
    CU_TRY(cudaEventCreate(&c->ev_k1));

    build_tensor_maps(c);   // failure is reported by the first encode-P submit

    CU_TRY(cudaStreamSynchronize(c->s_compute));
    return PFV_OK;
}

extern "C" int pfv_ctx_create(int device, uint32_t width, uint32_t height, const int32_t (*qtables)[64], uint32_t nq,
                              uint32_t nslots, uint32_t max_jobs, void *ext_stream, pfv_ctx **out)
{
    if (!out) return fail(PFV_ERR_BAD_ARG, "pfv_ctx_create: out is NULL");
    *out = nullptr;
    if (width == 0 || height == 0 || (width & 1) || (height & 1) || width > 65535 || height > 65535)
        return fail(PFV_ERR_BAD_ARG, "frame size %ux%u must be even, non-zero and fit u16 (src/frame.rs:13, src/enc.rs:195-196)",
                    width, height);
    if (!qtables || nq == 0 || nq > 256) return fail(PFV_ERR_BAD_ARG, "need 1..256 q-tables");
    if (nslots < 2) return fail(PFV_ERR_BAD_ARG, "need at least 2 frame slots");
    if (max_jobs == 0) return fail(PFV_ERR_BAD_ARG, "max_jobs must be > 0");
    for (uint32_t t = 0; t < nq; t++)
        for (int i = 0; i < 64; i++)
            if (qtables[t][i] < 0 || qtables[t][i] > 65535)
                return fail(PFV_ERR_BAD_ARG, "q-table %u entry %d = %d does not fit the u16 header field (src/enc.rs:202-216)",
                            t, i, qtables[t][i]);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(PFV_ERR_NO_DEVICE, "no CUDA device (%s); the engine has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(PFV_ERR_BAD_ARG, "device %d out of range (0..%d)", device, ndev - 1);

    pfv_ctx *c = new (std::nothrow) pfv_ctx();
    if (!c) return fail(PFV_ERR_NOMEM, "out of host memory");
    c->device = device;
    pfv_geometry_for(width, height, &c->geo);
    c->nq = nq;
    c->nslots = nslots;
    c->max_jobs = max_jobs;
    int rc = ctx_create_impl(c, qtables, ext_stream);
    if (rc != PFV_OK) {
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        pfv_ctx_destroy(c);
        memcpy(g_err, keep, sizeof(keep));
        return rc;
    }
    *out = c;
    return PFV_OK;
}

extern "C" int pfv_ctx_geometry(const pfv_ctx *c, pfv_geometry *out)
{
    if (!c || !out) return fail(PFV_ERR_BAD_ARG, "NULL argument");
    *out = c->geo;
    return PFV_OK;
}

extern "C" uint64_t pfv_ctx_launch_count(const pfv_ctx *c) { return c ? c->launches : 0; }

extern "C" int pfv_sync(pfv_ctx *c)
{
    if (!c) return fail(PFV_ERR_BAD_ARG, "NULL context");
    CU_TRY(ensure_device(c->device));
    CU_TRY(cudaStreamSynchronize(c->s_h2d));
    CU_TRY(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->s_compute));
    CU_TRY(cudaStreamSynchronize(c->s_compute));
    CU_TRY(cudaStreamSynchronize(c->s_d2h));
    if (*c->h_err) {
        const int bits = *c->h_err;
        *c->h_err = 0;
        CU_TRY(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->s_compute));
        CU_TRY(cudaStreamSynchronize(c->s_compute));
        if (bits & ERRBIT_BAD_MV)
            return fail(PFV_ERR_BAD_MV, "a motion vector pointed outside the padded plane (src/common.rs:258-259); "
                                        "the co-located block was used instead");
        return fail(PFV_ERR_CUDA, "device error bits 0x%x", bits);
    }
    return PFV_OK;
}

extern "C" int pfv_ctx_last_kernel_ms(pfv_ctx *c, float *ms_out)
{
    if (!c || !ms_out) return fail(PFV_ERR_BAD_ARG, "NULL argument");
    if (!c->want_kernel_time) {
        c->want_kernel_time = true;
        return fail(PFV_ERR_STATE, "kernel timing was off: it is on from now on, ask again after the next submit");
    }
    if (!c->have_kernel_time) return fail(PFV_ERR_STATE, "no kernel has been launched since timing was switched on");
    CU_TRY(ensure_device(c->device));
    CU_TRY(cudaEventSynchronize(c->ev_k1));
    CU_TRY(cudaEventElapsedTime(ms_out, c->ev_k0, c->ev_k1));
    return PFV_OK;
}

extern "C" int pfv_slot_reset(pfv_ctx *c, uint32_t slot)
{
    if (!c || slot >= c->nslots) return fail(PFV_ERR_BAD_ARG, "bad slot");
    CU_TRY(ensure_device(c->device));
    int rc = wait_slot_readers(c, &slot, 1);
    if (rc) return rc;
    return init_slot(c, slot, c->s_compute);
}

extern "C" int pfv_slot_read(pfv_ctx *c, uint32_t slot, uint8_t *frame_out)
{
    if (!c || slot >= c->nslots || !frame_out) return fail(PFV_ERR_BAD_ARG, "bad argument");
    int rc = pfv_sync(c);
    if (rc) return rc;
    CU_TRY(cudaMemcpy(frame_out, slot_ptr(c, slot), c->geo.frame_bytes, cudaMemcpyDeviceToHost));
    return PFV_OK;
}

extern "C" int pfv_slot_write(pfv_ctx *c, uint32_t slot, const uint8_t *frame_in)
{
    if (!c || slot >= c->nslots || !frame_in) return fail(PFV_ERR_BAD_ARG, "bad argument");
    int rc = pfv_sync(c);
    if (rc) return rc;
    CU_TRY(cudaMemcpy(slot_ptr(c, slot), frame_in, c->geo.frame_bytes, cudaMemcpyHostToDevice));
    return PFV_OK;
}

extern "C" int pfv_slot_device_ptr(pfv_ctx *c, uint32_t slot, void **out)
{
    if (!c || slot >= c->nslots || !out) return fail(PFV_ERR_BAD_ARG, "bad argument");
    *out = slot_ptr(c, slot);
    return PFV_OK;
}

// src/dec.rs:195-197: the visible crop of the three planes, on stream s
static int copy_visible(pfv_ctx *c, uint32_t slot, uint8_t *y, uint8_t *u, uint8_t *v, cudaStream_t s)
{
    const pfv_geometry &g = c->geo;
    const uint8_t *base = slot_ptr(c, slot);
    const size_t ny = (size_t)g.pw * g.ph, nc = (size_t)g.cpw * g.cph;
    if (y && u && v && g.pw == g.width && g.cpw == g.cwidth && u == y + ny && v == u + nc) {
        // rows are not padded and the caller's three planes sit where the slot's do: the visible planes are prefixes
        // of the padded ones, one copy moves them all (the few padding rows between them ride along)
        CU_TRY(cudaMemcpyAsync(y, base, ny + nc + (size_t)g.cwidth * g.cheight, cudaMemcpyDeviceToHost, s));
        return PFV_OK;
    }
    if (y) {
        if (g.pw == g.width) CU_TRY(cudaMemcpyAsync(y, base, (size_t)g.width * g.height, cudaMemcpyDeviceToHost, s));
        else CU_TRY(cudaMemcpy2DAsync(y, g.width, base, g.pw, g.width, g.height, cudaMemcpyDeviceToHost, s));
    }
    uint8_t *dst[2] = {u, v};
    for (int p = 0; p < 2; p++) {
        if (!dst[p]) continue;
        const uint8_t *src = base + ny + p * nc;
        if (g.cpw == g.cwidth)
            CU_TRY(cudaMemcpyAsync(dst[p], src, (size_t)g.cwidth * g.cheight, cudaMemcpyDeviceToHost, s));
        else
            CU_TRY(cudaMemcpy2DAsync(dst[p], g.cwidth, src, g.cpw, g.cwidth, g.cheight, cudaMemcpyDeviceToHost, s));
    }
    return PFV_OK;
}

extern "C" int pfv_slot_read_visible(pfv_ctx *c, uint32_t slot, uint8_t *y, uint8_t *u, uint8_t *v)
{
    if (!c || slot >= c->nslots) return fail(PFV_ERR_BAD_ARG, "bad slot");
    CU_TRY(ensure_device(c->device));
    // order after everything submitted so far on the compute stream; takes a submit id of its own so the
    // slot-reuse bookkeeping sees this read
    const uint64_t id = __atomic_add_fetch(&c->submit_id, 1, __ATOMIC_RELAXED);   // helper threads read it (pfv_ctx_wait_submit)
    cudaEvent_t ev = c->ev_d2h_ring[id % D2H_RING];
    CU_TRY(cudaEventRecord(ev, c->s_compute));
    CU_TRY(cudaStreamWaitEvent(c->s_d2h, ev, 0));
    int rc = copy_visible(c, slot, y, u, v, c->s_d2h);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(ev, c->s_d2h));
    c->slot_last_d2h[slot] = id;
    return PFV_OK;
}

// save_frame (src/lib.rs:365-395) of a slot's visible crop, on stream s, into device memory
static int convert_rgb(pfv_ctx *c, uint32_t slot, uint8_t *d_out, cudaStream_t s)
{
    const pfv_geometry &g = c->geo;
    const size_t ny = (size_t)g.pw * g.ph, nc = (size_t)g.cpw * g.cph;
    // (a batch of one: the batched kernels carry the fast 8-pixels-per-thread form for widths that are multiples of 8)
    CU_TRY(launch_yuv420_to_rgb_batch(c->d_pool, c->slot_stride, (uint32_t)ny, (uint32_t)(ny + nc), &slot, 1, g.width, g.height, g.pw, g.cpw,
                                      d_out, (size_t)g.width * g.height * 3, s));
    c->launches++;
    return PFV_OK;
}

extern "C" int pfv_slot_convert_rgb(pfv_ctx *c, uint32_t slot, void *rgb_device)
{
    if (!c || slot >= c->nslots || !rgb_device) return fail(PFV_ERR_BAD_ARG, "bad argument");
    CU_TRY(ensure_device(c->device));
    return convert_rgb(c, slot, static_cast<uint8_t *>(rgb_device), c->s_compute);
}

extern "C" int pfv_slots_convert_rgb(pfv_ctx *c, const uint32_t *slots, uint32_t n, void *rgb_device, size_t stride)
{
    if (!c || !slots || !rgb_device) return fail(PFV_ERR_BAD_ARG, "bad argument");
    const size_t bytes = (size_t)c->geo.width * c->geo.height * 3;
    if (stride < bytes) return fail(PFV_ERR_BAD_ARG, "stride %zu < %zu bytes of one RGB picture", stride, bytes);
    for (uint32_t i = 0; i < n; i++)
        if (slots[i] >= c->nslots) return fail(PFV_ERR_BAD_ARG, "slot %u out of range", slots[i]);
    if (n == 0) return PFV_OK;
    CU_TRY(ensure_device(c->device));
    const pfv_geometry &g = c->geo;
    const size_t ny = (size_t)g.pw * g.ph, nc = (size_t)g.cpw * g.cph;
    CU_TRY(launch_yuv420_to_rgb_batch(c->d_pool, c->slot_stride, (uint32_t)ny, (uint32_t)(ny + nc), slots, n, g.width, g.height, g.pw,
                                      g.cpw, static_cast<uint8_t *>(rgb_device), stride, c->s_compute));
    c->launches += (n + 63) / 64;
    return PFV_OK;
}

extern "C" int pfv_slot_read_rgb(pfv_ctx *c, uint32_t slot, uint8_t *rgb_host)
{
    if (!c || slot >= c->nslots || !rgb_host) return fail(PFV_ERR_BAD_ARG, "bad argument");
    CU_TRY(ensure_device(c->device));
    const size_t bytes = (size_t)c->geo.width * c->geo.height * 3;
    if (!c->d_rgb) {
        CU_TRY(cudaMalloc(&c->d_rgb, bytes));
        CU_TRY(cudaEventCreateWithFlags(&c->ev_rgb, cudaEventDisableTiming));
    }
    const uint64_t id = __atomic_add_fetch(&c->submit_id, 1, __ATOMIC_RELAXED);   // helper threads read it (pfv_ctx_wait_submit)
    cudaEvent_t ev = c->ev_d2h_ring[id % D2H_RING];
    CU_TRY(cudaStreamWaitEvent(c->s_compute, c->ev_rgb, 0));    // the previous picture has left the staging buffer
    int rc = convert_rgb(c, slot, c->d_rgb, c->s_compute);
    if (rc) return rc;
    CU_TRY(cudaEventRecord(ev, c->s_compute));
    CU_TRY(cudaStreamWaitEvent(c->s_d2h, ev, 0));
    CU_TRY(cudaMemcpyAsync(rgb_host, c->d_rgb, bytes, cudaMemcpyDeviceToHost, c->s_d2h));
    CU_TRY(cudaEventRecord(c->ev_rgb, c->s_d2h));
    CU_TRY(cudaEventRecord(ev, c->s_d2h));
    return PFV_OK;
}

// ---------------------------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------------------------
namespace {
// one decode job in internal form: dense (coeff) or sparse (mb_off/tok) coefficients
struct DecIn {
    uint32_t kind, flags, dst_slot, ref_slot;
    uint8_t  qidx[3];
    bool     sparse;
    bool     trusted;          // the offsets come from this library's own entropy decoder: no second walk over them
    const pfv_mbhdr *hdr;
    const int16_t   *coeff;
    const uint32_t  *mb_off, *tok;
    uint32_t ntok;
    uint8_t *out_y, *out_u, *out_v;
};
}  // namespace

static int decode_submit_impl(pfv_ctx *c, const DecIn *jobs, uint32_t njobs)
{
    const double tr0 = c->trace ? host_now() : 0;
    if (njobs > c->max_jobs) return fail(PFV_ERR_BAD_ARG, "%u jobs > max_jobs %u", njobs, c->max_jobs);
    const pfv_geometry &g = c->geo;
    bool any_sparse = false;
    for (uint32_t i = 0; i < njobs; i++) {
        const DecIn &j = jobs[i];
        if (j.kind != PFV_FRAME_I && j.kind != PFV_FRAME_P) return fail(PFV_ERR_BAD_ARG, "job %u: bad kind %u", i, j.kind);
        if (j.dst_slot >= c->nslots) return fail(PFV_ERR_BAD_ARG, "job %u: dst_slot %u out of range", i, j.dst_slot);
        if (j.sparse) {
            any_sparse = true;
            if (j.flags & ~(uint32_t)PFV_JOB_DENSE) return fail(PFV_ERR_BAD_ARG, "job %u: sparse jobs take host pointers only (no PFV_JOB_DEVICE_PTRS)", i);
            if (!j.mb_off || (j.ntok && !j.tok)) return fail(PFV_ERR_BAD_ARG, "job %u: mb_off/tok is NULL", i);
            if ((uint64_t)j.ntok > (uint64_t)g.nb * 256) return fail(PFV_ERR_BAD_ARG, "job %u: %u tokens > nb*256", i, j.ntok);
            if (j.mb_off[0] != 0 || j.mb_off[g.nb] != j.ntok)
                return fail(PFV_ERR_BAD_ARG, "job %u: mb_off[0] must be 0 and mb_off[nb] must equal ntok", i);
            // (the walk reads 4 (nb + 1) bytes another core has just written - 20 us at 1080p: the Decoder object, whose entropy
            // decoder produces the offsets and bounds every macroblock at 256 tokens by construction, skips it)
            for (uint32_t m = 0; m < g.nb && !j.trusted; m++)
                if (j.mb_off[m + 1] < j.mb_off[m] || j.mb_off[m + 1] - j.mb_off[m] > 256)
                    return fail(PFV_ERR_BAD_ARG, "job %u: mb_off is not a prefix of per-macroblock counts <= 256 (macroblock %u)", i, m);
        } else {
            if (!j.coeff) return fail(PFV_ERR_BAD_ARG, "job %u: coeff is NULL", i);
            if ((j.flags & PFV_JOB_DEVICE_PTRS) && (reinterpret_cast<uintptr_t>(j.coeff) & 15u))
                return fail(PFV_ERR_BAD_ARG, "job %u: device coefficient pointer must be 16-byte aligned", i);
        }
        for (int p = 0; p < 3; p++)
            if (j.qidx[p] >= c->nq)
                return fail(PFV_ERR_BAD_ARG, "job %u: q-table index %u >= %u (src/dec.rs:244-246 would panic)", i, j.qidx[p], c->nq);
        if (j.kind == PFV_FRAME_P) {
            if (j.ref_slot >= c->nslots) return fail(PFV_ERR_BAD_ARG, "job %u: ref_slot %u out of range", i, j.ref_slot);
            if (j.ref_slot == j.dst_slot)
                return fail(PFV_ERR_BAD_ARG, "job %u: ref_slot == dst_slot breaks the two-phase update (src/common.rs:498-521)", i);
            if (!j.hdr) return fail(PFV_ERR_BAD_ARG, "job %u: hdr is NULL", i);
        }
        const bool any = j.out_y || j.out_u || j.out_v;
        if (any && !(j.out_y && j.out_u && j.out_v)) return fail(PFV_ERR_BAD_ARG, "job %u: give all of out_y/u/v or none", i);
    }
    for (uint32_t i = 0; i < njobs; i++)
        if (jobs[i].kind == PFV_FRAME_P)
            for (uint32_t k = 0; k < njobs; k++)
                if (jobs[k].dst_slot == jobs[i].ref_slot)
                    return fail(PFV_ERR_BAD_ARG, "jobs %u and %u of one submit are dependent (ref_slot == dst_slot)", i, k);

    CU_TRY(ensure_device(c->device));
    bool any_dense_host = false;
    if (c->host_compact)
        for (uint32_t i = 0; i < njobs; i++) any_dense_host |= !jobs[i].sparse && !(jobs[i].flags & PFV_JOB_DEVICE_PTRS);
    if (any_sparse || any_dense_host) {
        int rc = ensure_sparse_staging(c);
        if (rc) return rc;
    }
    if (any_dense_host) {
        int rc = ensure_compact_staging(c);
        if (rc) return rc;
    }
    // (Measured and dropped, round 2: sending small submits - a 512x384 frame - down the compute stream alone, copies and
    // kernels in stream order without the cross-stream events, saves 8 driver calls per submit but loses the overlap of one
    // frame's copies with its neighbours' kernels: config 1 through the Decoder 14.5 k frames/s against 21-25 k.)
    cudaStream_t s_up = c->s_h2d, s_down = c->s_d2h;

    const uint64_t id = __atomic_add_fetch(&c->submit_id, 1, __ATOMIC_RELAXED);   // helper threads read it (pfv_ctx_wait_submit)
    Stage &st = c->st[id % STAGES];
    const double tr1 = c->trace ? host_now() : 0;
    CU_TRY(cudaEventSynchronize(st.ev_h2d));                    // pinned job table of this stage is free again
    const double tr2 = c->trace ? host_now() : 0;
    CU_TRY(cudaStreamWaitEvent(s_up, st.ev_kernel, 0));         // device buffers of this stage are free again
    if (st.d2h_used) {                                          // ... also for an encode submit that used the stage before
        CU_TRY(cudaStreamWaitEvent(s_up, st.ev_d2h, 0));
        st.d2h_used = false;
    }

    // dense HOST coefficients: compact them to tokens on the host pool (the stage's pinned token buffers are free: the
    // H2D copies that read them were waited for above) and carry on as sparse jobs
    std::vector<DecIn> compacted;
    if (any_dense_host) {
        compacted.assign(jobs, jobs + njobs);
        const uint32_t cap = g.nb * 128u;
        std::vector<uint32_t> counts(njobs, UINT32_MAX);
        parallel_for(*c->pool, njobs, [&](unsigned i) {
            const DecIn &j = compacted[i];
            if (j.sparse || (j.flags & PFV_JOB_DEVICE_PTRS)) return;
            counts[i] = compact_dense(j.coeff, j.kind == PFV_FRAME_P ? j.hdr : nullptr, g.nb,
                                      st.h_mboff + (size_t)i * (g.nb + 1), st.h_tok + (size_t)i * cap, cap);
        });
        for (uint32_t i = 0; i < njobs; i++) {
            if (counts[i] == UINT32_MAX) continue;              // sparse already, device resident, or too dense: unchanged
            DecIn &j = compacted[i];
            j.sparse = true;
            j.mb_off = st.h_mboff + (size_t)i * (g.nb + 1);
            j.tok = st.h_tok + (size_t)i * cap;
            j.ntok = counts[i];
        }
        jobs = compacted.data();
    }

    // job table: I jobs first, then P jobs, each kind sorted by q-index triple: jobs that share a triple form
    // one launch (the sub-block kernels take the dequantiser tables as kernel parameters)
    DecJob *tab = static_cast<DecJob *>(st.h_jobs);
    std::vector<uint32_t> order;
    order.reserve(njobs);
    uint32_t n_i = 0;
    // key frames whose sub-blocks mostly carry AC terms (PFV_JOB_DENSE, or a sparse job with > 6 tokens per sub-block) go to
    // the plain thread-per-sub-block kernel: measured 0.55 of the HBM roofline there against 0.42 through the staging kernel
    auto dense_hint = [&](uint32_t i) -> uint32_t {
        const DecIn &j = jobs[i];
        if (j.kind != PFV_FRAME_I) return 0u;
        if (j.sparse) return (uint64_t)j.ntok > (uint64_t)g.nb * 4u * 6u ? 1u : 0u;
        return (j.flags & PFV_JOB_DENSE) ? 1u : 0u;
    };
    auto qkey = [&](uint32_t i) { return dense_hint(i) << 24 | (uint32_t)jobs[i].qidx[0] << 16 | (uint32_t)jobs[i].qidx[1] << 8 | jobs[i].qidx[2]; };
    auto by_qkey = [&](uint32_t a, uint32_t b) { return qkey(a) < qkey(b); };
    for (uint32_t i = 0; i < njobs; i++) if (jobs[i].kind == PFV_FRAME_I) { order.push_back(i); n_i++; }
    std::stable_sort(order.begin(), order.end(), by_qkey);
    for (uint32_t i = 0; i < njobs; i++) if (jobs[i].kind == PFV_FRAME_P) order.push_back(i);
    std::stable_sort(order.begin() + n_i, order.end(), by_qkey);

    const size_t coeff_elems = (size_t)g.nb * 256;
    // merge H2D copies of buffers that are adjacent in host memory (a caller that decodes a batch into one
    // pinned arena gets one big copy instead of njobs small ones)
    const int16_t *run_src = nullptr; int16_t *run_dst = nullptr; size_t run_elems = 0;
    auto flush_run = [&]() -> int {
        if (run_elems) CU_TRY(cudaMemcpyAsync(run_dst, run_src, run_elems * sizeof(int16_t), cudaMemcpyHostToDevice, s_up));
        run_elems = 0;
        return PFV_OK;
    };
    uint32_t nsparse = 0;
    for (uint32_t k = 0; k < njobs; k++) {
        const DecIn &j = jobs[order[k]];
        DecJob &d = tab[k];
        const bool dev = (j.flags & PFV_JOB_DEVICE_PTRS) != 0;
        int16_t *d_coeff = st.d_coeff + (size_t)k * coeff_elems;
        pfv_mbhdr *d_hdr = st.d_hdr + (size_t)k * g.nb;
        if (dev) {
            d.coeff = j.coeff;
            d.hdr = j.hdr;
        } else {
            if (j.sparse) {
                uint32_t *d_mboff = st.d_pack + (size_t)k * c->pack_words;
                uint32_t *d_shdr = d_mboff + (g.nb + 1);
                uint32_t *d_tok = d_shdr + g.nb;
                const uint32_t *h_hdr32 = reinterpret_cast<const uint32_t *>(j.hdr);
                if (h_hdr32 && h_hdr32 == j.mb_off + (g.nb + 1) && (j.ntok == 0 || j.tok == h_hdr32 + g.nb)) {
                    // mb_off | headers | tokens adjacent in host memory: one copy
                    CU_TRY(cudaMemcpyAsync(d_mboff, j.mb_off, ((size_t)(g.nb + 1) + g.nb + j.ntok) * sizeof(uint32_t),
                                           cudaMemcpyHostToDevice, s_up));
                } else {
                    if (j.ntok) CU_TRY(cudaMemcpyAsync(d_tok, j.tok, (size_t)j.ntok * sizeof(uint32_t), cudaMemcpyHostToDevice, s_up));
                    CU_TRY(cudaMemcpyAsync(d_mboff, j.mb_off, (size_t)(g.nb + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s_up));
                    if (j.kind == PFV_FRAME_P)
                        CU_TRY(cudaMemcpyAsync(d_shdr, j.hdr, (size_t)g.nb * sizeof(pfv_mbhdr), cudaMemcpyHostToDevice, s_up));
                }
                SparseJob &sj = st.h_sjobs[nsparse++];
                sj.mb_off = d_mboff;
                sj.tok = d_tok;
                sj.hdr = j.kind == PFV_FRAME_P ? reinterpret_cast<const pfv_mbhdr *>(d_shdr) : nullptr;
                sj.coeff = d_coeff;
                d.coeff = d_coeff;
                d.hdr = sj.hdr;
                goto job_common;
            } else if (run_elems && j.coeff == run_src + run_elems && d_coeff == run_dst + run_elems) {
                run_elems += coeff_elems;
            } else {
                int rc = flush_run();
                if (rc) return rc;
                run_src = j.coeff; run_dst = d_coeff; run_elems = coeff_elems;
            }
            d.coeff = d_coeff;
            d.hdr = nullptr;
            if (j.kind == PFV_FRAME_P) {
                CU_TRY(cudaMemcpyAsync(d_hdr, j.hdr, (size_t)g.nb * sizeof(pfv_mbhdr), cudaMemcpyHostToDevice, s_up));
                d.hdr = d_hdr;
            }
        }
    job_common:
        d.dst = slot_ptr(c, j.dst_slot);
        d.ref = j.kind == PFV_FRAME_P ? slot_ptr(c, j.ref_slot) : nullptr;
        d.ref_slot = j.kind == PFV_FRAME_P ? (int32_t)j.ref_slot : 0;
        for (int p = 0; p < 3; p++) d.qt[p] = c->d_qt + j.qidx[p];
    }
    {
        int rc = flush_run();
        if (rc) return rc;
    }
    if (nsparse) CU_TRY(cudaMemcpyAsync(st.d_sjobs, st.h_sjobs, sizeof(SparseJob) * nsparse, cudaMemcpyHostToDevice, s_up));
    CU_TRY(cudaMemcpyAsync(st.d_jobs, tab, sizeof(DecJob) * njobs, cudaMemcpyHostToDevice, s_up));
    CU_TRY(cudaEventRecord(st.ev_h2d, s_up));

    // compute
    const double tr3 = c->trace ? host_now() : 0;
    CU_TRY(cudaStreamWaitEvent(c->s_compute, st.ev_h2d, 0));
    {
        std::vector<uint32_t> dsts(njobs);
        for (uint32_t i = 0; i < njobs; i++) dsts[i] = jobs[i].dst_slot;
        int rc = wait_slot_readers(c, dsts.data(), njobs);
        if (rc) return rc;
    }
    if (c->want_kernel_time) CU_TRY(cudaEventRecord(c->ev_k0, c->s_compute));
    if (nsparse) {
        CU_TRY(launch_expand_tokens(g.nb, st.d_sjobs, nsparse, c->s_compute));
        c->launches++;
    }
    const DecJob *d_tab = static_cast<const DecJob *>(st.d_jobs);
    auto sb_params = [&](uint32_t job_index) {
        SbParams P;
        P.g = c->fg;
        for (int p = 0; p < 3; p++) {
            memcpy(P.deq[p], &c->h_deq_scan[(size_t)jobs[job_index].qidx[p] * 64], 64 * sizeof(int32_t));
            P.cta_base[p] = c->cta_base[p];
        }
        P.cta_total = c->cta_total;
        P.tiles_per_warp = 1;
        return P;
    };
    if (n_i && c->decode_i_variant == 2) {
        CU_TRY(launch_decode(false, c->fg, d_tab, n_i, c->d_err, c->s_compute));
        c->launches++;
    } else {
        for (uint32_t a = 0; a < n_i;) {
            uint32_t b = a + 1;
            while (b < n_i && qkey(order[b]) == qkey(order[a])) b++;
            const SbParams P = sb_params(order[a]);
            if (c->decode_i_variant == 1) CU_TRY(launch_decode_i_sb(P, d_tab + a, b - a, c->s_compute));
            else if (dense_hint(order[a])) CU_TRY(launch_decode_i_direct(P, d_tab + a, b - a, c->s_compute));
            else CU_TRY(launch_decode_i_stream(P, d_tab + a, b - a, c->s_compute));
            c->launches++;
            a = b;
        }
    }
    if (njobs - n_i && (c->decode_p_variant == 2 || !c->have_tma)) {
        CU_TRY(launch_decode(true, c->fg, d_tab + n_i, njobs - n_i, c->d_err, c->s_compute));
        c->launches++;
    } else {
        for (uint32_t a = n_i; a < njobs;) {
            uint32_t b = a + 1;
            while (b < njobs && qkey(order[b]) == qkey(order[a])) b++;
            if (c->decode_p_variant == 0) {
                CU_TRY(launch_decode_p_fused(sb_params(order[a]), d_tab + a, b - a, c->d_err, c->tm_pf_luma, c->tm_pf_chroma, c->s_compute));
                c->launches++;
            } else {
                const uint32_t k0 = a - n_i, n = b - a;
                // list counts: cleared by the previous residual kernel itself when the whole group went out as one pair of
                // launches (self_clear); otherwise by a memset node here
                const bool self_clear = k0 == 0 && c->pcount_clean;
                if (!self_clear)
                    CU_TRY(cudaMemsetAsync(c->d_pcount + (size_t)k0 * 4, 0, (size_t)n * 4 * sizeof(uint32_t), c->s_compute));
                c->pcount_clean = k0 == 0 && n == njobs - n_i;
                CU_TRY(launch_decode_p_two_pass4(sb_params(order[a]), d_tab + a, n, c->d_plist + (size_t)k0 * c->geo.nb,
                                                 c->d_pcount + (size_t)k0 * 4, c->d_err, c->tm_win_luma, c->tm_win_chroma, c->s_compute,
                                                 c->pcount_clean ? c->d_pcount + (size_t)c->max_jobs * 4 : nullptr));
                c->launches += 2;
            }
            a = b;
        }
    }
    if (c->want_kernel_time) { CU_TRY(cudaEventRecord(c->ev_k1, c->s_compute)); c->have_kernel_time = true; }
    CU_TRY(cudaEventRecord(st.ev_kernel, c->s_compute));
    const double tr4 = c->trace ? host_now() : 0;

    // copy out
    bool any_out = false;
    for (uint32_t i = 0; i < njobs; i++) any_out |= jobs[i].out_y != nullptr;
    if (any_out) {
        // Pictures of a batch leave on TWO copy streams in turn: a copy's set-up (a few microseconds per cudaMemcpyAsync, ~5 % of
        // a 3 MB picture at PCIe speed) then overlaps its neighbour's transfer.  Small pictures or single jobs keep one stream
        // (three more driver calls per submit would cost more than they hide).
        uint32_t n_out = 0;
        for (uint32_t i = 0; i < njobs; i++) n_out += jobs[i].out_y ? 1u : 0u;
        // (Measured and dropped, visit zy: the whole batch as ONE strided copy - cudaMemcpy2DAsync with the pictures as rows - when
        // slots and host buffers are equally spaced: 14.2 k frames/s against 14.9 k with one copy per picture on two streams, and
        // two buffers that merely LOOK equally spaced may belong to different allocations, which the copy rejects.)
        const bool two = c->d2h_streams2 && n_out >= 2 && (size_t)g.width * g.height >= ((size_t)1 << 20);
        CU_TRY(cudaStreamWaitEvent(s_down, st.ev_kernel, 0));
        if (two) CU_TRY(cudaStreamWaitEvent(c->s_d2h2, st.ev_kernel, 0));
        uint32_t k = 0;
        for (uint32_t i = 0; i < njobs; i++) {
            const DecIn &j = jobs[i];
            if (!j.out_y) continue;
            int rc = copy_visible(c, j.dst_slot, j.out_y, j.out_u, j.out_v, (two && (k++ & 1u)) ? c->s_d2h2 : s_down);
            if (rc) return rc;
            c->slot_last_d2h[j.dst_slot] = id;
        }
        if (two) {                                              // everything that follows the first stream follows both
            CU_TRY(cudaEventRecord(c->ev_d2h_join, c->s_d2h2));
            CU_TRY(cudaStreamWaitEvent(s_down, c->ev_d2h_join, 0));
        }
    }
    CU_TRY(cudaEventRecord(c->ev_d2h_ring[id % D2H_RING], s_down));
    if (c->trace) {
        const double tr5 = host_now();
        c->t_dec_check += tr1 - tr0; c->t_dec_wait += tr2 - tr1; c->t_dec_copy += tr3 - tr2; c->t_dec_launch += tr4 - tr3; c->t_dec_d2h += tr5 - tr4;
        c->n_dec_submits++; c->n_dec_jobs += njobs;
    }
    return PFV_OK;
}

// A submit that fails after its id was issued leaves the id without a recorded completion event: mark it, and record
// the ring event anyway so that a later pfv_ctx_wait_submit(id) neither succeeds on the record of 64 submits ago nor
// blocks - it reports the failure.
static int finish_submit(pfv_ctx *c, uint64_t id_before, int rc)
{
    const uint64_t id = c->submit_id;
    if (rc != PFV_OK && id != id_before) {
        __atomic_store_n(&c->failed_ring[id % D2H_RING], id, __ATOMIC_RELEASE);
        cudaEventRecord(c->ev_d2h_ring[id % D2H_RING], c->s_d2h);
    }
    return rc;
}

extern "C" int pfv_decode_submit(pfv_ctx *c, const pfv_decode_job *jobs, uint32_t njobs)
{
    if (!c || !jobs) return fail(PFV_ERR_BAD_ARG, "NULL argument");
    if (njobs == 0) return PFV_OK;
    std::vector<DecIn> in(njobs);
    for (uint32_t i = 0; i < njobs; i++) {
        const pfv_decode_job &j = jobs[i];
        DecIn &d = in[i];
        d.kind = j.kind; d.flags = j.flags; d.dst_slot = j.dst_slot; d.ref_slot = j.ref_slot;
        memcpy(d.qidx, j.qidx, 3);
        d.sparse = false;
        d.trusted = false;
        d.hdr = j.hdr; d.coeff = j.coeff;
        d.mb_off = nullptr; d.tok = nullptr; d.ntok = 0;
        d.out_y = j.out_y; d.out_u = j.out_u; d.out_v = j.out_v;
    }
    const uint64_t before = c->submit_id;
    return finish_submit(c, before, decode_submit_impl(c, in.data(), njobs));
}

static int decode_submit_sparse_any(pfv_ctx *c, const pfv_decode_job_sparse *jobs, uint32_t njobs, bool trusted)
{
    if (!c || !jobs) return fail(PFV_ERR_BAD_ARG, "NULL argument");
    if (njobs == 0) return PFV_OK;
    std::vector<DecIn> in(njobs);
    for (uint32_t i = 0; i < njobs; i++) {
        const pfv_decode_job_sparse &j = jobs[i];
        DecIn &d = in[i];
        d.kind = j.kind; d.flags = j.flags; d.dst_slot = j.dst_slot; d.ref_slot = j.ref_slot;
        memcpy(d.qidx, j.qidx, 3);
        d.sparse = true;
        d.trusted = trusted;
        d.hdr = j.hdr; d.coeff = nullptr;
        d.mb_off = j.mb_off; d.tok = j.tok; d.ntok = j.ntok;
        d.out_y = j.out_y; d.out_u = j.out_u; d.out_v = j.out_v;
    }
    const uint64_t before = c->submit_id;
    return finish_submit(c, before, decode_submit_impl(c, in.data(), njobs));
}

extern "C" int pfv_decode_submit_sparse(pfv_ctx *c, const pfv_decode_job_sparse *jobs, uint32_t njobs)
{
    return decode_submit_sparse_any(c, jobs, njobs, false);
}

// (internal, pfv_internal.h) for the Decoder object: mb_off comes out of pfv_packet_decode, which bounds every macroblock's tokens
int pfv_decode_submit_sparse_trusted(pfv_ctx *c, const pfv_decode_job_sparse *jobs, uint32_t njobs)
{
    return decode_submit_sparse_any(c, jobs, njobs, true);
}

extern "C" uint64_t pfv_ctx_last_submit_id(const pfv_ctx *c) { return c ? __atomic_load_n(&c->submit_id, __ATOMIC_RELAXED) : 0; }

extern "C" int pfv_ctx_wait_submit(pfv_ctx *c, uint64_t id)
{
    if (!c) return fail(PFV_ERR_BAD_ARG, "NULL context");
    // may be called from a helper thread while the owning thread keeps submitting: only the id counter is read
    const uint64_t last = __atomic_load_n(&c->submit_id, __ATOMIC_RELAXED);
    if (id == 0 || id > last) return fail(PFV_ERR_BAD_ARG, "submit id %llu has not been issued", (unsigned long long)id);
    if (id + D2H_RING <= last)
        return fail(PFV_ERR_BAD_ARG, "submit id %llu is older than the %d most recent submits", (unsigned long long)id, D2H_RING);
    if (__atomic_load_n(&c->failed_ring[id % D2H_RING], __ATOMIC_ACQUIRE) == id)
        return fail(PFV_ERR_CUDA, "submit %llu failed after it was issued (see the error it returned)", (unsigned long long)id);
    CU_TRY(ensure_device(c->device));
    CU_TRY(cudaEventSynchronize(c->ev_d2h_ring[id % D2H_RING]));
    return PFV_OK;
}

// (internal, pfv_internal.h) The same for ONE thread that is about to hand the result to its caller (Decoder::advance_frame): the
// events block (cudaEventBlockingSync - a pool of waiting threads must not spin, see pfv_ctx_create), and a blocked thread wakes up
// tens of microseconds after its event; here the event is polled for up to `spin_s` seconds first.
int pfv_ctx_wait_submit_polling(pfv_ctx *c, uint64_t id, double spin_s)
{
    if (!c) return fail(PFV_ERR_BAD_ARG, "NULL context");
    const uint64_t last = __atomic_load_n(&c->submit_id, __ATOMIC_RELAXED);
    if (id == 0 || id > last || id + D2H_RING <= last) return pfv_ctx_wait_submit(c, id);       // (its error messages)
    if (__atomic_load_n(&c->failed_ring[id % D2H_RING], __ATOMIC_ACQUIRE) == id) return pfv_ctx_wait_submit(c, id);
    CU_TRY(ensure_device(c->device));
    const double t0 = host_now();
    for (;;) {
        const cudaError_t e = cudaEventQuery(c->ev_d2h_ring[id % D2H_RING]);
        if (e == cudaSuccess) return PFV_OK;
        if (e != cudaErrorNotReady) CU_TRY(e);
        if (host_now() - t0 > spin_s) break;
    }
    CU_TRY(cudaEventSynchronize(c->ev_d2h_ring[id % D2H_RING]));
    return PFV_OK;
}

// ---------------------------------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------------------------------
namespace {

// one frame of either encode entry point
struct EncIn {
    uint32_t kind, flags, dst_slot, ref_slot;
    float    px_err;
    const uint8_t *src_y, *src_u, *src_v;
    pfv_mbhdr *hdr_out;
    int16_t   *coeff_out;                              // dense seam
    bool       sparse;                                 // sparse seam: the three below, already as device-accessible addresses
    uint32_t   tok_cap;
    uint32_t  *mb_off_out, *tok_out, *stats_out;
};

// the address the device uses for memory the caller handed in; nullptr if the device cannot reach it (pageable host memory)
uint32_t *device_view(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type == cudaMemoryTypeUnregistered || !a.devicePointer) return nullptr;
    return static_cast<uint32_t *>(a.devicePointer);
}

int encode_submit_impl(pfv_ctx *c, const EncIn *jobs, uint32_t njobs)
{
    if (njobs > c->max_jobs) return fail(PFV_ERR_BAD_ARG, "%u jobs > max_jobs %u", njobs, c->max_jobs);
    if (c->nq < 4) return fail(PFV_ERR_BAD_ARG, "an encoder context needs the 4 q-tables of src/enc.rs:48-51");
    const pfv_geometry &g = c->geo;
    bool any_p = false, any_rgb = false;
    uint32_t n_sparse = 0;
    for (uint32_t i = 0; i < njobs; i++) {
        const EncIn &j = jobs[i];
        if (j.kind != PFV_FRAME_I && j.kind != PFV_FRAME_P) return fail(PFV_ERR_BAD_ARG, "job %u: bad kind %u", i, j.kind);
        if (j.dst_slot >= c->nslots) return fail(PFV_ERR_BAD_ARG, "job %u: dst_slot %u out of range", i, j.dst_slot);
        const bool rgb = (j.flags & PFV_JOB_SRC_RGB) != 0;
        any_rgb |= rgb;
        n_sparse += j.sparse ? 1u : 0u;
        if (!j.src_y || (!rgb && (!j.src_u || !j.src_v)) || (!j.sparse && !j.coeff_out))
            return fail(PFV_ERR_BAD_ARG, "job %u: NULL plane or coeff_out", i);
        if (j.kind == PFV_FRAME_P) {
            any_p = true;
            if (j.ref_slot >= c->nslots) return fail(PFV_ERR_BAD_ARG, "job %u: ref_slot %u out of range", i, j.ref_slot);
            if (j.ref_slot == j.dst_slot) return fail(PFV_ERR_BAD_ARG, "job %u: ref_slot == dst_slot", i);
            if (!j.hdr_out) return fail(PFV_ERR_BAD_ARG, "job %u: hdr_out is NULL", i);
            if (!(j.px_err >= 0.0f)) return fail(PFV_ERR_BAD_ARG, "job %u: px_err must be >= 0", i);
        }
    }
    for (uint32_t i = 0; i < njobs; i++)
        if (jobs[i].kind == PFV_FRAME_P)
            for (uint32_t k = 0; k < njobs; k++)
                if (jobs[k].dst_slot == jobs[i].ref_slot)
                    return fail(PFV_ERR_BAD_ARG, "jobs %u and %u of one submit are dependent (ref_slot == dst_slot)", i, k);
    if (any_p && !c->have_tma) return fail(PFV_ERR_CUDA, "encode-P needs TMA tensor maps: %s", c->tma_err);
    // the divisors of the tables an encoder uses must be non-zero (the reference clamps them to >= 1, src/enc.rs:48-51; its
    // `n / d` would panic on 0, src/dct.rs:95); the quantiser's reciprocal multiply is exact up to 65535
    if (!c->enc_tables_ok)
        return fail(PFV_ERR_BAD_ARG, "encoding needs q-tables 0..3 (intra_l, intra_c, inter_l, inter_c) with every divisor in 1..%d",
                    (int)QUANT_MAX_DIVISOR);

    CU_TRY(ensure_device(c->device));
    const double t0 = c->trace ? host_now() : 0;
    {
        int rc = ensure_src_staging(c);
        if (rc) return rc;
    }
    const size_t rgb_bytes = (size_t)g.width * g.height * 3;
    if (any_rgb && !c->st[0].d_rgb_src)
        for (int i = 0; i < STAGES; i++) CU_TRY(cudaMalloc(&c->st[i].d_rgb_src, rgb_bytes * c->max_jobs));
    // per job slot of the sparse seam: mb_off (nb+1), RLE entries (nb*256), statistics; every part 16-byte aligned
    const size_t tp_off_words = ((size_t)g.nb + 1 + 3) & ~(size_t)3;
    const size_t tp_words = tp_off_words + (size_t)g.nb * 256 + PFV_TOKSTATS_WORDS;
    if (n_sparse && !c->st[0].d_tokpack)
        for (int i = 0; i < STAGES; i++) {
            CU_TRY(cudaMalloc(&c->st[i].d_tokpack, tp_words * sizeof(uint32_t) * c->max_jobs));
            CU_TRY(cudaMalloc(&c->st[i].d_tjobs, sizeof(TokJob) * c->max_jobs));
            CU_TRY(cudaHostAlloc(&c->st[i].h_tjobs, sizeof(TokJob) * c->max_jobs, cudaHostAllocDefault));
        }
    const uint64_t id = __atomic_add_fetch(&c->submit_id, 1, __ATOMIC_RELAXED);   // helper threads read it (pfv_ctx_wait_submit)
    Stage &st = c->st[id % STAGES];
    const double t1 = c->trace ? host_now() : 0;
    CU_TRY(cudaEventSynchronize(st.ev_h2d));
    const double t2 = c->trace ? host_now() : 0;
    CU_TRY(cudaStreamWaitEvent(c->s_h2d, st.ev_kernel, 0));
    CU_TRY(cudaStreamWaitEvent(c->s_h2d, st.ev_d2h, 0));           // tok_store_kernel of the stage's previous use reads d_tjobs

    EncJob *tab = static_cast<EncJob *>(st.h_jobs);
    struct RgbConv { const uint8_t *rgb; uint8_t *planes; };
    std::vector<RgbConv> conv;
    std::vector<uint32_t> order;
    order.reserve(njobs);
    uint32_t n_i = 0;
    for (uint32_t i = 0; i < njobs; i++) if (jobs[i].kind == PFV_FRAME_I) { order.push_back(i); n_i++; }
    for (uint32_t i = 0; i < njobs; i++) if (jobs[i].kind == PFV_FRAME_P) order.push_back(i);

    const size_t ysz = (size_t)g.width * g.height, csz = (size_t)g.cwidth * g.cheight;
    const size_t coeff_elems = (size_t)g.nb * 256;
    uint32_t n_tok = 0, n_tok_key = 0;                             // (key frames come first in `order`)
    for (uint32_t k = 0; k < njobs; k++) {
        const EncIn &j = jobs[order[k]];
        EncJob &d = tab[k];
        const bool dev = (j.flags & PFV_JOB_DEVICE_PTRS) != 0;
        if (j.flags & PFV_JOB_SRC_RGB) {
            // load_frame + from_planes on the device: RGB in (host: copied; device: used in place), tight planes out
            uint8_t *s = st.d_src + (size_t)k * c->src_stride;
            const uint8_t *d_rgb = j.src_y;
            if (!dev) {
                uint8_t *stage_rgb = st.d_rgb_src + (size_t)k * rgb_bytes;
                CU_TRY(cudaMemcpyAsync(stage_rgb, j.src_y, rgb_bytes, cudaMemcpyHostToDevice, c->s_h2d));
                d_rgb = stage_rgb;
            }
            conv.push_back({d_rgb, s});
            for (int p = 0; p < 3; p++) d.src[p] = s + c->src_off[p];
            if (dev) { d.coeff = j.coeff_out; d.hdr = j.hdr_out; }
            else { d.coeff = st.d_coeff + (size_t)k * coeff_elems; d.hdr = st.d_hdr + (size_t)k * g.nb; }
        } else if (dev) {
            d.src[0] = j.src_y; d.src[1] = j.src_u; d.src[2] = j.src_v;
            d.coeff = j.coeff_out;
            d.hdr = j.hdr_out;
        } else {
            uint8_t *s = st.d_src + (size_t)k * c->src_stride;
            const bool packed = j.src_u == j.src_y + ysz && j.src_v == j.src_u + csz &&
                                c->src_off[1] == ysz && c->src_off[2] == ysz + csz;
            if (packed) {
                CU_TRY(cudaMemcpyAsync(s, j.src_y, ysz + 2 * csz, cudaMemcpyHostToDevice, c->s_h2d));
            } else {
                CU_TRY(cudaMemcpyAsync(s + c->src_off[0], j.src_y, ysz, cudaMemcpyHostToDevice, c->s_h2d));
                CU_TRY(cudaMemcpyAsync(s + c->src_off[1], j.src_u, csz, cudaMemcpyHostToDevice, c->s_h2d));
                CU_TRY(cudaMemcpyAsync(s + c->src_off[2], j.src_v, csz, cudaMemcpyHostToDevice, c->s_h2d));
            }
            for (int p = 0; p < 3; p++) d.src[p] = s + c->src_off[p];
            d.coeff = st.d_coeff + (size_t)k * coeff_elems;
            d.hdr = st.d_hdr + (size_t)k * g.nb;
        }
        d.dst = slot_ptr(c, j.dst_slot);
        d.ref = j.kind == PFV_FRAME_P ? slot_ptr(c, j.ref_slot) : nullptr;
        d.ref_slot = j.kind == PFV_FRAME_P ? (int32_t)j.ref_slot : 0;
        d.min_err = j.px_err * j.px_err * 256.0f;                  // src/common.rs:209 (f32, left to right)
        d.mb_cnt = nullptr;
        if (j.sparse) {
            // the dense coefficients stay in the stage's device buffer; only their RLE sequence leaves the device
            d.coeff = st.d_coeff + (size_t)k * coeff_elems;
            TokJob &t = st.h_tjobs[n_tok++];
            uint32_t *pack = st.d_tokpack + (size_t)k * tp_words;
            t.coeff = d.coeff;
            t.hdr = j.kind == PFV_FRAME_P ? d.hdr : nullptr;
            t.mb_off = pack;
            d.mb_cnt = pack + 1;                                   // the encode kernel leaves each macroblock's entry count here
            t.tok = pack + tp_off_words;
            t.stats = pack + tp_off_words + (size_t)g.nb * 256;
            t.out_tok = j.tok_out; t.out_stats = j.stats_out; t.out_mb_off = j.mb_off_out;
            t.tok_cap = j.tok_cap;
            // key frames: every macroblock carries coefficients - the tokenizer counts and emits them itself, one thread per
            // sub-block (pfv_kernels_tok.cu), and the encode kernel does not count
            t.padded = j.kind == PFV_FRAME_I ? 1u : 0u;
            if (t.padded) { d.mb_cnt = nullptr; n_tok_key++; }
        }
    }
    CU_TRY(cudaMemcpyAsync(st.d_jobs, tab, sizeof(EncJob) * njobs, cudaMemcpyHostToDevice, c->s_h2d));
    if (n_tok) CU_TRY(cudaMemcpyAsync(st.d_tjobs, st.h_tjobs, sizeof(TokJob) * n_tok, cudaMemcpyHostToDevice, c->s_h2d));
    CU_TRY(cudaEventRecord(st.ev_h2d, c->s_h2d));
    const double t3 = c->trace ? host_now() : 0;

    CU_TRY(cudaStreamWaitEvent(c->s_compute, st.ev_h2d, 0));
    CU_TRY(cudaStreamWaitEvent(c->s_compute, st.ev_d2h, 0));       // earlier D2H out of this stage's coeff/hdr buffers
    {
        std::vector<uint32_t> dsts(njobs);
        for (uint32_t i = 0; i < njobs; i++) dsts[i] = jobs[i].dst_slot;
        int rc = wait_slot_readers(c, dsts.data(), njobs);
        if (rc) return rc;
    }
    if (c->want_kernel_time) CU_TRY(cudaEventRecord(c->ev_k0, c->s_compute));
    for (const RgbConv &rc : conv) {
        CU_TRY(launch_rgb_to_yuv420(rc.rgb, g.width, g.height, rc.planes + c->src_off[0], rc.planes + c->src_off[1],
                                    rc.planes + c->src_off[2], c->s_compute));
        c->launches++;
    }
    const EncJob *d_tab = static_cast<const EncJob *>(st.d_jobs);
    const bool count = n_tok != 0;                                 // an entry point's jobs are all dense or all sparse (P frames only:
                                                                   // the key-frame tokenizer counts for itself)
    if (n_i) {
        if (c->encode_i_variant == 2) {
            CU_TRY(launch_encode_i(c->fg, d_tab, n_i, c->d_qt, false, c->s_compute));
        } else {
            EncSbParams P;
            P.g = c->fg;
            for (int t = 0; t < 2; t++) {                              // intra_l, intra_c (src/enc.rs:84-90)
                memcpy(P.encR[t], &c->h_enc_recip[(size_t)t * 64], 64 * sizeof(float));
                memcpy(P.deq[t], &c->h_deq_scan[(size_t)t * 64], 64 * sizeof(int32_t));
            }
            CU_TRY(launch_encode_i_persist(P, d_tab, n_i, c->d_work + 4, c->s_compute));
        }
        c->launches++;
    }
    if (njobs - n_i) {
        {
            const bool gen1 = c->encode_p_variant == 1;
            CU_TRY(launch_encode_p(c->fg, d_tab + n_i, njobs - n_i, c->d_qt, gen1 ? c->tm_luma : c->tm_ep2_luma, gen1 ? c->tm_chroma : c->tm_ep2_chroma,
                                   count, c->encode_p_variant, c->d_work, c->s_compute));
        }
        c->launches++;
    }
    if (n_tok) {
        CU_TRY(launch_tokenize(g.nb, st.d_tjobs, n_tok_key, n_tok, c->s_compute));
        c->launches += (n_tok_key ? 3u : 0u) + (n_tok > n_tok_key ? 2u : 0u);
    }
    if (c->want_kernel_time) { CU_TRY(cudaEventRecord(c->ev_k1, c->s_compute)); c->have_kernel_time = true; }
    CU_TRY(cudaEventRecord(st.ev_kernel, c->s_compute));
    const double t4 = c->trace ? host_now() : 0;

    bool any_host = false;
    for (uint32_t i = 0; i < njobs; i++) any_host |= (jobs[i].flags & PFV_JOB_DEVICE_PTRS) == 0;
    if (any_host || n_tok) {
        CU_TRY(cudaStreamWaitEvent(c->s_d2h, st.ev_kernel, 0));
        // the device writes each frame's RLE sequence itself (pinned host memory is reached over PCIe): the copy is as
        // long as the sequence, which no host-issued cudaMemcpyAsync could know at submit time
        if (n_tok) { CU_TRY(launch_token_store(g.nb, st.d_tjobs, n_tok, c->s_d2h)); c->launches++; }
        for (uint32_t k = 0; k < njobs; k++) {
            const EncIn &j = jobs[order[k]];
            if (j.flags & PFV_JOB_DEVICE_PTRS) continue;
            if (!j.sparse)
                CU_TRY(cudaMemcpyAsync(j.coeff_out, st.d_coeff + (size_t)k * coeff_elems, coeff_elems * sizeof(int16_t),
                                       cudaMemcpyDeviceToHost, c->s_d2h));
            if (j.kind == PFV_FRAME_P)
                CU_TRY(cudaMemcpyAsync(j.hdr_out, st.d_hdr + (size_t)k * g.nb, (size_t)g.nb * sizeof(pfv_mbhdr),
                                       cudaMemcpyDeviceToHost, c->s_d2h));
        }
    }
    CU_TRY(cudaEventRecord(st.ev_d2h, c->s_d2h));
    st.d2h_used = true;
    CU_TRY(cudaEventRecord(c->ev_d2h_ring[id % D2H_RING], c->s_d2h));
    if (c->trace) {
        const double t5 = host_now();
        c->t_enc_alloc += t1 - t0; c->t_enc_wait += t2 - t1; c->t_enc_copy += t3 - t2; c->t_enc_launch += t4 - t3; c->t_enc_d2h += t5 - t4;
        c->n_enc_submits++;
    }
    return PFV_OK;
}

}  // namespace

extern "C" int pfv_encode_submit(pfv_ctx *c, const pfv_encode_job *jobs, uint32_t njobs)
{
    if (!c || !jobs) return fail(PFV_ERR_BAD_ARG, "NULL argument");
    if (njobs == 0) return PFV_OK;
    std::vector<EncIn> in(njobs);
    for (uint32_t i = 0; i < njobs; i++) {
        const pfv_encode_job &j = jobs[i];
        in[i] = EncIn{j.kind, j.flags, j.dst_slot, j.ref_slot, j.px_err, j.src_y, j.src_u, j.src_v, j.hdr_out, j.coeff_out,
                      false, 0, nullptr, nullptr, nullptr};
    }
    const uint64_t before = c->submit_id;
    return finish_submit(c, before, encode_submit_impl(c, in.data(), njobs));
}

extern "C" int pfv_encode_submit_sparse(pfv_ctx *c, const pfv_encode_job_sparse *jobs, uint32_t njobs)
{
    if (!c || !jobs) return fail(PFV_ERR_BAD_ARG, "NULL argument");
    if (njobs == 0) return PFV_OK;
    CU_TRY(ensure_device(c->device));
    std::vector<EncIn> in(njobs);
    for (uint32_t i = 0; i < njobs; i++) {
        const pfv_encode_job_sparse &j = jobs[i];
        if (!j.tok_out || !j.stats_out) return fail(PFV_ERR_BAD_ARG, "job %u: NULL tok_out or stats_out", i);
        uint32_t *tok = device_view(j.tok_out), *stats = device_view(j.stats_out);
        uint32_t *mb_off = j.mb_off_out ? device_view(j.mb_off_out) : nullptr;
        if (!tok || !stats || (j.mb_off_out && !mb_off))
            return fail(PFV_ERR_BAD_ARG, "job %u: tok_out / stats_out / mb_off_out must be pinned host memory (pfv_host_alloc) or "
                                         "device memory: the device stores the RLE sequence itself", i);
        in[i] = EncIn{j.kind, j.flags, j.dst_slot, j.ref_slot, j.px_err, j.src_y, j.src_u, j.src_v, j.hdr_out, nullptr,
                      true, j.tok_cap, mb_off, tok, stats};
    }
    const uint64_t before = c->submit_id;
    return finish_submit(c, before, encode_submit_impl(c, in.data(), njobs));
}
