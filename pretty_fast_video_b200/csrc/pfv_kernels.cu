// pfv_kernels.cu — the four macroblock kernels of the engine (sm_100a).
//
//   decode_kernel<false>  decode-I : dequant + IDCT + clamp                (src/common.rs:477-496)
//   decode_kernel<true>   decode-P : motion-compensated copy (+ IDCT delta) (src/common.rs:498-521)
//   encode_i_kernel       encode-I : FDCT + quant, fused closed-loop recon  (src/enc.rs:84-97)
//   encode_p_kernel       encode-P : 4-level SSD block search on a TMA-staged window, skip test,
//                                    residual FDCT + quant, fused recon      (src/enc.rs:134-147)
//
// One warp = one macroblock (see pfv_device.cuh).  grid.y = job (frame) index, so one launch covers a
// whole batch of independent frames.
#include <cstdlib>
#include <type_traits>

#include "pfv_internal.h"
#include "pfv_device.cuh"
#include "pfv_tok.cuh"

namespace pfv {

constexpr int WARPS_PER_CTA = 8;
constexpr unsigned FULL = 0xffffffffu;

struct MbPos {
    int      p;        // plane 0..2
    uint32_t bx, by;   // pixel origin of the macroblock inside the padded plane
};

__device__ __forceinline__ const PlaneGeom &plane_of(const FrameGeom &g, int p)
{
    return p == 0 ? g.pl[0] : (p == 1 ? g.pl[1] : g.pl[2]);
}

__device__ __forceinline__ MbPos locate_mb(const FrameGeom &g, uint32_t m)
{
    MbPos r;
    r.p = (m >= g.pl[1].mb_base ? 1 : 0) + (m >= g.pl[2].mb_base ? 1 : 0);
    const PlaneGeom &pl = plane_of(g, r.p);
    uint32_t col;
    const uint32_t row = div_small(m - pl.mb_base, pl.bw, pl.rcp_bw, col);
    r.bx = col * 16u;
    r.by = row * 16u;
    return r;
}

// -------------------------------------------------------------------------------------------------
// decode
// -------------------------------------------------------------------------------------------------
// Each warp walks MPW macroblocks, CTA-interleaved so the 8 warps of a CTA always work on 8 consecutive
// macroblocks (their 16-byte row segments then form 128-byte runs in the plane).
template <bool INTER, int MPW>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4)
decode_kernel(const __grid_constant__ FrameGeom g, const DecJob *__restrict__ jobs, int *__restrict__ err)
{
    __shared__ WarpScratch scratch[WARPS_PER_CTA];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpScratch &ws = scratch[warp];
    const DecJob job = jobs[blockIdx.y];
    const int sb = lane >> 3, r = lane & 7;
    const uint32_t py = (uint32_t)(sb >> 1) * 8u + (uint32_t)r, px = (uint32_t)(sb & 1) * 8u;

    uint32_t gaddr[8];
    lane_gather_offsets(lane, gaddr);
    int32_t deq[8];
    int cur_plane = -1;

    const uint32_t m0 = blockIdx.x * (WARPS_PER_CTA * MPW) + warp;
#pragma unroll 1
    for (int it = 0; it < MPW; ++it) {
        const uint32_t m = m0 + it * WARPS_PER_CTA;
        if (m >= g.nb) break;
        const MbPos pos = locate_mb(g, m);
        const PlaneGeom &pl = plane_of(g, pos.p);

        bool coded = true;
        uint2 prev = make_uint2(0u, 0u);
        if (INTER) {
            const pfv_mbhdr h = job.hdr[m];
            coded = h.has_coeff != 0;
            int sx = (int)pos.bx + h.mx, sy = (int)pos.by + h.my;      // src/common.rs:255-256
            if (sx < 0 || sy < 0 || sx > (int)pl.pw - 16 || sy > (int)pl.ph - 16) {
                // reference: debug_assert / slice panic (src/common.rs:258-259).  Never read out of
                // bounds: flag the stream as bad and fall back to the co-located block.
                if (lane == 0) atomicOr(err, ERRBIT_BAD_MV);
                sx = (int)pos.bx;
                sy = (int)pos.by;
            }
            prev = ldg_u8x8_unaligned(job.ref + pl.off + (size_t)((uint32_t)sy + py) * pl.pw + (uint32_t)sx + px);
        }
        uint2 out = prev;                                                // src/common.rs:281-283
        if (coded) {
            const uint4 craw = __ldcs(reinterpret_cast<const uint4 *>(job.coeff + (size_t)m * 256) + lane);
            if (pos.p != cur_plane) {
                const QTables *q = pos.p == 0 ? job.qt[0] : (pos.p == 1 ? job.qt[1] : job.qt[2]);
                lane_load8(q->deqT, r, deq);
                cur_plane = pos.p;
            }
            stage_coeffs(ws, lane, craw);
            __syncwarp();
            int y[8];
            decode_mb_core(ws, lane, gaddr, deq, y);
            out = INTER ? apply_residual_row(y, prev) : pack_row_u8(y);
        }
        uint8_t *dst = job.dst + pl.off + (size_t)(pos.by + py) * pl.pw + pos.bx + px;
        *reinterpret_cast<uint2 *>(dst) = out;                           // src/common.rs:341-349 blit_block
    }
}

// -------------------------------------------------------------------------------------------------
// encode helpers
// -------------------------------------------------------------------------------------------------
// Row (x..x+7, y) of a tight vw x vh source plane, padded with the clear colour (src/common.rs:352-356).
__device__ __forceinline__ uint2 load_src_row(const uint8_t *__restrict__ src, const PlaneGeom &pl,
                                              uint32_t x, uint32_t y, bool fast)
{
    if (y >= pl.vh || x >= pl.vw) return make_uint2(pl.clear4, pl.clear4);
    const uint8_t *p = src + (size_t)y * pl.vw + x;
    if (fast) return __ldcs(reinterpret_cast<const uint2 *>(p));       // vw % 8 == 0 and base 8-byte aligned
    uint32_t w[2] = {0u, 0u};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t b = (x + k < pl.vw) ? (uint32_t)p[k] : (pl.clear4 & 0xffu);
        w[k >> 2] |= b << (8 * (k & 3));
    }
    return make_uint2(w[0], w[1]);
}

__device__ __forceinline__ int byte_of(uint2 v, int k)
{
    const uint32_t w = (k < 4) ? v.x : v.y;
    return (int)((w >> (8 * (k & 3))) & 0xffu);
}

// -------------------------------------------------------------------------------------------------
// encode-I
// -------------------------------------------------------------------------------------------------
template <int MPW, bool COUNT>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 4)
encode_i_kernel(const __grid_constant__ FrameGeom g, const EncJob *__restrict__ jobs,
                const QTables *__restrict__ qt)
{
    __shared__ WarpScratch scratch[WARPS_PER_CTA];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpScratch &ws = scratch[warp];
    const EncJob job = jobs[blockIdx.y];
    const int sb = lane >> 3, r = lane & 7;
    const uint32_t py = (uint32_t)(sb >> 1) * 8u + (uint32_t)r, px = (uint32_t)(sb & 1) * 8u;

    uint32_t gaddr[8];
    lane_gather_offsets(lane, gaddr);

    const uint32_t m0 = blockIdx.x * (WARPS_PER_CTA * MPW) + warp;
#pragma unroll 1
    for (int it = 0; it < MPW; ++it) {
        const uint32_t m = m0 + it * WARPS_PER_CTA;
        if (m >= g.nb) break;
        const MbPos pos = locate_mb(g, m);
        const PlaneGeom &pl = plane_of(g, pos.p);
        const uint8_t *src = pos.p == 0 ? job.src[0] : (pos.p == 1 ? job.src[1] : job.src[2]);
        const bool fast = ((reinterpret_cast<uintptr_t>(src) | pl.vw) & 7u) == 0;
        const QTables *q = qt + (pos.p == 0 ? 0 : 1);                   // intra_l, intra_c (src/enc.rs:84-90)

        const uint2 s = load_src_row(src, pl, pos.bx + px, pos.by + py, fast);
        int x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = (byte_of(s, k) - 128) * 256;   // src/common.rs:291
        uint32_t encM[8];
        int32_t scale[8];
        lane_load8(reinterpret_cast<const int32_t *>(q->encM), r, reinterpret_cast<int32_t(&)[8]>(encM));
#pragma unroll
        for (int k = 0; k < 8; ++k) scale[k] = c_scaleT[r * 8 + k];
        const uint4 craw = encode_mb_core(ws, lane, gaddr, encM, scale, x);
        __stcs(reinterpret_cast<uint4 *>(job.coeff + (size_t)m * 256) + lane, craw);
        if (COUNT) {                                                   // sparse seam: how many RLE entries this macroblock makes
            const uint32_t n = tok::warp_count(craw, (uint32_t)lane);
            if (lane == 0) job.mb_cnt[m] = n;
        }

        // closed-loop reconstruction (src/enc.rs:85,88,91: decode_plane of what was just encoded)
        int32_t deq[8];
        lane_load8(q->deqT, r, deq);
        int y[8];
        decode_mb_core(ws, lane, gaddr, deq, y);
        uint8_t *dst = job.dst + pl.off + (size_t)(pos.by + py) * pl.pw + pos.bx + px;
        *reinterpret_cast<uint2 *>(dst) = pack_row_u8(y);
    }
}

// -------------------------------------------------------------------------------------------------
// encode-P
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// 8 bytes at an arbitrary byte offset of the shared-memory search window
__device__ __forceinline__ uint2 lds_u8x8_unaligned(const uint8_t *win, uint32_t off)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(win + (off & ~3u));
    const uint32_t sh = (off & 3u) * 8u;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];
    uint2 r;
    r.x = __funnelshift_r(w0, w1, sh);
    r.y = __funnelshift_r(w1, w2, sh);
    return r;
}

// src/common.rs:125-139 over the whole macroblock: this lane's 8 pixels, then a warp sum.  Exact
// integer SSD (<= 256*255^2 < 2^24, so the reference's f32 sum is the same number).
__device__ __forceinline__ uint32_t warp_ssd(uint2 src, uint2 ref)
{
    const uint32_t d0 = __vabsdiffu4(src.x, ref.x);
    const uint32_t d1 = __vabsdiffu4(src.y, ref.y);
    uint32_t acc = __dp4a(d0, d0, 0u);
    acc = __dp4a(d1, d1, acc);
    return __reduce_add_sync(FULL, acc);
}

// One CTA = one tile of 8 horizontally adjacent macroblocks of one plane.  The tile's whole search
// window (block_search never moves further than 8+4+2+1 = 15 px, src/common.rs:154-204) is fetched
// from the reference slot by ONE 4-D TMA box {WIN_W, WIN_H, 1, 1}; the part outside the plane is
// zero-filled by the TMA unit and never visited (candidates there are skipped, src/common.rs:171,182).
template <int CTAS_PER_SM, bool COUNT>   // COUNT: sparse encode seam, leave each macroblock's RLE entry count in job.mb_cnt
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CTAS_PER_SM)
encode_p_kernel(const __grid_constant__ FrameGeom g, const EncJob *__restrict__ jobs,
                const QTables *__restrict__ qt,
                const __grid_constant__ CUtensorMap tm_luma, const __grid_constant__ CUtensorMap tm_chroma)
{
    __shared__ __align__(128) uint8_t stage[WIN_BYTES];                // where the TMA box lands (row pitch 176 B)
    // The search reads 32 different window rows per instruction.  With the 16-byte granular pitch of the TMA box
    // (44 words) those rows fall on only 8 bank groups: 2.5 wavefronts per LDS whatever the lane -> row mapping.  The box
    // is therefore re-pitched once per tile to 47 words (188 B): consecutive rows then start 15 banks apart and the
    // same reads take 1.25 wavefronts (simulated over all search states; ncu: the data pipe was at 80 %, half of it
    // bank conflicts).
    __shared__ __align__(16) uint8_t win[WINP_BYTES + 16];
    __shared__ WarpScratch scratch[WARPS_PER_CTA];
    __shared__ __align__(8) uint64_t bar;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpScratch &ws = scratch[warp];
    const EncJob job = jobs[blockIdx.y];

    // tile -> plane, macroblock row, first macroblock column
    const uint32_t tile = blockIdx.x;
    const int p = (tile >= g.pl[1].tile_base ? 1 : 0) + (tile >= g.pl[2].tile_base ? 1 : 0);
    const PlaneGeom &pl = plane_of(g, p);
    const uint32_t lt = tile - pl.tile_base;
    const uint32_t trow = lt / pl.tiles_per_row, tcol = lt - trow * pl.tiles_per_row;
    const int tile_x0 = (int)tcol * 128, by = (int)trow * 16;

    if (threadIdx.x == 0) {
        const uint32_t bar_a = smem_u32(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)WIN_BYTES) : "memory");
        const CUtensorMap *tm = (p == 0) ? &tm_luma : &tm_chroma;
        const int cx = tile_x0 - 16, cy = by - 15, cz = (p == 2) ? 1 : 0, cw = job.ref_slot;
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(smem_u32(stage)), "l"(tm), "r"(cx), "r"(cy), "r"(cz), "r"(cw), "r"(bar_a)
            : "memory");
    }

    const int sb = lane >> 3, r = lane & 7;
    const uint32_t py = (uint32_t)(sb >> 1) * 8u + (uint32_t)r, px = (uint32_t)(sb & 1) * 8u;
    const int bx = tile_x0 + warp * 16;
    const bool active = bx < (int)pl.pw;                               // ragged last tile of a row

    uint2 s = make_uint2(0u, 0u);
    uint32_t gaddr[8];
    if (active) {
        const uint8_t *src = p == 0 ? job.src[0] : (p == 1 ? job.src[1] : job.src[2]);
        const bool fast = ((reinterpret_cast<uintptr_t>(src) | pl.vw) & 7u) == 0;
        s = load_src_row(src, pl, (uint32_t)bx + px, (uint32_t)by + py, fast);
        lane_gather_offsets(lane, gaddr);
    }

    __syncthreads();                                                   // barrier init visible to all
    {
        const uint32_t bar_a = smem_u32(&bar);
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(bar_a), "r"(0u) : "memory");
        }
    }
    {   // re-pitch the window: 46 rows x 44 words -> pitch 47 words (one 16-byte load, four word stores per item)
        const uint4 *src128 = reinterpret_cast<const uint4 *>(stage);
        uint32_t *dst32 = reinterpret_cast<uint32_t *>(win);
        constexpr uint32_t CHUNKS = WIN_W / 16;
        for (uint32_t i = threadIdx.x; i < (uint32_t)WIN_H * CHUNKS; i += WARPS_PER_CTA * 32) {
            const uint32_t row = i / CHUNKS, col = i - row * CHUNKS;
            const uint4 v = src128[i];
            uint32_t *d = dst32 + row * (WINP_W / 4) + col * 4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
    }
    __syncthreads();
    if (!active) return;

    // window coordinates of this lane's 8 pixels for motion (0,0)
    const int wx0 = 16 + warp * 16 + (int)px, wy0 = 15 + (int)py;
    const int max_x = (int)pl.pw - 16, max_y = (int)pl.ph - 16;

    // src/common.rs:154-204, iteratively.  The centre of each level after the first is the previous
    // level's winner, whose error is already known (the reference recomputes the identical number).
    //
    // One pass per level evaluates all 8 neighbours at once: lane = (candidate g = lane >> 2, quarter q = lane & 3),
    // candidate g in the reference's visiting order (my outer, mx inner, centre skipped, :168-179), the lane owns
    // macroblock rows 4q .. 4q+3 (16 pixels each).  The 4 partial sums of a candidate meet by two xor-shuffles;
    // the winner is ONE redux.min over (ssd << 3 | g): the smallest error, and among equal errors the earliest
    // candidate - exactly what the sequential strict `<` of :189 selects.  Candidates outside the plane (:171,:182)
    // get the key 0xffffffff.  (The first version looped over the candidates with the whole warp on each one:
    // ~30 instructions per candidate, 1 300 per macroblock; this is ~100 per level.)
    int cx = bx, cy = by;                                              // current centre, plane coords
    uint32_t best = warp_ssd(s, lds_u8x8_unaligned(win, (uint32_t)(wy0 * WINP_W + wx0)));
    const int g8 = lane >> 2, q4 = lane & 3;
    const int mxg = (g8 == 0 || g8 == 3 || g8 == 5) ? -1 : ((g8 == 1 || g8 == 6) ? 0 : 1);
    const int myg = g8 < 3 ? -1 : (g8 < 5 ? 0 : 1);
    uint4 srow[4];                                                     // source rows 4q .. 4q+3 out of the (sub-block, row) layout
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int R = 4 * q4 + i;
        const int l0 = ((R >> 3) * 2) * 8 + (R & 7);                   // lane holding columns 0..7 of row R; +8: columns 8..15
        srow[i].x = __shfl_sync(FULL, s.x, l0);
        srow[i].y = __shfl_sync(FULL, s.y, l0);
        srow[i].z = __shfl_sync(FULL, s.x, l0 + 8);
        srow[i].w = __shfl_sync(FULL, s.y, l0 + 8);
    }
    // Levels 8 and 4 start from a centre that is a multiple of 4 away from the macroblock origin (itself a multiple of 16), so
    // all their candidates are word aligned in the window: 4 words per row, no funnel shifts.  Levels 2 and 1 are general.
    auto level = [&](const int step, auto aligned_tag) {
        constexpr bool ALIGNED = decltype(aligned_tag)::value;
        const int ox = cx + mxg * step, oy = cy + myg * step;
        const bool valid = ox >= 0 && ox <= max_x && oy >= 0 && oy <= max_y;
        // window offset of row q of the candidate block; invalid candidates still lie inside the (zero-filled) window
        const uint32_t base = (uint32_t)((15 + 4 * q4 + (oy - by)) * WINP_W + 16 + warp * 16 + (ox - bx));
        const uint32_t sh = (base & 3u) * 8u;
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(win + (base & ~3u));
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t *w = wp + i * (WINP_W / 4);
            uint32_t r0, r1, r2, r3;
            if (ALIGNED) {
                r0 = w[0]; r1 = w[1]; r2 = w[2]; r3 = w[3];
            } else {
                const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
                r0 = __funnelshift_r(w0, w1, sh); r1 = __funnelshift_r(w1, w2, sh);
                r2 = __funnelshift_r(w2, w3, sh); r3 = __funnelshift_r(w3, w4, sh);
            }
            const uint32_t d0 = __vabsdiffu4(srow[i].x, r0);
            const uint32_t d1 = __vabsdiffu4(srow[i].y, r1);
            const uint32_t d2 = __vabsdiffu4(srow[i].z, r2);
            const uint32_t d3 = __vabsdiffu4(srow[i].w, r3);
            acc = __dp4a(d0, d0, acc);
            acc = __dp4a(d1, d1, acc);
            acc = __dp4a(d2, d2, acc);
            acc = __dp4a(d3, d3, acc);
        }
        acc += __shfl_xor_sync(FULL, acc, 1);
        acc += __shfl_xor_sync(FULL, acc, 2);
        const uint32_t key = valid ? ((acc << 3) | (uint32_t)g8) : 0xffffffffu;
        const uint32_t kmin = __reduce_min_sync(FULL, key);
        int bdx = 0, bdy = 0;
        if (kmin != 0xffffffffu && (kmin >> 3) < best) {               // strict, src/common.rs:189
            best = kmin >> 3;
            const int src_lane = (int)(kmin & 7u) * 4;
            bdx = __shfl_sync(FULL, mxg, src_lane) * step;
            bdy = __shfl_sync(FULL, myg, src_lane) * step;
        }
        cx += bdx;
        cy += bdy;
    };
#pragma unroll 1
    for (int step = 8; step >= 4; step >>= 1) level(step, std::true_type{});
#pragma unroll 1
    for (int step = 2; step >= 1; step >>= 1) level(step, std::false_type{});
    const int mvx = cx - bx, mvy = cy - by;                            // |mv| <= 15
    const uint2 prev = lds_u8x8_unaligned(win, (uint32_t)((wy0 + mvy) * WINP_W + wx0 + mvx));
    const bool coded = !((float)best <= job.min_err);                  // src/common.rs:221

    const uint32_t m = pl.mb_base + trow * pl.bw + (uint32_t)(bx >> 4);
    if (lane == 0) {
        pfv_mbhdr h;
        h.mx = (int8_t)mvx;                                            // src/common.rs:222,235 `as i8`
        h.my = (int8_t)mvy;
        h.has_coeff = coded ? 1 : 0;
        h.reserved = 0;
        job.hdr[m] = h;
    }
    uint2 out = prev;
    if (coded) {
        const QTables *q = qt + (p == 0 ? 2 : 3);                      // inter_l, inter_c (src/enc.rs:134-140)
        int x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int d = byte_of(s, k) - byte_of(prev, k);            // src/common.rs:118-119 (|d| <= 255)
            x[k] = (d / 2) * 256;                                      // src/common.rs:304
        }
        uint32_t encM[8];
        int32_t scale[8];
        lane_load8(reinterpret_cast<const int32_t *>(q->encM), r, reinterpret_cast<int32_t(&)[8]>(encM));
#pragma unroll
        for (int k = 0; k < 8; ++k) scale[k] = c_scaleT[r * 8 + k];
        const uint4 craw = encode_mb_core(ws, lane, gaddr, encM, scale, x);
        __stcs(reinterpret_cast<uint4 *>(job.coeff + (size_t)m * 256) + lane, craw);
        if (COUNT) {                                                   // sparse seam: how many RLE entries this macroblock makes
            const uint32_t n = tok::warp_count(craw, (uint32_t)lane);
            if (lane == 0) job.mb_cnt[m] = n;
        }
        int32_t deq[8];
        lane_load8(q->deqT, r, deq);
        int y[8];
        decode_mb_core(ws, lane, gaddr, deq, y);
        out = apply_residual_row(y, prev);                             // src/common.rs:277
    } else if (COUNT && lane == 0) {
        job.mb_cnt[m] = 0;                                             // subblocks: None (src/enc.rs:357-358)
    }
    uint8_t *dst = job.dst + pl.off + (size_t)((uint32_t)by + py) * pl.pw + (uint32_t)bx + px;
    *reinterpret_cast<uint2 *>(dst) = out;
}

// -------------------------------------------------------------------------------------------------
// encode-P, second generation (the default): one WARP per tile of 8 macroblocks, column-strip search
// -------------------------------------------------------------------------------------------------
// ncu on encode_p_kernel: 749 warp instructions per macroblock, issue bound; ~400 of them the search (one warp per
// macroblock, lane = (candidate, row quarter): every candidate row is loaded on its own, 16 loads + 32 SSD instructions
// + a reduction per lane and level), ~55 the window re-pitch and the source-row shuffles, two CTA barriers per tile.
//
// Here a warp owns the whole tile and its own search window (ONE 160 x 46 TMA box per tile into the warp's buffer - no CTA
// barrier anywhere; the other warps of the SM cover the window's arrival, see EP2_WARPS).  Lane = (macroblock
// m = lane >> 2, column strip q = lane & 3): it keeps the 4-pixel-wide strip of its macroblock's 16 source rows in 16
// registers and evaluates ALL 8 candidates of a level on that strip.  For a fixed horizontal offset the three vertical
// candidates read the same window words 8 / 4 / 2 / 1 rows apart, so each loaded word feeds up to three candidates:
// 96 / 72 / ~110 loads per lane and level for 8 candidates instead of 128, and per macroblock (4 lanes instead of 32) a
// quarter of the instructions.  The 32 lanes of a load read 32 consecutive words of a window row (4 m + q): no bank
// conflicts at the TMA box's own pitch, so the re-pitch pass is gone.  Strip sums meet by two xor-shuffles per
// candidate; the winner is the minimum of (ssd << 3 | visiting index) as before (src/common.rs:154-204, strict `<`
// of :189 = earliest candidate among equals).  Exact integer SSD (src/common.rs:125-139).
//
// Skipped macroblocks (src/common.rs:221-222) write their predictor strip straight to the reconstruction slot (full
// 128-byte rows per warp store).  Coded ones (~15 %) go one at a time through the warp-wide transform of the first
// generation (encode_mb_core / decode_mb_core: residual, FDCT + quantiser, closed-loop reconstruction).
// Warps per CTA (one CTA per SM).  Each warp has ONE window buffer: the next tile's window is requested when the last lane is
// done with this one and the warp waits for it at the top of the next tile.  The first form kept two buffers per warp (the next
// window in flight during the search) and, at 2 x 8 KB + scratch per warp, 12 warps per SM.  Measured on 32 x 1080p (round 2,
// visits zd / ze): 12 warps x 2 buffers 132.2 k frames/s, 12 x 1 132.7 k - the window in flight buys nothing, the other warps
// cover the ~1 us of a window's arrival - and with the room for more warps 16 x 1 142.3 k, 18 x 1 141.3 k, 20 x 1 144.1 k
// (96 registers), 22 x 1 130.0 k (80 registers, spills).  16 = four per scheduler at 128 registers.
#ifndef PFV_EP2_NWARPS
#define PFV_EP2_NWARPS 16
#endif
constexpr int EP2_WARPS = PFV_EP2_NWARPS;
constexpr int EP2_MAX_JOBS = 64;
constexpr int EP2_WIN_STAGE = (EP2_WIN_BYTES + 127) & ~127;

struct __align__(128) Ep2Smem {
    uint8_t     win[EP2_WARPS][EP2_WIN_STAGE];
    WarpScratch scratch[EP2_WARPS];
    uint64_t    bar[EP2_WARPS];
    EncJob      job[EP2_MAX_JOBS];
};
static_assert(sizeof(Ep2Smem) + 1024 <= 227 * 1024, "the encode-P CTA must fit one SM");

// 4 bytes at an arbitrary byte offset of a shared-memory window
__device__ __forceinline__ uint32_t lds_u8x4_unaligned(const uint8_t *win, uint32_t off)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(win + (off & ~3u));
    return __funnelshift_r(w[0], w[1], (off & 3u) * 8u);
}

__device__ __forceinline__ uint32_t ssd4(uint32_t a, uint32_t b, uint32_t acc)
{
    const uint32_t d = __vabsdiffu4(a, b);
    return __dp4a(d, d, acc);
}

// 4 source bytes (x .. x+3, y) of a tight vw x vh plane, padded with the clear colour (src/common.rs:352-356)
__device__ __noinline__ uint32_t load_src4_ragged(const uint8_t *__restrict__ src, uint32_t vw, uint32_t vh, uint32_t clear4,
                                                  uint32_t x, uint32_t y)
{
    if (y >= vh) return clear4;
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) w |= ((x + k < vw) ? (uint32_t)src[(size_t)y * vw + x + k] : (clear4 & 0xffu)) << (8 * k);
    return w;
}

// One search level (src/common.rs:165-196 for one step): S = this lane's source strip, A = the sum of its squares, `base` =
// byte offset in the window of the strip's top-left pixel displaced by the current centre.  Returns the level's best key
// (ssd << 3 | g), 0xffffffff if no candidate lies inside the plane.  ALIGNED: every candidate is word aligned.
//
// The error of a candidate strip is taken apart, sum (a - b)^2 = sum a^2 + sum b^2 - 2 sum a b, all three exact integers
// (< 2^24 for a whole macroblock; src/common.rs:125-139 adds the same squares in f32, exact below 2^24 as well): the three
// vertical candidates of a column share every window word, so sum b^2 is ONE running dot product per loaded word (the
// candidates' sums are differences of its prefix), and each candidate adds one dot product with the source word - 4 multiply-
// add-pipe instructions per loaded word at most where |a - b| first (vabsdiff4 + dp4a per candidate) took 6, three of them on
// the ALU pipe, which ncu had as this kernel's busiest (52 % against 22 % for the multiply-add pipe, math-pipe throttle 0.35).
//
// The loop over the horizontal offset is NOT unrolled: straight-line code of this size runs at the speed of the
// instruction fetch (ncu on the encode-I kernel: a quarter of all stall samples "no instruction"), a loop body of
// ~120 instructions stays in the instruction cache.  For mx = 0 the middle candidate is the centre itself: it is
// computed with the others (its key is discarded) so that the three passes share one body (a separate two-candidate
// body for mx = 0 saves 4 % of the search's arithmetic and was measured 5 % SLOWER: 680 more instructions of code).
template <int STEP, bool ALIGNED>
__device__ __forceinline__ uint32_t search_level(const uint8_t *win, const uint32_t (&S)[16], uint32_t A, uint32_t base, uint32_t valid_mask)
{
    constexpr unsigned FULL = 0xffffffffu;
    uint32_t kmin = 0xffffffffu;
    // candidates in the reference's visiting order: g = 0,1,2: my = -1 (mx = -1,0,1); 3,4: my = 0 (mx = -1, 1); 5,6,7: my = +1
#pragma unroll 1
    for (int mx = -1; mx <= 1; ++mx) {
        uint32_t cu = 0, cm = 0, cd = 0;                          // sum a b of the three candidates
        uint32_t t = 0, t_s = 0, t_2s = 0, t_16 = 0, t_16s = 0;   // prefix of sum b^2 down the column, and where the candidates cut it
        const uint32_t col = base + (uint32_t)(mx * STEP) - (uint32_t)(STEP * EP2_WIN_W);     // top row of the my = -1 candidate
#pragma unroll
        for (int rho = 0; rho < 16 + 2 * STEP; ++rho) {
            const uint32_t off = col + (uint32_t)(rho * EP2_WIN_W);
            const uint32_t w = ALIGNED ? *reinterpret_cast<const uint32_t *>(win + off) : lds_u8x4_unaligned(win, off);
            if (rho == STEP) t_s = t;
            if (rho == 2 * STEP) t_2s = t;
            if (rho == 16) t_16 = t;
            if (rho == 16 + STEP) t_16s = t;
            t = __dp4a(w, w, t);
            if (rho < 16) cu = __dp4a(S[rho], w, cu);
            if (rho >= STEP && rho < 16 + STEP) cm = __dp4a(S[rho - STEP], w, cm);
            if (rho >= 2 * STEP) cd = __dp4a(S[rho - 2 * STEP], w, cd);
        }
        // this strip's share of the three errors (each a sum of squares itself: >= 0, < 2^22)
        uint32_t up = A + t_16 - 2u * cu;
        uint32_t mid = A + (t_16s - t_s) - 2u * cm;
        uint32_t dn = A + (t - t_2s) - 2u * cd;
        up += __shfl_xor_sync(FULL, up, 1);
        up += __shfl_xor_sync(FULL, up, 2);
        mid += __shfl_xor_sync(FULL, mid, 1);
        mid += __shfl_xor_sync(FULL, mid, 2);
        dn += __shfl_xor_sync(FULL, dn, 1);
        dn += __shfl_xor_sync(FULL, dn, 2);
        const uint32_t g_up = (uint32_t)(mx + 1), g_mid = mx < 0 ? 3u : 4u, g_dn = (uint32_t)(mx + 6);
        if ((valid_mask >> g_up) & 1u) kmin = min(kmin, (up << 3) | g_up);
        if (mx != 0 && ((valid_mask >> g_mid) & 1u)) kmin = min(kmin, (mid << 3) | g_mid);
        if ((valid_mask >> g_dn) & 1u) kmin = min(kmin, (dn << 3) | g_dn);
    }
    return kmin;
}

// The two fine levels (steps 2 and 1), where no candidate is word aligned.  ncu, round 2: the kernel's busiest unit is the
// shared-memory data pipe (52 M wavefronts per 32 x 1080p = 0.71 of its cycles) - every candidate word of these levels cost two
// loads (the two words it straddles), and after the first level the eight macroblocks of a warp read rows and columns of
// their own, so that a load is 2.2 - 2.5 wavefronts.  Here the 2 STEP + 4 bytes a row offers the three horizontal offsets are
// loaded ONCE (three words) and brought to the left-most candidate's alignment in registers (X0, X1: two funnel shifts); the
// candidate word of offset mx is then one more funnel shift by (mx + 1) STEP bytes (clamped: 4 bytes = the second register).
// Loads per row 6 -> 3, instructions per row 9 -> 8.
template <int STEP>
__device__ __forceinline__ uint32_t search_level_fine(const uint8_t *win, const uint32_t (&S)[16], uint32_t A, uint32_t base, uint32_t valid_mask)
{
    static_assert(STEP == 1 || STEP == 2, "2 STEP + 4 bytes must fit two words");
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int ROWS = 16 + 2 * STEP;
    const uint32_t col0 = base - (uint32_t)STEP - (uint32_t)(STEP * EP2_WIN_W);    // first byte of the top row of candidate (-1, -1)
    const uint32_t *wp = reinterpret_cast<const uint32_t *>(win + (col0 & ~3u));
    const uint32_t sh = (col0 & 3u) * 8u;                                      // (EP2_WIN_W % 4 == 0: the same for every row)
    uint32_t X0[ROWS], X1[ROWS];
#pragma unroll
    for (int rho = 0; rho < ROWS; ++rho) {
        const uint32_t w0 = wp[rho * (EP2_WIN_W / 4)], w1 = wp[rho * (EP2_WIN_W / 4) + 1], w2 = wp[rho * (EP2_WIN_W / 4) + 2];
        X0[rho] = __funnelshift_r(w0, w1, sh);
        X1[rho] = __funnelshift_r(w1, w2, sh);
    }
    uint32_t kmin = 0xffffffffu;
#pragma unroll 1
    for (int mx = -1; mx <= 1; ++mx) {
        const uint32_t shv = (uint32_t)((mx + 1) * STEP) * 8u;                 // 0, 8 STEP, 16 STEP <= 32
        uint32_t cu = 0, cm = 0, cd = 0;
        uint32_t t = 0, t_s = 0, t_2s = 0, t_16 = 0, t_16s = 0;
#pragma unroll
        for (int rho = 0; rho < ROWS; ++rho) {
            const uint32_t w = __funnelshift_rc(X0[rho], X1[rho], shv);
            if (rho == STEP) t_s = t;
            if (rho == 2 * STEP) t_2s = t;
            if (rho == 16) t_16 = t;
            if (rho == 16 + STEP) t_16s = t;
            t = __dp4a(w, w, t);
            if (rho < 16) cu = __dp4a(S[rho], w, cu);
            if (rho >= STEP && rho < 16 + STEP) cm = __dp4a(S[rho - STEP], w, cm);
            if (rho >= 2 * STEP) cd = __dp4a(S[rho - 2 * STEP], w, cd);
        }
        uint32_t up = A + t_16 - 2u * cu;
        uint32_t mid = A + (t_16s - t_s) - 2u * cm;
        uint32_t dn = A + (t - t_2s) - 2u * cd;
        up += __shfl_xor_sync(FULL, up, 1);
        up += __shfl_xor_sync(FULL, up, 2);
        mid += __shfl_xor_sync(FULL, mid, 1);
        mid += __shfl_xor_sync(FULL, mid, 2);
        dn += __shfl_xor_sync(FULL, dn, 1);
        dn += __shfl_xor_sync(FULL, dn, 2);
        const uint32_t g_up = (uint32_t)(mx + 1), g_mid = mx < 0 ? 3u : 4u, g_dn = (uint32_t)(mx + 6);
        if ((valid_mask >> g_up) & 1u) kmin = min(kmin, (up << 3) | g_up);
        if (mx != 0 && ((valid_mask >> g_mid) & 1u)) kmin = min(kmin, (mid << 3) | g_mid);
        if ((valid_mask >> g_dn) & 1u) kmin = min(kmin, (dn << 3) | g_dn);
    }
    return kmin;
}

template <bool COUNT>
__global__ void __launch_bounds__(EP2_WARPS * 32, 1)
encode_p2_kernel(const __grid_constant__ FrameGeom g, const EncJob *__restrict__ jobs, uint32_t njobs,
                 const QTables *__restrict__ qt, float rcp_tiles, uint32_t *__restrict__ work,
                 const __grid_constant__ CUtensorMap tm_luma, const __grid_constant__ CUtensorMap tm_chroma)
{
    extern __shared__ __align__(128) unsigned char ep2_raw[];
    Ep2Smem &sm = *reinterpret_cast<Ep2Smem *>(ep2_raw);
    constexpr unsigned FULL = 0xffffffffu;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    for (uint32_t i = threadIdx.x; i < njobs; i += blockDim.x) sm.job[i] = jobs[i];
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&sm.bar[warp])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t nitems = njobs * g.total_tiles;
    const uint32_t nwarps = blockDim.x >> 5;                      // EP2_WARPS (fewer only as a tuning experiment)
    const uint32_t stride = gridDim.x * nwarps;
    const uint32_t m8 = lane >> 2, q = lane & 3u;
    WarpScratch &ws = sm.scratch[warp];

    struct Item { uint32_t job; int p; uint32_t trow; int tile_x0, by; };
    auto item_of = [&](uint32_t it) {
        Item r;
        uint32_t tile;
        r.job = div_small(it, g.total_tiles, rcp_tiles, tile);
        r.p = (tile >= g.pl[1].tile_base ? 1 : 0) + (tile >= g.pl[2].tile_base ? 1 : 0);
        const PlaneGeom &pl = plane_of(g, r.p);
        const uint32_t lt = tile - pl.tile_base;
        uint32_t tcol;
        r.trow = div_small(lt, pl.tiles_per_row, pl.rcp_tiles_per_row, tcol);
        r.tile_x0 = (int)tcol * 128;
        r.by = (int)r.trow * 16;
        return r;
    };
    auto issue = [&](const Item &it) {                           // lane 0: the tile's search window into the warp's buffer
        const uint32_t bar_a = smem_u32(&sm.bar[warp]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)EP2_WIN_BYTES) : "memory");
        const CUtensorMap *tm = (it.p == 0) ? &tm_luma : &tm_chroma;
        const int cx = it.tile_x0 - 16, cy = it.by - 15, cz = (it.p == 2) ? 1 : 0, cw = sm.job[it.job].ref_slot;
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(smem_u32(sm.win[warp])), "l"(tm), "r"(cx), "r"(cy), "r"(cz), "r"(cw), "r"(bar_a)
            : "memory");
    };
    // this lane's source strip of an item: bytes (bx + 4q .. +3, by + r), r = 0..15
    auto load_strip = [&](const Item &it, uint32_t (&S)[16]) {
        const PlaneGeom &pl = plane_of(g, it.p);
        const EncJob &job = sm.job[it.job];
        const uint8_t *src = it.p == 0 ? job.src[0] : (it.p == 1 ? job.src[1] : job.src[2]);
        const uint32_t x = (uint32_t)it.tile_x0 + m8 * 16u + q * 4u;
        const bool fast = ((reinterpret_cast<uintptr_t>(src) | pl.vw) & 3u) == 0;
        if (fast && (uint32_t)it.by + 16u <= pl.vh && (uint32_t)it.tile_x0 + 128u <= pl.vw) {      // the whole tile is picture (most are)
            const uint8_t *s0 = src + (size_t)(uint32_t)it.by * pl.vw + x;
#pragma unroll
            for (int r = 0; r < 16; ++r) S[r] = __ldcs(reinterpret_cast<const uint32_t *>(s0 + (size_t)r * pl.vw));
            return;
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const uint32_t y = (uint32_t)it.by + (uint32_t)r;
            if (fast) S[r] = (y < pl.vh && x < pl.vw) ? __ldcs(reinterpret_cast<const uint32_t *>(src + (size_t)y * pl.vw + x)) : pl.clear4;
            else      S[r] = load_src4_ragged(src, pl.vw, pl.vh, pl.clear4, x, y);
        }
    };

    // Tiles are handed out in order by a device-wide counter (work[0]): a warp's first tile is its static one, every further
    // one the next that nobody has taken.  With the static walk (tile += warps of the grid) the SMs finished up to 17 % apart
    // (ncu: sm__cycles_active min / avg / max 457 k / 499 k / 548 k) because a tile costs what its coded macroblocks cost.
    // The fetch-and-add for the tile after next is issued at the top of an iteration and its result first used at the bottom.
    // work[1] counts the warps that are done; the last one zeroes both words for the next launch.
    auto grab = [&]() -> uint32_t {                             // lane 0's value counts
        return lane == 0 ? stride + atomicAdd(&work[0], 1u) : 0u;
    };
    auto retire = [&]() {
        if (lane == 0 && atomicAdd(&work[1], 1u) == stride - 1u) { work[0] = 0u; work[1] = 0u; }
    };
    uint32_t it = blockIdx.x * nwarps + warp;
    if (it >= nitems) { retire(); return; }
    uint32_t it_next = __shfl_sync(FULL, grab(), 0);
    Item cur = item_of(it);
    if (lane == 0) issue(cur);
    uint32_t S[16];
    load_strip(cur, S);
    uint32_t k = 0;
#pragma unroll 1
    for (; it < nitems; ++k) {
        const bool has_next = it_next < nitems;
        Item nxt = cur;
        if (has_next) nxt = item_of(it_next);
        const uint32_t grabbed = has_next ? grab() : 0xffffffffu;   // in flight until the bottom of the iteration
        const PlaneGeom &pl = plane_of(g, cur.p);
        const EncJob &job = sm.job[cur.job];
        const uint8_t *win = sm.win[warp];
        {
            const uint32_t bar_a = smem_u32(&sm.bar[warp]);
            const uint32_t parity = k & 1u;
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done) : "r"(bar_a), "r"(parity) : "memory");
            }
        }
        const int bx = cur.tile_x0 + (int)m8 * 16, by = cur.by;
        const bool active = bx < (int)pl.pw;                      // ragged last tile of a row
        const int max_x = (int)pl.pw - 16, max_y = (int)pl.ph - 16;
        // window byte offset of this lane's strip at motion (0, 0)
        const uint32_t o0 = (uint32_t)(15 * EP2_WIN_W + 16) + m8 * 16u + q * 4u;

        // src/common.rs:154-204, iteratively; the centre's error at every level after the first is the previous winner's
        int cx = 0, cy = 0;
        uint32_t best, A = 0;                                      // A: sum of squares of this lane's source strip
#pragma unroll
        for (int r = 0; r < 16; ++r) A = __dp4a(S[r], S[r], A);
        {
            uint32_t a = 0;
#pragma unroll
            for (int r = 0; r < 16; ++r) a = ssd4(S[r], *reinterpret_cast<const uint32_t *>(win + o0 + r * EP2_WIN_W), a);
            a += __shfl_xor_sync(FULL, a, 1);
            a += __shfl_xor_sync(FULL, a, 2);
            best = a;
        }
        // which of the 8 candidates of a level lie inside the plane (src/common.rs:171,182): three column tests and three row
        // tests combined (bit g = candidate g in visiting order: my = -1: g 0..2, my = 0: g 3 (mx = -1), 4 (mx = +1), my = +1: g 5..7)
        // (a macroblock at least 15 pixels from every edge has them all, at every level: most tiles skip the tests)
        const bool all_inside = __all_sync(FULL, bx >= 15 && bx + 15 <= max_x && by >= 15 && by + 15 <= max_y);
        auto valid_mask_of = [&](int step) {
            if (all_inside) return 0xffu;
            uint32_t vx = 0, vy = 0;
#pragma unroll
            for (int d = -1; d <= 1; ++d) {
                const int ox = bx + cx + d * step, oy = by + cy + d * step;
                vx |= (ox >= 0 && ox <= max_x) ? (1u << (d + 1)) : 0u;
                vy |= (oy >= 0 && oy <= max_y) ? (1u << (d + 1)) : 0u;
            }
            return ((vy & 1u) ? vx : 0u) | ((vy & 2u) ? ((vx & 1u) | ((vx >> 1) & 2u)) << 3 : 0u) | ((vy & 4u) ? vx << 5 : 0u);
        };
        auto take = [&](uint32_t kmin, int step) {
            if (kmin != 0xffffffffu && (kmin >> 3) < best) {        // strict, src/common.rs:189
                best = kmin >> 3;
                const uint32_t gi = kmin & 7u;
                const int mx = (int)((0x9224u >> (2u * gi)) & 3u) - 1;   // mx + 1 = 0,1,2,0,2,0,1,2 for g = 0..7, two bits each
                const int my = gi < 3u ? -1 : (gi < 5u ? 0 : 1);
                cx += mx * step;
                cy += my * step;
            }
        };
        take(search_level<8, true>(win, S, A, o0, valid_mask_of(8)), 8);
        take(search_level<4, true>(win, S, A, (uint32_t)((int)o0 + cy * EP2_WIN_W + cx), valid_mask_of(4)), 4);
        take(search_level_fine<2>(win, S, A, (uint32_t)((int)o0 + cy * EP2_WIN_W + cx), valid_mask_of(2)), 2);
        take(search_level_fine<1>(win, S, A, (uint32_t)((int)o0 + cy * EP2_WIN_W + cx), valid_mask_of(1)), 1);

        const bool coded = active && !((float)best <= job.min_err);    // src/common.rs:221
        const uint32_t m = pl.mb_base + cur.trow * pl.bw + (uint32_t)(bx >> 4);
        if (active && q == 0) {
            pfv_mbhdr h;
            h.mx = (int8_t)cx;                                     // src/common.rs:222,235 `as i8`
            h.my = (int8_t)cy;
            h.has_coeff = coded ? 1 : 0;
            h.reserved = 0;
            job.hdr[m] = h;
            if (COUNT && !coded) job.mb_cnt[m] = 0;                // subblocks: None (src/enc.rs:357-358)
        }
        // source strip of the next tile: in flight during the rest of this one
        uint32_t Sn[16];
        if (has_next) load_strip(nxt, Sn);

        if (active && !coded) {                                    // the predictor is the reconstruction (src/common.rs:281-283)
            const uint32_t po = (uint32_t)((int)o0 + cy * EP2_WIN_W + cx);
            uint8_t *dst = job.dst + pl.off + (size_t)(uint32_t)by * pl.pw + (uint32_t)bx + q * 4u;
            // eight rows' loads first, then their shifts and stores (row by row every funnel shift waited for its own
            // shared-memory round trip: ncu, ~50 stall samples on each of the 16)
            const uint32_t *wbase = reinterpret_cast<const uint32_t *>(win + (po & ~3u));   // EP2_WIN_W % 4 == 0: one shift for all rows
            const uint32_t sh = (po & 3u) * 8u;
#pragma unroll
            for (int r0 = 0; r0 < 16; r0 += 8) {
                uint32_t a[8], b[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) { a[r] = wbase[(r0 + r) * (EP2_WIN_W / 4)]; b[r] = wbase[(r0 + r) * (EP2_WIN_W / 4) + 1]; }
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    *reinterpret_cast<uint32_t *>(dst + (size_t)(r0 + r) * pl.pw) = __funnelshift_r(a[r], b[r], sh);
            }
        }
        // coded macroblocks, one at a time, the whole warp on each (lane = (sub-block, row) as in pfv_device.cuh)
        uint32_t todo = __ballot_sync(FULL, coded && q == 0);
        if (todo) {
            const int sbk = (int)(lane >> 3), r8 = (int)(lane & 7u);
            const uint32_t py = (uint32_t)(sbk >> 1) * 8u + (uint32_t)r8, px = (uint32_t)(sbk & 1) * 8u;
            uint32_t gaddr[8];
            lane_gather_offsets((int)lane, gaddr);
            const QTables *qq = qt + (cur.p == 0 ? 2 : 3);         // inter_l, inter_c (src/enc.rs:134-140)
            uint32_t encM[8];
            int32_t scale[8], deq[8];
            lane_load8(reinterpret_cast<const int32_t *>(qq->encM), r8, reinterpret_cast<int32_t(&)[8]>(encM));
            lane_load8(qq->deqT, r8, deq);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) scale[kk] = c_scaleT[r8 * 8 + kk];
            const uint8_t *src = cur.p == 0 ? job.src[0] : (cur.p == 1 ? job.src[1] : job.src[2]);
            const bool fast8 = ((reinterpret_cast<uintptr_t>(src) | pl.vw) & 7u) == 0;
            // this lane's 8 source pixels of the first coded macroblock; inside the loop the NEXT one's are fetched before
            // the current one is transformed (the load's latency was fully exposed once per coded macroblock: ncu, 5 % of
            // all stall samples on its first use)
            uint2 s8n = load_src_row(src, pl, (uint32_t)(cur.tile_x0 + ((__ffs((int)todo) - 1) >> 2) * 16) + px, (uint32_t)by + py, fast8);
#pragma unroll 1
            while (todo) {
                const int l0 = __ffs((int)todo) - 1;
                todo &= todo - 1u;
                const int mcx = __shfl_sync(FULL, cx, l0), mcy = __shfl_sync(FULL, cy, l0);
                const uint32_t mm = __shfl_sync(FULL, m, l0);
                const int mbx = cur.tile_x0 + (l0 >> 2) * 16;
                const uint2 s8 = s8n;
                if (todo) s8n = load_src_row(src, pl, (uint32_t)(cur.tile_x0 + ((__ffs((int)todo) - 1) >> 2) * 16) + px, (uint32_t)by + py, fast8);
                const uint2 prev = lds_u8x8_unaligned(win, (uint32_t)((15 + (int)py + mcy) * EP2_WIN_W + 16 + (l0 >> 2) * 16 + (int)px + mcx));
                int x[8];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const int d = byte_of(s8, kk) - byte_of(prev, kk);   // src/common.rs:118-119 (|d| <= 255)
                    x[kk] = (d / 2) * 256;                               // src/common.rs:304
                }
                const uint4 craw = encode_mb_core(ws, (int)lane, gaddr, encM, scale, x);
                __stcs(reinterpret_cast<uint4 *>(job.coeff + (size_t)mm * 256) + lane, craw);
                if (COUNT) {
                    const uint32_t n = tok::warp_count(craw, lane);
                    if (lane == 0) job.mb_cnt[mm] = n;
                }
                int y[8];
                decode_mb_core(ws, (int)lane, gaddr, deq, y);
                uint8_t *dst = job.dst + pl.off + (size_t)((uint32_t)by + py) * pl.pw + (uint32_t)mbx + px;
                *reinterpret_cast<uint2 *>(dst) = apply_residual_row(y, prev);   // src/common.rs:277
            }
        }
        __syncwarp();                                              // every lane is done with the window
        if (has_next && lane == 0) issue(nxt);
#pragma unroll
        for (int r = 0; r < 16; ++r) S[r] = Sn[r];
        cur = nxt;
        it = it_next;
        it_next = __shfl_sync(FULL, grabbed, 0);
    }
    retire();
}

// -------------------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------------------
constexpr int DEC_MPW = 2;
constexpr int ENC_MPW = 1;

cudaError_t launch_decode(bool inter, const FrameGeom &g, const DecJob *d_jobs, uint32_t njobs,
                          int *d_err, cudaStream_t s)
{
    const uint32_t per_cta = WARPS_PER_CTA * DEC_MPW;
    dim3 grid((g.nb + per_cta - 1) / per_cta, njobs, 1), block(WARPS_PER_CTA * 32, 1, 1);
    if (inter) decode_kernel<true, DEC_MPW><<<grid, block, 0, s>>>(g, d_jobs, d_err);
    else       decode_kernel<false, DEC_MPW><<<grid, block, 0, s>>>(g, d_jobs, d_err);
    return cudaGetLastError();
}

cudaError_t launch_encode_i(const FrameGeom &g, const EncJob *d_jobs, uint32_t njobs,
                            const QTables *d_qt, bool count, cudaStream_t s)
{
    const uint32_t per_cta = WARPS_PER_CTA * ENC_MPW;
    dim3 grid((g.nb + per_cta - 1) / per_cta, njobs, 1), block(WARPS_PER_CTA * 32, 1, 1);
    if (count) encode_i_kernel<ENC_MPW, true><<<grid, block, 0, s>>>(g, d_jobs, d_qt);
    else       encode_i_kernel<ENC_MPW, false><<<grid, block, 0, s>>>(g, d_jobs, d_qt);
    return cudaGetLastError();
}

cudaError_t launch_encode_p(const FrameGeom &g, const EncJob *d_jobs, uint32_t njobs,
                            const QTables *d_qt, const CUtensorMap &tm_luma, const CUtensorMap &tm_chroma,
                            bool count, int variant, uint32_t *d_work, cudaStream_t s)
{
    if (variant == 1) {
        // first generation: one CTA per tile, one warp per macroblock; compiled for 4 resident CTAs per SM (64 registers):
        // measured on 32 x 1080p 3: 77.5 k, 4: 84.8 k, 5: 81.5 k, 6: 83.2 k frames/s
        dim3 grid(g.total_tiles, njobs, 1), block(WARPS_PER_CTA * 32, 1, 1);
        if (count) encode_p_kernel<4, true><<<grid, block, 0, s>>>(g, d_jobs, d_qt, tm_luma, tm_chroma);
        else       encode_p_kernel<4, false><<<grid, block, 0, s>>>(g, d_jobs, d_qt, tm_luma, tm_chroma);
        return cudaGetLastError();
    }
    static bool attr_done[64] = {};           // cudaFuncSetAttribute is per DEVICE: one process may own contexts on several
    const int smem = (int)sizeof(Ep2Smem);
    if (first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(encode_p2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(encode_p2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    const float rcp_tiles = 1.0f / (float)g.total_tiles;
    for (uint32_t j0 = 0; j0 < njobs; j0 += EP2_MAX_JOBS) {
        const uint32_t n = njobs - j0 < (uint32_t)EP2_MAX_JOBS ? njobs - j0 : (uint32_t)EP2_MAX_JOBS;
        const uint32_t items = n * g.total_tiles;
        static const int warps_env = getenv("PFV_EP2_WARPS") ? atoi(getenv("PFV_EP2_WARPS")) : EP2_WARPS;   // tuning experiment
        const uint32_t nw = warps_env >= 1 && warps_env <= EP2_WARPS ? (uint32_t)warps_env : (uint32_t)EP2_WARPS;
        uint32_t ctas = (items + nw - 1) / nw;
        if (ctas > 148u) ctas = 148u;
        if (count) encode_p2_kernel<true><<<ctas, nw * 32, smem, s>>>(g, d_jobs + j0, n, d_qt, rcp_tiles, d_work, tm_luma, tm_chroma);
        else       encode_p2_kernel<false><<<ctas, nw * 32, smem, s>>>(g, d_jobs + j0, n, d_qt, rcp_tiles, d_work, tm_luma, tm_chroma);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace pfv
