// pfv_kernels_p.cu — decode-P as two dense kernels (the default for P frames; sm_100a).
//
// VideoPlane::decode_plane_delta_into (src/common.rs:498-521) is "for every macroblock: fetch the motion-
// compensated block of the OLD plane (get_block, :327-339); if it carries coefficients, add the decoded residual
// (decode_block_delta :254-285, apply_residuals :98-104); write it".  In real streams most macroblocks are
// skipped, so the work is two very different populations.  Fusing them in one kernel (decode_p_stream_kernel,
// decode_sbw_kernel<true>) ties a copy that wants 64 resident warps per SM to a transform that needs 128+
// registers per thread; ncu showed ~7 warps per SM and 20 % issue utilisation.  Here instead:
//
//   mc_copy_kernel      every macroblock: predictor -> destination slot, fully coalesced 128-byte row stores,
//                       ~50 registers, full occupancy; coded macroblocks are appended to a per-(frame, plane)
//                       list with one warp-aggregated atomic.
//   residual_sb_kernel  one thread per coded 8x8 sub-block of those lists (so warps are full): coefficients,
//                       register-resident IDCT (or the DC-only shortcut), residual added to the predictor that
//                       mc_copy_kernel left at the block's own, aligned position.
//
// The second kernel re-reads 256 B per CODED macroblock (mostly from L2); nothing else is touched twice.
#include <stdlib.h>

#include "pfv_internal.h"
#include "pfv_device.cuh"
#include "pfv_sb.cuh"

namespace pfv {

static cudaError_t launch_residual(const SbParams &P, const DecJob *d_jobs, uint32_t njobs, const uint32_t *d_lists,
                                   uint32_t *d_counts, uint32_t *d_done, cudaStream_t s);

constexpr int MC_WARPS = 8;

// lane = (macroblock of the tile) * 4 + row group; a lane moves rows rg, rg+4, rg+8, rg+12 of its macroblock, so
// every store instruction of the warp covers four complete 128-byte lines of the destination plane.
__global__ void __launch_bounds__(MC_WARPS * 32)
mc_copy_kernel(const __grid_constant__ FrameGeom g, const __grid_constant__ McTiles T, const DecJob *__restrict__ jobs,
               uint32_t *__restrict__ lists, uint32_t *__restrict__ counts, int *__restrict__ err)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t tile_g = blockIdx.x * MC_WARPS + (threadIdx.x >> 5);
    if (tile_g >= T.total) return;
    const int p = (tile_g >= T.base[1] ? 1 : 0) + (tile_g >= T.base[2] ? 1 : 0);
    const PlaneGeom &pl = p == 0 ? g.pl[0] : (p == 1 ? g.pl[1] : g.pl[2]);
    const uint32_t tile = tile_g - (p == 0 ? T.base[0] : (p == 1 ? T.base[1] : T.base[2]));
    const uint32_t nmb = pl.bw * pl.bh;
    const uint32_t lm = tile * 8u + (lane >> 2), rg = lane & 3u;
    const bool valid = lm < nmb;
    const DecJob job = jobs[blockIdx.y];

    bool coded = false;
    if (valid) {
        const uint32_t hw = __ldg(reinterpret_cast<const uint32_t *>(job.hdr) + pl.mb_base + lm);   // {mx, my, has_coeff, 0}
        coded = ((hw >> 16) & 0xffu) != 0u;
        uint32_t col;
        const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
        const int bx = (int)col * 16, by = (int)row * 16;
        int sx = bx + (int)(int8_t)(hw & 0xffu), sy = by + (int)(int8_t)((hw >> 8) & 0xffu);       // src/common.rs:255-256
        if (sx < 0 || sy < 0 || sx > (int)pl.pw - 16 || sy > (int)pl.ph - 16) {
            // reference: debug_assert / slice panic (src/common.rs:258-259).  Never read out of bounds: the
            // stream is flagged bad and the co-located block is used.
            if (rg == 0) atomicOr(err, ERRBIT_BAD_MV);
            sx = bx;
            sy = by;
        }
        const uint32_t i0 = ((uint32_t)sx & 15u) >> 2, sh = ((uint32_t)sx & 3u) * 8u;
        const uint8_t *src = job.ref + pl.off + (size_t)((uint32_t)sy + rg) * pl.pw + ((uint32_t)sx & ~15u);
        uint8_t *dst = job.dst + pl.off + (size_t)((uint32_t)by + rg) * pl.pw + (uint32_t)bx;
        uint4 lo[4], hi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // the two aligned 16-byte chunks that cover the row's 16 unaligned bytes (the frame pool has slack
            // behind its last row for the second one)
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src + (size_t)(4 * i) * pl.pw);
            lo[i] = __ldg(s4);
            hi[i] = __ldg(s4 + 1);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t w[8] = {lo[i].x, lo[i].y, lo[i].z, lo[i].w, hi[i].x, hi[i].y, hi[i].z, hi[i].w};
            uint32_t u[6], t[5];
#pragma unroll
            for (int j = 0; j < 6; ++j) u[j] = (i0 & 2u) ? w[j + 2] : w[j];
#pragma unroll
            for (int j = 0; j < 5; ++j) t[j] = (i0 & 1u) ? u[j + 1] : u[j];
            uint4 o;
            o.x = __funnelshift_r(t[0], t[1], sh);
            o.y = __funnelshift_r(t[1], t[2], sh);
            o.z = __funnelshift_r(t[2], t[3], sh);
            o.w = __funnelshift_r(t[3], t[4], sh);
            __stcg(reinterpret_cast<uint4 *>(dst + (size_t)(4 * i) * pl.pw), o);   // blit_block, src/common.rs:341-349
        }
    }
    // coded macroblocks of the tile -> the (frame, plane) list; order inside a list does not matter
    const uint32_t vote = __ballot_sync(0xffffffffu, coded && rg == 0);
    if (vote) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&counts[blockIdx.y * 4u + (uint32_t)p], (uint32_t)__popc(vote));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (coded && rg == 0)
            lists[(size_t)blockIdx.y * g.nb + pl.mb_base + base + (uint32_t)__popc(vote & ((1u << lane) - 1u))] = lm;
    }
}

// One thread = one coded 8x8 sub-block (thread index -> entry of the plane's list).  CTAs past the end of a list
// exit at once; the grid is sized for the worst case (every macroblock coded).
__global__ void __launch_bounds__(SB_THREADS, 3)
residual_sb_kernel(const __grid_constant__ SbParams P, const DecJob *__restrict__ jobs,
                   const uint32_t *__restrict__ lists, const uint32_t *__restrict__ counts)
{
    const uint32_t cta = blockIdx.x;
    const int p = (cta >= P.cta_base[1] ? 1 : 0) + (cta >= P.cta_base[2] ? 1 : 0);
    const PlaneGeom &pl = p == 0 ? P.g.pl[0] : (p == 1 ? P.g.pl[1] : P.g.pl[2]);
    const uint32_t e = (cta - (p == 0 ? P.cta_base[0] : (p == 1 ? P.cta_base[1] : P.cta_base[2]))) * SB_MBS_PER_CTA +
                       (threadIdx.x >> 2);
    if (e >= counts[blockIdx.y * 4u + (uint32_t)p]) return;
    const uint32_t lm = lists[(size_t)blockIdx.y * P.g.nb + pl.mb_base + e];
    const int sb = (int)(threadIdx.x & 3u);
    const DecJob job = jobs[blockIdx.y];
    const int32_t *deq = p == 0 ? P.deq[0] : (p == 1 ? P.deq[1] : P.deq[2]);

    const uint4 *src = reinterpret_cast<const uint4 *>(job.coeff + ((size_t)(pl.mb_base + lm) * 256 + sb * 64));
    uint4 raw[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) raw[k] = __ldcs(src + k);
    uint32_t col;
    const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
    uint8_t *dst = job.dst + pl.off + (size_t)(row * 16u + (uint32_t)(sb >> 1) * 8u) * pl.pw + col * 16u + (uint32_t)(sb & 1) * 8u;
    uint2 prev[8];                                            // the predictor mc_copy_kernel stored here
#pragma unroll
    for (int r = 0; r < 8; ++r) prev[r] = __ldcg(reinterpret_cast<const uint2 *>(dst + (size_t)r * pl.pw));

    uint32_t ac = raw[0].x & 0xffff0000u;
    ac |= raw[0].y | raw[0].z | raw[0].w;
#pragma unroll
    for (int k = 1; k < 8; ++k) ac |= raw[k].x | raw[k].y | raw[k].z | raw[k].w;
    if (ac == 0u) {
        // DC only: both IDCT passes collapse to the DC term (see pfv_sb.cuh), one clamped delta for the sub-block
        const int c0 = (int)(int16_t)(raw[0].x & 0xffffu);
        if (c0 == 0) return;                                  // d = 128, delta = 0: the predictor stands
        const int v = (c0 * deq[0] + (128 << 8)) >> 8;
        const int delta = (min(max(v, 0), 255) - 128) * 2;    // src/common.rs:101
        const uint32_t pos4 = (uint32_t)max(delta, 0) * 0x01010101u;
        const uint32_t neg4 = (uint32_t)min(max(-delta, 0), 255) * 0x01010101u;
#pragma unroll
        for (int r = 0; r < 8; ++r)
            __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pl.pw),
                   make_uint2(add_delta_sat4(prev[r].x, pos4, neg4), add_delta_sat4(prev[r].y, pos4, neg4)));
        return;
    }
    int m[64];
    unpack_dequant(raw, deq, m);
    idct8x8_regs(m);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = m[r * 8 + i];
        __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pl.pw), apply_residual_row(y, prev[r]));   // src/common.rs:277
    }
}

// -------------------------------------------------------------------------------------------------
// v2 of the two kernels (the default).  ncu on v1 (profiles/prof_decode_p_two_s4a.summary.csv):
//   mc_copy_kernel      51 us / 32 frames, DRAM at 39 %, warps 86 % active and stalled on the long scoreboard: the
//                       per-lane row fetches touch 32 different 128-byte lines per load instruction, 272 L1 tag
//                       look-ups per 8-macroblock tile - the L1 t-stage, not HBM, was the limit;
//   residual_sb_kernel  41 us, of which most is launching ~10 700 CTAs that find their list empty while holding the
//                       register file of a 142-register kernel (3 CTAs per SM).
// v2: (a) the predictor tile is staged with warp-wide 16-byte cp.async where lane = (row, half), i.e. one request
// covers the 16 rows of ONE macroblock (16-20 lines instead of 2 x 32), double buffered one tile ahead, and the
// lanes then pick their rows out of shared memory for the same full-line stores as before: ~160 tag look-ups per
// tile; warps are persistent and walk tiles with a stride.  (b) the residual kernel is persistent too: one CTA per
// resident slot, each walks the compacted lists of all (frame, plane) pairs in chunks of 32 macroblocks.
// -------------------------------------------------------------------------------------------------
constexpr int MC2_WARPS = 8;
constexpr int MC2_MB_STRIDE = 16 * 32 + 16;                   // bytes per staged macroblock (+16: macroblocks start 4 banks apart)
constexpr int MC2_BUF = 8 * MC2_MB_STRIDE;                    // one tile

__device__ __forceinline__ void mc2_cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

struct McItem { uint32_t job, p, tile; };

__global__ void __launch_bounds__(MC2_WARPS * 32)
mc_copy2_kernel(const __grid_constant__ FrameGeom g, const __grid_constant__ McTiles T, const DecJob *__restrict__ jobs,
                uint32_t njobs, uint32_t *__restrict__ lists, uint32_t *__restrict__ counts, int *__restrict__ err)
{
    extern __shared__ __align__(16) unsigned char mc2_smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned char *buf0 = mc2_smem + (size_t)warp * 2 * MC2_BUF;
    const uint32_t nwarps = gridDim.x * MC2_WARPS;
    const uint32_t nitems = njobs * T.total;
    uint32_t item = blockIdx.x * MC2_WARPS + warp;
    if (item >= nitems) return;

    auto decode_item = [&](uint32_t it) {
        McItem r;
        r.job = it / T.total;
        const uint32_t tg = it - r.job * T.total;
        r.p = (tg >= T.base[1] ? 1u : 0u) + (tg >= T.base[2] ? 1u : 0u);
        r.tile = tg - (r.p == 0 ? T.base[0] : (r.p == 1 ? T.base[1] : T.base[2]));
        return r;
    };
    auto plane = [&](uint32_t p) -> const PlaneGeom & { return p == 0 ? g.pl[0] : (p == 1 ? g.pl[1] : g.pl[2]); };
    // header word of "this lane's" macroblock (lane >> 2) of an item; 0 when the macroblock is past the plane
    auto load_hw = [&](const McItem &it) -> uint32_t {
        const PlaneGeom &pl = plane(it.p);
        const uint32_t lm = it.tile * 8u + (lane >> 2);
        if (lm >= pl.bw * pl.bh) return 0u;
        return __ldg(reinterpret_cast<const uint32_t *>(jobs[it.job].hdr) + pl.mb_base + lm);
    };
    struct Where { int sx, sy, bx, by; bool bad; };
    auto locate = [&](const PlaneGeom &pl, uint32_t lm, uint32_t hw) {
        Where w;
        uint32_t col;
        const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
        w.bx = (int)col * 16; w.by = (int)row * 16;
        w.sx = w.bx + (int)(int8_t)(hw & 0xffu); w.sy = w.by + (int)(int8_t)((hw >> 8) & 0xffu);   // src/common.rs:255-256
        w.bad = w.sx < 0 || w.sy < 0 || w.sx > (int)pl.pw - 16 || w.sy > (int)pl.ph - 16;
        if (w.bad) { w.sx = w.bx; w.sy = w.by; }             // src/common.rs:258-259: never read out of bounds
        return w;
    };
    // stage the 8 predictors of an item: lane = (row, half) of each macroblock in turn
    auto stage = [&](const McItem &it, uint32_t hw, unsigned char *buf) {
        const PlaneGeom &pl = plane(it.p);
        const uint32_t nmb = pl.bw * pl.bh;
        const uint32_t lm = it.tile * 8u + (lane >> 2);
        const uint8_t *src = jobs[it.job].ref + pl.off;
        if (lm < nmb) {
            const Where w = locate(pl, lm, hw);
            src += (size_t)w.sy * pl.pw + ((uint32_t)w.sx & ~15u);
        }
        const unsigned long long sp = reinterpret_cast<unsigned long long>(src);
        const uint32_t row = lane >> 1, half = lane & 1u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned long long pj = __shfl_sync(0xffffffffu, sp, 4 * j);
            if (it.tile * 8u + (uint32_t)j < nmb)
                mc2_cp_async16(buf + j * MC2_MB_STRIDE + row * 32u + half * 16u,
                               reinterpret_cast<const uint8_t *>(pj) + (size_t)row * pl.pw + half * 16u);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    McItem cur = decode_item(item);
    uint32_t hw_cur = load_hw(cur);
    stage(cur, hw_cur, buf0);
    uint32_t next_item = item + nwarps;
    McItem nxt = cur;
    uint32_t hw_next = 0;
    if (next_item < nitems) { nxt = decode_item(next_item); hw_next = load_hw(nxt); }

    uint32_t parity = 0;
#pragma unroll 1
    for (;;) {
        unsigned char *buf = buf0 + parity * MC2_BUF;
        const bool has_next = next_item < nitems;
        McItem nn = nxt;
        uint32_t hw_nn = 0;
        if (has_next) {
            stage(nxt, hw_next, buf0 + (parity ^ 1u) * MC2_BUF);
            const uint32_t n2 = next_item + nwarps;
            if (n2 < nitems) { nn = decode_item(n2); hw_nn = load_hw(nn); }
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();

        // ---- this tile: rows out of shared memory, full-line stores ----
        const PlaneGeom &pl = plane(cur.p);
        const uint32_t lm = cur.tile * 8u + (lane >> 2), rg = lane & 3u;
        const bool valid = lm < pl.bw * pl.bh;
        bool coded = false;
        if (valid) {
            coded = ((hw_cur >> 16) & 0xffu) != 0u;
            const Where w = locate(pl, lm, hw_cur);
            if (w.bad && rg == 0) atomicOr(err, ERRBIT_BAD_MV);
            const uint32_t i0 = ((uint32_t)w.sx & 15u) >> 2, sh = ((uint32_t)w.sx & 3u) * 8u;
            uint8_t *dst = jobs[cur.job].dst + pl.off + (size_t)((uint32_t)w.by + rg) * pl.pw + (uint32_t)w.bx;
            const unsigned char *mb = buf + (lane >> 2) * MC2_MB_STRIDE + rg * 32u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 lo = *reinterpret_cast<const uint4 *>(mb + i * 128);
                const uint4 hi = *reinterpret_cast<const uint4 *>(mb + i * 128 + 16);
                const uint32_t wd[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                uint32_t u[6], t[5];
#pragma unroll
                for (int j = 0; j < 6; ++j) u[j] = (i0 & 2u) ? wd[j + 2] : wd[j];
#pragma unroll
                for (int j = 0; j < 5; ++j) t[j] = (i0 & 1u) ? u[j + 1] : u[j];
                uint4 o;
                o.x = __funnelshift_r(t[0], t[1], sh);
                o.y = __funnelshift_r(t[1], t[2], sh);
                o.z = __funnelshift_r(t[2], t[3], sh);
                o.w = __funnelshift_r(t[3], t[4], sh);
                __stcg(reinterpret_cast<uint4 *>(dst + (size_t)(4 * i) * pl.pw), o);   // blit_block, src/common.rs:341-349
            }
        }
        const uint32_t vote = __ballot_sync(0xffffffffu, coded && rg == 0);
        if (vote) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&counts[cur.job * 4u + cur.p], (uint32_t)__popc(vote));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (coded && rg == 0)
                lists[(size_t)cur.job * g.nb + pl.mb_base + base + (uint32_t)__popc(vote & ((1u << lane) - 1u))] = lm;
        }
        if (!has_next) break;
        __syncwarp();                                         // every lane is done with buf before it is refilled
        cur = nxt; hw_cur = hw_next;
        next_item += nwarps;
        nxt = nn; hw_next = hw_nn;
        parity ^= 1u;
    }
}

#ifndef RS2_CTAS_PER_SM
#define RS2_CTAS_PER_SM 4    // 3 -> 4 (128 registers, no spills): 25.4 -> 24.3 us per 32 frames
#endif
// Persistent residual kernel: chunk = 32 consecutive entries of one (frame, plane) list; CTA c takes chunks
// c, c + gridDim.x, ...  The chunk table (prefix over njobs * 3 lists) is rebuilt by every CTA from the counts.
__global__ void __launch_bounds__(SB_THREADS, RS2_CTAS_PER_SM)
residual_sb2_kernel(const __grid_constant__ SbParams P, const DecJob *__restrict__ jobs, uint32_t njobs,
                    const uint32_t *__restrict__ lists, uint32_t *__restrict__ counts, uint32_t *__restrict__ done)
{
    extern __shared__ uint32_t pre[];                         // njobs * 3 + 1
    const uint32_t nl = njobs * 3u;
    for (uint32_t i = threadIdx.x; i < nl; i += SB_THREADS) {
        const uint32_t j = i / 3u, p = i - j * 3u;
        pre[i + 1] = (counts[j * 4u + p] + SB_MBS_PER_CTA - 1) / SB_MBS_PER_CTA;
    }
    if (threadIdx.x == 0) pre[0] = 0;
    __syncthreads();
    if (threadIdx.x == 0)
        for (uint32_t i = 0; i < nl; ++i) pre[i + 1] += pre[i];
    __syncthreads();
    const uint32_t total = pre[nl];
    const int sb = (int)(threadIdx.x & 3u);

    // chunk -> (frame, plane, list entry of this thread); ok = the entry exists
    struct Loc { uint32_t j, p, e; bool ok; };
    auto locate = [&](uint32_t chunk) {
        Loc r = {0u, 0u, 0u, false};
        if (chunk >= total) return r;
        uint32_t lo = 0, hi = nl;                             // largest i with pre[i] <= chunk
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (pre[mid] <= chunk) lo = mid; else hi = mid;
        }
        r.j = lo / 3u; r.p = lo - r.j * 3u;
        r.e = (chunk - pre[lo]) * SB_MBS_PER_CTA + (threadIdx.x >> 2);
        r.ok = r.e < counts[r.j * 4u + r.p];
        return r;
    };
    auto list_entry = [&](const Loc &l) -> uint32_t {
        const PlaneGeom &pl = l.p == 0 ? P.g.pl[0] : (l.p == 1 ? P.g.pl[1] : P.g.pl[2]);
        return l.ok ? lists[(size_t)l.j * P.g.nb + pl.mb_base + l.e] : 0u;
    };
    // Software pipeline: the list entry of the NEXT chunk is loaded while this one is transformed, and its
    // coefficient line is prefetched into L2 as soon as the entry is known, so the next iteration's loads find
    // their data on chip (ncu: this kernel spent a third of its time on exposed DRAM latency at 10 warps per SM).
    Loc cur = locate(blockIdx.x);
    uint32_t lm_cur = list_entry(cur);
#pragma unroll 1
    for (uint32_t chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
        const Loc nxt = locate(chunk + gridDim.x);
        const uint32_t lm_next = list_entry(nxt);
        const Loc here = cur;
        const uint32_t lm = lm_cur;
        cur = nxt;
        lm_cur = lm_next;
        if (!here.ok) continue;
        const uint32_t j = here.j, p = here.p;
        const PlaneGeom &pl = p == 0 ? P.g.pl[0] : (p == 1 ? P.g.pl[1] : P.g.pl[2]);
        const DecJob &job = jobs[j];
        const int32_t *deq = p == 0 ? P.deq[0] : (p == 1 ? P.deq[1] : P.deq[2]);

        const uint4 *src = reinterpret_cast<const uint4 *>(job.coeff + ((size_t)(pl.mb_base + lm) * 256 + sb * 64));
        uint4 raw[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) raw[k] = __ldcs(src + k);
        uint32_t col;
        const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
        uint8_t *dst = job.dst + pl.off + (size_t)(row * 16u + (uint32_t)(sb >> 1) * 8u) * pl.pw + col * 16u + (uint32_t)(sb & 1) * 8u;
        uint2 prev[8];                                        // the predictor the copy kernel stored here
#pragma unroll
        for (int r = 0; r < 8; ++r) prev[r] = __ldcg(reinterpret_cast<const uint2 *>(dst + (size_t)r * pl.pw));

        uint32_t ac = raw[0].x & 0xffff0000u;
        ac |= raw[0].y | raw[0].z | raw[0].w;
#pragma unroll
        for (int k = 1; k < 8; ++k) ac |= raw[k].x | raw[k].y | raw[k].z | raw[k].w;
        if (nxt.ok) {
            const PlaneGeom &pn = nxt.p == 0 ? P.g.pl[0] : (nxt.p == 1 ? P.g.pl[1] : P.g.pl[2]);
            const void *nsrc = jobs[nxt.j].coeff + ((size_t)(pn.mb_base + lm_next) * 256 + sb * 64);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nsrc));
        }
        // A sub-block without AC terms would collapse to one clamped delta (see pfv_sb.cuh), but its lane sits in a warp
        // whose other lanes run the full transform anyway (ncu: the 267-instruction shortcut was taken by ~2 lanes in 75 %
        // of the warp passes = 11 % of all issue slots for nothing).  Only a warp with no AC terms at all, and a lane with
        // nothing to add, leave early; the transform of a DC-only block gives the same pixels.
        if (__ballot_sync(__activemask(), ac != 0u) == 0u && (raw[0].x & 0xffffu) == 0u) continue;
        int m[64];
        unpack_dequant(raw, deq, m);
        idct8x8_regs(m);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = m[r * 8 + i];
            __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pl.pw), apply_residual_row(y, prev[r]));   // src/common.rs:277
        }
    }
    // The last CTA to finish clears the list counts for the next batch (done != nullptr): saves the memset node the
    // host would otherwise put in front of every copy kernel.
    if (done) {
        __shared__ uint32_t last;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            last = atomicAdd(done, 1u) == gridDim.x - 1 ? 1u : 0u;
        }
        __syncthreads();
        if (last) {
            for (uint32_t i = threadIdx.x; i < njobs * 4u; i += SB_THREADS) counts[i] = 0u;
            if (threadIdx.x == 0) *done = 0u;
        }
    }
}

// -------------------------------------------------------------------------------------------------
// v3 of the copy kernel: the predictor of every macroblock is ONE TMA box.  ncu on v1/v2 showed the L1 data pipe
// (l1tex__data_pipe_lsu_wavefronts 82 %) as the limit: a motion-compensated 16x16 block is 16 rows in 16 different
// 128-byte lines at an arbitrary byte offset, i.e. 32 (v1) or 17+8 (v2) LSU wavefronts per macroblock just to get it
// on chip.  A tensor-map copy fetches the 32 x 16 byte box that covers the block (the innermost TMA coordinate must
// be 16-byte aligned - tools/exp/tma_box.cu - so the box starts at x & ~15; get_block, src/common.rs:327-339)
// into shared memory without touching the LSU; the lanes then pick their unaligned 16 bytes per row out of shared
// memory (2 LDS.128 + funnel shifts) and write full 128-byte lines: ~80 LSU wavefronts per tile of 8 instead of 272.
// -------------------------------------------------------------------------------------------------
constexpr int MC3_WARPS = 4;
constexpr int MC3_STAGES = 3;
constexpr int MC3_MB = 16 * 32;                               // one staged macroblock: 16 rows of 32 bytes
constexpr int MC3_TILE = 8 * MC3_MB;

struct __align__(128) Mc3Smem {
    unsigned char buf[MC3_WARPS][MC3_STAGES][MC3_TILE];
    uint64_t bar[MC3_WARPS][MC3_STAGES];
};

__device__ __forceinline__ uint32_t mc3_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(MC3_WARPS * 32)
mc_copy3_kernel(const __grid_constant__ FrameGeom g, const __grid_constant__ McTiles T, const DecJob *__restrict__ jobs,
                uint32_t njobs, uint32_t *__restrict__ lists, uint32_t *__restrict__ counts, int *__restrict__ err,
                const __grid_constant__ CUtensorMap tm_luma, const __grid_constant__ CUtensorMap tm_chroma)
{
    extern __shared__ __align__(128) unsigned char mc3_raw[];
    Mc3Smem &sm = *reinterpret_cast<Mc3Smem *>(mc3_raw);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t nwarps = gridDim.x * MC3_WARPS;
    const uint32_t nitems = njobs * T.total;
    const uint32_t first = blockIdx.x * MC3_WARPS + warp;
    if (first >= nitems) return;

    if (lane == 0) {
#pragma unroll
        for (int s2 = 0; s2 < MC3_STAGES; ++s2)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mc3_smem(&sm.bar[warp][s2])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    auto plane = [&](uint32_t p) -> const PlaneGeom & { return p == 0 ? g.pl[0] : (p == 1 ? g.pl[1] : g.pl[2]); };
    struct Item { uint32_t job, p, tile, hw; };
    // decode an item index and fetch the header word of "this lane's" macroblock (lane >> 2)
    auto fetch = [&](uint32_t it) {
        Item r;
        r.job = it / T.total;
        const uint32_t tg = it - r.job * T.total;
        r.p = (tg >= T.base[1] ? 1u : 0u) + (tg >= T.base[2] ? 1u : 0u);
        r.tile = tg - (r.p == 0 ? T.base[0] : (r.p == 1 ? T.base[1] : T.base[2]));
        const PlaneGeom &pl = plane(r.p);
        const uint32_t lm = r.tile * 8u + (lane >> 2);
        r.hw = 0u;
        if (lm < pl.bw * pl.bh) r.hw = __ldg(reinterpret_cast<const uint32_t *>(jobs[r.job].hdr) + pl.mb_base + lm);
        return r;
    };
    struct Where { int sx, sy, bx, by; bool bad; };
    auto locate = [&](const PlaneGeom &pl, uint32_t lm, uint32_t hw) {
        Where w;
        uint32_t col;
        const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
        w.bx = (int)col * 16; w.by = (int)row * 16;
        w.sx = w.bx + (int)(int8_t)(hw & 0xffu); w.sy = w.by + (int)(int8_t)((hw >> 8) & 0xffu);   // src/common.rs:255-256
        w.bad = w.sx < 0 || w.sy < 0 || w.sx > (int)pl.pw - 16 || w.sy > (int)pl.ph - 16;
        if (w.bad) { w.sx = w.bx; w.sy = w.by; }             // src/common.rs:258-259: never read out of bounds
        return w;
    };
    // lane 0 issues one box per macroblock of the tile into stage st (TMA instructions run on the uniform datapath:
    // one elected thread); the source coordinates come from the lanes that hold the macroblocks' header words
    auto issue = [&](const Item &it, uint32_t st) {
        const PlaneGeom &pl = plane(it.p);
        const uint32_t nmb = pl.bw * pl.bh;
        const uint32_t nvalid = min(8u, nmb - it.tile * 8u);
        const uint32_t bar = mc3_smem(&sm.bar[warp][st]);
        uint32_t xy = 0;
        if ((lane >> 2) < nvalid) {
            const Where w = locate(pl, it.tile * 8u + (lane >> 2), it.hw);
            xy = ((uint32_t)w.sy << 16) | ((uint32_t)w.sx & ~15u);
        }
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nvalid * (uint32_t)MC3_MB) : "memory");
        const CUtensorMap *tm = it.p == 0 ? &tm_luma : &tm_chroma;
        const int cz = it.p == 2 ? 1 : 0, cw = jobs[it.job].ref_slot;
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {
            const uint32_t c = __shfl_sync(0xffffffffu, xy, 4 * j);
            if (lane == 0 && j < nvalid) {
                const int cx = (int)(c & 0xffffu), cy = (int)(c >> 16);
                asm volatile(
                    "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
                    "[%0], [%1, {%2, %3, %4, %5}], [%6];"
                    ::"r"(mc3_smem(&sm.buf[warp][st][j * (uint32_t)MC3_MB])), "l"(tm), "r"(cx), "r"(cy), "r"(cz), "r"(cw), "r"(bar)
                    : "memory");
            }
        }
    };

    // prologue: fill the pipeline
    Item ring[MC3_STAGES];
    uint32_t nissued = 0;
#pragma unroll
    for (int s2 = 0; s2 < MC3_STAGES; ++s2) {
        const uint32_t it = first + (uint32_t)s2 * nwarps;
        if (it < nitems) { ring[s2] = fetch(it); issue(ring[s2], (uint32_t)s2); nissued++; }
    }

    uint32_t k = 0;                                           // tiles consumed
#pragma unroll 1
    for (uint32_t it = first; it < nitems; it += nwarps, ++k) {
        const uint32_t st = k % MC3_STAGES;
        // which ring entry is current: entries are rotated explicitly to keep them in registers
        const Item cur = ring[0];
        {   // wait for the stage
            const uint32_t bar = mc3_smem(&sm.bar[warp][st]);
            const uint32_t parity = (k / MC3_STAGES) & 1u;
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            }
        }
        const PlaneGeom &pl = plane(cur.p);
        const uint32_t lm = cur.tile * 8u + (lane >> 2), rg = lane & 3u;
        const bool valid = lm < pl.bw * pl.bh;
        bool coded = false;
        uint4 lo[4], hi[4];
        if (valid) {
            const unsigned char *mb = sm.buf[warp][st] + (lane >> 2) * (uint32_t)MC3_MB + rg * 32u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                lo[i] = *reinterpret_cast<const uint4 *>(mb + i * 128);
                hi[i] = *reinterpret_cast<const uint4 *>(mb + i * 128 + 16);
            }
        }
        __syncwarp();                                         // all lanes have read the stage: it may be refilled
        // prefetch the header of the tile MC3_STAGES ahead and refill this stage with it
        const uint32_t ahead = it + (uint32_t)MC3_STAGES * nwarps;
        Item nx = cur;
        const bool refill = ahead < nitems;
        if (refill) { nx = fetch(ahead); issue(nx, st); }
        if (valid) {
            coded = ((cur.hw >> 16) & 0xffu) != 0u;
            const Where w = locate(pl, lm, cur.hw);
            if (w.bad && rg == 0) atomicOr(err, ERRBIT_BAD_MV);
            uint8_t *dst = jobs[cur.job].dst + pl.off + (size_t)((uint32_t)w.by + rg) * pl.pw + (uint32_t)w.bx;
            const uint32_t i0 = ((uint32_t)w.sx & 15u) >> 2, sh = ((uint32_t)w.sx & 3u) * 8u;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t wd[8] = {lo[i].x, lo[i].y, lo[i].z, lo[i].w, hi[i].x, hi[i].y, hi[i].z, hi[i].w};
                uint32_t u[6], t[5];
#pragma unroll
                for (int j = 0; j < 6; ++j) u[j] = (i0 & 2u) ? wd[j + 2] : wd[j];
#pragma unroll
                for (int j = 0; j < 5; ++j) t[j] = (i0 & 1u) ? u[j + 1] : u[j];
                uint4 o;
                o.x = __funnelshift_r(t[0], t[1], sh);
                o.y = __funnelshift_r(t[1], t[2], sh);
                o.z = __funnelshift_r(t[2], t[3], sh);
                o.w = __funnelshift_r(t[3], t[4], sh);
                __stcg(reinterpret_cast<uint4 *>(dst + (size_t)(4 * i) * pl.pw), o);   // blit_block, src/common.rs:341-349
            }
        }
        const uint32_t vote = __ballot_sync(0xffffffffu, coded && rg == 0);
        if (vote) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&counts[cur.job * 4u + cur.p], (uint32_t)__popc(vote));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (coded && rg == 0)
                lists[(size_t)cur.job * g.nb + pl.mb_base + base + (uint32_t)__popc(vote & ((1u << lane) - 1u))] = lm;
        }
#pragma unroll
        for (int s2 = 0; s2 + 1 < MC3_STAGES; ++s2) ring[s2] = ring[s2 + 1];
        ring[MC3_STAGES - 1] = nx;
    }
}

cudaError_t launch_decode_p_two_pass3(const SbParams &P, const DecJob *d_jobs, uint32_t njobs,
                                      uint32_t *d_lists, uint32_t *d_counts, int *d_err,
                                      const CUtensorMap &tm_luma, const CUtensorMap &tm_chroma, cudaStream_t s)
{
    static bool attr_done[64] = {};           // cudaFuncSetAttribute is per DEVICE: one process may own contexts on several
    const int mc_smem = (int)sizeof(Mc3Smem);
    if (first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(mc_copy3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mc_smem);
        if (e != cudaSuccess) return e;
    }
    const FrameGeom &g = P.g;
    McTiles T;
    uint32_t t = 0;
    for (int p = 0; p < 3; p++) {
        T.base[p] = t;
        t += (g.pl[p].bw * g.pl[p].bh + 7u) / 8u;
    }
    T.total = t;
    {
        static const int cps_env = getenv("PFV_MC3_CTAS_PER_SM") ? atoi(getenv("PFV_MC3_CTAS_PER_SM")) : 0;
        const uint32_t items = njobs * T.total;
        uint32_t ctas = (items + MC3_WARPS - 1) / MC3_WARPS;
        const uint32_t resident = 148u * (uint32_t)(cps_env > 0 ? cps_env : 4);
        if (ctas > resident) ctas = resident;
        mc_copy3_kernel<<<ctas, MC3_WARPS * 32, mc_smem, s>>>(g, T, d_jobs, njobs, d_lists, d_counts, d_err, tm_luma, tm_chroma);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    uint32_t ctas = P.cta_total * njobs;                      // worst case: every macroblock coded
    if (ctas > 148u * 3u) ctas = 148u * 3u;
    residual_sb2_kernel<<<ctas, SB_THREADS, (njobs * 3 + 1) * sizeof(uint32_t), s>>>(P, d_jobs, njobs, d_lists, d_counts, nullptr);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------
// v4 (the default): window staging + list-free residual pass.
// ncu on v3 (profiles/prof_p3c.summary.csv): 58 us; neither L1 (42 %) nor L2 (53 %) saturated, but 25 M warp
// instructions (8 TMA issues per tile from one lane, barrier polling, header latency in the issue path) and
// 300 MB of L2->SM traffic for 100 MB of useful predictor bytes: every 16-byte row segment arrives as a 32-byte
// box row = 1.5 sectors, so even a perfect per-macroblock fetch would sit on the L2 port limit at ~35 us.
// v4 fetches, per CTA, ONE box: the 176 x 94 byte window that contains every possible predictor (|mv| <= 15,
// src/common.rs:154-204; validated range src/common.rs:258-259) of a block of 8 x 4 macroblocks, in full 16-byte
// aligned rows: 517 B of L2 traffic per macroblock instead of 768, one TMA instruction per 32 macroblocks instead
// of 32.  Warp w of the CTA owns macroblock row w of the block; lanes pick their unaligned rows out of the window
// and write full 128-byte lines.  The coded-macroblock lists (and their atomics and memset) are gone: the residual
// kernel compacts the coded macroblocks of a contiguous header range itself (ballot + prefix in shared memory).
// -------------------------------------------------------------------------------------------------
constexpr int MC4_ROWS = 4;                                   // macroblock rows per window = warps per CTA
constexpr int MC4_WIN_W = 176;                                // 16 + 8*16 + 15, rounded up to 16
constexpr int MC4_WIN_H = MC4_ROWS * 16 + 30;
constexpr int MC4_WIN_BYTES = MC4_WIN_W * MC4_WIN_H;          // 16 544
constexpr int MC4_STAGE = (MC4_WIN_BYTES + 127) & ~127;
constexpr int MC4_STAGES = 2;

struct __align__(128) Mc4Smem {
    unsigned char win[MC4_STAGES][MC4_STAGE];
    uint64_t bar[MC4_STAGES];
};

struct McWin {                                                // window items of one frame
    uint32_t base[3];                                         // first item of each plane
    uint32_t tiles_x[3];                                      // windows per row of windows
    uint32_t total;
};

__global__ void __launch_bounds__(MC4_ROWS * 32)
mc_copy4_kernel(const __grid_constant__ FrameGeom g, const __grid_constant__ McWin W, const DecJob *__restrict__ jobs,
                uint32_t njobs, uint32_t *__restrict__ lists, uint32_t *__restrict__ counts, int *__restrict__ err,
                const __grid_constant__ CUtensorMap tm_luma, const __grid_constant__ CUtensorMap tm_chroma,
                uint32_t *__restrict__ jobdone)
{
    extern __shared__ __align__(128) unsigned char mc4_raw[];
    Mc4Smem &sm = *reinterpret_cast<Mc4Smem *>(mc4_raw);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t nitems = njobs * W.total;
    if (blockIdx.x >= nitems) return;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s2 = 0; s2 < MC4_STAGES; ++s2)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mc3_smem(&sm.bar[s2])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    struct Item { uint32_t job, p, gy, tx; };
    auto decode_item = [&](uint32_t it) {
        // frame-interleaved order: CTAs that run at the same time work on DIFFERENT frames, so their list-count atomics
        // hit njobs * 3 different addresses instead of the same handful
        // (live mode, jobdone != nullptr: frame-major instead, so that frames complete one after the other and the
        // residual kernel running next to this one can start on frame 0 while the later frames are still copied)
        Item r;
        uint32_t wi;
        if (jobdone) { r.job = it / W.total; wi = it - r.job * W.total; }
        else { wi = it / njobs; r.job = it - wi * njobs; }
        r.p = (wi >= W.base[1] ? 1u : 0u) + (wi >= W.base[2] ? 1u : 0u);
        const uint32_t li = wi - (r.p == 0 ? W.base[0] : (r.p == 1 ? W.base[1] : W.base[2]));
        const uint32_t txs = r.p == 0 ? W.tiles_x[0] : (r.p == 1 ? W.tiles_x[1] : W.tiles_x[2]);
        r.gy = li / txs;
        r.tx = li - r.gy * txs;
        return r;
    };
    auto plane = [&](uint32_t p) -> const PlaneGeom & { return p == 0 ? g.pl[0] : (p == 1 ? g.pl[1] : g.pl[2]); };
    auto issue = [&](const Item &it, uint32_t st) {          // thread 0 only
        const uint32_t bar = mc3_smem(&sm.bar[st]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)MC4_WIN_BYTES) : "memory");
        const CUtensorMap *tm = it.p == 0 ? &tm_luma : &tm_chroma;
        const int cx = (int)it.tx * 128 - 16, cy = (int)it.gy * (MC4_ROWS * 16) - 15, cz = it.p == 2 ? 1 : 0, cw = jobs[it.job].ref_slot;
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(mc3_smem(sm.win[st])), "l"(tm), "r"(cx), "r"(cy), "r"(cz), "r"(cw), "r"(bar)
            : "memory");
    };
    // header word of this lane's macroblock (row = gy*4 + warp, column = tx*8 + (lane >> 2)); bit 31 set = no such macroblock
    auto load_hw = [&](const Item &it) -> uint32_t {
        const PlaneGeom &pl = plane(it.p);
        const uint32_t row = it.gy * MC4_ROWS + warp, col = it.tx * 8u + (lane >> 2);
        if (row >= pl.bh || col >= pl.bw) return 0x80000000u;
        return __ldg(reinterpret_cast<const uint32_t *>(jobs[it.job].hdr) + pl.mb_base + row * pl.bw + col);
    };

    // coded macroblocks of this warp's tile -> the (frame, plane) list.  The atomic that reserves the list range is issued
    // one window AHEAD of its use, so its round trip hides behind a whole iteration.
    auto reserve = [&](const Item &it, uint32_t hw, uint32_t &vote) -> uint32_t {
        const bool coded = !(hw & 0x80000000u) && ((hw >> 16) & 0xffu) != 0u && (lane & 3u) == 0u;
        vote = __ballot_sync(0xffffffffu, coded);
        uint32_t base = 0;
        if (vote && lane == 0) base = atomicAdd(&counts[it.job * 4u + it.p], (uint32_t)__popc(vote));
        return base;
    };

    // Software pipeline over windows i (being copied), i+1 (TMA in flight, list range being reserved) and i+2 (headers
    // being loaded): no load or atomic is consumed in the iteration that issues it.
    Item cur = decode_item(blockIdx.x);
    if (threadIdx.x == 0) issue(cur, 0);
    uint32_t hw_cur = load_hw(cur);
    uint32_t vote = 0;
    uint32_t list_base = reserve(cur, hw_cur, vote);
    Item nxt = cur;
    uint32_t hw_next = 0x80000000u;
    if (blockIdx.x + gridDim.x < nitems) { nxt = decode_item(blockIdx.x + gridDim.x); hw_next = load_hw(nxt); }
    uint32_t k = 0;
#pragma unroll 1
    for (uint32_t it = blockIdx.x; it < nitems; it += gridDim.x, ++k) {
        const uint32_t st = k & 1u;
        const bool has_next = it + gridDim.x < nitems;
        uint32_t vote_next = 0, list_base_next = 0;
        if (has_next) {
            if (threadIdx.x == 0) issue(nxt, st ^ 1u);       // that stage was released by the barrier at the end of the previous iteration
            list_base_next = reserve(nxt, hw_next, vote_next);
        }
        Item nn = nxt;
        uint32_t hw_nn = 0x80000000u;
        if (it + 2 * gridDim.x < nitems) { nn = decode_item(it + 2 * gridDim.x); hw_nn = load_hw(nn); }
        {
            const uint32_t bar = mc3_smem(&sm.bar[st]);
            const uint32_t parity = (k >> 1) & 1u;
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done) : "r"(bar), "r"(parity) : "memory");
            }
        }
        if (!(hw_cur & 0x80000000u)) {
            const PlaneGeom &pl = plane(cur.p);
            const uint32_t rg = lane & 3u, mb = lane >> 2;
            const int bx = (int)(cur.tx * 8u + mb) * 16, by = (int)(cur.gy * MC4_ROWS + warp) * 16;
            int mvx = (int)(int8_t)(hw_cur & 0xffu), mvy = (int)(int8_t)((hw_cur >> 8) & 0xffu);   // src/common.rs:255-256
            const int sx = bx + mvx, sy = by + mvy;
            if (sx < 0 || sy < 0 || sx > (int)pl.pw - 16 || sy > (int)pl.ph - 16) {
                // reference: debug_assert / slice panic (src/common.rs:258-259).  Never follow it: the stream is
                // flagged bad and the co-located block is used.
                if (rg == 0) atomicOr(err, ERRBIT_BAD_MV);
                mvx = 0; mvy = 0;
            }
            uint8_t *dst = jobs[cur.job].dst + pl.off + (size_t)((uint32_t)by + rg) * pl.pw + (uint32_t)bx;
            if (mvx < -16 || mvx > 15 || mvy < -15 || mvy > 15) {
                // Legal for the reference decoder (7-bit vectors, src/dec.rs:367-368) but outside the staged window -
                // its own encoder never searches further than +-15 (src/common.rs:154-204).  Rare: fetch from global.
                const uint8_t *gsrc = jobs[cur.job].ref + pl.off + (size_t)((uint32_t)(by + mvy) + rg) * pl.pw + (uint32_t)(bx + mvx);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint2 a = ldg_u8x8_unaligned(gsrc + (size_t)(4 * i) * pl.pw);
                    const uint2 b = ldg_u8x8_unaligned(gsrc + (size_t)(4 * i) * pl.pw + 8);
                    __stcg(reinterpret_cast<uint4 *>(dst + (size_t)(4 * i) * pl.pw), make_uint4(a.x, a.y, b.x, b.y));
                }
            } else {
                const uint32_t wx = (uint32_t)(16 + (int)mb * 16 + mvx), wy = (uint32_t)(15 + (int)warp * 16 + mvy) + rg;
                const unsigned char *src = sm.win[st] + wy * MC4_WIN_W + (wx & ~15u);
                const uint32_t i0 = (wx & 15u) >> 2, sh = (wx & 3u) * 8u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 lo = *reinterpret_cast<const uint4 *>(src + i * 4 * MC4_WIN_W);
                    const uint4 hi = *reinterpret_cast<const uint4 *>(src + i * 4 * MC4_WIN_W + 16);
                    const uint32_t wd[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                    uint32_t u[6], t[5];
#pragma unroll
                    for (int j = 0; j < 6; ++j) u[j] = (i0 & 2u) ? wd[j + 2] : wd[j];
#pragma unroll
                    for (int j = 0; j < 5; ++j) t[j] = (i0 & 1u) ? u[j + 1] : u[j];
                    uint4 o;
                    o.x = __funnelshift_r(t[0], t[1], sh);
                    o.y = __funnelshift_r(t[1], t[2], sh);
                    o.z = __funnelshift_r(t[2], t[3], sh);
                    o.w = __funnelshift_r(t[3], t[4], sh);
                    __stcg(reinterpret_cast<uint4 *>(dst + (size_t)(4 * i) * pl.pw), o);   // blit_block, src/common.rs:341-349
                }
            }
        }
        if (vote) {
            list_base = __shfl_sync(0xffffffffu, list_base, 0);
            const bool coded = !(hw_cur & 0x80000000u) && ((hw_cur >> 16) & 0xffu) != 0u && (lane & 3u) == 0u;
            if (coded) {
                const PlaneGeom &pl = plane(cur.p);
                const uint32_t lm = (cur.gy * MC4_ROWS + warp) * pl.bw + cur.tx * 8u + (lane >> 2);
                lists[(size_t)cur.job * g.nb + pl.mb_base + list_base + (uint32_t)__popc(vote & ((1u << lane) - 1u))] = lm;
            }
        }
        if (jobdone) __threadfence();                         // this thread's stores of the window are visible device-wide ...
        __syncthreads();                                      // the whole CTA is done with this stage
        if (jobdone && threadIdx.x == 0) atomicAdd(&jobdone[cur.job], 1u);   // ... before the window is counted as done
        cur = nxt; hw_cur = hw_next;
        nxt = nn; hw_next = hw_nn;
        vote = vote_next;
        list_base = list_base_next;
    }
}

// Residual pass without lists: a CTA walks a contiguous range of 256-macroblock chunks of the frames' header
// arrays, compacts the coded macroblocks of a chunk into shared memory (ballot + prefix) and runs full warps of
// sub-blocks over them; what is left over (< 32 macroblocks) is carried into the next chunk of the same plane.
constexpr uint32_t RS3_CHUNK = 256;
__global__ void __launch_bounds__(SB_THREADS, 3)
residual_sb3_kernel(const __grid_constant__ SbParams P, const DecJob *__restrict__ jobs, uint32_t njobs, uint32_t chunks_per_cta)
{
    __shared__ uint32_t list[RS3_CHUNK + SB_MBS_PER_CTA];
    __shared__ uint32_t wcount[SB_THREADS / 32];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int sb = (int)(threadIdx.x & 3u);
    uint32_t cpp[3], cpf = 0;                                 // chunks per plane / per frame
#pragma unroll
    for (int p = 0; p < 3; ++p) { cpp[p] = (P.g.pl[p].bw * P.g.pl[p].bh + RS3_CHUNK - 1) / RS3_CHUNK; cpf += cpp[p]; }
    const uint32_t total = cpf * njobs;
    const uint32_t c_begin = blockIdx.x * chunks_per_cta, c_end = min(c_begin + chunks_per_cta, total);

    uint32_t have = 0;                                        // entries in list[] (uniform across the CTA)
    uint32_t cur_j = 0xffffffffu, cur_p = 0;

    auto run = [&](uint32_t j, uint32_t p, uint32_t first, uint32_t count) {
        // process list[first .. first+count) of (job j, plane p): thread -> (entry, sub-block)
        const PlaneGeom &pl = p == 0 ? P.g.pl[0] : (p == 1 ? P.g.pl[1] : P.g.pl[2]);
        const int32_t *deq = p == 0 ? P.deq[0] : (p == 1 ? P.deq[1] : P.deq[2]);
        const DecJob &job = jobs[j];
#pragma unroll 1
        for (uint32_t e = first + (threadIdx.x >> 2); e < first + count; e += SB_MBS_PER_CTA) {
            const uint32_t lm = list[e];
            const uint4 *src = reinterpret_cast<const uint4 *>(job.coeff + ((size_t)(pl.mb_base + lm) * 256 + sb * 64));
            uint4 raw[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) raw[k] = __ldcs(src + k);
            uint32_t col;
            const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
            uint8_t *dst = job.dst + pl.off + (size_t)(row * 16u + (uint32_t)(sb >> 1) * 8u) * pl.pw + col * 16u + (uint32_t)(sb & 1) * 8u;
            uint2 prev[8];                                    // the predictor mc_copy4_kernel stored here
#pragma unroll
            for (int r = 0; r < 8; ++r) prev[r] = __ldcg(reinterpret_cast<const uint2 *>(dst + (size_t)r * pl.pw));
            uint32_t ac = raw[0].x & 0xffff0000u;
            ac |= raw[0].y | raw[0].z | raw[0].w;
#pragma unroll
            for (int k = 1; k < 8; ++k) ac |= raw[k].x | raw[k].y | raw[k].z | raw[k].w;
            if (ac == 0u) {
                // DC only: both IDCT passes collapse to the DC term (see pfv_sb.cuh), one clamped delta for the sub-block
                const int c0 = (int)(int16_t)(raw[0].x & 0xffffu);
                if (c0 == 0) continue;                        // d = 128, delta = 0: the predictor stands
                const int v = (c0 * deq[0] + (128 << 8)) >> 8;
                const int delta = (min(max(v, 0), 255) - 128) * 2;    // src/common.rs:101
                const uint32_t pos4 = (uint32_t)max(delta, 0) * 0x01010101u;
                const uint32_t neg4 = (uint32_t)min(max(-delta, 0), 255) * 0x01010101u;
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pl.pw),
                           make_uint2(add_delta_sat4(prev[r].x, pos4, neg4), add_delta_sat4(prev[r].y, pos4, neg4)));
                continue;
            }
            int m[64];
            unpack_dequant(raw, deq, m);
            idct8x8_regs(m);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                int y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = m[r * 8 + i];
                __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pl.pw), apply_residual_row(y, prev[r]));   // src/common.rs:277
            }
        }
    };

#pragma unroll 1
    for (uint32_t c = c_begin; c < c_end; ++c) {
        const uint32_t j = c / cpf;
        uint32_t ci = c - j * cpf;
        const uint32_t p = (ci >= cpp[0] ? 1u : 0u) + (ci >= cpp[0] + cpp[1] ? 1u : 0u);
        ci -= p == 0 ? 0u : (p == 1 ? cpp[0] : cpp[0] + cpp[1]);
        if (j != cur_j || p != cur_p) {                       // plane or frame changes: flush the carry
            if (have) run(cur_j, cur_p, 0, have);
            __syncthreads();
            have = 0; cur_j = j; cur_p = p;
        }
        const PlaneGeom &pl = p == 0 ? P.g.pl[0] : (p == 1 ? P.g.pl[1] : P.g.pl[2]);
        const uint32_t nmb = pl.bw * pl.bh;
        const uint32_t *hdr32 = reinterpret_cast<const uint32_t *>(jobs[j].hdr) + pl.mb_base;
        // two macroblocks per thread: ci*256 + tid and + 128
        uint32_t base_cnt = have;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint32_t lm = ci * RS3_CHUNK + (uint32_t)half * SB_THREADS + threadIdx.x;
            const bool coded = lm < nmb && ((__ldg(hdr32 + min(lm, nmb - 1)) >> 16) & 0xffu) != 0u;
            const uint32_t vote = __ballot_sync(0xffffffffu, coded);
            if (lane == 0) wcount[warp] = (uint32_t)__popc(vote);
            __syncthreads();
            uint32_t off = base_cnt, tot = 0;
#pragma unroll
            for (uint32_t w = 0; w < SB_THREADS / 32; ++w) { const uint32_t n = wcount[w]; if (w < warp) off += n; tot += n; }
            if (coded) list[off + (uint32_t)__popc(vote & ((1u << lane) - 1u))] = lm;
            base_cnt += tot;
            __syncthreads();
        }
        have = base_cnt;
        const uint32_t full = have & ~(SB_MBS_PER_CTA - 1);   // whole passes of 32 macroblocks
        if (full) {
            run(j, p, 0, full);
            __syncthreads();
            const uint32_t rest = have - full;                // move the remainder to the front
            uint32_t v = 0;
            if (threadIdx.x < rest) v = list[full + threadIdx.x];
            __syncthreads();
            if (threadIdx.x < rest) list[threadIdx.x] = v;
            __syncthreads();
            have = rest;
        }
    }
    if (have) run(cur_j, cur_p, 0, have);
}

// listless = true pairs the window copy with residual_sb3_kernel (no lists, no memset; kept for comparison: its
// contiguous chunk ranges balance badly when the coded macroblocks cluster), false with the list-driven persistent
// residual_sb2_kernel (the default).
cudaError_t launch_decode_p_two_pass4(const SbParams &P, const DecJob *d_jobs, uint32_t njobs, uint32_t *d_lists, uint32_t *d_counts,
                                      bool listless, int *d_err, const CUtensorMap &tm_luma, const CUtensorMap &tm_chroma, cudaStream_t s,
                                      cudaEvent_t after_copy, uint32_t *d_done)
{
    static bool attr_done[64] = {};           // cudaFuncSetAttribute is per DEVICE: one process may own contexts on several
    const int mc_smem = (int)sizeof(Mc4Smem);
    if (first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(mc_copy4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mc_smem);
        if (e != cudaSuccess) return e;
    }
    const FrameGeom &g = P.g;
    McWin W;
    uint32_t t = 0;
    for (int p = 0; p < 3; p++) {
        W.base[p] = t;
        W.tiles_x[p] = (g.pl[p].bw + 7u) / 8u;
        t += W.tiles_x[p] * ((g.pl[p].bh + MC4_ROWS - 1) / MC4_ROWS);
    }
    W.total = t;
    {
        static const int cps_env = getenv("PFV_MC4_CTAS_PER_SM") ? atoi(getenv("PFV_MC4_CTAS_PER_SM")) : 0;
        uint32_t ctas = njobs * W.total;
        const uint32_t resident = 148u * (uint32_t)(cps_env > 0 ? cps_env : 6);
        if (ctas > resident) ctas = resident;
        mc_copy4_kernel<<<ctas, MC4_ROWS * 32, mc_smem, s>>>(g, W, d_jobs, njobs, d_lists, d_counts, d_err, tm_luma, tm_chroma, nullptr);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (after_copy && (e = cudaEventRecord(after_copy, s)) != cudaSuccess) return e;
    }
    if (!listless) return launch_residual(P, d_jobs, njobs, d_lists, d_counts, d_done, s);
    uint32_t cpf = 0;
    for (int p = 0; p < 3; p++) cpf += (g.pl[p].bw * g.pl[p].bh + RS3_CHUNK - 1) / RS3_CHUNK;
    const uint32_t total = cpf * njobs;
    const uint32_t resident = 148u * 3u;
    const uint32_t per_cta = (total + resident - 1) / resident;
    const uint32_t ctas = (total + per_cta - 1) / per_cta;
    residual_sb3_kernel<<<ctas, SB_THREADS, 0, s>>>(P, d_jobs, njobs, per_cta);
    return cudaGetLastError();
}

// -------------------------------------------------------------------------------------------------
// residual pass with sub-block compaction (PFV_RESIDUAL_VARIANT=4; NOT the default: measured 27.4 us against 22.4 us
// for residual_sb2_kernel per 32 frames - the shared-memory round trip and the three CTA barriers per step cost more
// than the idle lanes they remove).  ncu's source view of residual_sb2_kernel: the transform body
// ran with 21 of 32 lanes on average - a third of the sub-blocks of coded macroblocks carry no coefficient at all and
// their lanes just idle through ~1 700 instructions.  Here a CTA takes 64 listed macroblocks per step, every thread
// looks at two sub-blocks, the ones with any coefficient are appended to a shared-memory ring (warp-aggregated
// shared-memory atomic, 128-byte entries, chunk order swizzled like the decode-I ring), and only as many warps as
// there are entries run the transform.  Empty sub-blocks need nothing: the copy kernel already stored the predictor.
// -------------------------------------------------------------------------------------------------
constexpr uint32_t RS4_MBS = 64;                              // listed macroblocks per CTA step (256 sub-blocks)
struct __align__(16) Rs4Smem {
    uint4    coef[RS4_MBS * 4 * 8];                           // entry s keeps 16-byte chunk k at [s*8 + (k ^ (s & 7))]
    uint32_t id[RS4_MBS * 4];                                 // (macroblock inside the plane << 2) | sub-block
    uint32_t n;
};

__global__ void __launch_bounds__(SB_THREADS, 4)
residual_sb4_kernel(const __grid_constant__ SbParams P, const DecJob *__restrict__ jobs, uint32_t njobs,
                    const uint32_t *__restrict__ lists, const uint32_t *__restrict__ counts)
{
    extern __shared__ __align__(16) unsigned char rs4_raw[];
    Rs4Smem &sm = *reinterpret_cast<Rs4Smem *>(rs4_raw);
    uint32_t *pre = reinterpret_cast<uint32_t *>(rs4_raw + sizeof(Rs4Smem));   // njobs * 3 + 1
    const uint32_t nl = njobs * 3u;
    for (uint32_t i = threadIdx.x; i < nl; i += SB_THREADS) {
        const uint32_t j = i / 3u, p = i - j * 3u;
        pre[i + 1] = (counts[j * 4u + p] + RS4_MBS - 1) / RS4_MBS;
    }
    if (threadIdx.x == 0) { pre[0] = 0; sm.n = 0; }
    __syncthreads();
    if (threadIdx.x == 0)
        for (uint32_t i = 0; i < nl; ++i) pre[i + 1] += pre[i];
    __syncthreads();
    const uint32_t total = pre[nl];
    const uint32_t lane = threadIdx.x & 31u;
    const int sb = (int)(threadIdx.x & 3u);

#pragma unroll 1
    for (uint32_t chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
        uint32_t lo = 0, hi = nl;                             // largest i with pre[i] <= chunk
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (pre[mid] <= chunk) lo = mid; else hi = mid;
        }
        const uint32_t j = lo / 3u, p = lo - j * 3u;
        const PlaneGeom &pl = p == 0 ? P.g.pl[0] : (p == 1 ? P.g.pl[1] : P.g.pl[2]);
        const DecJob &job = jobs[j];
        const int32_t *deq = p == 0 ? P.deq[0] : (p == 1 ? P.deq[1] : P.deq[2]);
        const uint32_t cnt = counts[j * 4u + p];
        const uint32_t e0 = (chunk - pre[lo]) * RS4_MBS;

        // ---- A: look at two sub-blocks per thread, queue the ones that carry coefficients ----
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint32_t e = e0 + (uint32_t)half * (SB_THREADS / 4) + (threadIdx.x >> 2);
            bool need = false;
            uint4 raw[8];
            uint32_t lm = 0;
            if (e < cnt) {
                lm = lists[(size_t)j * P.g.nb + pl.mb_base + e];
                const uint4 *src = reinterpret_cast<const uint4 *>(job.coeff + ((size_t)(pl.mb_base + lm) * 256 + sb * 64));
#pragma unroll
                for (int k = 0; k < 8; ++k) raw[k] = __ldcs(src + k);
                uint32_t any = 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) any |= raw[k].x | raw[k].y | raw[k].z | raw[k].w;
                need = any != 0u;
            }
            const uint32_t vote = __ballot_sync(0xffffffffu, need);
            uint32_t base = 0;
            if (vote && lane == 0) base = atomicAdd(&sm.n, (uint32_t)__popc(vote));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (need) {
                const uint32_t slot = base + (uint32_t)__popc(vote & ((1u << lane) - 1u));
#pragma unroll
                for (int k = 0; k < 8; ++k) sm.coef[slot * 8u + ((uint32_t)k ^ (slot & 7u))] = raw[k];
                sm.id[slot] = (lm << 2) | (uint32_t)sb;
            }
        }
        __syncthreads();
        const uint32_t n = sm.n;
        __syncthreads();
        if (threadIdx.x == 0) sm.n = 0;                       // next step's appends come after the barrier below

        // ---- B: full warps over the queued sub-blocks ----
#pragma unroll 1
        for (uint32_t slot = threadIdx.x; slot < n; slot += SB_THREADS) {
            uint4 r2[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) r2[k] = sm.coef[slot * 8u + ((uint32_t)k ^ (slot & 7u))];
            const uint32_t id = sm.id[slot];
            const uint32_t lm = id >> 2, s2 = id & 3u;
            uint32_t col;
            const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
            uint8_t *dst = job.dst + pl.off + (size_t)(row * 16u + (s2 >> 1) * 8u) * pl.pw + col * 16u + (s2 & 1u) * 8u;
            uint2 prev[8];                                    // the predictor the copy kernel stored here
#pragma unroll
            for (int r = 0; r < 8; ++r) prev[r] = __ldcg(reinterpret_cast<const uint2 *>(dst + (size_t)r * pl.pw));
            int m[64];
            unpack_dequant(r2, deq, m);
            idct8x8_regs(m);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                int y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = m[r * 8 + i];
                __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pl.pw), apply_residual_row(y, prev[r]));   // src/common.rs:277
            }
        }
        __syncthreads();                                      // the ring may be refilled
    }
}

// d_done != nullptr: the kernel clears d_counts when it is done (the caller then skips its memset)
static cudaError_t launch_residual(const SbParams &P, const DecJob *d_jobs, uint32_t njobs, const uint32_t *d_lists,
                                   uint32_t *d_counts, uint32_t *d_done, cudaStream_t s)
{
    static const int v_env = getenv("PFV_RESIDUAL_VARIANT") ? atoi(getenv("PFV_RESIDUAL_VARIANT")) : 2;
    if (v_env != 4) {
        uint32_t ctas = P.cta_total * njobs;                  // worst case: every macroblock coded
        if (ctas > 148u * RS2_CTAS_PER_SM) ctas = 148u * RS2_CTAS_PER_SM;
        residual_sb2_kernel<<<ctas, SB_THREADS, (njobs * 3 + 1) * sizeof(uint32_t), s>>>(P, d_jobs, njobs, d_lists, d_counts, d_done);
        return cudaGetLastError();
    }
    static bool attr_done[64] = {};           // cudaFuncSetAttribute is per DEVICE: one process may own contexts on several
    const size_t smem = sizeof(Rs4Smem) + (njobs * 3 + 1) * sizeof(uint32_t);
    if (first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(residual_sb4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
    }
    uint32_t ctas = (P.cta_total * njobs + 1) / 2;
    if (ctas > 148u * 4u) ctas = 148u * 4u;
    residual_sb4_kernel<<<ctas, SB_THREADS, smem, s>>>(P, d_jobs, njobs, d_lists, d_counts);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && d_done) e = cudaMemsetAsync(d_counts, 0, (size_t)njobs * 4 * sizeof(uint32_t), s);   // this variant does not clear them itself
    return e;
}

// -------------------------------------------------------------------------------------------------
// "live" decode-P: the copy kernel and the residual kernel run AT THE SAME TIME on two streams and share the SMs
// (copy: 32 registers, memory bound; residual: 128 registers, issue bound).  The copy kernel walks the frames one
// after the other and counts finished windows per frame (jobdone); a residual CTA waits until a frame is complete,
// then takes 32-macroblock chunks of its lists from a per-frame ticket counter until there are none left, and moves
// on to the next frame.  Grids are sized so that both kernels are resident together (4 + 2 CTAs per SM); the wait is
// bounded: a frame that does not complete within ~0.2 s flags ERRBIT_TIMEOUT instead of hanging the GPU.
// The last residual CTA clears every counter for the next batch.
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(SB_THREADS, 2)
residual_live_kernel(const __grid_constant__ SbParams P, const DecJob *__restrict__ jobs, uint32_t njobs,
                     const uint32_t *lists, uint32_t *counts, uint32_t *jobdone, uint32_t *rtake, uint32_t *done,
                     uint32_t windows_per_job, int *__restrict__ err)
{
    __shared__ uint32_t s_chunk, s_ok;
    const int sb = (int)(threadIdx.x & 3u);
#pragma unroll 1
    for (uint32_t j = 0; j < njobs; ++j) {
        if (threadIdx.x == 0) {
            uint32_t spins = 0, ok = 1;
            while (ld_acquire_u32(&jobdone[j]) < windows_per_job) {
                __nanosleep(100);
                if (++spins > 2000000u) { atomicOr(err, ERRBIT_TIMEOUT); ok = 0; break; }
            }
            s_ok = ok;
        }
        __syncthreads();
        if (!s_ok) break;
        const uint32_t c0 = __ldcg(&counts[j * 4u + 0]), c1 = __ldcg(&counts[j * 4u + 1]), c2 = __ldcg(&counts[j * 4u + 2]);
        const uint32_t n0 = (c0 + SB_MBS_PER_CTA - 1) / SB_MBS_PER_CTA, n1 = (c1 + SB_MBS_PER_CTA - 1) / SB_MBS_PER_CTA,
                       n2 = (c2 + SB_MBS_PER_CTA - 1) / SB_MBS_PER_CTA;
        const uint32_t total = n0 + n1 + n2;
        const DecJob &job = jobs[j];
#pragma unroll 1
        for (;;) {
            __syncthreads();                                  // everyone has read the previous ticket
            if (threadIdx.x == 0) s_chunk = atomicAdd(&rtake[j], 1u);
            __syncthreads();
            const uint32_t chunk = s_chunk;
            if (chunk >= total) break;
            const uint32_t p = (chunk >= n0 ? 1u : 0u) + (chunk >= n0 + n1 ? 1u : 0u);
            const uint32_t cbase = p == 0 ? 0u : (p == 1 ? n0 : n0 + n1), cnt = p == 0 ? c0 : (p == 1 ? c1 : c2);
            const PlaneGeom &pl = p == 0 ? P.g.pl[0] : (p == 1 ? P.g.pl[1] : P.g.pl[2]);
            const int32_t *deq = p == 0 ? P.deq[0] : (p == 1 ? P.deq[1] : P.deq[2]);
            const uint32_t e = (chunk - cbase) * SB_MBS_PER_CTA + (threadIdx.x >> 2);
            if (e >= cnt) continue;
            const uint32_t lm = __ldcg(&lists[(size_t)j * P.g.nb + pl.mb_base + e]);
            const uint4 *src = reinterpret_cast<const uint4 *>(job.coeff + ((size_t)(pl.mb_base + lm) * 256 + sb * 64));
            uint4 raw[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) raw[k] = __ldcs(src + k);
            uint32_t col;
            const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
            uint8_t *dst = job.dst + pl.off + (size_t)(row * 16u + (uint32_t)(sb >> 1) * 8u) * pl.pw + col * 16u + (uint32_t)(sb & 1) * 8u;
            uint2 prev[8];                                    // the predictor the copy kernel stored here
#pragma unroll
            for (int r = 0; r < 8; ++r) prev[r] = __ldcg(reinterpret_cast<const uint2 *>(dst + (size_t)r * pl.pw));
            uint32_t any = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) any |= raw[k].x | raw[k].y | raw[k].z | raw[k].w;
            if (__ballot_sync(__activemask(), any != 0u) == 0u) continue;   // nothing to add in this whole warp
            int m[64];
            unpack_dequant(raw, deq, m);
            idct8x8_regs(m);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                int y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] = m[r * 8 + i];
                __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pl.pw), apply_residual_row(y, prev[r]));   // src/common.rs:277
            }
        }
    }
    // the last CTA to leave clears the counters of this batch
    __shared__ uint32_t last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(done, 1u) == gridDim.x - 1 ? 1u : 0u;
    }
    __syncthreads();
    if (last) {
        for (uint32_t i = threadIdx.x; i < njobs * 4u; i += SB_THREADS) counts[i] = 0u;
        for (uint32_t i = threadIdx.x; i < njobs; i += SB_THREADS) { jobdone[i] = 0u; rtake[i] = 0u; }
        if (threadIdx.x == 0) *done = 0u;
    }
}

// ctl: [done (4 words) | jobdone (max_jobs) | rtake (max_jobs)], zero on entry and on exit
cudaError_t launch_decode_p_live(const SbParams &P, const DecJob *d_jobs, uint32_t njobs, uint32_t *d_lists, uint32_t *d_counts,
                                 uint32_t *d_ctl, uint32_t max_jobs, int *d_err, const CUtensorMap &tm_luma, const CUtensorMap &tm_chroma,
                                 cudaStream_t s_copy, cudaStream_t s_resid)
{
    static bool attr_done[64] = {};
    const int mc_smem = (int)sizeof(Mc4Smem);
    if (first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(mc_copy4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mc_smem);
        if (e != cudaSuccess) return e;
    }
    const FrameGeom &g = P.g;
    McWin W;
    uint32_t t = 0;
    for (int p = 0; p < 3; p++) {
        W.base[p] = t;
        W.tiles_x[p] = (g.pl[p].bw + 7u) / 8u;
        t += W.tiles_x[p] * ((g.pl[p].bh + MC4_ROWS - 1) / MC4_ROWS);
    }
    W.total = t;
    static const int cc_env = getenv("PFV_LIVE_COPY_CTAS") ? atoi(getenv("PFV_LIVE_COPY_CTAS")) : 0;
    static const int rc_env = getenv("PFV_LIVE_RESID_CTAS") ? atoi(getenv("PFV_LIVE_RESID_CTAS")) : 0;
    uint32_t *jobdone = d_ctl + 4, *rtake = d_ctl + 4 + max_jobs;
    uint32_t ctas = njobs * W.total;
    const uint32_t copy_res = 148u * (uint32_t)(cc_env > 0 ? cc_env : 4);
    if (ctas > copy_res) ctas = copy_res;
    mc_copy4_kernel<<<ctas, MC4_ROWS * 32, mc_smem, s_copy>>>(g, W, d_jobs, njobs, d_lists, d_counts, d_err, tm_luma, tm_chroma, jobdone);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const uint32_t rctas = 148u * (uint32_t)(rc_env > 0 ? rc_env : 2);
    residual_live_kernel<<<rctas, SB_THREADS, 0, s_resid>>>(P, d_jobs, njobs, d_lists, d_counts, jobdone, rtake, d_ctl, W.total, d_err);
    return cudaGetLastError();
}

cudaError_t launch_decode_p_two_pass2(const SbParams &P, const DecJob *d_jobs, uint32_t njobs,
                                      uint32_t *d_lists, uint32_t *d_counts, int *d_err, cudaStream_t s)
{
    static bool attr_done[64] = {};           // cudaFuncSetAttribute is per DEVICE: one process may own contexts on several
    const int mc_smem = MC2_WARPS * 2 * MC2_BUF;
    if (first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(mc_copy2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mc_smem);
        if (e != cudaSuccess) return e;
    }
    const FrameGeom &g = P.g;
    McTiles T;
    uint32_t t = 0;
    for (int p = 0; p < 3; p++) {
        T.base[p] = t;
        t += (g.pl[p].bw * g.pl[p].bh + 7u) / 8u;
    }
    T.total = t;
    {
        const uint32_t items = njobs * T.total;
        uint32_t ctas = (items + MC2_WARPS - 1) / MC2_WARPS;
        const uint32_t resident = 148u * 3u;                  // 3 CTAs of 66 KB shared memory per SM
        if (ctas > resident) ctas = resident;
        mc_copy2_kernel<<<ctas, MC2_WARPS * 32, mc_smem, s>>>(g, T, d_jobs, njobs, d_lists, d_counts, d_err);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    uint32_t ctas = P.cta_total * njobs;                      // worst case: every macroblock coded
    if (ctas > 148u * 3u) ctas = 148u * 3u;
    residual_sb2_kernel<<<ctas, SB_THREADS, (njobs * 3 + 1) * sizeof(uint32_t), s>>>(P, d_jobs, njobs, d_lists, d_counts, nullptr);
    return cudaGetLastError();
}

// d_lists: njobs * nb entries, d_counts: njobs * 4 entries, zeroed by the caller on the same stream.
// P.cta_base / cta_total: CTAs of 32 macroblocks per plane (worst case: every macroblock coded).
cudaError_t launch_decode_p_two_pass(const SbParams &P, const DecJob *d_jobs, uint32_t njobs,
                                     uint32_t *d_lists, uint32_t *d_counts, int *d_err, cudaStream_t s)
{
    {
        const FrameGeom &g = P.g;
        McTiles T;
        uint32_t t = 0;
        for (int p = 0; p < 3; p++) {
            T.base[p] = t;
            t += (g.pl[p].bw * g.pl[p].bh + 7u) / 8u;
        }
        T.total = t;
        dim3 grid((T.total + MC_WARPS - 1) / MC_WARPS, njobs, 1), block(MC_WARPS * 32, 1, 1);
        mc_copy_kernel<<<grid, block, 0, s>>>(g, T, d_jobs, d_lists, d_counts, d_err);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    dim3 grid(P.cta_total, njobs, 1), block(SB_THREADS, 1, 1);
    residual_sb_kernel<<<grid, block, 0, s>>>(P, d_jobs, d_lists, d_counts);
    return cudaGetLastError();
}

}  // namespace pfv
