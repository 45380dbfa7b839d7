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
#include "pfv_internal.h"
#include "pfv_device.cuh"
#include "pfv_sb.cuh"

namespace pfv {

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
