// pfv_kernels_p.cu — decode-P as two dense kernels (the fallback pair for P frames; sm_100a).
//
// VideoPlane::decode_plane_delta_into (src/common.rs:498-521) is "for every macroblock: fetch the motion-
// compensated block of the OLD plane (get_block, :327-339); if it carries coefficients, add the decoded residual
// (decode_block_delta :254-285, apply_residuals :98-104); write it".  In real streams most macroblocks are
// skipped, so the work is two very different populations:
//
//   mc_copy4_kernel      every macroblock: predictor -> destination slot out of ONE TMA window per 8 x 4 macroblocks,
//                        fully coalesced 128-byte row stores, 32 registers; coded macroblocks are appended to a
//                        per-(frame, plane) list with one warp-aggregated atomic.
//   residual_sb2_kernel  one thread per coded 8x8 sub-block of those lists: coefficients, register-resident IDCT,
//                        residual added to the predictor that the copy kernel left at the block's own position.
//
// The fused kernel of pfv_kernels_pf.cu (the default) does both in one pass; this pair stays as the simple,
// independently written second implementation (PFV_DECODE_P_VARIANT=win).  The generations before it (per-lane
// global gathers, cp.async tiles, per-macroblock TMA boxes, list-free and concurrent residual passes) were all
// measured slower - their numbers are in profiles/README.md and DESIGN.md - and have been removed.
#include <stdlib.h>

#include "pfv_internal.h"
#include "pfv_device.cuh"
#include "pfv_sb.cuh"

namespace pfv {

__device__ __forceinline__ uint32_t mc3_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

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
// Window copy.  Per CTA ONE TMA box: the 176 x 94 byte window that contains every possible predictor (|mv| <= 15,
// src/common.rs:154-204; validated range src/common.rs:258-259) of a block of 8 x 4 macroblocks, in full 16-byte
// aligned rows (a TMA box cannot start at an unaligned byte, and per-macroblock 32 x 16 boxes cost 1.5 sectors per
// 16-byte row: 300 MB of L2->SM traffic for 100 MB useful).  Warp w of the CTA owns macroblock row w of the block;
// lanes pick their unaligned rows out of the window and write full 128-byte lines.
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

__global__ void __launch_bounds__(MC4_ROWS * 32)
mc_copy4_kernel(const __grid_constant__ FrameGeom g, const __grid_constant__ McWin W, const DecJob *__restrict__ jobs,
                uint32_t njobs, uint32_t *__restrict__ lists, uint32_t *__restrict__ counts, int *__restrict__ err,
                const __grid_constant__ CUtensorMap tm_luma, const __grid_constant__ CUtensorMap tm_chroma)
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
        Item r;
        uint32_t wi;
        wi = it / njobs;
        r.job = it - wi * njobs;
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
        __syncthreads();                                      // the whole CTA is done with this stage
        cur = nxt; hw_cur = hw_next;
        nxt = nn; hw_next = hw_nn;
        vote = vote_next;
        list_base = list_base_next;
    }
}

cudaError_t launch_decode_p_two_pass4(const SbParams &P, const DecJob *d_jobs, uint32_t njobs, uint32_t *d_lists, uint32_t *d_counts,
                                      int *d_err, const CUtensorMap &tm_luma, const CUtensorMap &tm_chroma, cudaStream_t s,
                                      uint32_t *d_done)
{
    static bool attr_done[64] = {};           // cudaFuncSetAttribute is per DEVICE: one process may own contexts on several
    const int mc_smem = (int)sizeof(Mc4Smem);
    if (first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(mc_copy4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mc_smem);
        if (e != cudaSuccess) return e;
    }
    const FrameGeom &g = P.g;
    const McWin W = make_mc_windows(g, MC4_ROWS);
    uint32_t ctas = njobs * W.total;
    if (ctas > 148u * 6u) ctas = 148u * 6u;
    mc_copy4_kernel<<<ctas, MC4_ROWS * 32, mc_smem, s>>>(g, W, d_jobs, njobs, d_lists, d_counts, d_err, tm_luma, tm_chroma);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    // d_done != nullptr: the residual kernel clears d_counts when it is done (the caller then skips its memset)
    uint32_t rctas = P.cta_total * njobs;                     // worst case: every macroblock coded
    if (rctas > 148u * RS2_CTAS_PER_SM) rctas = 148u * RS2_CTAS_PER_SM;
    residual_sb2_kernel<<<rctas, SB_THREADS, (njobs * 3 + 1) * sizeof(uint32_t), s>>>(P, d_jobs, njobs, d_lists, d_counts, d_done);
    return cudaGetLastError();
}

}  // namespace pfv
