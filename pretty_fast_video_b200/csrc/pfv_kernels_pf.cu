// pfv_kernels_pf.cu — decode-P in ONE warp-specialised kernel (the default for P frames; sm_100a).
//
// VideoPlane::decode_plane_delta_into (src/common.rs:498-521): every macroblock fetches its motion-compensated
// predictor from the OLD plane (get_block, :327-339); a coded one adds the decoded residual (decode_block_delta
// :254-285, apply_residuals :98-104).  The two populations want opposite things: the copy is memory bound and needs
// few registers, the residual transform is issue bound (exact integer IDCT, ~1 300 instructions per 8x8) and needs
// 128.  As two kernels (pfv_kernels_p.cu) they run back to back, 40 us + 25 us per 32 x 1080p, and the coded blocks
// cross L2 twice.  Here a persistent CTA (2 per SM) has two halves that never meet at a CTA barrier:
//
//   warps 0-3  COPY.  Per item ONE TMA box: the 176 x 94 window holding every possible predictor (|mv| <= 15,
//              src/common.rs:154-204) of 8 x 4 macroblocks, three windows in flight.  Warp w owns macroblock row w;
//              lane = macroblock*4 + row group picks its unaligned rows out of the window.  Skipped macroblocks
//              (src/common.rs:281-283) go straight to the destination as full 128-byte lines.  A CODED macroblock
//              is handed over instead: its predictor is written into a slot of a shared-memory ring and its 512 B of
//              coefficients are fetched into the same slot by one bulk copy (cp.async.bulk), both counted on the slot
//              group's mbarrier.
//   warps 4-7  TRANSFORM.  Warp x takes ring groups x, x+4, ... : 8 macroblocks = 32 sub-blocks = one full warp,
//              whatever windows, planes or frames they came from.  One thread per 8x8 sub-block: dequantise,
//              register-resident IDCT, residual on the predictor out of the ring, 8 row stores; then the group is
//              released to the copy half (second mbarrier).
//
// A transform lane is (sub-block s = lane >> 3, slot j = lane & 7).  The ring keeps the coefficients of slot j at
// coef[j * 528] exactly as the bulk copy delivered them (scan order, no re-arranging pass) and the predictor of
// (s, j) at pred[lane * 80]: with those pitches the 128-bit reads of 8 consecutive lanes fall into 8 different bank
// groups.  The dequantiser tables are read from shared memory (a group may mix planes, so they cannot be
// constant-bank operands).
#include <stdlib.h>

#include "pfv_internal.h"
#include "pfv_device.cuh"
#include "pfv_sb.cuh"

namespace pfv {

// PF_ROWS (pfv_internal.h) = macroblock rows per window = warps per copy pipeline
constexpr int PF_WIN_BYTES = PF_WIN_W * PF_WIN_H;
constexpr int PF_STAGE = (PF_WIN_BYTES + 127) & ~127;
#ifndef PFV_PF_PIPES
#define PFV_PF_PIPES 4
#endif
#ifndef PFV_PF_STAGES
#define PFV_PF_STAGES 2
#endif
#ifndef PFV_PF_NG
#define PFV_PF_NG 13
#endif
constexpr int PF_PIPES = PFV_PF_PIPES;                       // independent copy pipelines per CTA (each: PF_ROWS warps, its own windows)
constexpr int PF_STAGES = PFV_PF_STAGES;                     // windows in flight per pipeline
constexpr int PF_COPY_WARPS = PF_PIPES * PF_ROWS;
constexpr int PF_NG = PFV_PF_NG;                             // ring groups; the ring must hold more macroblocks (104) than the
constexpr int PF_RING_MB = PF_NG * 8;                        // pipelines can have unfinished at once (4 x 24): see the empty-wait
constexpr int PF_COEF_PITCH = 528;                           // bytes per slot: 512 + 16 (8 slots -> 8 different bank groups)
constexpr int PF_PRED_PITCH = 80;                            // bytes per sub-block: 64 + 16
constexpr int PF_CTAS_PER_SM = 1;

struct __align__(16) PfGroup {
    unsigned char coef[8 * PF_COEF_PITCH];
    unsigned char pred[32 * PF_PRED_PITCH];                  // sub-block s of slot j: row r at [(s*8 + j)*80 + r*8]
    uint4         id[8];                                     // {byte offset of the macroblock in a frame slot, job, plane, valid}
};

constexpr int PF_MAX_JOBS = 64;                              // frames per launch (longer batches go out as several launches)
constexpr int PF_PIPE_ITEMS = 84;                            // windows one pipeline walks per launch (its slice of the item table)

// One window of a CTA's walk, decoded once at kernel start: the first version re-derived (frame, plane, window row, window
// column) from the item index three times per window in every thread - ~150 of the copy half's ~370 instructions.
struct __align__(16) PfItemRec {
    uint32_t job_p;                                          // job | plane << 16 | macroblock rows in the window << 20 | columns << 24
    uint32_t bx0_by0;                                        // pixel origin of the window's first macroblock: x | y << 16
    uint32_t hdr0;                                           // index of that macroblock in the frame's header / coefficient arrays
    uint32_t dst0;                                           // byte offset of its top-left pixel inside a frame slot
    uint32_t pw, bw;                                         // the plane's padded width in pixels / in macroblocks
    uint32_t max_xy;                                         // largest legal predictor origin: (pw - 16) | (ph - 16) << 16
    uint32_t plane_off;                                      // byte offset of the plane inside a frame slot
};

struct PfJob {                                               // what the kernel needs of a DecJob, kept in shared memory: every
    const int16_t  *coeff;                                   // field is read once per window, and a dependent global load per
    const uint32_t *hdr;                                     // window (pointer, then data) was what the copy half waited for
    uint8_t        *dst;
    const uint8_t  *ref;
    int32_t         ref_slot, pad;
};

struct __align__(128) PfSmem {
    unsigned char win[PF_PIPES][PF_STAGES][PF_STAGE];
    PfGroup  grp[PF_NG];
    PfItemRec item[PF_PIPES][PF_PIPE_ITEMS];
    PfJob    job[PF_MAX_JOBS];
    int32_t  deq[3][64];
    uint64_t win_full[PF_PIPES][PF_STAGES];
    uint64_t win_empty[PF_PIPES][PF_STAGES];                 // the pipeline's four warps have read the stage (4 arrivals)
    uint64_t grp_full[PF_NG];                                // 8 arrivals (one per slot) + the slots' coefficient bytes
    uint64_t grp_empty[PF_NG];                               // the transform warp has taken the group into registers
    uint32_t tail;                                           // ring slots handed out so far
    uint32_t total_groups;                                   // 0xffffffff until the copy half is done (written and polled with atomics)
};

static_assert(PF_CTAS_PER_SM * (sizeof(PfSmem) + 1024) <= 227 * 1024, "the fused decode-P kernel's CTAs must fit one SM");
static_assert(PF_RING_MB > PF_PIPES * PF_ROWS * 8, "the ring must be larger than what the copy pipelines can have unfinished at once");

__device__ __forceinline__ bool bar_try(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return done != 0;
}

// the same with a suspend-time hint (ns): an idle warp sleeps in the barrier unit instead of spinning through issue slots
__device__ __forceinline__ bool bar_try_sleepy(uint64_t *bar, uint32_t parity, uint32_t ns)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_addr(bar)), "r"(parity), "r"(ns) : "memory");
    return done != 0;
}

__device__ __forceinline__ void bar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void bar_arrive_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// 16 bytes at byte offset x of a window row (x + 24 <= PF_WIN_W): three 8-byte loads, one select level, four funnel shifts
__device__ __forceinline__ uint4 win_row16(const unsigned char *row, uint32_t x)
{
    const uint2 *q = reinterpret_cast<const uint2 *>(row + (x & ~7u));
    const uint2 a = q[0], b = q[1], c = q[2];
    const bool odd = (x & 4u) != 0u;
    const uint32_t t0 = odd ? a.y : a.x, t1 = odd ? b.x : a.y, t2 = odd ? b.y : b.x, t3 = odd ? c.x : b.y, t4 = odd ? c.y : c.x;
    const uint32_t sh = (x & 3u) * 8u;
    uint4 o;
    o.x = __funnelshift_r(t0, t1, sh);
    o.y = __funnelshift_r(t1, t2, sh);
    o.z = __funnelshift_r(t2, t3, sh);
    o.w = __funnelshift_r(t3, t4, sh);
    return o;
}

// the dequantiser out of shared memory (src/dct.rs:78-83, tables by scan position); see unpack_dequant
__device__ __forceinline__ void unpack_dequant_smem(const uint4 (&raw)[8], const int32_t *deq, int (&m)[64])
{
    constexpr int zz[64] = PFV_ZIGZAG_INIT;
    const int4 *dq = reinterpret_cast<const int4 *>(deq);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int4 d = dq[k];
        const uint4 &q = raw[k >> 1];
        const uint32_t w0 = (k & 1) ? q.z : q.x, w1 = (k & 1) ? q.w : q.y;
        m[zz[4 * k + 0]] = __dp2a_lo((int)w0, 0x00000001, 0) * d.x;
        m[zz[4 * k + 1]] = __dp2a_hi((int)w0, 0x01000000, 0) * d.y;
        m[zz[4 * k + 2]] = __dp2a_lo((int)w1, 0x00000001, 0) * d.z;
        m[zz[4 * k + 3]] = __dp2a_hi((int)w1, 0x01000000, 0) * d.w;
    }
}

template <int PF_XF_WARPS>
__global__ void __launch_bounds__((PF_COPY_WARPS + PF_XF_WARPS) * 32, PF_CTAS_PER_SM)
decode_p_fused_kernel(const __grid_constant__ SbParams P, const __grid_constant__ McWin W, const DecJob *__restrict__ jobs,
                      uint32_t njobs, int *__restrict__ err,
                      const __grid_constant__ CUtensorMap tm_luma, const __grid_constant__ CUtensorMap tm_chroma)
{
    extern __shared__ __align__(128) unsigned char pf_raw[];
    PfSmem &sm = *reinterpret_cast<PfSmem *>(pf_raw);
    constexpr uint32_t PF_THREADS = (PF_COPY_WARPS + PF_XF_WARPS) * 32;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const FrameGeom &g = P.g;
    const uint32_t nitems = njobs * W.total;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < PF_PIPES * PF_STAGES; ++s) { bar_init(&sm.win_full[0][0] + s, 1); bar_init(&sm.win_empty[0][0] + s, PF_ROWS); }
#pragma unroll
        for (int s = 0; s < PF_NG; ++s) { bar_init(&sm.grp_full[s], 8); bar_init(&sm.grp_empty[s], 1); }
        sm.tail = 0;
        sm.total_groups = 0xffffffffu;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = threadIdx.x; i < 3 * 64; i += PF_THREADS) (&sm.deq[0][0])[i] = (&P.deq[0][0])[i];
    for (uint32_t i = threadIdx.x; i < njobs; i += PF_THREADS) {
        const DecJob &j = jobs[i];
        sm.job[i] = PfJob{j.coeff, reinterpret_cast<const uint32_t *>(j.hdr), j.dst, j.ref, j.ref_slot, 0};
    }
    // (Measured and dropped, round 2: handing the windows out dynamically - a device-wide counter, the pipeline's leader decoding
    // the next record five windows ahead into a small ring.  It evens out the SMs (sm__cycles_active min / avg / max 93 k / 100 k /
    // 102 k instead of 80 k / 86 k / 95 k) but the leader's extra work per window costs more than the balance returns: 56.8 us per
    // 32 x 1080p against 52.1 us.)
    // (Measured and dropped as well, visit zx: the CTA's OWN windows handed to its four pipelines one at a time from a shared-memory
    // counter - no global atomic, no record to decode, the leader publishes the index to its pipeline two windows ahead.  603 k ->
    // 591 k frames/s at 32 frames per launch, 684 k -> 669 k at 64, 174 k -> 169 k at 3840x2160: the pipelines of an SM do not finish
    // far enough apart to pay for the hand-shake; what is uneven is the SMs among each other.)
    // The windows of the launch are dealt out to the copy pipelines (3 per CTA) in frame-interleaved order (pipelines that run at
    // the same time work on DIFFERENT frames): pipeline gp takes items gp, gp + npipes, ...; item -> (window wi = item / njobs,
    // job = item % njobs)
    const uint32_t npipes = gridDim.x * PF_PIPES;
    for (uint32_t t = threadIdx.x; t < (uint32_t)(PF_PIPES * PF_PIPE_ITEMS); t += PF_THREADS) {
        const uint32_t pipe = t / PF_PIPE_ITEMS, i = t - pipe * PF_PIPE_ITEMS;
        const uint32_t it = blockIdx.x * PF_PIPES + pipe + i * npipes;
        if (it >= nitems) continue;
        // (Measured and dropped, round 2: rotating the frame index by the window index - job = (item + item / njobs) % njobs - so that a
        // pipeline, whose stride through the items is 592 = 16 mod 32, walks all frames of the launch instead of two.  The SMs finish
        // as far apart as before (sm__cycles_active min / avg / max 80 k / 87 k / 99 k): it is not the frames' content.)
        const uint32_t wi = it / njobs, job = it - wi * njobs;
        const uint32_t p = (wi >= W.base[1] ? 1u : 0u) + (wi >= W.base[2] ? 1u : 0u);
        const PlaneGeom &pl = p == 0 ? g.pl[0] : (p == 1 ? g.pl[1] : g.pl[2]);
        const uint32_t li = wi - (p == 0 ? W.base[0] : (p == 1 ? W.base[1] : W.base[2]));
        const uint32_t txs = p == 0 ? W.tiles_x[0] : (p == 1 ? W.tiles_x[1] : W.tiles_x[2]);
        const uint32_t gy = li / txs, tx = li - gy * txs;
        const uint32_t rows = min((uint32_t)PF_ROWS, pl.bh - gy * PF_ROWS), cols = min(8u, pl.bw - tx * 8u);
        PfItemRec r;
        r.job_p = job | p << 16 | rows << 20 | cols << 24;
        r.bx0_by0 = (tx * 128u) | (gy * (PF_ROWS * 16u)) << 16;
        r.hdr0 = pl.mb_base + gy * PF_ROWS * pl.bw + tx * 8u;
        r.dst0 = pl.off + gy * (PF_ROWS * 16u) * pl.pw + tx * 128u;
        r.pw = pl.pw; r.bw = pl.bw;
        r.max_xy = (pl.pw - 16u) | (pl.ph - 16u) << 16;
        r.plane_off = pl.off;
        sm.item[pipe][i] = r;
    }
    __syncthreads();

    auto plane = [&](uint32_t p) -> const PlaneGeom & { return p == 0 ? g.pl[0] : (p == 1 ? g.pl[1] : g.pl[2]); };

    if (warp >= PF_COPY_WARPS) {
        // ================================ TRANSFORM half ================================
        const uint32_t sb = lane >> 3, sj = lane & 7u;
#pragma unroll 1
        for (uint32_t G = warp - PF_COPY_WARPS;; G += PF_XF_WARPS) {
            const uint32_t rgp = G % PF_NG, par = (G / PF_NG) & 1u;
            bool stop = false;
            while (!bar_try_sleepy(&sm.grp_full[rgp], par, 2000u)) {
                if (atomicOr(&sm.total_groups, 0u) <= G) { stop = true; break; }
            }
            if (stop) break;
            const PfGroup &grp = sm.grp[rgp];
            const uint4 id = grp.id[sj];
            uint4 raw[8], pv[4];
#pragma unroll
            for (int k = 0; k < 8; ++k) raw[k] = *reinterpret_cast<const uint4 *>(grp.coef + sj * PF_COEF_PITCH + sb * 128u + 16 * k);
#pragma unroll
            for (int k = 0; k < 4; ++k) pv[k] = *reinterpret_cast<const uint4 *>(grp.pred + lane * PF_PRED_PITCH + 16 * k);
            const uint32_t p = min(id.z, 2u);
            const int32_t *deq = sm.deq[p];
            int m[64];
            unpack_dequant_smem(raw, deq, m);
            __syncwarp();
            if (lane == 0) bar_arrive(&sm.grp_empty[rgp]);     // everything of the group is in registers now
            const bool valid = id.w != 0u;
            const uint32_t pw = plane(p).pw;
            uint8_t *dst = nullptr;
            if (valid) dst = sm.job[id.y].dst + id.x + (size_t)((sb >> 1) * 8u) * pw + (sb & 1u) * 8u;
            idct8x8_regs(m);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int r = 2 * k + h;
                    int y[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = m[r * 8 + i];
                    const uint2 o = apply_residual_row(y, h == 0 ? make_uint2(pv[k].x, pv[k].y) : make_uint2(pv[k].z, pv[k].w));   // src/common.rs:277
                    if (valid) __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pw), o);
                }
            }
        }
        return;
    }

    // ==================================== COPY half ====================================
    const uint32_t pipe = warp / PF_ROWS, wrow = warp % PF_ROWS;           // this warp's pipeline and its macroblock row of the window
    const uint32_t mb = lane >> 2, rg = lane & 3u;
    const uint32_t gp = blockIdx.x * PF_PIPES + pipe;
    const uint32_t nmine = gp < nitems ? min((uint32_t)PF_PIPE_ITEMS, (nitems - gp + npipes - 1) / npipes) : 0u;
    const PfItemRec *items = sm.item[pipe];
    const bool leader = wrow == 0 && lane == 0;
    auto issue = [&](const PfItemRec &it, uint32_t st) {       // the pipeline's leader only
        bar_arrive_tx(&sm.win_full[pipe][st], (uint32_t)PF_WIN_BYTES);
        const uint32_t p = (it.job_p >> 16) & 3u;
        const CUtensorMap *tm = p == 0 ? &tm_luma : &tm_chroma;
        const int cx = (int)(it.bx0_by0 & 0xffffu) - 16, cy = (int)(it.bx0_by0 >> 16) - 15, cz = p == 2 ? 1 : 0;
        const int cw = sm.job[it.job_p & 0xffffu].ref_slot;
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
            "[%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(smem_addr(sm.win[pipe][st])), "l"(tm), "r"(cx), "r"(cy), "r"(cz), "r"(cw), "r"(smem_addr(&sm.win_full[pipe][st]))
            : "memory");
    };
    // (Measured and dropped, round 2: the same box as an L2 prefetch - cp.async.bulk.prefetch.tensor - two windows beyond the
    // staged ones.  The wait for the window's TMA is the copy half's largest single stall, but with the prefetch the kernel
    // took 57 us per 32 frames instead of 52: the prefetches compete with the demand loads for the same DRAM queues.)
    // header word of this lane's macroblock (row `warp`, column `mb` of the window); bit 31 set = no such macroblock
    auto load_hw = [&](const PfItemRec &it) -> uint32_t {
        if (wrow >= ((it.job_p >> 20) & 15u) || mb >= (it.job_p >> 24)) return 0x80000000u;
        return __ldg(sm.job[it.job_p & 0xffffu].hdr + it.hdr0 + wrow * it.bw + mb);
    };

    // Software pipeline over windows k (being copied), k+1 (TMA in flight) and k+2 (TMA issued at the end of iteration k), headers
    // loaded two windows ahead: no load is consumed in the iteration that issues it.
    if (leader)
        for (uint32_t s = 0; s < (uint32_t)PF_STAGES && s < nmine; ++s) issue(items[s], s);
    uint32_t hw_cur = nmine > 0 ? load_hw(items[0]) : 0x80000000u;
    uint32_t hw_n1 = nmine > 1 ? load_hw(items[1]) : 0x80000000u;
#pragma unroll 1
    for (uint32_t k = 0; k < nmine; ++k) {
        const uint32_t st = k % PF_STAGES;
        const PfItemRec cur = items[k];
        uint32_t hw_n2 = 0x80000000u;
        if (k + 2u < nmine) hw_n2 = load_hw(items[k + 2u]);

        const uint32_t cjob = cur.job_p & 0xffffu, cp = (cur.job_p >> 16) & 3u;
        const PfJob &job = sm.job[cjob];
        const bool exists = !(hw_cur & 0x80000000u);
        const bool coded = exists && ((hw_cur >> 16) & 0xffu) != 0u;
        const int bx = (int)((cur.bx0_by0 & 0xffffu) + mb * 16u), by = (int)((cur.bx0_by0 >> 16) + wrow * 16u);
        int mvx = (int)(int8_t)(hw_cur & 0xffu), mvy = (int)(int8_t)((hw_cur >> 8) & 0xffu);   // src/common.rs:255-256
        if (exists) {
            const int sx = bx + mvx, sy = by + mvy;
            if (sx < 0 || sy < 0 || sx > (int)(cur.max_xy & 0xffffu) || sy > (int)(cur.max_xy >> 16)) {
                // reference: debug_assert / slice panic (src/common.rs:258-259).  Never follow it: the stream is
                // flagged bad and the co-located block is used.
                if (rg == 0) atomicOr(err, ERRBIT_BAD_MV);
                mvx = 0; mvy = 0;
            }
        }
        // ring slots for this warp's coded macroblocks (one warp-aggregated shared-memory atomic)
        const uint32_t vote = __ballot_sync(0xffffffffu, coded && rg == 0);
        uint32_t e = 0;
        if (vote) {
            if (lane == 0) e = atomicAdd(&sm.tail, (uint32_t)__popc(vote));
            e = __shfl_sync(0xffffffffu, e, 0) + (uint32_t)__popc(vote & ((1u << (lane & ~3u)) - 1u));
        }
        const uint32_t slot = e % PF_RING_MB, rgp = slot >> 3, sj = slot & 7u;
        PfGroup &grp = sm.grp[rgp];
        if (coded) bar_wait(&sm.grp_empty[rgp], ((e / PF_RING_MB) & 1u) ^ 1u);   // the slot's previous tenant has been taken out

        bar_wait(&sm.win_full[pipe][st], (k / PF_STAGES) & 1u);
        const uint32_t pw = cur.pw;
        const uint32_t mb_off = cur.dst0 + wrow * 16u * pw + mb * 16u;           // this macroblock's top-left pixel in a frame slot
        if (exists) {
            uint8_t *dst = job.dst + mb_off + (size_t)rg * pw;
            const bool in_window = mvx >= -16 && mvx <= 15 && mvy >= -15 && mvy <= 15;
            // Vectors beyond +-15 are legal for the reference decoder (7-bit vectors, src/dec.rs:367-368) but outside the
            // staged window - its own encoder never searches further (src/common.rs:154-204).  Rare: fetch from global.
            const uint32_t wx = (uint32_t)(16 + (int)mb * 16 + mvx), wy = (uint32_t)(15 + (int)wrow * 16 + mvy) + rg;
            const unsigned char *wline = sm.win[pipe][st] + wy * PF_WIN_W;
            // all four rows are fetched before anything is stored: the predictor stores below go to shared memory too, and
            // the compiler must assume they alias the window (it kept every row's loads behind the previous row's stores and
            // each row waited for its own shared-memory round trip)
            uint4 o[4];
            if (in_window) {
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = win_row16(wline + i * 4 * PF_WIN_W, wx);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint8_t *gsrc = job.ref + cur.plane_off + (size_t)((uint32_t)(by + mvy) + rg + 4u * (uint32_t)i) * pw + (uint32_t)(bx + mvx);
                    const uint2 a = ldg_u8x8_unaligned(gsrc);
                    const uint2 b = ldg_u8x8_unaligned(gsrc + 8);
                    o[i] = make_uint4(a.x, a.y, b.x, b.y);
                }
            }
            if (coded) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    // row R = rg + 4i of the macroblock: left half -> sub-block (R >> 3) * 2, right half -> the next one
                    const uint32_t R = rg + 4u * (uint32_t)i, L = (R >> 3) * 16u + sj;
                    unsigned char *pp = grp.pred + L * PF_PRED_PITCH + (R & 7u) * 8u;
                    *reinterpret_cast<uint2 *>(pp) = make_uint2(o[i].x, o[i].y);
                    *reinterpret_cast<uint2 *>(pp + 8 * PF_PRED_PITCH) = make_uint2(o[i].z, o[i].w);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) __stcg(reinterpret_cast<uint4 *>(dst + (size_t)(4 * i) * pw), o[i]);   // blit_block, src/common.rs:341-349
            }
        }
        if (coded && rg == 0) grp.id[sj] = make_uint4(mb_off, cjob, cp, 1u);
        __syncwarp();                                           // the four lanes of a macroblock have written its predictor
        if (coded && rg == 0) {
            bar_arrive_tx(&sm.grp_full[rgp], 512u);
            bulk_copy_g2s(grp.coef + sj * PF_COEF_PITCH, job.coeff + (size_t)(cur.hdr0 + wrow * cur.bw + mb) * 256, 512u, &sm.grp_full[rgp]);
        }
        // this warp is done with the window stage; the stage is refilled once all four are (no CTA or pipeline barrier: a
        // warp that finishes early goes on to the next window, only the leader's lane 0 waits for the slowest one)
        if (lane == 0) bar_arrive(&sm.win_empty[pipe][st]);
        if (leader && k + PF_STAGES < nmine) {
            bar_wait(&sm.win_empty[pipe][st], (k / PF_STAGES) & 1u);
            issue(items[k + PF_STAGES], st);
        }
        hw_cur = hw_n1; hw_n1 = hw_n2;
    }

    // the copy half is done: complete the last, partly filled group with empty slots and tell the transform half where to stop
    asm volatile("bar.sync %0, %1;" ::"n"(1 + PF_PIPES), "n"(PF_COPY_WARPS * 32) : "memory");
    if (threadIdx.x == 0) {
        const uint32_t tail = *reinterpret_cast<volatile uint32_t *>(&sm.tail);
        const uint32_t rem = (8u - (tail & 7u)) & 7u;
        for (uint32_t q = 0; q < rem; ++q) {
            const uint32_t e = tail + q, slot = e % PF_RING_MB, rgp = slot >> 3;
            bar_wait(&sm.grp_empty[rgp], ((e / PF_RING_MB) & 1u) ^ 1u);
            sm.grp[rgp].id[slot & 7u] = make_uint4(0u, 0u, 0u, 0u);
            bar_arrive(&sm.grp_full[rgp]);
        }
        __threadfence_block();
        atomicExch(&sm.total_groups, (tail + 7u) >> 3);
    }
}

template <int XF>
static cudaError_t launch_decode_p_fused_t(const SbParams &P, const DecJob *d_jobs, uint32_t njobs, int *d_err,
                                           const CUtensorMap &tm_luma, const CUtensorMap &tm_chroma, cudaStream_t s)
{
    static bool attr_done[64] = {};           // cudaFuncSetAttribute is per DEVICE: one process may own contexts on several
    const int smem = (int)sizeof(PfSmem);
    if (first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(decode_p_fused_kernel<XF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    const McWin W = make_mc_windows(P.g, PF_ROWS);
    // frames per launch: at most PF_MAX_JOBS, and no more than lets every pipeline's item table hold its share of the windows
    const uint32_t max_pipes = 148u * PF_CTAS_PER_SM * PF_PIPES;
    uint32_t per_launch = (uint32_t)(((uint64_t)PF_PIPE_ITEMS * max_pipes) / W.total);
    per_launch = per_launch < 1u ? 1u : (per_launch > (uint32_t)PF_MAX_JOBS ? (uint32_t)PF_MAX_JOBS : per_launch);
    for (uint32_t j0 = 0; j0 < njobs; j0 += per_launch) {
        const uint32_t n = njobs - j0 < per_launch ? njobs - j0 : per_launch;
        uint32_t ctas = (n * W.total + PF_PIPES - 1) / PF_PIPES;
        if (ctas > 148u * PF_CTAS_PER_SM) ctas = 148u * PF_CTAS_PER_SM;
        if ((n * W.total + ctas * PF_PIPES - 1) / (ctas * PF_PIPES) > (uint32_t)PF_PIPE_ITEMS) return cudaErrorInvalidConfiguration;   // (a frame of > 37 000 windows)
        decode_p_fused_kernel<XF><<<ctas, (PF_COPY_WARPS + XF) * 32, smem, s>>>(P, W, d_jobs + j0, n, d_err, tm_luma, tm_chroma);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_decode_p_fused(const SbParams &P, const DecJob *d_jobs, uint32_t njobs, int *d_err,
                                  const CUtensorMap &tm_luma, const CUtensorMap &tm_chroma, cudaStream_t s)
{
    static const int xf_env = getenv("PFV_PF_XF") ? atoi(getenv("PFV_PF_XF")) : 0;   // tuning aid: transform warps per CTA
    if (xf_env == 4) return launch_decode_p_fused_t<4>(P, d_jobs, njobs, d_err, tm_luma, tm_chroma, s);
    if (xf_env == 8) return launch_decode_p_fused_t<8>(P, d_jobs, njobs, d_err, tm_luma, tm_chroma, s);
    return launch_decode_p_fused_t<6>(P, d_jobs, njobs, d_err, tm_luma, tm_chroma, s);
}

}  // namespace pfv
