// pfv_kernels_sb.cu — "register-resident sub-block" kernels (sm_100a).
//
// One THREAD owns one 8x8 sub-block: its 64 coefficients live in registers, both 1-D passes run without any
// exchange, the zig-zag permutation is a compile-time register renaming and the (de)quantiser tables are
// kernel parameters, i.e. constant-bank operands of the multiplies.  A warp covers 8 consecutive macroblocks
// (lane = mb*4 + sub-block), so one 8-byte row store of the warp forms two full 128-byte lines of the plane.
// Compared with the warp-per-macroblock kernels of pfv_kernels.cu this removes every shared-memory round trip
// and amortises the addressing over 8 macroblocks (ncu: 255 -> ~150 warp instructions per macroblock).
//
// Arithmetic is the reference's (see pfv_device.cuh for the file:line map).
#include <stdlib.h>

#include "pfv_internal.h"
#include "pfv_device.cuh"
#include "pfv_sb.cuh"

namespace pfv {

constexpr int SB_WARPS = SB_THREADS / 32;     // 4 warps = 32 macroblocks per CTA, never straddling a plane

__global__ void __launch_bounds__(SB_WARPS * 32, 4)
decode_i_sb_kernel(const __grid_constant__ SbParams P, const DecJob *__restrict__ jobs)
{
    constexpr int zz[64] = PFV_ZIGZAG_INIT;
    // plane from the CTA index alone: everything derived from it stays in the uniform datapath
    const uint32_t cta = blockIdx.x;
    const int p = (cta >= P.cta_base[1] ? 1 : 0) + (cta >= P.cta_base[2] ? 1 : 0);
    const PlaneGeom &pl = p == 0 ? P.g.pl[0] : (p == 1 ? P.g.pl[1] : P.g.pl[2]);
    const uint32_t lm = (cta - (p == 0 ? P.cta_base[0] : (p == 1 ? P.cta_base[1] : P.cta_base[2]))) * (SB_WARPS * 8) +
                        (threadIdx.x >> 2);                  // macroblock inside the plane
    if (lm >= pl.bw * pl.bh) return;
    const int sb = threadIdx.x & 3;
    const DecJob job = jobs[blockIdx.y];

    const uint4 *src = reinterpret_cast<const uint4 *>(job.coeff + ((size_t)(pl.mb_base + lm) * 256 + sb * 64));
    uint4 raw[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) raw[k] = __ldcs(src + k);

    const int32_t *deq = p == 0 ? P.deq[0] : (p == 1 ? P.deq[1] : P.deq[2]);
    int m[64];
#pragma unroll
    for (int s = 0; s < 64; ++s) {
        const uint4 &q = raw[s >> 3];
        const uint32_t w = ((s >> 1) & 3) == 0 ? q.x : ((s >> 1) & 3) == 1 ? q.y : ((s >> 1) & 3) == 2 ? q.z : q.w;
        const int c = (s & 1) ? ((int)w >> 16) : (int)(int16_t)(w & 0xffffu);
        m[zz[s]] = c * deq[s];                               // src/dct.rs:78-83 (tables by scan position)
    }
    idct8x8_regs(m);

    uint32_t col;
    const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
    uint8_t *dst = job.dst + pl.off + (size_t)(row * 16u + (uint32_t)(sb >> 1) * 8u) * pl.pw + col * 16u + (uint32_t)(sb & 1) * 8u;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        uint2 o;
        o.x = pack4_sat_u8(m[r * 8 + 0], m[r * 8 + 1], m[r * 8 + 2], m[r * 8 + 3]);
        o.y = pack4_sat_u8(m[r * 8 + 4], m[r * 8 + 5], m[r * 8 + 6], m[r * 8 + 7]);
        __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pl.pw), o);
    }
}


// -------------------------------------------------------------------------------------------------
// decode, "classify, compact, transform" (the default for I and P frames)
// -------------------------------------------------------------------------------------------------
// The exact integer IDCT costs ~1300 instructions per sub-block, which at 48 960 sub-blocks per 1080p frame is
// more issue time than the frame's HBM time.  Real streams are sparse: most sub-blocks carry only a DC
// coefficient, and for those both passes collapse exactly (src/dct.rs:241-293 with v[1..7] = 0 returns v[0] in
// every output; all-zero columns stay zero): every pixel is clamp((c0 * deq0 + 32768) >> 8).
//
// Each WARP streams over `tiles_per_warp` consecutive tiles of 8 macroblocks (lane = macroblock*4 + sub-block):
//   A. load the tile (8 x 16 B per lane; P frames: header, motion-compensated predictor), finish the sub-blocks
//      that need no transform on the spot (DC-only, skipped) and append the others to the warp's private
//      shared-memory ring (ballot + prefix, no atomics, no CTA barrier);
//   B. whenever the ring holds 32 entries, run the full register-resident transform on them: a full warp.
// What is left at the end (< 32 entries per warp) is pooled across the CTA's four warps and flushed.
// Results do not depend on the path taken; tests/test_gpu_parity.py drives dense, sparse and mixed inputs.
constexpr int SBW_WARPS = 4;
constexpr int SBW_RING = 64;                // slots per warp: up to 31 carried + 32 new

struct __align__(16) WarpRing {
    uint4    coef[SBW_RING * 8];            // slot s keeps 16-byte chunk k at [s*8 + (k ^ (s & 7))]: conflict-free both ways
    uint32_t id[SBW_RING];                  // (macroblock inside the plane << 2) | sub-block
    uint32_t hw[SBW_RING];                  // P frames: the macroblock's header word {mx, my, has_coeff, 0}
};

struct SbWhere {
    uint8_t       *dst;                     // top-left of the 8x8 in the destination slot
    const uint8_t *ref;                     // P: top-left of the motion-compensated 8x8 in the reference slot
    bool           bad_mv;
};

template <bool INTER>
__device__ __forceinline__ SbWhere sb_where(const DecJob &job, const PlaneGeom &pl, uint32_t lm, int sb, uint32_t hw)
{
    SbWhere w;
    uint32_t col;
    const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
    const uint32_t bx = col * 16u, by = row * 16u;
    const uint32_t oy = (uint32_t)(sb >> 1) * 8u, ox = (uint32_t)(sb & 1) * 8u;
    w.dst = job.dst + pl.off + (size_t)(by + oy) * pl.pw + bx + ox;
    w.ref = nullptr;
    w.bad_mv = false;
    if (INTER) {
        int sx = (int)bx + (int)(int8_t)(hw & 0xffu), sy = (int)by + (int)(int8_t)((hw >> 8) & 0xffu);   // src/common.rs:255-256
        if (sx < 0 || sy < 0 || sx > (int)pl.pw - 16 || sy > (int)pl.ph - 16) {
            // reference: debug_assert / slice panic (src/common.rs:258-259).  Never read out of bounds: the
            // stream is flagged bad and the co-located block is used.
            w.bad_mv = true;
            sx = (int)bx;
            sy = (int)by;
        }
        w.ref = job.ref + pl.off + (size_t)((uint32_t)sy + oy) * pl.pw + (uint32_t)sx + ox;
    }
    return w;
}

// Phase B for one ring entry per lane: full transform (+ residual on the re-fetched predictor for P frames).
template <bool INTER>
__device__ __forceinline__ void transform_entry(const WarpRing &ring, uint32_t slot, const DecJob &job,
                                                const PlaneGeom &pl, const int32_t *deq)
{
    uint4 r2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r2[k] = ring.coef[slot * 8u + ((uint32_t)k ^ (slot & 7u))];
    const uint32_t id = ring.id[slot];
    const SbWhere w = sb_where<INTER>(job, pl, id >> 2, (int)(id & 3u), INTER ? ring.hw[slot] : 0u);
    int m[64];
    unpack_dequant(r2, deq, m);
    idct8x8_regs(m);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        uint2 o;
        if (INTER) {
            int y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = m[r * 8 + i];
            o = apply_residual_row(y, ldg_u8x8_unaligned(w.ref + (size_t)r * pl.pw));   // src/common.rs:277
        } else {
            o.x = pack4_sat_u8(m[r * 8 + 0], m[r * 8 + 1], m[r * 8 + 2], m[r * 8 + 3]);
            o.y = pack4_sat_u8(m[r * 8 + 4], m[r * 8 + 5], m[r * 8 + 6], m[r * 8 + 7]);
        }
        __stcg(reinterpret_cast<uint2 *>(w.dst + (size_t)r * pl.pw), o);
    }
}

template <bool INTER>
__global__ void __launch_bounds__(SBW_WARPS * 32, 4)
decode_sbw_kernel(const __grid_constant__ SbParams P, const DecJob *__restrict__ jobs, int *__restrict__ err)
{
    __shared__ WarpRing rings[SBW_WARPS];
    __shared__ uint32_t left_head[SBW_WARPS], left_cnt[SBW_WARPS];

    const uint32_t cta = blockIdx.x;
    const int p = (cta >= P.cta_base[1] ? 1 : 0) + (cta >= P.cta_base[2] ? 1 : 0);
    const PlaneGeom &pl = p == 0 ? P.g.pl[0] : (p == 1 ? P.g.pl[1] : P.g.pl[2]);
    const int32_t *deq = p == 0 ? P.deq[0] : (p == 1 ? P.deq[1] : P.deq[2]);
    const uint32_t nmb = pl.bw * pl.bh, ntiles = (nmb + 7u) / 8u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tile_begin = ((cta - (p == 0 ? P.cta_base[0] : (p == 1 ? P.cta_base[1] : P.cta_base[2]))) * SBW_WARPS + warp) * P.tiles_per_warp;
    const uint32_t tile_end = min(tile_begin + P.tiles_per_warp, ntiles);
    const DecJob job = jobs[blockIdx.y];
    const int sb = (int)(lane & 3u);
    WarpRing &ring = rings[warp];
    const uint32_t *hdr32 = INTER ? reinterpret_cast<const uint32_t *>(job.hdr) + pl.mb_base : nullptr;

    uint32_t head = 0, tail = 0;                              // warp-uniform ring positions
    uint32_t hnext = 0;
    if (INTER && tile_begin < tile_end && tile_begin * 8u + (lane >> 2) < nmb) hnext = __ldg(hdr32 + tile_begin * 8u + (lane >> 2));

#pragma unroll 1
    for (uint32_t tile = tile_begin; tile < tile_end; ++tile) {
        const uint32_t lm = tile * 8u + (lane >> 2);
        const bool valid = lm < nmb;
        const uint32_t hw = hnext;                            // {mx, my, has_coeff, 0} (src/dec.rs:9-13)
        if (INTER && tile + 1 < tile_end && lm + 8u < nmb) hnext = __ldg(hdr32 + lm + 8u);

        // ---- A: load, classify, finish what needs no transform ----
        bool general = false;
        uint4 raw[8];
        if (valid) {
            const SbWhere w = sb_where<INTER>(job, pl, lm, sb, hw);
            if (INTER && w.bad_mv) atomicOr(err, ERRBIT_BAD_MV);
            uint2 prev[8];
            if (INTER) {
#pragma unroll
                for (int r = 0; r < 8; ++r) prev[r] = ldg_u8x8_unaligned(w.ref + (size_t)r * pl.pw);   // get_block, src/common.rs:327-339
            }
            const bool coded = !INTER || ((hw >> 16) & 0xffu) != 0u;
            uint32_t pos4 = 0u, neg4 = 0u, dc4 = 0u;
            if (coded) {
                const uint4 *src = reinterpret_cast<const uint4 *>(job.coeff + ((size_t)(pl.mb_base + lm) * 256 + sb * 64));
#pragma unroll
                for (int k = 0; k < 8; ++k) raw[k] = __ldcs(src + k);
                uint32_t ac = raw[0].x & 0xffff0000u;
                ac |= raw[0].y | raw[0].z | raw[0].w;
#pragma unroll
                for (int k = 1; k < 8; ++k) ac |= raw[k].x | raw[k].y | raw[k].z | raw[k].w;
                general = ac != 0u;
                const int c0 = (int)(int16_t)(raw[0].x & 0xffffu);
                const int v = (c0 * deq[0] + (128 << 8)) >> 8;                           // both passes collapse to the DC term
                if (INTER) {
                    const int delta = (min(max(v, 0), 255) - 128) * 2;                   // src/common.rs:101
                    pos4 = (uint32_t)max(delta, 0) * 0x01010101u;
                    neg4 = (uint32_t)min(max(-delta, 0), 255) * 0x01010101u;
                } else {
                    dc4 = pack4_sat_u8(v, v, v, v);
                }
            }
            if (!general) {
                if (INTER && (pos4 | neg4) != 0u) {                                      // DC-only residual
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        prev[r].x = add_delta_sat4(prev[r].x, pos4, neg4);
                        prev[r].y = add_delta_sat4(prev[r].y, pos4, neg4);
                    }
                }
#pragma unroll
                for (int r = 0; r < 8; ++r)                                              // skipped: the copy, src/common.rs:281-283
                    __stcg(reinterpret_cast<uint2 *>(w.dst + (size_t)r * pl.pw), INTER ? prev[r] : make_uint2(dc4, dc4));
            }
        }
        const uint32_t vote = __ballot_sync(0xffffffffu, general);
        if (general) {
            const uint32_t slot = (tail + (uint32_t)__popc(vote & ((1u << lane) - 1u))) & (SBW_RING - 1);
#pragma unroll
            for (int k = 0; k < 8; ++k) ring.coef[slot * 8u + ((uint32_t)k ^ (slot & 7u))] = raw[k];
            ring.id[slot] = (lm << 2) | (uint32_t)sb;
            if (INTER) ring.hw[slot] = hw;
        }
        tail += (uint32_t)__popc(vote);
        __syncwarp();

        // ---- B: a full warp of queued sub-blocks ----
        if (tail - head >= 32u) {
            transform_entry<INTER>(ring, (head + lane) & (SBW_RING - 1), job, pl, deq);
            head += 32u;
            __syncwarp();
        }
    }

    // ---- flush: pool what the four warps have left ----
    if (lane == 0) { left_head[warp] = head; left_cnt[warp] = tail - head; }
    __syncthreads();
    uint32_t pre[SBW_WARPS + 1];
    pre[0] = 0;
#pragma unroll
    for (int w = 0; w < SBW_WARPS; ++w) pre[w + 1] = pre[w] + left_cnt[w];
#pragma unroll 1
    for (uint32_t c = warp * 32u; c < pre[SBW_WARPS]; c += SBW_WARPS * 32u) {
        const uint32_t e = c + lane;
        if (e < pre[SBW_WARPS]) {
            int w = 0;
#pragma unroll
            for (int k = 1; k < SBW_WARPS; ++k) w += e >= pre[k] ? 1 : 0;
            transform_entry<INTER>(rings[w], (left_head[w] + (e - pre[w])) & (SBW_RING - 1), job, pl, deq);
        }
    }
}


// -------------------------------------------------------------------------------------------------
// decode-I with TMA-staged tiles (the default for key frames)
// -------------------------------------------------------------------------------------------------
// Same classify/compact/transform structure, but the 4 KB coefficient tile of a warp (8 macroblocks x 512 B,
// contiguous in the dense layout) is fetched by ONE bulk async copy (cp.async.bulk, the TMA unit) into a
// per-warp double buffer, completion signalled on a per-stage mbarrier.  While a warp classifies or transforms
// tile i, tiles i+1 and i+2 are in flight without holding registers or LSU slots, which is what the load phase
// of decode_sbw_kernel<false> lacked (ncu: 25 % issue utilisation, long-scoreboard stalls, 45 % DRAM).
// Lanes read their 128 B from the stage with the chunk order rotated by (lane & 7) so the 128-byte-strided
// reads are bank-conflict free; the rotation is undone by address arithmetic when an entry is queued.
constexpr int STG_TILE_BYTES = 8 * 512;

// STAGES tiles in flight per warp, RING queue slots per warp.  <2, 64> (the default): the queue always has room for a
// whole tile (31 carried + 32 new).  <3, 48> and <3, 40> (PFV_DECODE_I_STAGES=3 / 4): a third tile in flight per warp
// inside the same 3-CTAs-per-SM shared-memory budget; the queue may then be too small for a tile's general
// sub-blocks, in which case the ones that do not fit wait for a transform pass and are re-read from the stage (kept
// until then).  Measured on the config-2 stream: 0.81 of roofline for <2, 64> against 0.63 / 0.61 - textured regions
// are dense, so the overflow path is taken all the time there; kept selectable, parity-tested, not used.
template <int STAGES, int RING>
struct __align__(128) StreamSmemT {
    uint4    stage[SBW_WARPS][STAGES][STG_TILE_BYTES / 16];
    uint4    coef[SBW_WARPS][RING * 8];
    uint32_t id[SBW_WARPS][RING];
    uint64_t bar[SBW_WARPS][STAGES];
    uint32_t left_head[SBW_WARPS], left_cnt[SBW_WARPS];
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    const uint32_t b = smem_addr(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(b) : "memory");
}

__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t b = smem_addr(bar);
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(b), "r"(parity) : "memory");
    }
}

template <int RING>
__device__ __forceinline__ uint32_t ring_slot(uint32_t u) { return (RING & (RING - 1)) == 0 ? (u & (uint32_t)(RING - 1)) : (u % (uint32_t)RING); }

__device__ __forceinline__ void transform_entry_i(const uint4 *coef, const uint32_t *idv, uint32_t slot, const DecJob &job,
                                                  const PlaneGeom &pl, const int32_t *deq)
{
    uint4 r2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r2[k] = coef[slot * 8u + ((uint32_t)k ^ (slot & 7u))];
    const uint32_t id = idv[slot];
    const SbWhere w = sb_where<false>(job, pl, id >> 2, (int)(id & 3u), 0u);
    int m[64];
    unpack_dequant(r2, deq, m);
    idct8x8_regs(m);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        uint2 o;
        o.x = pack4_sat_u8(m[r * 8 + 0], m[r * 8 + 1], m[r * 8 + 2], m[r * 8 + 3]);
        o.y = pack4_sat_u8(m[r * 8 + 4], m[r * 8 + 5], m[r * 8 + 6], m[r * 8 + 7]);
        __stcg(reinterpret_cast<uint2 *>(w.dst + (size_t)r * pl.pw), o);
    }
}

template <int STAGES, int RING>
__global__ void __launch_bounds__(SBW_WARPS * 32, 3)
decode_i_stream_kernel(const __grid_constant__ SbParams P, const DecJob *__restrict__ jobs)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StreamSmemT<STAGES, RING> &sm = *reinterpret_cast<StreamSmemT<STAGES, RING> *>(smem_raw);

    const uint32_t cta = blockIdx.x;
    const int p = (cta >= P.cta_base[1] ? 1 : 0) + (cta >= P.cta_base[2] ? 1 : 0);
    const PlaneGeom &pl = p == 0 ? P.g.pl[0] : (p == 1 ? P.g.pl[1] : P.g.pl[2]);
    const int32_t *deq = p == 0 ? P.deq[0] : (p == 1 ? P.deq[1] : P.deq[2]);
    const uint32_t nmb = pl.bw * pl.bh, ntiles = (nmb + 7u) / 8u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tile_begin = ((cta - (p == 0 ? P.cta_base[0] : (p == 1 ? P.cta_base[1] : P.cta_base[2]))) * SBW_WARPS + warp) * P.tiles_per_warp;
    const uint32_t tile_end = min(tile_begin + P.tiles_per_warp, ntiles);
    const uint32_t ntl = tile_end > tile_begin ? tile_end - tile_begin : 0u;
    const DecJob job = jobs[blockIdx.y];
    const int sb = (int)(lane & 3u);
    const uint32_t rot = lane & 7u;
    uint4 *ring = sm.coef[warp];
    uint32_t *ring_id = sm.id[warp];
    const char *plane_coeff = reinterpret_cast<const char *>(job.coeff + (size_t)pl.mb_base * 256);

    auto issue = [&](uint32_t i) {                            // lane 0 only: tile (tile_begin + i) into stage i % STAGES
        const uint32_t tile = tile_begin + i;
        const uint32_t mbs = min(8u, nmb - tile * 8u);
        bulk_load(sm.stage[warp][i % STAGES], plane_coeff + (size_t)tile * STG_TILE_BYTES, mbs * 512u,
                  &sm.bar[warp][i % STAGES]);
    };
    if (lane == 0) {
#pragma unroll
        for (int s2 = 0; s2 < STAGES; ++s2)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&sm.bar[warp][s2])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (uint32_t i = 0; i < (uint32_t)STAGES; ++i)
            if (i < ntl) issue(i);
    }
    __syncwarp();

    uint32_t head = 0, tail = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < ntl; ++i) {
        const uint32_t tile = tile_begin + i;
        const uint32_t lm = tile * 8u + (lane >> 2);
        const bool valid = lm < nmb;
        const uint4 *stg = sm.stage[warp][i % STAGES];

        // ---- A: take this lane's 128 B out of the stage (chunk k ^ rot lands in raw[k]) ----
        bar_wait(&sm.bar[warp][i % STAGES], (i / STAGES) & 1u);
        uint4 raw[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) raw[k] = stg[lane * 8u + ((uint32_t)k ^ rot)];

        uint32_t ac = 0u, w0 = 0u;                            // w0: the word that holds the DC coefficient (chunk 0)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool is0 = rot == (uint32_t)k;
            w0 = is0 ? raw[k].x : w0;
            ac |= (is0 ? (raw[k].x & 0xffff0000u) : raw[k].x) | raw[k].y | raw[k].z | raw[k].w;
        }
        const bool general = valid && ac != 0u;
        const uint32_t vote = __ballot_sync(0xffffffffu, general);  // (also: every lane has read the stage)
        const uint32_t n = (uint32_t)__popc(vote), rank = (uint32_t)__popc(vote & ((1u << lane) - 1u));
        const uint32_t room = (uint32_t)RING - (tail - head);
        const bool overflow = RING < 63 && n > room;          // compile-time false for the 64-slot queue
        if (!overflow && lane == 0 && i + STAGES < ntl) issue(i + STAGES);   // the stage may be refilled
        const bool fits = general && (!overflow || rank < room);
        if (fits) {
            const uint32_t slot = ring_slot<RING>(tail + rank);
            const uint32_t y = rot ^ (slot & 7u);             // raw[k] is chunk k ^ rot; the ring keeps chunk j at j ^ (slot & 7)
#pragma unroll
            for (int k = 0; k < 8; ++k) ring[slot * 8u + ((uint32_t)k ^ y)] = raw[k];
            ring_id[slot] = (lm << 2) | (uint32_t)sb;
        } else if (valid && !general) {
            const int c0 = (int)(int16_t)(w0 & 0xffffu);
            const int v = (c0 * deq[0] + (128 << 8)) >> 8;   // both passes collapse to the DC term
            const uint32_t dc4 = pack4_sat_u8(v, v, v, v);
            const SbWhere w = sb_where<false>(job, pl, lm, sb, 0u);
#pragma unroll
            for (int r = 0; r < 8; ++r) __stcg(reinterpret_cast<uint2 *>(w.dst + (size_t)r * pl.pw), make_uint2(dc4, dc4));
        }
        if (overflow) {
            // the queue is full (RING entries): transform 32 of them, then queue the sub-blocks that had to wait,
            // re-reading them from the stage, and only then let the stage go
            tail += room;
            __syncwarp();
            transform_entry_i(ring, ring_id, ring_slot<RING>(head + lane), job, pl, deq);
            head += 32u;
            __syncwarp();
            if (general && rank >= room) {
                const uint32_t slot = ring_slot<RING>(tail + (rank - room));
                const uint32_t y = rot ^ (slot & 7u);
#pragma unroll
                for (int k = 0; k < 8; ++k) ring[slot * 8u + ((uint32_t)k ^ y)] = stg[lane * 8u + ((uint32_t)k ^ rot)];
                ring_id[slot] = (lm << 2) | (uint32_t)sb;
            }
            tail += n - room;
            __syncwarp();
            if (lane == 0 && i + STAGES < ntl) issue(i + STAGES);
        } else {
            tail += n;
            __syncwarp();
        }

        // ---- B: a full warp of queued sub-blocks ----
        if (tail - head >= 32u) {
            transform_entry_i(ring, ring_id, ring_slot<RING>(head + lane), job, pl, deq);
            head += 32u;
            __syncwarp();
        }
    }

    // ---- flush: pool what the four warps have left ----
    if (lane == 0) { sm.left_head[warp] = head; sm.left_cnt[warp] = tail - head; }
    __syncthreads();
    uint32_t pre[SBW_WARPS + 1];
    pre[0] = 0;
#pragma unroll
    for (int w = 0; w < SBW_WARPS; ++w) pre[w + 1] = pre[w] + sm.left_cnt[w];
#pragma unroll 1
    for (uint32_t c = warp * 32u; c < pre[SBW_WARPS]; c += SBW_WARPS * 32u) {
        const uint32_t e = c + lane;
        if (e < pre[SBW_WARPS]) {
            int w = 0;
#pragma unroll
            for (int k = 1; k < SBW_WARPS; ++k) w += e >= pre[k] ? 1 : 0;
            transform_entry_i(sm.coef[w], sm.id[w], ring_slot<RING>(sm.left_head[w] + (e - pre[w])), job, pl, deq);
        }
    }
}


// -------------------------------------------------------------------------------------------------
// decode-P with cp.async-staged predictors (the default for P frames; src/common.rs:498-521 -> :254-285)
// -------------------------------------------------------------------------------------------------
// ncu on decode_sbw_kernel<true> showed the motion-compensated fetch done as 24 scalar, arbitrarily aligned
// LDG.32 per lane is bound by L1 wavefronts, not by bytes.  Here a warp stages the predictor of each of its 8
// macroblocks with ONE warp-wide 16-byte cp.async (lane = row*2 + half: the two aligned 16-byte chunks that
// cover the row's 16 unaligned bytes; get_block, src/common.rs:327-339) into a double-buffered shared tile,
// one tile ahead of use, and lanes then take their unaligned 8 bytes per row from shared memory.
//   A. skipped macroblocks (src/common.rs:281-283) store the copy; coded sub-blocks are appended to the warp's
//      ring: predictor (64 B), id, and their 128 B of coefficients fetched by cp.async straight into the slot;
//   B. whenever 32 coded sub-blocks are queued: DC-only ones apply one clamped delta, the others run the full
//      transform + residual (src/common.rs:98-104).
constexpr int PST_MB_STRIDE = 16 * 32 + 16;                   // bytes: 16 rows x 32 B, +16 so macroblocks start 4 banks apart
struct __align__(16) PWarpSmem {
    uint4    ref[2][8 * PST_MB_STRIDE / 16];                  // double-buffered predictor tile
    uint4    coef[SBW_RING * 8];                              // slot s keeps chunk k at [s*8 + (k ^ (s & 7))]
    uint4    prev[SBW_RING * 4];                              // rows 2k,2k+1 at [s*4 + (k ^ ((s >> 1) & 3))]
    uint32_t id[SBW_RING];
};
struct __align__(16) PStreamSmem {
    PWarpSmem w[SBW_WARPS];
    uint32_t  left_head[SBW_WARPS], left_cnt[SBW_WARPS];
};

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 8 bytes at byte offset `off` of a shared-memory row (off + 8 <= 32)
__device__ __forceinline__ uint2 lds_u8x8(const unsigned char *row, uint32_t off)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(row + (off & ~3u));
    const uint32_t sh = (off & 3u) * 8u;
    const uint32_t w0 = w[0], w1 = w[1], w2 = w[2];             // off <= 23, so all three words lie inside the 32-byte row
    uint2 r;
    r.x = __funnelshift_r(w0, w1, sh);
    r.y = __funnelshift_r(w1, w2, sh);
    return r;
}

__device__ __forceinline__ void transform_entry_p(PWarpSmem &ws, uint32_t slot, const DecJob &job, const PlaneGeom &pl,
                                                  const int32_t *deq)
{
    uint4 r2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r2[k] = ws.coef[slot * 8u + ((uint32_t)k ^ (slot & 7u))];
    const uint32_t id = ws.id[slot];
    const SbWhere w = sb_where<false>(job, pl, id >> 2, (int)(id & 3u), 0u);
    uint32_t ac = r2[0].x & 0xffff0000u;
    ac |= r2[0].y | r2[0].z | r2[0].w;
#pragma unroll
    for (int k = 1; k < 8; ++k) ac |= r2[k].x | r2[k].y | r2[k].z | r2[k].w;
    if (ac == 0u) {                                           // DC only: one delta for the whole sub-block
        const int c0 = (int)(int16_t)(r2[0].x & 0xffffu);
        const int v = (c0 * deq[0] + (128 << 8)) >> 8;
        const int delta = (min(max(v, 0), 255) - 128) * 2;    // src/common.rs:101
        const uint32_t pos4 = (uint32_t)max(delta, 0) * 0x01010101u;
        const uint32_t neg4 = (uint32_t)min(max(-delta, 0), 255) * 0x01010101u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 pv = ws.prev[slot * 4u + ((uint32_t)k ^ ((slot >> 1) & 3u))];
            __stcg(reinterpret_cast<uint2 *>(w.dst + (size_t)(2 * k) * pl.pw),
                   make_uint2(add_delta_sat4(pv.x, pos4, neg4), add_delta_sat4(pv.y, pos4, neg4)));
            __stcg(reinterpret_cast<uint2 *>(w.dst + (size_t)(2 * k + 1) * pl.pw),
                   make_uint2(add_delta_sat4(pv.z, pos4, neg4), add_delta_sat4(pv.w, pos4, neg4)));
        }
        return;
    }
    int m[64];
    unpack_dequant(r2, deq, m);
    idct8x8_regs(m);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint4 pv = ws.prev[slot * 4u + ((uint32_t)k ^ ((slot >> 1) & 3u))];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = 2 * k + h;
            int y[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) y[i] = m[r * 8 + i];
            const uint2 o = apply_residual_row(y, h == 0 ? make_uint2(pv.x, pv.y) : make_uint2(pv.z, pv.w));   // src/common.rs:277
            __stcg(reinterpret_cast<uint2 *>(w.dst + (size_t)r * pl.pw), o);
        }
    }
}

__global__ void __launch_bounds__(SBW_WARPS * 32, 2)
decode_p_stream_kernel(const __grid_constant__ SbParams P, const DecJob *__restrict__ jobs, int *__restrict__ err)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PStreamSmem &sm = *reinterpret_cast<PStreamSmem *>(smem_raw);

    const uint32_t cta = blockIdx.x;
    const int p = (cta >= P.cta_base[1] ? 1 : 0) + (cta >= P.cta_base[2] ? 1 : 0);
    const PlaneGeom &pl = p == 0 ? P.g.pl[0] : (p == 1 ? P.g.pl[1] : P.g.pl[2]);
    const int32_t *deq = p == 0 ? P.deq[0] : (p == 1 ? P.deq[1] : P.deq[2]);
    const uint32_t nmb = pl.bw * pl.bh, ntiles = (nmb + 7u) / 8u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tile_begin = ((cta - (p == 0 ? P.cta_base[0] : (p == 1 ? P.cta_base[1] : P.cta_base[2]))) * SBW_WARPS + warp) * P.tiles_per_warp;
    const uint32_t tile_end = min(tile_begin + P.tiles_per_warp, ntiles);
    const DecJob job = jobs[blockIdx.y];
    const int sb = (int)(lane & 3u);
    PWarpSmem &ws = sm.w[warp];
    const uint32_t *hdr32 = reinterpret_cast<const uint32_t *>(job.hdr) + pl.mb_base;
    const uint8_t *ref_plane = job.ref + pl.off;

    // Per-lane view of "its" macroblock of a tile: header word -> 16-byte aligned source of the predictor rows
    // and the byte offset of the block inside the 32-byte staged rows.
    struct Mc { const uint8_t *src; uint32_t off; };
    auto locate = [&](uint32_t lm, uint32_t hw) {
        uint32_t col;
        const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
        const int bx = (int)col * 16, by = (int)row * 16;
        int sx = bx + (int)(int8_t)(hw & 0xffu), sy = by + (int)(int8_t)((hw >> 8) & 0xffu);   // src/common.rs:255-256
        if (sx < 0 || sy < 0 || sx > (int)pl.pw - 16 || sy > (int)pl.ph - 16) {
            atomicOr(err, ERRBIT_BAD_MV);                     // src/common.rs:258-259: never read out of bounds
            sx = bx;
            sy = by;
        }
        Mc m;
        m.off = (uint32_t)sx & 15u;
        m.src = ref_plane + (size_t)sy * pl.pw + ((uint32_t)sx & ~15u);
        return m;
    };
    // all lanes: stage the predictors of tile `tile` (headers in hw, one per lane's macroblock) into buffer b
    auto stage_tile = [&](uint32_t tile, uint32_t hw, int b) -> uint32_t {
        const uint32_t lm = tile * 8u + (lane >> 2);
        Mc mc = {ref_plane, 0u};
        if (lm < nmb) mc = locate(lm, hw);
        const unsigned long long sp = reinterpret_cast<unsigned long long>(mc.src);
        const uint32_t row = lane >> 1, half = lane & 1u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const unsigned long long pj = __shfl_sync(0xffffffffu, sp, 4 * j);
            if (tile * 8u + (uint32_t)j < nmb)
                cp_async16(reinterpret_cast<unsigned char *>(ws.ref[b]) + j * PST_MB_STRIDE + row * 32u + half * 16u,
                           reinterpret_cast<const uint8_t *>(pj) + (size_t)row * pl.pw + half * 16u);
        }
        cp_async_commit();
        return mc.off;
    };

    uint32_t head = 0, tail = 0;
    uint32_t hw_cur = 0, hw_next = 0, off_cur = 0, off_next = 0;
    if (tile_begin < tile_end) {
        if (tile_begin * 8u + (lane >> 2) < nmb) hw_cur = __ldg(hdr32 + tile_begin * 8u + (lane >> 2));
        if (tile_begin + 1 < tile_end && (tile_begin + 1) * 8u + (lane >> 2) < nmb)
            hw_next = __ldg(hdr32 + (tile_begin + 1) * 8u + (lane >> 2));
        off_cur = stage_tile(tile_begin, hw_cur, 0);
        cp_async_commit();                                    // empty group: keeps the R(t), C(t-1), R(t+1) commit pattern uniform
    }

#pragma unroll 1
    for (uint32_t tile = tile_begin; tile < tile_end; ++tile) {
        const int b = (int)((tile - tile_begin) & 1u);
        const uint32_t lm = tile * 8u + (lane >> 2);
        const bool valid = lm < nmb;
        uint32_t hw_next2 = 0;
        if (tile + 1 < tile_end) {
            off_next = stage_tile(tile + 1, hw_next, b ^ 1);                       // one tile ahead
            if (tile + 2 < tile_end && lm + 16u < nmb) hw_next2 = __ldg(hdr32 + lm + 16u);
            cp_async_wait<2>();       // commit order is R(t), C(t-1), R(t+1): R(t) must have landed, the two newer may be in flight
        } else {
            cp_async_wait<1>();       // R(t), C(t-1)
        }
        __syncwarp();

        // ---- A ----
        const bool coded = valid && ((hw_cur >> 16) & 0xffu) != 0u;
        uint2 prev[8];
        if (valid) {
            const unsigned char *rows = reinterpret_cast<const unsigned char *>(ws.ref[b]) + (lane >> 2) * PST_MB_STRIDE +
                                        (uint32_t)(sb >> 1) * 8u * 32u;
            const uint32_t off = off_cur + (uint32_t)(sb & 1) * 8u;
#pragma unroll
            for (int r = 0; r < 8; ++r) prev[r] = lds_u8x8(rows + r * 32, off);
        }
        const uint32_t vote = __ballot_sync(0xffffffffu, coded);
        if (coded) {
            const uint32_t slot = (tail + (uint32_t)__popc(vote & ((1u << lane) - 1u))) & (SBW_RING - 1);
            const uint4 *src = reinterpret_cast<const uint4 *>(job.coeff + ((size_t)(pl.mb_base + lm) * 256 + sb * 64));
#pragma unroll
            for (int k = 0; k < 8; ++k) cp_async16(&ws.coef[slot * 8u + ((uint32_t)k ^ (slot & 7u))], src + k);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                ws.prev[slot * 4u + ((uint32_t)k ^ ((slot >> 1) & 3u))] =
                    make_uint4(prev[2 * k].x, prev[2 * k].y, prev[2 * k + 1].x, prev[2 * k + 1].y);
            ws.id[slot] = (lm << 2) | (uint32_t)sb;
        } else if (valid) {                                                          // skipped: src/common.rs:281-283
            uint32_t col;
            const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
            uint8_t *dst = job.dst + pl.off + (size_t)(row * 16u + (uint32_t)(sb >> 1) * 8u) * pl.pw + col * 16u + (uint32_t)(sb & 1) * 8u;
#pragma unroll
            for (int r = 0; r < 8; ++r) __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pl.pw), prev[r]);
        }
        cp_async_commit();
        tail += (uint32_t)__popc(vote);
        hw_cur = hw_next; hw_next = hw_next2; off_cur = off_next;

        // ---- B ----
        if (tail - head >= 32u) {
            cp_async_wait<0>();
            __syncwarp();
            transform_entry_p(ws, (head + lane) & (SBW_RING - 1), job, pl, deq);
            head += 32u;
        }
        __syncwarp();
    }

    // ---- flush ----
    cp_async_wait<0>();
    if (lane == 0) { sm.left_head[warp] = head; sm.left_cnt[warp] = tail - head; }
    __syncthreads();
    uint32_t pre[SBW_WARPS + 1];
    pre[0] = 0;
#pragma unroll
    for (int w = 0; w < SBW_WARPS; ++w) pre[w + 1] = pre[w] + sm.left_cnt[w];
#pragma unroll 1
    for (uint32_t c = warp * 32u; c < pre[SBW_WARPS]; c += SBW_WARPS * 32u) {
        const uint32_t e = c + lane;
        if (e < pre[SBW_WARPS]) {
            int w = 0;
#pragma unroll
            for (int k = 1; k < SBW_WARPS; ++k) w += e >= pre[k] ? 1 : 0;
            transform_entry_p(sm.w[w], (sm.left_head[w] + (e - pre[w])) & (SBW_RING - 1), job, pl, deq);
        }
    }
}

cudaError_t launch_decode_i_sb(const SbParams &P, const DecJob *d_jobs, uint32_t njobs, cudaStream_t s)
{
    dim3 grid(P.cta_total, njobs, 1), block(SB_WARPS * 32, 1, 1);
    decode_i_sb_kernel<<<grid, block, 0, s>>>(P, d_jobs);
    return cudaGetLastError();
}

// `P` arrives with g and deq filled; the work split is completed here.  tiles_per_warp trades the per-warp
// remainder (one partly filled transform pass per warp at the end) against having enough warps to fill the
// chip: aim at ~2 waves of the 148 SMs x 16 resident warps, at most 16 tiles per warp.
static void sbw_split(SbParams &P, uint32_t njobs, uint32_t waves_x_warps, uint32_t max_tpw)
{
    uint32_t tiles = 0;
    for (int p = 0; p < 3; p++) tiles += (P.g.pl[p].bw * P.g.pl[p].bh + 7u) / 8u;
    const uint64_t total = (uint64_t)tiles * njobs;
    uint32_t tpw = (uint32_t)(total / waves_x_warps);
    tpw = tpw < 1 ? 1 : (tpw > max_tpw ? max_tpw : tpw);
    static const int tpw_env = getenv("PFV_TILES_PER_WARP") ? atoi(getenv("PFV_TILES_PER_WARP")) : 0;   // tuning aid
    if (tpw_env > 0) tpw = (uint32_t)tpw_env;
    P.tiles_per_warp = tpw;
    uint32_t cta = 0;
    for (int p = 0; p < 3; p++) {
        P.cta_base[p] = cta;
        const uint32_t ntiles = (P.g.pl[p].bw * P.g.pl[p].bh + 7u) / 8u;
        cta += (ntiles + SBW_WARPS * tpw - 1) / (SBW_WARPS * tpw);
    }
    P.cta_total = cta;
}

cudaError_t launch_decode_sbw(bool inter, SbParams P, const DecJob *d_jobs, uint32_t njobs, int *d_err, cudaStream_t s)
{
    sbw_split(P, njobs, 2u * 148u * 16u, 16u);
    dim3 grid(P.cta_total, njobs, 1), block(SBW_WARPS * 32, 1, 1);
    if (inter) decode_sbw_kernel<true><<<grid, block, 0, s>>>(P, d_jobs, d_err);
    else       decode_sbw_kernel<false><<<grid, block, 0, s>>>(P, d_jobs, d_err);
    return cudaGetLastError();
}

// Job coefficient pointers must be 16-byte aligned (bulk copies); pfv_decode_submit checks.
template <int STAGES, int RING>
static cudaError_t launch_decode_i_stream_t(SbParams P, const DecJob *d_jobs, uint32_t njobs, cudaStream_t s)
{
    static bool attr_done[64] = {};           // cudaFuncSetAttribute is per DEVICE: one process may own contexts on several
    const int smem = (int)sizeof(StreamSmemT<STAGES, RING>);
    if (first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(decode_i_stream_kernel<STAGES, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    sbw_split(P, njobs, 6u * 148u * 12u, 16u);               // ~6 waves of 148 SMs x 12 resident warps
    dim3 grid(P.cta_total, njobs, 1), block(SBW_WARPS * 32, 1, 1);
    decode_i_stream_kernel<STAGES, RING><<<grid, block, smem, s>>>(P, d_jobs);
    return cudaGetLastError();
}

cudaError_t launch_decode_i_stream(SbParams P, const DecJob *d_jobs, uint32_t njobs, cudaStream_t s)
{
    static const int stages_env = getenv("PFV_DECODE_I_STAGES") ? atoi(getenv("PFV_DECODE_I_STAGES")) : 2;
    if (stages_env == 3) return launch_decode_i_stream_t<3, 48>(P, d_jobs, njobs, s);
    if (stages_env == 4) return launch_decode_i_stream_t<3, 40>(P, d_jobs, njobs, s);
    return launch_decode_i_stream_t<2, 64>(P, d_jobs, njobs, s);
}

cudaError_t launch_decode_p_stream(SbParams P, const DecJob *d_jobs, uint32_t njobs, int *d_err, cudaStream_t s)
{
    static bool attr_done[64] = {};           // cudaFuncSetAttribute is per DEVICE: one process may own contexts on several
    if (first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(decode_p_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)sizeof(PStreamSmem));
        if (e != cudaSuccess) return e;
    }
    sbw_split(P, njobs, 6u * 148u * 8u, 16u);                 // ~6 waves of 148 SMs x 8 resident warps
    dim3 grid(P.cta_total, njobs, 1), block(SBW_WARPS * 32, 1, 1);
    decode_p_stream_kernel<<<grid, block, sizeof(PStreamSmem), s>>>(P, d_jobs, d_err);
    return cudaGetLastError();
}

}  // namespace pfv
