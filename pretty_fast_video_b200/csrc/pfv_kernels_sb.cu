// pfv_kernels_sb.cu — "register-resident sub-block" kernels (sm_100a).
//
// One THREAD owns one 8x8 sub-block: its 64 coefficients live in registers, both 1-D passes run without any
// exchange, the zig-zag permutation is a compile-time register renaming and the (de)quantiser tables are
// kernel parameters, i.e. constant-bank operands of the multiplies.  A warp covers 8 consecutive macroblocks
// (lane = mb*4 + sub-block), so one 8-byte row store of the warp forms two full 128-byte lines of the plane.
// Compared with the warp-per-macroblock kernels of pfv_kernels.cu this removes every shared-memory round trip
// and amortises the addressing over 8 macroblocks (ncu: 255 -> ~150 warp instructions per macroblock).
//
//   decode_i_sb_kernel      the plain form: load, transform, store (PFV_DECODE_I_VARIANT=sb; second implementation)
//   decode_i_stream_kernel  the default for key frames: TMA-staged tiles, classify / compact / transform
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
// decode-I for frames the caller flags as dense (PFV_JOB_DENSE): every sub-block takes the full transform, so there is
// nothing to classify - but the plain kernel above stalls on its eight coefficient loads at the start of every CTA (ncu:
// long-scoreboard 1.8 of 4.4 stall cycles per issue, 13 of 16 warps resident on average because CTAs are short-lived).
// Here a warp walks tiles of 8 macroblocks; the NEXT tile's 4 KB are copied global -> shared by cp.async (16 bytes per lane
// and instruction, no registers held) while this one is transformed.  Chunk k of sub-block L lands at
// L*128 + ((k ^ (L & 7)) << 4): the copy instructions write and the transforming lanes read 16-byte columns that differ
// inside every quarter-warp, i.e. both sides are bank-conflict free, and lane L finds chunk k in the SAME register
// whatever L is (a lane-dependent rotation would need the ring of the streaming kernel to undo it).
// -------------------------------------------------------------------------------------------------
constexpr int DIR_WARPS = 4;
void sbw_split(SbParams &P, uint32_t njobs, uint32_t warps_per_cta, uint32_t resident_ctas, uint32_t max_tpw, uint32_t start_cost, uint32_t forced_tpw);

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst_smem)), "l"(src) : "memory");
}

__global__ void __launch_bounds__(DIR_WARPS * 32, 4)
decode_i_direct_kernel(const __grid_constant__ SbParams P, const DecJob *__restrict__ jobs)
{
    __shared__ uint4 stage[DIR_WARPS][2][256];                  // per warp: two tiles of 8 macroblocks x 512 B
    const uint32_t cta = blockIdx.x;
    const int p = (cta >= P.cta_base[1] ? 1 : 0) + (cta >= P.cta_base[2] ? 1 : 0);
    const PlaneGeom &pl = p == 0 ? P.g.pl[0] : (p == 1 ? P.g.pl[1] : P.g.pl[2]);
    const int32_t *deq = p == 0 ? P.deq[0] : (p == 1 ? P.deq[1] : P.deq[2]);
    const uint32_t nmb = pl.bw * pl.bh, ntiles = (nmb + 7u) / 8u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tile_begin = ((cta - (p == 0 ? P.cta_base[0] : (p == 1 ? P.cta_base[1] : P.cta_base[2]))) * DIR_WARPS + warp) * P.tiles_per_warp;
    const uint32_t tile_end = min(tile_begin + P.tiles_per_warp, ntiles);
    if (tile_begin >= tile_end) return;
    const DecJob job = jobs[blockIdx.y];
    const uint4 *plane_coeff = reinterpret_cast<const uint4 *>(job.coeff + (size_t)pl.mb_base * 256);
    // copy side: chunk c = j*32 + lane of a tile is chunk (lane & 7) of sub-block j*4 + (lane >> 3)
    const uint32_t wr = (lane >> 3) * 8u + ((lane & 7u) ^ (lane >> 3));       // + j*32: (j*4 + (lane >> 3)) & 7 == ((lane >> 3) + 4*(j & 1)) & 7
    auto issue = [&](uint32_t tile) {
        uint4 *stg = stage[warp][tile & 1u];
        const uint4 *src = plane_coeff + (size_t)tile * 256u + lane;
        const uint32_t mbs = min(8u, nmb - tile * 8u);
#pragma unroll
        for (uint32_t j = 0; j < 8u; ++j)
            if (j < mbs) cp_async16(stg + j * 32u + (wr ^ ((j & 1u) << 2)), src + j * 32u);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(tile_begin);
#pragma unroll 1
    for (uint32_t tile = tile_begin; tile != tile_end; ++tile) {
        if (tile + 1u != tile_end) {
            issue(tile + 1u);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();                                           // every lane's copies of this tile have landed
        const uint4 *stg = stage[warp][tile & 1u];
        uint4 raw[8];
#pragma unroll
        for (uint32_t k = 0; k < 8u; ++k) raw[k] = stg[lane * 8u + (k ^ (lane & 7u))];
        const uint32_t lm = tile * 8u + (lane >> 2);
        if (lm < nmb) {
            int m[64];
            unpack_dequant(raw, deq, m);
            idct8x8_regs(m);
            uint8_t *dst = sb_dst(job.dst, pl, lm, (int)(lane & 3u));
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                uint2 o;
                o.x = pack4_sat_u8(m[r * 8 + 0], m[r * 8 + 1], m[r * 8 + 2], m[r * 8 + 3]);
                o.y = pack4_sat_u8(m[r * 8 + 4], m[r * 8 + 5], m[r * 8 + 6], m[r * 8 + 7]);
                __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pl.pw), o);
            }
        }
        __syncwarp();                                           // the stage is refilled two iterations from now
    }
}

// -------------------------------------------------------------------------------------------------
// decode-I, "classify, compact, transform" with TMA-staged tiles (the default for key frames)
// -------------------------------------------------------------------------------------------------
// The exact integer IDCT costs ~1300 instructions per sub-block, which at 48 960 sub-blocks per 1080p frame is
// more issue time than the frame's HBM time.  Real streams are sparse: most sub-blocks carry only a DC
// coefficient, and for those both passes collapse exactly (src/dct.rs:241-293 with v[1..7] = 0 returns v[0] in
// every output; all-zero columns stay zero): every pixel is clamp((c0 * deq0 + 32768) >> 8).
//
// Each WARP streams over `tiles_per_warp` consecutive tiles of 8 macroblocks (lane = macroblock*4 + sub-block).
// The 4 KB coefficient tile (8 macroblocks x 512 B, contiguous in the dense layout) is fetched by ONE bulk async
// copy (cp.async.bulk, the TMA unit) into a per-warp double buffer, completion signalled on a per-stage mbarrier:
// while a warp classifies or transforms tile i, tiles i+1 and i+2 are in flight without holding registers or LSU
// slots.
//   A. take the tile out of the stage, finish the sub-blocks that need no transform on the spot (DC-only) and
//      append the others to the warp's private shared-memory ring (ballot + prefix, no atomics, no CTA barrier);
//   B. whenever the ring holds 32 entries, run the full register-resident transform on them: a full warp.
// What is left at the end (< 32 entries per warp) is pooled across the CTA's four warps and flushed.
// Lanes read their 128 B from the stage with the chunk order rotated by (lane & 7) so the 128-byte-strided reads
// are bank-conflict free; the rotation is undone by address arithmetic when an entry is queued.
// Results do not depend on the path taken; tests/test_gpu_parity.py drives dense, sparse and mixed inputs.
constexpr int STG_TILE_BYTES = 8 * 512;
constexpr int STG_STAGES = 2;

template <int WARPS>
struct __align__(128) StreamSmem {
    uint4    stage[WARPS][STG_STAGES][STG_TILE_BYTES / 16];
    uint4    coef[WARPS][SBW_RING * 8];
    uint32_t id[WARPS][SBW_RING];
    uint64_t bar[WARPS][STG_STAGES];
    uint32_t left_head[WARPS], left_cnt[WARPS];
};

// WARPS x CTAS = resident warps per SM the kernel is compiled for.  Each warp owns 16 KB of shared memory (two stages
// + the ring), so at most 14 warps fit an SM; the transform is issue bound with 4-cycle dependent chains, so resident
// warps count as long as the 64-value sub-block stays in registers.
// POOL: what the warps have left at the end is pooled across the CTA and transformed by full warps (a second copy of
// the transform behind a CTA barrier); otherwise each warp drains its own ring through the loop's one transform site
// with the idle lanes masked off (half the code, no barrier, a partly filled last pass per warp).
template <int WARPS, int CTAS, bool POOL>
__global__ void __launch_bounds__(WARPS * 32, CTAS)
decode_i_stream_kernel(const __grid_constant__ SbParams P, const DecJob *__restrict__ jobs)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StreamSmem<WARPS> &sm = *reinterpret_cast<StreamSmem<WARPS> *>(smem_raw);

    const uint32_t cta = blockIdx.x;
    const int p = (cta >= P.cta_base[1] ? 1 : 0) + (cta >= P.cta_base[2] ? 1 : 0);
    const PlaneGeom &pl = p == 0 ? P.g.pl[0] : (p == 1 ? P.g.pl[1] : P.g.pl[2]);
    const int32_t *deq = p == 0 ? P.deq[0] : (p == 1 ? P.deq[1] : P.deq[2]);
    const uint32_t nmb = pl.bw * pl.bh, ntiles = (nmb + 7u) / 8u;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t tile_begin = ((cta - (p == 0 ? P.cta_base[0] : (p == 1 ? P.cta_base[1] : P.cta_base[2]))) * WARPS + warp) * P.tiles_per_warp;
    const uint32_t tile_end = min(tile_begin + P.tiles_per_warp, ntiles);
    const uint32_t ntl = tile_end > tile_begin ? tile_end - tile_begin : 0u;
    const DecJob job = jobs[blockIdx.y];
    const int sb = (int)(lane & 3u);
    const uint32_t rot = lane & 7u;
    uint4 *ring = sm.coef[warp];
    uint32_t *ring_id = sm.id[warp];
    const char *plane_coeff = reinterpret_cast<const char *>(job.coeff + (size_t)pl.mb_base * 256);

    auto issue = [&](uint32_t i) {                            // lane 0 only: tile (tile_begin + i) into stage i % STG_STAGES
        const uint32_t tile = tile_begin + i;
        const uint32_t mbs = min(8u, nmb - tile * 8u);
        bulk_load(sm.stage[warp][i % STG_STAGES], plane_coeff + (size_t)tile * STG_TILE_BYTES, mbs * 512u,
                  &sm.bar[warp][i % STG_STAGES]);
    };
    if (lane == 0) {
#pragma unroll
        for (int s2 = 0; s2 < STG_STAGES; ++s2) bar_init(&sm.bar[warp][s2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#pragma unroll
        for (uint32_t i = 0; i < (uint32_t)STG_STAGES; ++i)
            if (i < ntl) issue(i);
    }
    __syncwarp();

    uint32_t head = 0, tail = 0;
#pragma unroll 1
    for (uint32_t i = 0; i < ntl; ++i) {
        const uint32_t tile = tile_begin + i;
        const uint32_t lm = tile * 8u + (lane >> 2);
        const bool valid = lm < nmb;
        const uint4 *stg = sm.stage[warp][i % STG_STAGES];

        // ---- A: take this lane's 128 B out of the stage (chunk k ^ rot lands in raw[k]) ----
        bar_wait(&sm.bar[warp][i % STG_STAGES], (i / STG_STAGES) & 1u);
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
        if (lane == 0 && i + STG_STAGES < ntl) issue(i + STG_STAGES);   // the stage may be refilled
        if (general) {
            const uint32_t slot = (tail + (uint32_t)__popc(vote & ((1u << lane) - 1u))) & (SBW_RING - 1);
            const uint32_t y = rot ^ (slot & 7u);             // raw[k] is chunk k ^ rot; the ring keeps chunk j at j ^ (slot & 7)
#pragma unroll
            for (int k = 0; k < 8; ++k) ring[slot * 8u + ((uint32_t)k ^ y)] = raw[k];
            ring_id[slot] = (lm << 2) | (uint32_t)sb;
        } else if (valid) {
            store_dc_only(sb_dst(job.dst, pl, lm, sb), pl.pw, (int)(int16_t)(w0 & 0xffffu), deq[0]);
        }
        tail += (uint32_t)__popc(vote);
        __syncwarp();

        // ---- B: a full warp of queued sub-blocks ----
        const bool drain = !POOL && i + 1u == ntl;
#pragma unroll 1
        while (tail - head >= 32u || (drain && tail != head)) {
            if (lane < tail - head) transform_entry_i(ring, ring_id, (head + lane) & (SBW_RING - 1), job.dst, pl, deq);
            head += min(32u, tail - head);
            __syncwarp();
        }
    }
    if (POOL) flush_rings_i<WARPS>(sm.coef, sm.id, sm.left_head, sm.left_cnt, head, tail, warp, lane, job.dst, pl, deq);
}

cudaError_t launch_decode_i_sb(const SbParams &P, const DecJob *d_jobs, uint32_t njobs, cudaStream_t s)
{
    dim3 grid(P.cta_total, njobs, 1), block(SB_WARPS * 32, 1, 1);
    decode_i_sb_kernel<<<grid, block, 0, s>>>(P, d_jobs);
    return cudaGetLastError();
}

cudaError_t launch_decode_i_direct(SbParams P, const DecJob *d_jobs, uint32_t njobs, cudaStream_t s)
{
    static const int tpw_env = getenv("PFV_DECODE_I_TPW") ? atoi(getenv("PFV_DECODE_I_TPW")) : 0;   // tuning aid: tiles per warp
    sbw_split(P, njobs, DIR_WARPS, 148u * 4u, 16u, 0u, tpw_env > 0 ? (uint32_t)tpw_env : 0u);   // (__launch_bounds__: 4 CTAs per SM)
    dim3 grid(P.cta_total, njobs, 1), block(DIR_WARPS * 32, 1, 1);
    decode_i_direct_kernel<<<grid, block, 0, s>>>(P, d_jobs);
    return cudaGetLastError();
}

// `P` arrives with g and deq filled; the work split is completed here.  A warp walks `tiles_per_warp` consecutive tiles, a CTA
// never straddles a plane; `resident_ctas` of them run at a time.  Every CTA lives about as long as its warps have tiles, so the
// launch takes ceil(CTAs / resident_ctas) x (tiles_per_warp + start_cost) tile times (start_cost: what a CTA's start is worth in tile
// times - about one for the streaming kernel, whose bulk-copy pipeline has to fill, nothing measurable for the direct one): the
// split takes the tiles_per_warp (<= max_tpw) with the smallest product, the larger one among equals (fewer CTA starts, fewer partly filled transform passes at the warps' ends).
// The first form aimed at "about six waves" (tiles_per_warp = tiles / (6 x resident warps), rounded down) and for 64 x 1080p got
// 9 tiles per warp = 2 880 CTAs = 6.49 waves of 444: the seventh wave was half empty and its planes' last CTAs nearly so (ncu:
// sm__cycles_active avg 169 k of 191 k elapsed).  8 tiles per warp are 3 072 CTAs = 6.92 waves of full CTAs.  Measured (visit zj,
// 64 x 1080p per launch, frames/s by tiles per warp): config-2 stream 4: 593 k, 8: 625 k, 9: 618 k, 10: 612 k, 12: 603 k, 15: 593 k;
// all-dense frames through the direct kernel 3: 480 k, 5: 480 k, 9: 470 k, 11: 452 k, 16: 465 k.
void sbw_split(SbParams &P, uint32_t njobs, uint32_t warps_per_cta, uint32_t resident_ctas, uint32_t max_tpw, uint32_t start_cost, uint32_t forced_tpw)
{
    uint32_t ntiles[3];
    for (int p = 0; p < 3; p++) ntiles[p] = (P.g.pl[p].bw * P.g.pl[p].bh + 7u) / 8u;
    auto ctas_per_frame = [&](uint32_t tpw) {
        uint32_t n = 0;
        for (int p = 0; p < 3; p++) n += (ntiles[p] + warps_per_cta * tpw - 1) / (warps_per_cta * tpw);
        return n;
    };
    uint32_t best = 1;
    uint64_t best_cost = ~0ull;
    for (uint32_t tpw = max_tpw; tpw >= 1; --tpw) {
        const uint64_t ctas = (uint64_t)ctas_per_frame(tpw) * njobs;
        const uint64_t cost = ((ctas + resident_ctas - 1) / resident_ctas) * (tpw + start_cost);
        if (cost < best_cost) { best_cost = cost; best = tpw; }
    }
    if (forced_tpw >= 1 && forced_tpw <= max_tpw) best = forced_tpw;
    P.tiles_per_warp = best;
    uint32_t cta = 0;
    for (int p = 0; p < 3; p++) {
        P.cta_base[p] = cta;
        cta += (ntiles[p] + warps_per_cta * best - 1) / (warps_per_cta * best);
    }
    P.cta_total = cta;
}

// Job coefficient pointers must be 16-byte aligned (bulk copies); pfv_decode_submit checks.
template <int WARPS, int CTAS, bool POOL>
static cudaError_t launch_decode_i_stream_t(SbParams P, const DecJob *d_jobs, uint32_t njobs, cudaStream_t s)
{
    static bool attr_done[64] = {};           // cudaFuncSetAttribute is per DEVICE: one process may own contexts on several
    const int smem = (int)sizeof(StreamSmem<WARPS>);
    if (first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(decode_i_stream_kernel<WARPS, CTAS, POOL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    static const int tpw_env = getenv("PFV_DECODE_I_TPW") ? atoi(getenv("PFV_DECODE_I_TPW")) : 0;   // tuning aid: tiles per warp
    sbw_split(P, njobs, WARPS, 148u * (uint32_t)CTAS, 16u, 1u, tpw_env > 0 ? (uint32_t)tpw_env : 0u);
    dim3 grid(P.cta_total, njobs, 1), block(WARPS * 32, 1, 1);
    decode_i_stream_kernel<WARPS, CTAS, POOL><<<grid, block, smem, s>>>(P, d_jobs);
    return cudaGetLastError();
}

cudaError_t launch_decode_i_stream(SbParams P, const DecJob *d_jobs, uint32_t njobs, cudaStream_t s)
{
    // (CTAs of 6 or 7 warps, 2 per SM, measured slower: 0.80 / 0.61 of the roofline against 0.81 on the config-2 stream)
    // per-warp drain (POOL = false) measured 577 k frames/s on the config-2 stream against 550-562 k with the pooled flush
    return launch_decode_i_stream_t<4, 3, false>(P, d_jobs, njobs, s);
}

}  // namespace pfv
