// pfv_kernels_tok.cu — device side of the sparse ENCODE transport (SURVEY §8 f2 "on encode" + f4 "GPU-side RLE", sm_100a).
//
// The reference's entropy stage walks every macroblock's 256 dense coefficients on the host (rle_encode,
// src/rle.rs:9-39, called per macroblock from src/enc.rs:256-262 / :370-378), counts the symbols (update_table,
// src/rle.rs:41-47) and only then writes bits.  A dense 1080p frame is 6.27 MB of i16 over PCIe and through a host
// scan; >90 % of it is zeros.  Here the run-length pass runs on the device and only the RLE sequence crosses PCIe:
//
// P frames (few macroblocks carry coefficients):
//   (encode-P kernel)  counts the entries of each coded macroblock while its coefficients are still in registers (pfv_tok.cuh)
//   tok_scan_kernel    one CTA per frame: exclusive prefix sum -> mb_off[nb+1]; clears the frame's statistics
//   tok_emit_kernel    a warp takes 8 consecutive macroblocks, skips the ones without coefficients by one ballot over their
//                      headers, compacts the non-zero coefficients of the others and writes their entries at their final
//                      offsets, in stream order, + the two symbol histograms
//   tok_store_kernel   copies exactly `ntok` entries (+ statistics, + mb_off on request) to the caller's buffers with
//                      16-byte stores; the destination may be pinned HOST memory (zero-copy over PCIe), so the
//                      transfer size follows the data without a host round trip
//
// Key frames (every macroblock carries coefficients, ~32 entries each at quality 5): the macroblock-at-a-time walk above costs
// 390 warp instructions per macroblock (ncu: 170 us per 32 x 1080p, issue 83 %, ALU pipe 69 % - instruction bound) and the
// counting inside the encode kernel another 40 us.  Instead:
//   tok_emit_sb_kernel one THREAD per 8x8 sub-block, a warp = 8 macroblocks: 64-bit non-zero mask, entry counts and the
//                      lane's offset inside its macroblock from the mask (sb_runs, pfv_dct.cuh), then every lane walks ITS
//                      non-zeros and writes their entries into the macroblock's padded slot tok[m * 256 ..] (256 entries are
//                      the most a macroblock can make); leaves the count in mb_off[m + 1] and adds to the statistics
//   tok_scan_kernel    as above, without clearing the statistics
//   tok_store_kernel   gathers the slots to their final offsets on the way out (and clears the staging statistics)
//
// Entry format (what pfv_packet_encode_tokens takes): run | size << 4 | uint16(value) << 16, exactly one word per
// RLESequence {num_zeroes, coeff_size, coeff} (src/rle.rs:3-7).
#include <cstdlib>

#include "pfv_internal.h"
#include "pfv_tok.cuh"
#include "pfv_dct.cuh"

namespace pfv {

using tok::FULL;
using tok::escapes_of;

constexpr int TOK_WARPS = 2;            // no CTA-wide step: a warp that is done gives its slot back (see tok_emit_kernel)
constexpr uint32_t TOK_CHUNK = 8;       // macroblocks per warp

// mb_off[m+1] holds the count of macroblock m on entry, the inclusive sum on exit; mb_off[0] = 0.
constexpr int SCAN_THREADS = 1024, SCAN_ITEMS = 12;      // 12 288 counts per iteration: a 1080p frame (12 240) in one
template <bool ZERO_STATS>
__global__ void __launch_bounds__(SCAN_THREADS)
tok_scan_kernel(uint32_t nb, const TokJob *__restrict__ jobs)
{
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    __shared__ uint32_t carry_s;
    const TokJob job = jobs[blockIdx.x];
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    uint32_t *cnt = job.mb_off + 1;
    if (t == 0) { job.mb_off[0] = 0; carry_s = 0; }
    if (ZERO_STATS && t < PFV_TOKSTATS_WORDS) job.stats[t] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += SCAN_THREADS * SCAN_ITEMS) {
        const uint32_t i0 = base + t * SCAN_ITEMS;
        uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            v[k] = (i0 + k < nb) ? cnt[i0 + k] : 0u;
            sum += v[k];
            v[k] = sum;                                              // inclusive inside the thread
        }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(FULL, incl, d);
            if ((int)lane >= d) incl += u;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        const uint32_t carry = carry_s;
        if (warp == 0) {
            uint32_t w = wsum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(FULL, w, d);
                if ((int)lane >= d) w += u;
            }
            wsum[lane] = w;                                          // inclusive over warps
        }
        __syncthreads();
        const uint32_t before = carry + (warp ? wsum[warp - 1] : 0u) + (incl - sum);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (i0 + k < nb) cnt[i0 + k] = before + v[k];
        __syncthreads();
        if (t == 0) carry_s = carry + wsum[SCAN_THREADS / 32 - 1];
        __syncthreads();
    }
    if (t == 0) job.stats[PFV_TOKSTATS_NTOK] = carry_s;
}

// Symbol statistics without contended atomics.  Nearly every entry of a macroblock carries one of two or three symbols (run 0/1,
// size 2/3), so 32 lanes adding to one shared-memory histogram serialise 32-fold.  Instead a lane counts its own entries of one
// macroblock in two 64-bit registers (16 bins x 4 bits: a lane makes at most 9 entries per macroblock), folds them after each
// macroblock into four registers of 8 x 8 bits, and the warp adds those up every <= 16 macroblocks by a reduce-scatter (16
// shuffles) that leaves word k of the 16 x (2 x 16 bit) totals in lanes 2k, 2k+1; 16 lanes then add two bins each to the
// frame's counters (only the bins that occurred: a handful per warp).
struct LaneHist {
    uint64_t acc[4];     // [0] run bins 0,2,..,14  [1] run bins 1,3,..,15  [2] size bins even  [3] size bins odd; 8 bits each
};

__device__ __forceinline__ void hist_fold(LaneHist &h, uint64_t r64, uint64_t s64)
{
    constexpr uint64_t M = 0x0f0f0f0f0f0f0f0full;
    h.acc[0] += r64 & M;
    h.acc[1] += (r64 >> 4) & M;
    h.acc[2] += s64 & M;
    h.acc[3] += (s64 >> 4) & M;
}

__device__ __forceinline__ void hist_flush(LaneHist &h, uint32_t lane, uint32_t *hist)
{
    uint32_t w[16];                                                  // word a*4+j: byte 2j of acc[a] | byte 2j+1 << 16
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const uint32_t lo = (uint32_t)h.acc[a], hi = (uint32_t)(h.acc[a] >> 32);
        w[a * 4 + 0] = __byte_perm(lo, 0u, 0x4140);
        w[a * 4 + 1] = __byte_perm(lo, 0u, 0x4342);
        w[a * 4 + 2] = __byte_perm(hi, 0u, 0x4140);
        w[a * 4 + 3] = __byte_perm(hi, 0u, 0x4342);
        h.acc[a] = 0;
    }
    uint32_t v8[8], v4[4], v2[2];
    const bool b4 = lane & 16u, b3 = lane & 8u, b2 = lane & 4u, b1 = lane & 2u;
#pragma unroll
    for (int i = 0; i < 8; ++i) v8[i] = (b4 ? w[8 + i] : w[i]) + __shfl_xor_sync(FULL, b4 ? w[i] : w[8 + i], 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) v4[i] = (b3 ? v8[4 + i] : v8[i]) + __shfl_xor_sync(FULL, b3 ? v8[i] : v8[4 + i], 8);
#pragma unroll
    for (int i = 0; i < 2; ++i) v2[i] = (b2 ? v4[2 + i] : v4[i]) + __shfl_xor_sync(FULL, b2 ? v4[i] : v4[2 + i], 4);
    uint32_t v = (b1 ? v2[1] : v2[0]) + __shfl_xor_sync(FULL, b1 ? v2[0] : v2[1], 2);
    v += __shfl_xor_sync(FULL, v, 1);
    if ((lane & 1u) == 0) {                                          // this lane pair holds word k = lane >> 1 = a*4 + j
        const uint32_t k = lane >> 1, a = k >> 2, j = k & 3u;
        const uint32_t bin = (a >= 2 ? 16u : 0u) + 4u * j + (a & 1u);   // byte 2j of acc[a] counts bin 2*(2j) + (a & 1)
        if (v & 0xffffu) atomicAdd(&hist[bin], v & 0xffffu);
        if (v >> 16) atomicAdd(&hist[bin + 2u], v >> 16);
    }
}

// One warp = `chunk` consecutive macroblocks; grid.x = chunks / TOK_WARPS, grid.y = frame.
//
// Per macroblock: (1) a lane keeps the non-zero ones of its 8 coefficients and the warp compacts them, in order, into a
// shared-memory list of (position, value) - one popc, one 5-step scan, <= 8 short predicated stores; (2) the list is walked 32
// entries at a time, one entry per lane: run = distance to the previous list entry, escapes = (run-1)/15, a second scan turns
// entries-per-lane into output offsets, and neighbouring lanes write neighbouring words of the sequence.  (The first version
// let every lane walk its own 8 coefficients with the whole entry logic inside 8 divergent steps: ~800 warp instructions per
// macroblock, 136 us per 32 P frames; ncu, profiles/README.md.)
__global__ void __launch_bounds__(TOK_WARPS * 32)
tok_emit_kernel(uint32_t nb, const TokJob *__restrict__ jobs)
{
    // Warps are independent (no CTA barrier, statistics added straight to the frame's counters): with a shared histogram and a
    // final __syncthreads the fast warps of a CTA sat at the barrier holding their slots (9.6 stalled warps per issue slot,
    // 84 us per 32 P frames) while the warps with a cluster of coded macroblocks finished.
    __shared__ uint32_t list[TOK_WARPS][256];                        // (position << 16) | uint16(value), in coefficient order
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const TokJob job = jobs[blockIdx.y];
    // a warp walks its macroblocks one after the other (~0.4 us each): short chunks keep the longest chain short where the
    // coded macroblocks of a P frame cluster (32 per warp measured 136 us per 32 frames, most of it waiting for the slowest warp)
    constexpr uint32_t chunk = TOK_CHUNK;
    const uint32_t m0 = (blockIdx.x * TOK_WARPS + warp) * chunk;
    if (m0 < nb) {
        // which of this chunk's macroblocks carry coefficients: `subblocks: None` writes nothing (src/enc.rs:357-358)
        const uint32_t mine = m0 + lane;
        const bool in_chunk = lane < chunk && mine < nb;
        const bool has = in_chunk && (job.hdr == nullptr || job.hdr[mine].has_coeff != 0);
        uint32_t todo = __ballot_sync(FULL, has);
        const uint32_t my_off = in_chunk ? job.mb_off[mine] : 0u;    // where each macroblock's entries start
        uint4 raw_next = make_uint4(0u, 0u, 0u, 0u);
        if (todo) raw_next = __ldcs(reinterpret_cast<const uint4 *>(job.coeff + (size_t)(m0 + (uint32_t)(__ffs((int)todo) - 1)) * 256) + lane);
        LaneHist lh;
        lh.acc[0] = lh.acc[1] = lh.acc[2] = lh.acc[3] = 0;
        uint32_t folded = 0, esc_total = 0, range_bad = 0;
        uint32_t *L = list[warp];
        while (todo) {
            const int j = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const uint4 raw = raw_next;
            if (todo)                                                // the next macroblock's coefficients travel while this one is walked
                raw_next = __ldcs(reinterpret_cast<const uint4 *>(job.coeff + (size_t)(m0 + (uint32_t)(__ffs((int)todo) - 1)) * 256) + lane);
            // (1) compaction
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
            uint32_t nz = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) nz |= (((w[k >> 1] >> (16 * (k & 1))) & 0xffffu) != 0u ? 1u : 0u) << k;
            const uint32_t cnt = (uint32_t)__popc(nz);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t u = __shfl_up_sync(FULL, incl, d);
                if ((int)lane >= d) incl += u;
            }
            const uint32_t T = __shfl_sync(FULL, incl, 31);          // non-zero coefficients of the macroblock
            uint32_t r = incl - cnt;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (nz & (1u << k)) L[r++] = ((8u * lane + (uint32_t)k) << 16) | ((w[k >> 1] >> (16 * (k & 1))) & 0xffffu);
            __syncwarp();
            // (2) entries, 32 list items at a time
            uint32_t *out = job.tok + __shfl_sync(FULL, my_off, j);
            uint64_t r64 = 0, s64 = 0;                               // this lane's entries of this macroblock, 16 bins x 4 bits
            for (uint32_t t0 = 0; t0 < T; t0 += 32u) {
                const uint32_t i = t0 + lane;
                const bool act = i < T;
                const uint32_t e = act ? L[i] : 0u;
                const int pos = (int)(e >> 16);
                const int pprev = (act && i > 0u) ? (int)(L[i - 1u] >> 16) : -1;
                int run = pos - pprev - 1;
                const int esc = act ? escapes_of(run) : 0;           // `while run > 15 { push(15,0,0); run -= 15 }` (src/rle.rs:18-21)
                const uint32_t ntk = act ? (uint32_t)esc + 1u : 0u;
                uint32_t inc2 = ntk;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t u = __shfl_up_sync(FULL, inc2, d);
                    if ((int)lane >= d) inc2 += u;
                }
                if (act) {
                    uint32_t *o = out + (inc2 - ntk);
                    for (int q = 0; q < esc; ++q) *o++ = 15u;        // {15, 0, 0}
                    run -= 15 * esc;
                    const int v = (int)(int16_t)(e & 0xffffu);
                    const uint32_t a = (uint32_t)(v < 0 ? -v : v) & 0xffffu;     // val.abs() as u16 (src/rle.rs:23)
                    const uint32_t size = (32u - (uint32_t)__clz((int)a)) + 1u;   // (16 - leading_zeros) + 1 (src/rle.rs:24)
                    range_bad |= size > 15u ? PFV_TOKFLAG_RANGE : 0u;              // does not fit a 4-bit symbol (src/rle.rs:43)
                    *o = (uint32_t)run | (size << 4) | (e << 16);
                    r64 += 1ull << (run << 2);
                    s64 += 1ull << ((size & 15u) << 2);
                    esc_total += (uint32_t)esc;
                }
                out += __shfl_sync(FULL, inc2, 31);
            }
            if (lane == 0) {                                         // the tail of the macroblock (src/rle.rs:31-38)
                int run = 255 - (T ? (int)(L[T - 1u] >> 16) : -1);
                if (run > 0) {
                    const int esc = escapes_of(run);
                    for (int q = 0; q < esc; ++q) *out++ = 15u;
                    run -= 15 * esc;
                    esc_total += (uint32_t)esc;
                    *out = (uint32_t)run;                            // {run, 0, 0}
                    r64 += 1ull << (run << 2);
                    s64 += 1ull;
                }
            }
            __syncwarp();                                            // the list is rewritten by the next macroblock
            hist_fold(lh, r64, s64);
            if (++folded == 16u) { hist_flush(lh, lane, job.stats); folded = 0; }   // 16 x 9 entries fit the 8-bit fields
        }
        if (folded) hist_flush(lh, lane, job.stats);
        // escapes {15, 0, 0} can come 17 to an entry: counted apart, one add per warp
        esc_total = __reduce_add_sync(FULL, esc_total);
        if (lane == 0 && esc_total) {
            atomicAdd(job.stats + 15, esc_total);
            atomicAdd(job.stats + 16, esc_total);
        }
        if (__any_sync(FULL, range_bad != 0) && lane == 0) atomicOr(job.stats + PFV_TOKSTATS_FLAGS, PFV_TOKFLAG_RANGE);
    }
}

// ---------------------------------------------------------------------------------------------------
// key frames: one thread per sub-block (see the top of this file)
// ---------------------------------------------------------------------------------------------------
constexpr int TSB_WARPS = 4;
constexpr int TSB_PITCH = 144;                                       // bytes per sub-block in the value stage: 128 + 16 (conflict free)
constexpr uint32_t TSB_TILES = 3;                                    // tiles of 8 macroblocks per warp (3 x 64 entries per lane and
                                                                     // symbol still fit the 8-bit fields of LaneHist)
constexpr uint32_t TSB_SLOT = 64;                                    // entries of a macroblock staged in shared memory (the rest, rare,
                                                                     // goes straight to global memory)

struct __align__(16) TsbSmem {
    uint4    val[TSB_WARPS][32 * TSB_PITCH / 16];                    // the tile's coefficients, looked up by position during the walk
    uint32_t ent[TSB_WARPS][8][TSB_SLOT];                            // the tile's entries, macroblock by macroblock
};

__global__ void __launch_bounds__(TSB_WARPS * 32)
tok_emit_sb_kernel(uint32_t nb, const TokJob *__restrict__ jobs)
{
    __shared__ TsbSmem sm;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, sb = lane & 3u, mbi = lane >> 2;
    const uint32_t tile0 = (blockIdx.x * TSB_WARPS + warp) * TSB_TILES;
    if (tile0 * 8u >= nb) return;
    const TokJob job = jobs[blockIdx.y];
    unsigned char *mine = reinterpret_cast<unsigned char *>(sm.val[warp]) + lane * TSB_PITCH;
    uint32_t *slot = sm.ent[warp][mbi];
    LaneHist lh;
    lh.acc[0] = lh.acc[1] = lh.acc[2] = lh.acc[3] = 0;
    uint32_t esc_total = 0, range_bad = 0;

    // the first tile's coefficients; inside the loop the NEXT tile's are fetched before this one is walked
    auto fetch = [&](uint32_t tile, uint4 (&v)[8]) {
        const uint32_t mm = tile * 8u + mbi;
        const uint4 *src = reinterpret_cast<const uint4 *>(job.coeff + ((size_t)min(mm, nb - 1u) * 256 + sb * 64));
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = __ldcs(src + k);
    };
    uint4 nxt[8];
    fetch(tile0, nxt);
#pragma unroll 1
    for (uint32_t t = 0; t < TSB_TILES; ++t) {
        const uint32_t m0 = (tile0 + t) * 8u;
        if (m0 >= nb) break;
        const uint32_t m = m0 + mbi;
        const bool valid = m < nb;
        // the sub-block's 64 coefficients: kept in shared memory for the walk (values are looked up by position there;
        // registers cannot be indexed), their non-zero mask and run bookkeeping from registers
        uint32_t w[32];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint4 v = nxt[k];
            if (!valid) v = make_uint4(0u, 0u, 0u, 0u);
            *reinterpret_cast<uint4 *>(mine + 16 * k) = v;
            w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
        }
        if (t + 1u < TSB_TILES && (tile0 + t + 1u) * 8u < nb) fetch(tile0 + t + 1u, nxt);
        const SbRuns r = sb_runs(w);
        // the macroblock's four lanes: last non-zero position before this sub-block, entries per lane, offsets (src/rle.rs:9-39)
        const int own_last = r.first >= 0 ? (int)(64u * sb) + r.last : -1;
        int prev = -1;
#pragma unroll
        for (int d = 1; d <= 3; ++d) {
            const int u = __shfl_up_sync(FULL, own_last, d, 4);
            if ((int)sb >= d) prev = max(prev, u);
        }
        const int tail_run = 255 - max(prev, own_last);              // lane 3 of the group: zeros after the last coefficient
        const bool tail = sb == 3u && valid && tail_run > 0;
        uint32_t n = r.inner;
        if (r.first >= 0) n += rle_escapes((int)(64u * sb) + r.first - prev - 1);
        if (tail) n += 1u + rle_escapes(tail_run);
        uint32_t incl = n;
        {
            uint32_t u = __shfl_up_sync(FULL, incl, 1, 4);
            if (sb >= 1u) incl += u;
            u = __shfl_up_sync(FULL, incl, 2, 4);
            if (sb >= 2u) incl += u;
        }
        const uint32_t total = __shfl_sync(FULL, incl, 3, 4);
        if (valid && sb == 0u) job.mb_off[m + 1u] = total;           // the scan turns the counts into offsets
        __syncwarp();                                                // (the value stage; the entry slots of the previous tile are out)

        // Entries go to the macroblock's slot in shared memory and leave as coalesced stores below: written straight to global
        // memory every store instruction of the walk touched 32 different sectors (ncu on the first version: 148 us per 32
        // frames at 39 % issue, the LSU the limit).  The walk itself is 32-bit arithmetic only: the 64-bit mask is taken half by
        // half, the symbol counters are 8 bins x 4 bits per word.
        uint32_t *gout = job.tok + (size_t)m * 256u;
        uint32_t at = incl - n;                                      // entry index inside the macroblock
        auto put = [&](uint32_t e) {
            if (at < TSB_SLOT) slot[at] = e; else gout[at] = e;
            ++at;
        };
        uint32_t rlo = 0, rhi = 0, slo = 0, shi = 0, pending = 0;
        auto count = [&](uint32_t run, uint32_t size) {
            const uint32_t rb = 1u << ((run & 7u) * 4u), sbit = 1u << ((size & 7u) * 4u);
            rlo += run < 8u ? rb : 0u; rhi += run < 8u ? 0u : rb;
            slo += size < 8u ? sbit : 0u; shi += size < 8u ? 0u : sbit;
            if (++pending == 15u) {
                hist_fold(lh, ((uint64_t)rhi << 32) | rlo, ((uint64_t)shi << 32) | slo);
                rlo = rhi = slo = shi = 0; pending = 0;
            }
        };
        uint32_t bits = (uint32_t)r.mask, more = (uint32_t)(r.mask >> 32);
        int base = (int)(64u * sb), pp = prev;
        const unsigned char *vals = mine;
        while (bits | more) {
            if (bits == 0u) { bits = more; more = 0u; base += 32; vals += 64; }
            const int p = __ffs((int)bits) - 1;
            bits &= bits - 1u;
            const int pos = base + p;
            int run = pos - pp - 1;
            pp = pos;
            const int esc = (int)rle_escapes(run);                   // `while run > 15 { push(15,0,0); run -= 15 }` (src/rle.rs:18-21)
            for (int q = 0; q < esc; ++q) put(15u);
            run -= 15 * esc;
            const uint32_t e16 = *reinterpret_cast<const uint16_t *>(vals + 2 * p);
            const int v = (int)(int16_t)e16;
            const uint32_t a = (uint32_t)(v < 0 ? -v : v) & 0xffffu; // val.abs() as u16 (src/rle.rs:23)
            const uint32_t size = (32u - (uint32_t)__clz((int)a)) + 1u;   // (16 - leading_zeros) + 1 (src/rle.rs:24)
            range_bad |= size > 15u ? PFV_TOKFLAG_RANGE : 0u;
            put((uint32_t)run | (size << 4) | (e16 << 16));
            esc_total += (uint32_t)esc;
            count((uint32_t)run, size & 15u);
        }
        if (tail) {                                                  // the tail of the macroblock (src/rle.rs:31-38)
            int run = tail_run;
            const int esc = (int)rle_escapes(run);
            for (int q = 0; q < esc; ++q) put(15u);
            run -= 15 * esc;
            esc_total += (uint32_t)esc;
            put((uint32_t)run);                                      // {run, 0, 0}
            count((uint32_t)run, 0u);
        }
        hist_fold(lh, ((uint64_t)rhi << 32) | rlo, ((uint64_t)shi << 32) | slo);
        __syncwarp();
        // the slots out: macroblock j's first min(total, TSB_SLOT) entries, 32 at a time
#pragma unroll
        for (uint32_t j = 0; j < 8u; ++j) {
            const uint32_t nj = min(__shfl_sync(FULL, total, 4 * j), TSB_SLOT);
            uint32_t *g = job.tok + (size_t)(m0 + j) * 256u;
#pragma unroll
            for (uint32_t i = 0; i < TSB_SLOT; i += 32u)
                if (i + lane < nj) g[i + lane] = sm.ent[warp][j][i + lane];
        }
    }
    __syncwarp();
    hist_flush(lh, lane, job.stats);                                 // (a lane makes at most 3 x 64 entries of one symbol)
    esc_total = __reduce_add_sync(FULL, esc_total);
    if (lane == 0 && esc_total) {
        atomicAdd(job.stats + 15, esc_total);
        atomicAdd(job.stats + 16, esc_total);
    }
    if (__any_sync(FULL, range_bad != 0) && lane == 0) atomicOr(job.stats + PFV_TOKSTATS_FLAGS, PFV_TOKFLAG_RANGE);
}

constexpr int STORE_THREADS = 256;
constexpr uint32_t STORE_MBS = 64;                                   // macroblocks a CTA packs at a time (key frames)
constexpr uint32_t STORE_PACK = 4096;                                // ... if they make no more entries than this (else piece by piece)
__global__ void __launch_bounds__(STORE_THREADS)
tok_store_kernel(uint32_t nb, const TokJob *__restrict__ jobs)
{
    const TokJob job = jobs[blockIdx.y];
    const uint32_t ntok = job.stats[PFV_TOKSTATS_NTOK];
    const uint32_t n = min(ntok, job.tok_cap);
    const uint32_t tid = blockIdx.x * STORE_THREADS + threadIdx.x, nthr = gridDim.x * STORE_THREADS;
    if (job.padded) {
        // Key frames: macroblock m's entries from its slot to [mb_off[m], mb_off[m + 1]) of the sequence.  A CTA takes
        // STORE_MBS consecutive macroblocks at a time, packs their entries in shared memory (a warp per macroblock) and writes
        // the packed run with consecutive threads on consecutive words - the destination is usually pinned host memory, and
        // macroblock-sized pieces (~128 B at odd offsets) cost a tenth of the encode leg's end-to-end rate over PCIe.
        __shared__ uint32_t pack_s[STORE_PACK];
        const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        for (uint32_t m0 = blockIdx.x * STORE_MBS; m0 < nb; m0 += gridDim.x * STORE_MBS) {
            const uint32_t m1 = min(m0 + STORE_MBS, nb);
            const uint32_t o0 = job.mb_off[m0], o1 = min(job.mb_off[m1], n);
            const bool fits = o1 > o0 && o1 - o0 <= STORE_PACK;
            for (uint32_t mb = m0 + warp; mb < m1; mb += STORE_THREADS / 32) {
                const uint32_t o = job.mb_off[mb], e = min(job.mb_off[mb + 1u], n);
                const uint32_t *src = job.tok + (size_t)mb * 256u;
                for (uint32_t i = o + lane; i < e; i += 32u) {
                    if (fits) pack_s[i - o0] = src[i - o]; else job.out_tok[i] = src[i - o];
                }
            }
            __syncthreads();
            if (fits)
                for (uint32_t i = o0 + threadIdx.x; i < o1; i += STORE_THREADS) job.out_tok[i] = pack_s[i - o0];
            __syncthreads();
        }
    } else if ((reinterpret_cast<uintptr_t>(job.out_tok) & 15u) == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(job.tok);
        uint4 *d4 = reinterpret_cast<uint4 *>(job.out_tok);
        const uint32_t n4 = n >> 2;
        for (uint32_t i = tid; i < n4; i += nthr) d4[i] = s4[i];
        for (uint32_t i = (n4 << 2) + tid; i < n; i += nthr) job.out_tok[i] = job.tok[i];
    } else {
        for (uint32_t i = tid; i < n; i += nthr) job.out_tok[i] = job.tok[i];
    }
    if (job.out_mb_off)
        for (uint32_t i = tid; i <= nb; i += nthr) job.out_mb_off[i] = job.mb_off[i];
    if (blockIdx.x == 0 && threadIdx.x < PFV_TOKSTATS_WORDS) {
        uint32_t v = job.stats[threadIdx.x];
        if (threadIdx.x == PFV_TOKSTATS_FLAGS && ntok > job.tok_cap) v |= PFV_TOKFLAG_OVERFLOW;
        job.out_stats[threadIdx.x] = v;
    }
}

__global__ void tok_zero_kernel(const TokJob *__restrict__ jobs)
{
    if (threadIdx.x < PFV_TOKSTATS_WORDS) jobs[blockIdx.x].stats[threadIdx.x] = 0;
}

// Key frames (the first n_key jobs): statistics cleared, entries emitted into the padded slots with their counts, then the scan.
// P frames: the per-macroblock entry counts are already in mb_off[1..] (written by the encode kernel, EncJob::mb_cnt).
cudaError_t launch_tokenize(uint32_t nb, const TokJob *d_jobs, uint32_t n_key, uint32_t njobs, cudaStream_t s)
{
    if (n_key) {
        tok_zero_kernel<<<n_key, 64, 0, s>>>(d_jobs);
        const uint32_t tiles = (nb + 7u) / 8u;
        dim3 grid((tiles + TSB_WARPS * TSB_TILES - 1) / (TSB_WARPS * TSB_TILES), n_key, 1);
        tok_emit_sb_kernel<<<grid, TSB_WARPS * 32, 0, s>>>(nb, d_jobs);
        tok_scan_kernel<false><<<n_key, SCAN_THREADS, 0, s>>>(nb, d_jobs);
    }
    if (njobs > n_key) {
        const uint32_t np = njobs - n_key;
        tok_scan_kernel<true><<<np, SCAN_THREADS, 0, s>>>(nb, d_jobs + n_key);
        const uint32_t chunks = (nb + TOK_CHUNK - 1u) / TOK_CHUNK;
        dim3 grid((chunks + TOK_WARPS - 1) / TOK_WARPS, np, 1), block(TOK_WARPS * 32, 1, 1);
        tok_emit_kernel<<<grid, block, 0, s>>>(nb, d_jobs + n_key);
    }
    return cudaGetLastError();
}

cudaError_t launch_token_store(uint32_t nb, const TokJob *d_jobs, uint32_t njobs, cudaStream_t s)
{
    // few CTAs: the sink is PCIe (or a neighbouring HBM buffer), and the kernel shares the GPU with the next submit's work
    const char *e = getenv("PFV_TOK_STORE_CTAS");
    const int v = e ? atoi(e) : 0;
    dim3 grid((v >= 1 && v <= 1024) ? (unsigned)v : 32u, njobs, 1);
    tok_store_kernel<<<grid, STORE_THREADS, 0, s>>>(nb, d_jobs);
    return cudaGetLastError();
}

}  // namespace pfv
