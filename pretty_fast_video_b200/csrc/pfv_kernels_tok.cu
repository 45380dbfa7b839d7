// pfv_kernels_tok.cu — device side of the sparse ENCODE transport (SURVEY §8 f2 "on encode" + f4 "GPU-side RLE", sm_100a).
//
// The reference's entropy stage walks every macroblock's 256 dense coefficients on the host (rle_encode,
// src/rle.rs:9-39, called per macroblock from src/enc.rs:256-262 / :370-378), counts the symbols (update_table,
// src/rle.rs:41-47) and only then writes bits.  A dense 1080p frame is 6.27 MB of i16 over PCIe and through a host
// scan; >90 % of it is zeros.  Here the run-length pass runs on the device and only the RLE sequence crosses PCIe:
//
//   tok_count_kernel   one warp per macroblock: number of RLE entries it produces
//   tok_scan_kernel    one CTA per frame: exclusive prefix sum -> mb_off[nb+1]; clears the frame's statistics
//   tok_emit_kernel    one warp per macroblock: the entries themselves, in stream order, + the two symbol histograms
//   tok_store_kernel   copies exactly `ntok` entries (+ statistics, + mb_off on request) to the caller's buffers with
//                      16-byte stores; the destination may be pinned HOST memory (zero-copy over PCIe), so the
//                      transfer size follows the data without a host round trip
//
// Entry format (what pfv_packet_encode_tokens takes): run | size << 4 | uint16(value) << 16, exactly one word per
// RLESequence {num_zeroes, coeff_size, coeff} (src/rle.rs:3-7).
#include "pfv_internal.h"

namespace pfv {

constexpr int TOK_WARPS = 8;
constexpr unsigned FULL = 0xffffffffu;

// escapes a zero run of `run` costs before its final entry: `while run > 15 { push(15,0,0); run -= 15 }` (src/rle.rs:18-21)
__device__ __forceinline__ int escapes_of(int run)
{
    return (run - 1) / 15;                                           // run = 0 -> 0 (C division truncates)
}

struct LaneCoeffs {
    int      v[8];      // this lane's coefficients: macroblock positions 8*lane .. 8*lane+7
    uint32_t nz;        // bit k set = v[k] != 0
    int      prev;      // position of the last non-zero coefficient before 8*lane (-1: none)
    int      last;      // position of the macroblock's last non-zero coefficient (-1: none), all lanes
};

__device__ __forceinline__ LaneCoeffs load_lane(const int16_t *__restrict__ mb, uint32_t lane)
{
    LaneCoeffs c;
    const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(mb) + lane);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    c.nz = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        c.v[k] = (int)(int16_t)(w[k >> 1] >> (16 * (k & 1)));
        c.nz |= (c.v[k] != 0 ? 1u : 0u) << k;
    }
    int incl = c.nz ? (int)(8u * lane) + (31 - __clz((int)c.nz)) : -1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, d);
        if ((int)lane >= d) incl = max(incl, t);
    }
    c.prev = __shfl_up_sync(FULL, incl, 1);
    if (lane == 0) c.prev = -1;
    c.last = __shfl_sync(FULL, incl, 31);
    return c;
}

// RLE entries this lane produces: its non-zero coefficients with the escapes in front of them; lane 31 also owns the
// tail of the macroblock (src/rle.rs:31-38)
__device__ __forceinline__ uint32_t lane_count(const LaneCoeffs &c, uint32_t lane)
{
    uint32_t n = 0;
    int p = c.prev;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (c.nz & (1u << k)) {
            const int pos = (int)(8u * lane) + k;
            n += 1u + (uint32_t)escapes_of(pos - p - 1);
            p = pos;
        }
    if (lane == 31) {
        const int run = 255 - c.last;
        if (run > 0) n += 1u + (uint32_t)escapes_of(run);
    }
    return n;
}

__device__ __forceinline__ bool mb_coded(const TokJob &job, uint32_t m)
{
    return job.hdr == nullptr || job.hdr[m].has_coeff != 0;          // subblocks: None writes nothing (src/enc.rs:357-358)
}

__global__ void __launch_bounds__(TOK_WARPS * 32)
tok_count_kernel(uint32_t nb, const TokJob *__restrict__ jobs)
{
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t m = blockIdx.x * TOK_WARPS + warp;
    if (m >= nb) return;
    const TokJob job = jobs[blockIdx.y];
    uint32_t total = 0;
    if (mb_coded(job, m)) {
        const LaneCoeffs c = load_lane(job.coeff + (size_t)m * 256, lane);
        total = __reduce_add_sync(FULL, lane_count(c, lane));
    }
    if (lane == 0) job.mb_off[m + 1] = total;
}

// mb_off[m+1] holds the count of macroblock m on entry, the inclusive sum on exit; mb_off[0] = 0.
constexpr int SCAN_THREADS = 1024, SCAN_ITEMS = 4;
__global__ void __launch_bounds__(SCAN_THREADS)
tok_scan_kernel(uint32_t nb, const TokJob *__restrict__ jobs)
{
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    __shared__ uint32_t carry_s;
    const TokJob job = jobs[blockIdx.x];
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    uint32_t *cnt = job.mb_off + 1;
    if (t == 0) { job.mb_off[0] = 0; carry_s = 0; }
    if (t < PFV_TOKSTATS_WORDS) job.stats[t] = 0;
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

__global__ void __launch_bounds__(TOK_WARPS * 32)
tok_emit_kernel(uint32_t nb, const TokJob *__restrict__ jobs)
{
    __shared__ uint32_t hist[32];                                    // [0..15] num_zeroes symbols, [16..31] coeff_size symbols
    __shared__ uint32_t bad;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (threadIdx.x < 32) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    const uint32_t m = blockIdx.x * TOK_WARPS + warp;
    const TokJob job = jobs[blockIdx.y];
    if (m < nb && mb_coded(job, m)) {
        const LaneCoeffs c = load_lane(job.coeff + (size_t)m * 256, lane);
        const uint32_t n = lane_count(c, lane);
        uint32_t incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(FULL, incl, d);
            if ((int)lane >= d) incl += u;
        }
        uint32_t *out = job.tok + job.mb_off[m] + (incl - n);
        int p = c.prev;
        uint32_t esc_total = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (c.nz & (1u << k)) {
                const int pos = (int)(8u * lane) + k;
                int run = pos - p - 1;
                const int esc = escapes_of(run);
                for (int e = 0; e < esc; ++e) *out++ = 15u;          // {15, 0, 0}
                run -= 15 * esc;
                esc_total += (uint32_t)esc;
                const int v = c.v[k];
                const uint32_t a = (uint32_t)(v < 0 ? -v : v) & 0xffffu;     // val.abs() as u16 (src/rle.rs:23)
                const uint32_t size = (32u - (uint32_t)__clz((int)a)) + 1u;   // (16 - leading_zeros) + 1 (src/rle.rs:24)
                if (size > 15u) atomicOr(&bad, PFV_TOKFLAG_RANGE);             // does not fit a 4-bit symbol (src/rle.rs:43)
                *out++ = (uint32_t)run | (size << 4) | ((uint32_t)(uint16_t)v << 16);
                atomicAdd(&hist[run], 1u);
                atomicAdd(&hist[16u + (size & 15u)], 1u);
                p = pos;
            }
        if (lane == 31) {
            int run = 255 - c.last;
            if (run > 0) {
                const int esc = escapes_of(run);
                for (int e = 0; e < esc; ++e) *out++ = 15u;
                run -= 15 * esc;
                esc_total += (uint32_t)esc;
                *out++ = (uint32_t)run;                              // {run, 0, 0} (src/rle.rs:36-38)
                atomicAdd(&hist[run], 1u);
                atomicAdd(&hist[16], 1u);
            }
        }
        esc_total = __reduce_add_sync(FULL, esc_total);
        if (lane == 0 && esc_total) {
            atomicAdd(&hist[15], esc_total);
            atomicAdd(&hist[16], esc_total);
        }
    }
    __syncthreads();
    if (threadIdx.x < 32 && hist[threadIdx.x]) atomicAdd(job.stats + threadIdx.x, hist[threadIdx.x]);
    if (threadIdx.x == 32 && bad) atomicOr(job.stats + PFV_TOKSTATS_FLAGS, bad);
}

constexpr int STORE_THREADS = 256;
__global__ void __launch_bounds__(STORE_THREADS)
tok_store_kernel(uint32_t nb, const TokJob *__restrict__ jobs)
{
    const TokJob job = jobs[blockIdx.y];
    const uint32_t ntok = job.stats[PFV_TOKSTATS_NTOK];
    const uint32_t n = min(ntok, job.tok_cap);
    const uint32_t tid = blockIdx.x * STORE_THREADS + threadIdx.x, nthr = gridDim.x * STORE_THREADS;
    if ((reinterpret_cast<uintptr_t>(job.out_tok) & 15u) == 0) {
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

cudaError_t launch_tokenize(uint32_t nb, const TokJob *d_jobs, uint32_t njobs, cudaStream_t s)
{
    dim3 grid((nb + TOK_WARPS - 1) / TOK_WARPS, njobs, 1), block(TOK_WARPS * 32, 1, 1);
    tok_count_kernel<<<grid, block, 0, s>>>(nb, d_jobs);
    tok_scan_kernel<<<njobs, SCAN_THREADS, 0, s>>>(nb, d_jobs);
    tok_emit_kernel<<<grid, block, 0, s>>>(nb, d_jobs);
    return cudaGetLastError();
}

cudaError_t launch_token_store(uint32_t nb, const TokJob *d_jobs, uint32_t njobs, cudaStream_t s)
{
    // few CTAs: the sink is PCIe (or a neighbouring HBM buffer), and the kernel shares the GPU with the next submit's work
    dim3 grid(32, njobs, 1);
    tok_store_kernel<<<grid, STORE_THREADS, 0, s>>>(nb, d_jobs);
    return cudaGetLastError();
}

}  // namespace pfv
