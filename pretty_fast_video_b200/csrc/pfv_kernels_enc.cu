// pfv_kernels_enc.cu — encode-I with one THREAD per 8x8 sub-block (the default for key frames; sm_100a).
//
// Encoder::encode_iframe (src/enc.rs:84-97) per plane: encode_plane (src/common.rs:351-386 -> encode_block :141-152 ->
// encode_subblock :287-298 -> fdct / DctMatrix8x8::encode, src/dct.rs:176-239, :88-99) and then decode_plane of what
// was just encoded (src/common.rs:423-446), blitted into prev_frame - the closed loop.
//
// The first-generation kernel (pfv_kernels.cu, one warp per macroblock, transposes and the zig-zag through shared
// memory) ran at 0.18 of the HBM roofline.  Here, like the decode side: a thread keeps its sub-block's 64 values in
// registers, both forward passes (in fp32, where they are exact: pfv_dct.cuh) and the quantiser (four FMA-pipe instructions
// per coefficient) run without any exchange, the zig-zag is a compile-time register renaming, the divisors are reciprocals
// in the constant bank.  A warp walks tiles of 8 consecutive macroblocks (lane = macroblock*4 + sub-block): its 8-byte source
// loads cover full 128-byte lines of the tight source plane (the rows of the next tile are fetched while this one is
// transformed) and the tile's coefficients leave through a shared-memory stage as 8 fully coalesced 512-byte stores.
//
// The reconstruction is the decode-I problem again: a sub-block whose AC terms all quantised to zero reconstructs to
// one flat value (see pfv_sb.cuh) and is stored on the spot; the others are queued in the warp's shared-memory ring
// and inverse-transformed 32 at a time by full warps (classify / compact / transform, as decode_i_stream_kernel).
//
// The kernel is PERSISTENT (3 CTAs per SM for the whole launch): chunks of ENC_CHUNK consecutive tiles are handed out in
// order by a device-wide counter.  Its first form was a grid of short-lived CTAs, 16 tiles per warp (ncu, 64 x 1080p):
//   * a warp's ring of queued sub-blocks now lives across chunks, frames and planes of a class: it is drained ONCE per plane
//     class and warp (1 776 x 2 partly filled transform passes per launch instead of 10 240 - 9 % of all instructions);
//   * nobody waits for the last wave of CTAs (sm__cycles_active min / avg / max was 389 k / 420 k / 452 k of 458 k elapsed).
// 0.53 -> 0.59 of the HBM roofline.  The loop exists in two copies, specialised per plane class (luma / chroma): the
// quantiser's reciprocals and the dequantiser's multipliers are then compile-time offsets into the kernel parameters, i.e.
// constant-bank operands of the multiplies (selected at run time every one of the 128 table reads of a sub-block was a
// uniform load instruction of its own), and with all luma chunks of a launch handed out before all chroma chunks an SM runs
// one copy at a time (each is ~30 KB of code: the instruction cache).  The inverse transform is the rolled one
// (idct8x8_regs_rolled): 64 register moves per pass buy 420 instructions of footprint.
//
#include <stdlib.h>

#include "pfv_internal.h"
#include "pfv_device.cuh"
#include "pfv_sb.cuh"

namespace pfv {

constexpr int ENC_WARPS = 4;
// Resident CTAs per SM (the kernel is persistent: that many per SM are launched).  Measured on 64 x 1080p, frames/s: 1: 110 k,
// 2: 163 k, 3: 309 k (157 registers), 4: 288 k (128 registers, 8 bytes of spills), 5: 197 k (96 registers, spills).
constexpr int ENC_CTAS_PER_SM = 3;

constexpr int ENC_OUT_PITCH = 144;                           // bytes per sub-block in the output stage: 128 + 16

// A lane's 128 B of coefficients are 128 B apart from its neighbour's in the dense layout: stored straight from registers,
// every store instruction of the warp touched 32 different lines (ncu: the stores' source registers were what the next
// instructions waited for).  The tile goes through a per-warp stage instead - written sub-block by sub-block (pitch
// 144 B: conflict free), read back 512 contiguous bytes at a time - and leaves as 8 fully coalesced 512-byte stores.
struct __align__(16) EncPersistSmem {
    uint4    coef[ENC_WARPS][SBW_RING * 8];
    uint2    id[ENC_WARPS][SBW_RING];                          // {macroblock in plane << 2 | sub-block, job << 2 | plane}
    uint4    out[ENC_WARPS][32 * ENC_OUT_PITCH / 16];
};

// Rows y0 .. y0+7, bytes x0 .. x0+7 of a tight vw x vh source plane that are NOT all inside the plane (or whose rows are not
// 8-byte aligned): byte by byte, padded with the clear colour (src/common.rs:352-356).  Rare (the bottom and right edges)
// and out of line: the kernel's loop must stay small enough for the instruction cache.
__device__ __noinline__ void load_src_sb_edge(const uint8_t *__restrict__ src, uint32_t vw, uint32_t vh, uint32_t clear,
                                              uint32_t x0, uint32_t y0, uint2 *rows)
{
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
        const uint32_t y = y0 + (uint32_t)r;
        uint32_t w[2] = {0u, 0u};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t x = x0 + (uint32_t)k;
            const uint32_t b = (y < vh && x < vw) ? (uint32_t)src[(size_t)y * vw + x] : (clear & 0xffu);
            w[k >> 2] |= b << (8 * (k & 3));
        }
        rows[r] = make_uint2(w[0], w[1]);
    }
}

// the eight 8-byte rows of sub-block `sb` of macroblock `lm` of a tight source plane
__device__ __forceinline__ void load_src_sb(const uint8_t *__restrict__ src, const PlaneGeom &pl, uint32_t lm, uint32_t sb,
                                            bool aligned, uint2 (&rows)[8])
{
    uint32_t col;
    const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
    const uint32_t x0 = col * 16u + (sb & 1u) * 8u, y0 = row * 16u + (sb >> 1) * 8u;
    if (aligned && x0 + 8u <= pl.vw && y0 + 8u <= pl.vh) {        // vw % 8 == 0 and base 8-byte aligned: all 64 bytes exist
        const uint8_t *q = src + (size_t)y0 * pl.vw + x0;
#pragma unroll
        for (int r = 0; r < 8; ++r) rows[r] = __ldcs(reinterpret_cast<const uint2 *>(q + (size_t)r * pl.vw));
    } else {
        uint2 tmp[8];                                             // (its address escapes: keep `rows` itself in registers)
        load_src_sb_edge(src, pl.vw, pl.vh, pl.clear4, x0, y0, tmp);
#pragma unroll
        for (int r = 0; r < 8; ++r) rows[r] = tmp[r];
    }
}

// the last tile of a plane when it holds fewer than 8 macroblocks: the coalesced copy out of the stage, predicated
__device__ __noinline__ void store_partial_tile(const unsigned char *stg, uint4 *dstc, uint32_t tile_mbs, uint32_t lane)
{
#pragma unroll 1
    for (uint32_t j = 0; j < 8u; ++j) {
        const uint32_t c = j * 32u + lane;
        if ((c >> 5) < tile_mbs) __stcs(dstc + c, *reinterpret_cast<const uint4 *>(stg + (c >> 3) * 144u + (c & 7u) * 16u));
    }
}

constexpr uint32_t ENC_CHUNK = 2;                              // tiles per chunk (what the slowest warp can finish after the others: ~5 us)

// Chunks are numbered luma first (all frames), then chroma: a warp runs the luma copy of the loop until the counter hands it a
// chroma chunk, drains, and carries on in the chroma copy.  Ring entries carry their frame and plane (a second word).
struct EncChunks {
    uint32_t nl, nc;              // chunks per luma plane / per chroma plane
    float    rcp_nl, rcp_nc;
    uint32_t luma_total, total;   // njobs * nl, njobs * (nl + 2 nc)
};

struct EncTilePos {               // one tile's place: which frame, which plane, which tile of it
    uint32_t job, p, tile;
};

// chunk index -> first tile.  ntiles: tiles of a plane of the chunk's class.
template <int PC>
__device__ __forceinline__ EncTilePos enc_chunk_pos(const EncChunks &C, uint32_t chunk)
{
    EncTilePos t;
    if (PC == 0) {
        uint32_t ct;
        t.job = div_small(chunk, C.nl, C.rcp_nl, ct);
        t.p = 0;
        t.tile = ct * ENC_CHUNK;
    } else {
        uint32_t ct;
        const uint32_t pj = div_small(chunk - C.luma_total, C.nc, C.rcp_nc, ct);     // job * 2 + (plane - 1)
        t.job = pj >> 1;
        t.p = 1u + (pj & 1u);
        t.tile = ct * ENC_CHUNK;
    }
    return t;
}

__device__ __forceinline__ void transform_entry_job(const uint4 *ring, const uint2 *idv, uint32_t slot, const EncJob *__restrict__ jobs,
                                                    const FrameGeom &g, const int32_t *deq)
{
    uint4 r2[8];
    ring_get(ring, slot, r2);
    const uint2 id = idv[slot];
    const PlaneGeom &pl = (id.y & 3u) == 0u ? g.pl[0] : ((id.y & 3u) == 1u ? g.pl[1] : g.pl[2]);
    uint8_t *dst = sb_dst(jobs[id.y >> 2].dst, pl, id.x >> 2, (int)(id.x & 3u));
    int m[64];
    unpack_dequant_transposed(r2, deq, m);
    idct8x8_regs_rolled(m);                                          // (one copy of the 1-D transforms: the instruction cache)
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        uint2 o;
        o.x = pack4_sat_u8(m[r * 8 + 0], m[r * 8 + 1], m[r * 8 + 2], m[r * 8 + 3]);
        o.y = pack4_sat_u8(m[r * 8 + 4], m[r * 8 + 5], m[r * 8 + 6], m[r * 8 + 7]);
        __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pl.pw), o);
    }
}

// All chunks of plane class PC this warp gets, starting with `chunk` (which is of this class); `next` is the chunk after it,
// already taken from the counter.  Returns with `chunk` = the first chunk of another class (or >= C.total), `next` the one after.
template <int PC>
__device__ __forceinline__ void encode_i_class(const EncSbParams &P, const EncChunks &C, const EncJob *__restrict__ jobs,
                                               EncPersistSmem &sm, uint32_t *work, const uint32_t first_dynamic,
                                               uint32_t &chunk, uint32_t &next)
{
    const float *encR = P.encR[PC];
    const int32_t *deq = P.deq[PC];
    const PlaneGeom &plc = PC == 0 ? P.g.pl[0] : P.g.pl[1];     // everything but `off` / `mb_base` is the same for U and V
    const uint32_t nmb = plc.bw * plc.bh, ntiles = (nmb + 7u) / 8u;
    const uint32_t class_end = PC == 0 ? C.luma_total : C.total;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, sb = lane & 3u;
    uint4 *ring = sm.coef[warp];
    uint2 *ring_id = sm.id[warp];
    unsigned char *stg = reinterpret_cast<unsigned char *>(sm.out[warp]);
    const unsigned char *stg_rd = stg + (lane >> 3) * ENC_OUT_PITCH + (lane & 7u) * 16u;

    auto plane_of = [&](uint32_t p) -> const PlaneGeom & { return PC == 0 ? P.g.pl[0] : (p == 1u ? P.g.pl[1] : P.g.pl[2]); };
    // what the loop needs of a frame's job record, kept in registers and re-read (from global memory) only where a warp moves
    // on to another chunk - one tile before the first use (read where they are used, ncu had every tile wait for them).
    // (Measured and dropped: taking the counter two chunks ahead and prefetching the next chunk's job record into L1 at the
    // start of a chunk, and an L2 prefetch of the tile after next as the grid form had it - 210.6 us per 64 frames became
    // 217.6 and 224.7: what ncu shows as long-scoreboard stalls at the bottom of a tile is the next tile's rows being moved
    // into the loop-carried registers, and more address arithmetic in front of them only delays those loads.)
    struct JobRegs { const uint8_t *src; int16_t *coeff; uint8_t *dst; };
    auto job_regs = [&](const EncTilePos &t) {
        const EncJob &j = jobs[t.job];
        JobRegs r;
        r.src = PC == 0 ? j.src[0] : (t.p == 1u ? j.src[1] : j.src[2]);
        r.coeff = j.coeff; r.dst = j.dst;
        return r;
    };
    auto fetch = [&](const EncTilePos &t, const uint8_t *src, uint2 (&rows)[8]) {
        const PlaneGeom &pl = plane_of(t.p);
        load_src_sb(src, pl, min(t.tile * 8u + (lane >> 2), nmb - 1u), sb, ((reinterpret_cast<uintptr_t>(src) | pl.vw) & 7u) == 0, rows);
    };

    EncTilePos cur = enc_chunk_pos<PC>(C, chunk);
    JobRegs jr = job_regs(cur);
    uint32_t left = min(ENC_CHUNK, ntiles - cur.tile);         // tiles of the current chunk not yet done (including `cur`)
    uint2 nxt[8];
    fetch(cur, jr.src, nxt);
    uint32_t head = 0, tail = 0;
    bool more = true;
#pragma unroll 1
    while (more || tail != head) {
        if (more) {
            // where the NEXT tile is: the same chunk, or the first tile of the chunk this warp holds next (if it is of this class)
            EncTilePos nt = cur;
            JobRegs njr = jr;
            uint32_t grabbed = 0xffffffffu;
            bool have_next = true;
            if (left > 1u) {
                nt.tile = cur.tile + 1u;
            } else {
                // last tile of the chunk: take the chunk after `next` now (whatever its class - the other class's loop needs a
                // successor too), its number is in flight until the bottom of this tile
                if (lane == 0 && next < C.total) grabbed = first_dynamic + atomicAdd(work, 1u);
                if (next < class_end) { nt = enc_chunk_pos<PC>(C, next); njr = job_regs(nt); }
                else have_next = false;
            }
            const PlaneGeom &pl = plane_of(cur.p);
            const uint32_t tile = cur.tile;
            const uint32_t lm = tile * 8u + (lane >> 2);
            const bool valid = lm < nmb;
            float y[64];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                float v[8];
                fdct8_f32_row_of_bytes(nxt[r].x, nxt[r].y, v);
#pragma unroll
                for (int c = 0; c < 8; ++c) y[r * 8 + c] = v[c];
            }
            // (unconditional - the last tile of all fetches itself again: a conditional fetch made the compiler keep a second copy
            // of the 16 registers, 32 moves per tile, and wait for the fetch at the end of the iteration that issued it)
            fetch(nt, njr.src, nxt);
            uint32_t w[32];
            fdct8x8_f32_columns(y);
            quantise_sb_f32(y, encR, w);
            uint32_t ac = w[0] & 0xffff0000u;
#pragma unroll
            for (int i = 1; i < 32; ++i) ac |= w[i];
            {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    *reinterpret_cast<uint4 *>(stg + lane * ENC_OUT_PITCH + 16 * k) = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
                __syncwarp();
                const uint32_t tile_mbs = min(8u, nmb - tile * 8u);
                uint4 *dstc = reinterpret_cast<uint4 *>(jr.coeff + (size_t)(pl.mb_base + tile * 8u) * 256);
                if (tile_mbs == 8u) {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        __stcs(dstc + j * 32 + lane, *reinterpret_cast<const uint4 *>(stg_rd + j * 4 * ENC_OUT_PITCH));
                } else {
                    store_partial_tile(stg, dstc, tile_mbs, lane);
                }
                __syncwarp();
            }
            const bool general = valid && ac != 0u;
            const uint32_t vote = __ballot_sync(0xffffffffu, general);
            if (general) {
                const uint32_t slot = (tail + (uint32_t)__popc(vote & ((1u << lane) - 1u))) & (SBW_RING - 1);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    ring[slot * 8u + ((uint32_t)k ^ (slot & 7u))] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
                ring_id[slot] = make_uint2((lm << 2) | sb, (cur.job << 2) | cur.p);
            } else if (valid) {
                store_dc_only(sb_dst(jr.dst, pl, lm, (int)sb), pl.pw, (int)(int16_t)(w[0] & 0xffffu), deq[0]);
            }
            tail += (uint32_t)__popc(vote);
            __syncwarp();
            // advance
            if (left > 1u) {
                --left;
            } else {
                chunk = next;
                next = __shfl_sync(0xffffffffu, grabbed, 0);    // (0xffffffff when nothing was taken: past the end)
                left = have_next ? min(ENC_CHUNK, ntiles - nt.tile) : 0u;
            }
            cur = nt;
            jr = njr;
            more = have_next;
        }
        const uint32_t queued = tail - head;                    // at most 63: 31 carried + 32 new
        if (queued >= 32u || (!more && queued != 0u)) {
            if (lane < queued) transform_entry_job(ring, ring_id, (head + lane) & (SBW_RING - 1), jobs, P.g, deq);
            head += min(32u, queued);
            __syncwarp();
        }
    }
}

__global__ void __launch_bounds__(ENC_WARPS * 32, ENC_CTAS_PER_SM)
encode_i_persist_kernel(const __grid_constant__ EncSbParams P, const __grid_constant__ EncChunks C, const EncJob *__restrict__ jobs,
                        uint32_t *__restrict__ work)
{
    extern __shared__ __align__(16) unsigned char encp_raw[];
    EncPersistSmem &sm = *reinterpret_cast<EncPersistSmem *>(encp_raw);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t nwarps = gridDim.x * ENC_WARPS;
    // the first chunk of a warp is static, the second already comes from the counter (work[0]); work[1] counts retired warps
    uint32_t chunk = blockIdx.x * ENC_WARPS + warp;
    uint32_t next = 0xffffffffu;
    if (chunk < C.total) {
        uint32_t g0 = 0;
        if (lane == 0) g0 = nwarps + atomicAdd(&work[0], 1u);
        next = __shfl_sync(0xffffffffu, g0, 0);
        if (chunk < C.luma_total) encode_i_class<0>(P, C, jobs, sm, &work[0], nwarps, chunk, next);
        if (chunk < C.total)      encode_i_class<1>(P, C, jobs, sm, &work[0], nwarps, chunk, next);
    }
    if (lane == 0 && atomicAdd(&work[1], 1u) == nwarps - 1u) { work[0] = 0u; work[1] = 0u; }
}

// (The sparse encode seam needs nothing from this kernel: the key-frame tokenizer counts a macroblock's run-length entries
// itself, pfv_kernels_tok.cu.  A variant that counted them here while the coefficients were in registers cost 40 us per 32
// frames on a 114 us kernel.)
cudaError_t launch_encode_i_persist(EncSbParams P, const EncJob *d_jobs, uint32_t njobs, uint32_t *d_work, cudaStream_t s)
{
    const uint32_t nty = (P.g.pl[0].bw * P.g.pl[0].bh + 7u) / 8u, ntc = (P.g.pl[1].bw * P.g.pl[1].bh + 7u) / 8u;
    EncChunks C;
    C.nl = (nty + ENC_CHUNK - 1) / ENC_CHUNK;
    C.nc = (ntc + ENC_CHUNK - 1) / ENC_CHUNK;
    C.rcp_nl = 1.0f / (float)C.nl;
    C.rcp_nc = 1.0f / (float)C.nc;
    if ((uint64_t)njobs * (C.nl + 2ull * C.nc) >= (1ull << 24)) return cudaErrorInvalidConfiguration;   // (div_small's range)
    C.luma_total = njobs * C.nl;
    C.total = njobs * (C.nl + 2u * C.nc);
    P.tiles_per_warp = 0; P.cta_total = 0;
    for (int p = 0; p < 3; p++) P.cta_base[p] = 0;
    static bool attr_done[64] = {};           // cudaFuncSetAttribute is per DEVICE: one process may own contexts on several
    const int smem = (int)sizeof(EncPersistSmem);
    if (first_use_on_device(attr_done)) {
        cudaError_t e = cudaFuncSetAttribute(encode_i_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
    }
    uint32_t ctas = (C.total + ENC_WARPS - 1) / ENC_WARPS;
    if (ctas > 148u * (uint32_t)ENC_CTAS_PER_SM) ctas = 148u * (uint32_t)ENC_CTAS_PER_SM;
    encode_i_persist_kernel<<<ctas, ENC_WARPS * 32, smem, s>>>(P, C, d_jobs, d_work);
    return cudaGetLastError();
}

}  // namespace pfv
