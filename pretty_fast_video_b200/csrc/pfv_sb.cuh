// pfv_sb.cuh — pieces shared by the register-resident sub-block kernels (one thread = one 8x8 sub-block).
//
// The DC-only shortcut used by all of them: with v[1..7] = 0 the 1-D inverse transform (src/dct.rs:241-293)
// returns v[0] in every output (b0 = b1 = c0, every other term is 0 or 0/k), so a sub-block whose 63 AC
// coefficients are zero decodes to clamp((c0 * deq0 + 32768) >> 8) in all 64 pixels — exactly, in wrapping
// arithmetic, what the two full passes produce.
#pragma once
#include "pfv_device.cuh"

namespace pfv {

constexpr int SB_THREADS = 128;             // threads per CTA of the dense sub-block kernels (= SB_MBS_PER_CTA * 4)

// The two int16 halves of a word are sign-extended by IDP.2A (a 16x8-bit dot product with the constant bytes (1, 0)):
// it issues on the multiply pipe, where the transform kernels have slack, instead of PRMT/SHF on the ALU pipe.
__device__ __forceinline__ void unpack_dequant(const uint4 (&raw)[8], const int32_t *deq, int (&m)[64])
{
    constexpr int zz[64] = PFV_ZIGZAG_INIT;
#pragma unroll
    for (int s = 0; s < 64; ++s) {
        const uint4 &q = raw[s >> 3];
        const uint32_t w = ((s >> 1) & 3) == 0 ? q.x : ((s >> 1) & 3) == 1 ? q.y : ((s >> 1) & 3) == 2 ? q.z : q.w;
        const int c = (s & 1) ? __dp2a_hi((int)w, 0x01000000, 0) : __dp2a_lo((int)w, 0x00000001, 0);
        m[zz[s]] = c * deq[s];                               // src/dct.rs:78-83 (tables by scan position)
    }
}

// ... into the TRANSPOSED register layout idct8x8_regs_rolled takes (t[c * 8 + r])
__device__ __forceinline__ void unpack_dequant_transposed(const uint4 (&raw)[8], const int32_t *deq, int (&t)[64])
{
    constexpr int zz[64] = PFV_ZIGZAG_INIT;
#pragma unroll
    for (int s = 0; s < 64; ++s) {
        const uint4 &q = raw[s >> 3];
        const uint32_t w = ((s >> 1) & 3) == 0 ? q.x : ((s >> 1) & 3) == 1 ? q.y : ((s >> 1) & 3) == 2 ? q.z : q.w;
        const int c = (s & 1) ? __dp2a_hi((int)w, 0x01000000, 0) : __dp2a_lo((int)w, 0x00000001, 0);
        t[(zz[s] & 7) * 8 + (zz[s] >> 3)] = c * deq[s];
    }
}

// out = clamp(prev + delta) on four packed pixels; pos4/neg4 = max(delta,0) / min(max(-delta,0),255) in every byte
// (src/common.rs:100-102; delta is in [-256, 254], and prev - 255 already clamps to 0 for every prev)
__device__ __forceinline__ uint32_t add_delta_sat4(uint32_t prev, uint32_t pos4, uint32_t neg4)
{
    return __vsubus4(__vaddus4(prev, pos4), neg4);            // one of pos4/neg4 is zero
}

// -------------------------------------------------------------------------------------------------
// pieces of the streaming ("classify, compact, transform") kernels: decode-I (pfv_kernels_sb.cu) and encode-I
// (pfv_kernels_enc.cu).  A warp walks tiles of 8 macroblocks (lane = macroblock*4 + sub-block); sub-blocks that
// need the full inverse transform are queued in a per-warp shared-memory ring until 32 are there.
// -------------------------------------------------------------------------------------------------
constexpr int SBW_RING = 64;                // slots per warp: up to 31 carried + 32 new

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one bulk async copy (TMA unit) global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    const uint32_t b = smem_addr(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(b) : "memory");
}

__device__ __forceinline__ void bar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
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

// where sub-block `sb` of macroblock `lm` (index inside its plane) lives in the destination slot
__device__ __forceinline__ uint8_t *sb_dst(uint8_t *slot, const PlaneGeom &pl, uint32_t lm, int sb)
{
    uint32_t col;
    const uint32_t row = div_small(lm, pl.bw, pl.rcp_bw, col);
    return slot + pl.off + (size_t)(row * 16u + (uint32_t)(sb >> 1) * 8u) * pl.pw + col * 16u + (uint32_t)(sb & 1) * 8u;
}

// ring slot s keeps 16-byte chunk k of its sub-block at [s*8 + (k ^ (s & 7))]: conflict-free for writers (one slot per
// lane, consecutive slots) and readers alike
__device__ __forceinline__ void ring_get(const uint4 *ring, uint32_t slot, uint4 (&r2)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) r2[k] = ring[slot * 8u + ((uint32_t)k ^ (slot & 7u))];
}

// full inverse transform of one queued intra sub-block (src/common.rs:313-325) and its 8 row stores
__device__ __forceinline__ void transform_entry_i(const uint4 *ring, const uint32_t *idv, uint32_t slot, uint8_t *slot_base,
                                                  const PlaneGeom &pl, const int32_t *deq)
{
    uint4 r2[8];
    ring_get(ring, slot, r2);
    const uint32_t id = idv[slot];
    uint8_t *dst = sb_dst(slot_base, pl, id >> 2, (int)(id & 3u));
    int m[64];
    unpack_dequant(r2, deq, m);
    idct8x8_regs(m);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        uint2 o;
        o.x = pack4_sat_u8(m[r * 8 + 0], m[r * 8 + 1], m[r * 8 + 2], m[r * 8 + 3]);
        o.y = pack4_sat_u8(m[r * 8 + 4], m[r * 8 + 5], m[r * 8 + 6], m[r * 8 + 7]);
        __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pl.pw), o);
    }
}

// a sub-block with no AC term: both passes collapse to the DC term (see the top of this file)
__device__ __forceinline__ void store_dc_only(uint8_t *dst, uint32_t pw, int c0, int deq0)
{
    const int v = (c0 * deq0 + (128 << 8)) >> 8;
    const uint32_t dc4 = pack4_sat_u8(v, v, v, v);
#pragma unroll
    for (int r = 0; r < 8; ++r) __stcg(reinterpret_cast<uint2 *>(dst + (size_t)r * pw), make_uint2(dc4, dc4));
}

// End of a streaming kernel: pool what the CTA's warps still have queued (< 32 entries each) and transform it with
// full warps.  `coef`/`id` are the [WARPS] arrays of rings; left_head/left_cnt shared scratch.
template <int WARPS, typename RingCoef, typename RingId>
__device__ __forceinline__ void flush_rings_i(RingCoef &coef, RingId &id, uint32_t *left_head, uint32_t *left_cnt,
                                              uint32_t head, uint32_t tail, uint32_t warp, uint32_t lane,
                                              uint8_t *slot_base, const PlaneGeom &pl, const int32_t *deq)
{
    if (lane == 0) { left_head[warp] = head; left_cnt[warp] = tail - head; }
    __syncthreads();
    uint32_t pre[WARPS + 1];
    pre[0] = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) pre[w + 1] = pre[w] + left_cnt[w];
#pragma unroll 1
    for (uint32_t c = warp * 32u; c < pre[WARPS]; c += WARPS * 32u) {
        const uint32_t e = c + lane;
        if (e < pre[WARPS]) {
            int w = 0;
#pragma unroll
            for (int k = 1; k < WARPS; ++k) w += e >= pre[k] ? 1 : 0;
            transform_entry_i(coef[w], id[w], (left_head[w] + (e - pre[w])) & (SBW_RING - 1), slot_base, pl, deq);
        }
    }
}

}  // namespace pfv
