// pfv_kernels_fmt.cu — colour / format steps next to the macroblock path (sm_100a).
//
//   rgb_to_yuv420_kernel   packed RGB8 -> tight Y, U, V planes: the test helper load_frame (src/lib.rs:337-363: JPEG
//                          YCbCr in f32, evaluated left to right, `as u8`) followed by VideoFrame::from_planes
//                          (src/frame.rs:51-60), whose `reduce` (src/common.rs:523-536) keeps the chroma sample of the
//                          top-left pixel of every 2x2 block.  One thread per 2x2 block.
//   yuv420_to_rgb_batch*   a frame slot's visible crop (src/dec.rs:195-197) -> packed RGB8: save_frame
//                          (src/lib.rs:365-395) with its `double` (nearest-neighbour chroma, src/common.rs:538-556).
//                          One thread per 8 horizontal pixels (or per 4 when the width is not a multiple of 8).
//
// The reference does this arithmetic in f32 without fused multiply-adds; __fmul_rn / __fadd_rn / __fsub_rn keep nvcc
// from contracting, so the bytes are identical to the CPU's.
#include "pfv_internal.h"

namespace pfv {

// Rust `f as u8`: truncation toward zero, saturating, NaN -> 0
__device__ __forceinline__ uint32_t f32_as_u8(float f) { return (uint32_t)min(max(__float2int_rz(f), 0), 255); }

__global__ void __launch_bounds__(256)
rgb_to_yuv420_kernel(const uint8_t *__restrict__ rgb, uint32_t w, uint32_t h, uint8_t *__restrict__ yp,
                     uint8_t *__restrict__ up, uint8_t *__restrict__ vp)
{
    const uint32_t cw = w / 2, ch = h / 2;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cw * ch) return;
    const uint32_t cy = i / cw, cx = i - cy * cw;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        const size_t row = (size_t)(cy * 2 + dy) * w + cx * 2;
        const uint8_t *p = rgb + row * 3;
        uint32_t yy[2];
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const float r = (float)p[dx * 3], g = (float)p[dx * 3 + 1], b = (float)p[dx * 3 + 2];
            const float fy = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, b));
            yy[dx] = f32_as_u8(fy);
            if (dy == 0 && dx == 0) {                                  // reduce: sample (2*ix, 2*iy)
                const float fu = __fadd_rn(__fsub_rn(__fsub_rn(128.0f, __fmul_rn(0.168736f, r)), __fmul_rn(0.331264f, g)), __fmul_rn(0.5f, b));
                const float fv = __fsub_rn(__fsub_rn(__fadd_rn(128.0f, __fmul_rn(0.5f, r)), __fmul_rn(0.418688f, g)), __fmul_rn(0.081312f, b));
                up[i] = (uint8_t)f32_as_u8(fu);
                vp[i] = (uint8_t)f32_as_u8(fv);
            }
        }
        *reinterpret_cast<uint16_t *>(yp + row) = (uint16_t)(yy[0] | (yy[1] << 8));   // w is even: 2-byte aligned
    }
}

// A batch of slots in ONE launch (grid.y = picture; a single picture is a batch of one): one launch per picture left this step
// launch bound at 1.1 TB/s.  4 pixels per thread: one 32-bit luma load, two 16-bit chroma loads (plane pitches are multiples of 16), three
// 32-bit stores.  Picture i lands at out_base + i * out_stride.
struct RgbBatch {
    uint32_t slot[64];
};
__global__ void __launch_bounds__(256)
yuv420_to_rgb_batch_kernel(const uint8_t *__restrict__ pool, size_t slot_stride, uint32_t off_u, uint32_t off_v,
                           const __grid_constant__ RgbBatch B, uint32_t w, uint32_t h, uint32_t pw, uint32_t cpw,
                           uint8_t *__restrict__ out_base, size_t out_stride)
{
    const uint32_t qw = (w + 3) / 4;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= qw * h) return;
    const uint32_t y = i / qw, x = (i - y * qw) * 4;
    const uint32_t n = min(4u, w - x);
    const uint8_t *frame = pool + (size_t)B.slot[blockIdx.y] * slot_stride;
    // the loads may touch up to 2 bytes past the visible width: still inside the padded row (pw >= w rounded up to 16)
    const uint32_t y4 = *reinterpret_cast<const uint32_t *>(frame + (size_t)y * pw + x);
    const uint32_t u2 = *reinterpret_cast<const uint16_t *>(frame + off_u + (size_t)(y >> 1) * cpw + (x >> 1));
    const uint32_t v2 = *reinterpret_cast<const uint16_t *>(frame + off_v + (size_t)(y >> 1) * cpw + (x >> 1));
    uint8_t out[12];
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) {
        const float fy = (float)((y4 >> (8 * k)) & 0xffu);
        const float fu = __fsub_rn((float)((u2 >> (8 * (k >> 1))) & 0xffu), 128.0f), fv = __fsub_rn((float)((v2 >> (8 * (k >> 1))) & 0xffu), 128.0f);
        const float r = __fadd_rn(fy, __fmul_rn(1.402f, fv));
        const float g = __fsub_rn(__fsub_rn(fy, __fmul_rn(0.344136f, fu)), __fmul_rn(0.714136f, fv));
        const float b = __fadd_rn(fy, __fmul_rn(1.772f, fu));
        out[k * 3] = (uint8_t)f32_as_u8(r);
        out[k * 3 + 1] = (uint8_t)f32_as_u8(g);
        out[k * 3 + 2] = (uint8_t)f32_as_u8(b);
    }
    uint8_t *dst = out_base + (size_t)blockIdx.y * out_stride + ((size_t)y * w + x) * 3;
    if (n == 4 && ((reinterpret_cast<uintptr_t>(dst) & 3u) == 0)) {
        uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
#pragma unroll
        for (int k = 0; k < 3; ++k)
            __stcs(d32 + k, (uint32_t)out[4 * k] | ((uint32_t)out[4 * k + 1] << 8) | ((uint32_t)out[4 * k + 2] << 16) | ((uint32_t)out[4 * k + 3] << 24));
    } else {
        for (uint32_t k = 0; k < n * 3; ++k) dst[k] = out[k];
    }
}

// The batch step again for pictures whose width is a multiple of 8 (and a 16-byte aligned destination): 8 pixels per thread.
// ncu on the 4-pixel kernel (64 x 1080p per launch, 0.39 of the HBM roofline): a thread's three 32-bit stores lie 12 bytes from its
// neighbour's, so every store instruction of a warp touches all twelve sectors of the warp's 384 bytes a third at a time, and the 20
// conversions per thread (I2F, F2I) run on the quarter-rate XU pipe.  Here
//   * byte -> float is a byte permute into the mantissa of 2^23 and one subtraction, float -> byte (Rust's saturating, truncating
//     `as u8`, src/lib.rs:384-386) an addition of 1.5 * 2^23 rounded toward zero and a saturating pack of four: no conversion
//     instruction at all, the arithmetic in between is the reference's, operation by operation (the products with the chroma
//     sample are shared by the two pixels that share the sample);
//   * a warp's 768 output bytes are contiguous (the picture is tight, so a group's bytes follow its predecessor's even across the
//     end of a row): they go through 768 bytes of shared memory and leave as 16-byte stores of consecutive lanes.
__device__ __forceinline__ float byte_as_f32(uint32_t w, uint32_t k)
{
    return __fsub_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u | k)), 8388608.0f);   // 2^23 + p, exactly
}
// floor(f) as an integer without a conversion instruction: f + 1.5 * 2^23 rounded toward zero (the sum is positive, so that is the
// floor) has floor(f) in its mantissa, biased.  Rust's `f as u8` (saturating, toward zero, src/lib.rs:384-386) is the floor clamped
// to 0..255: for f >= 0 the floor IS the truncation, every f < 0 ends at 0 either way.  The clamp comes with the pack (I2IP.SAT).
__device__ __forceinline__ int f32_floor_i32(float f)
{
    return (int)(__float_as_uint(__fadd_rz(f, 12582912.0f)) - 0x4B400000u);       // |f| < 2^22 here (|f| < 512)
}
// d = (c << 16) | (sat_u8(a) << 8) | sat_u8(b)   (PTX cvt.pack, SASS I2IP)
__device__ __forceinline__ uint32_t pack_sat_u8x2(int a, int b, uint32_t c)
{
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t pack4_sat(int p0, int p1, int p2, int p3)     // p0 in the lowest byte
{
    return pack_sat_u8x2(p1, p0, pack_sat_u8x2(p3, p2, 0u));
}

constexpr int RGB8_WARPS = 8;
__global__ void __launch_bounds__(RGB8_WARPS * 32)
yuv420_to_rgb_batch8_kernel(const uint8_t *__restrict__ pool, size_t slot_stride, uint32_t off_u, uint32_t off_v,
                            const __grid_constant__ RgbBatch B, uint32_t gw, float rcp_gw, uint32_t ngroups, uint32_t pw, uint32_t cpw,
                            uint8_t *__restrict__ out_base, size_t out_stride)
{
    __shared__ __align__(16) uint8_t stage[RGB8_WARPS][768];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t g0 = (blockIdx.x * RGB8_WARPS + warp) * 32u;          // this warp's first group of 8 pixels
    if (g0 >= ngroups) return;
    const uint32_t g = g0 + lane;
    const bool valid = g < ngroups;
    const uint8_t *frame = pool + (size_t)B.slot[blockIdx.y] * slot_stride;
    uint32_t w[6] = {0u, 0u, 0u, 0u, 0u, 0u};
    if (valid) {
        uint32_t y = (uint32_t)(__uint2float_rz(g) * rcp_gw);                // g / gw (g < 2^24: launch_yuv420_to_rgb_batch checks)
        int xg = (int)(g - y * gw);
        if (xg < 0) { --y; xg += (int)gw; } else if (xg >= (int)gw) { ++y; xg -= (int)gw; }
        const uint2 y8 = __ldcs(reinterpret_cast<const uint2 *>(frame + (size_t)y * pw + (uint32_t)xg * 8u));
        const uint32_t u4 = __ldg(reinterpret_cast<const uint32_t *>(frame + off_u + (size_t)(y >> 1) * cpw + (uint32_t)xg * 4u));
        const uint32_t v4 = __ldg(reinterpret_cast<const uint32_t *>(frame + off_v + (size_t)(y >> 1) * cpw + (uint32_t)xg * 4u));
        int o[24];
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            // u - 128 in one subtraction: (2^23 + u) - (2^23 + 128), both integers below 2^24 - the same value as (u as f32) - 128.0
            const float fu = __fsub_rn(__uint_as_float(__byte_perm(u4, 0x4B000000u, 0x7540u | j)), 8388736.0f);
            const float fv = __fsub_rn(__uint_as_float(__byte_perm(v4, 0x4B000000u, 0x7540u | j)), 8388736.0f);
            const float rv = __fmul_rn(1.402f, fv), gu = __fmul_rn(0.344136f, fu), gv = __fmul_rn(0.714136f, fv), bu = __fmul_rn(1.772f, fu);
#pragma unroll
            for (uint32_t k = 0; k < 2; ++k) {
                const uint32_t px = 2u * j + k;
                const float fy = byte_as_f32(px < 4u ? y8.x : y8.y, px & 3u);
                o[px * 3 + 0] = f32_floor_i32(__fadd_rn(fy, rv));
                o[px * 3 + 1] = f32_floor_i32(__fsub_rn(__fsub_rn(fy, gu), gv));
                o[px * 3 + 2] = f32_floor_i32(__fadd_rn(fy, bu));
            }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) w[i] = pack4_sat(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    }
    uint8_t *dst = out_base + (size_t)blockIdx.y * out_stride + (size_t)g0 * 24u;
    if (g0 + 32u <= ngroups) {
        uint2 *st = reinterpret_cast<uint2 *>(stage[warp] + lane * 24u);      // (24-byte pitch, 8-byte stores: conflict free)
        st[0] = make_uint2(w[0], w[1]);
        st[1] = make_uint2(w[2], w[3]);
        st[2] = make_uint2(w[4], w[5]);
        __syncwarp();
        const uint4 *rd = reinterpret_cast<const uint4 *>(stage[warp]);
        __stcs(reinterpret_cast<uint4 *>(dst) + lane, rd[lane]);
        if (lane < 16u) __stcs(reinterpret_cast<uint4 *>(dst) + 32u + lane, rd[32u + lane]);
    } else if (valid) {                                                      // the picture's last, partly filled warp
        uint2 *d8 = reinterpret_cast<uint2 *>(dst + lane * 24u);
        d8[0] = make_uint2(w[0], w[1]);
        d8[1] = make_uint2(w[2], w[3]);
        d8[2] = make_uint2(w[4], w[5]);
    }
}

cudaError_t launch_yuv420_to_rgb_batch(const uint8_t *d_pool, size_t slot_stride, uint32_t off_u, uint32_t off_v, const uint32_t *slots,
                                       uint32_t n, uint32_t w, uint32_t h, uint32_t pw, uint32_t cpw, uint8_t *d_out, size_t out_stride,
                                       cudaStream_t s)
{
    const uint32_t per = ((w + 3) / 4) * h;
    const bool by8 = w % 8u == 0 && out_stride % 16u == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15u) == 0 && (uint64_t)(w / 8u) * h < (1ull << 24);
    for (uint32_t i0 = 0; i0 < n; i0 += 64) {
        RgbBatch B;
        const uint32_t m = n - i0 < 64u ? n - i0 : 64u;
        for (uint32_t i = 0; i < m; i++) B.slot[i] = slots[i0 + i];
        if (by8) {
            const uint32_t gw = w / 8u, ngroups = gw * h;
            dim3 grid8((ngroups + RGB8_WARPS * 32 - 1) / (RGB8_WARPS * 32), m, 1);
            yuv420_to_rgb_batch8_kernel<<<grid8, RGB8_WARPS * 32, 0, s>>>(d_pool, slot_stride, off_u, off_v, B, gw, 1.0f / (float)gw, ngroups, pw, cpw,
                                                                         d_out + (size_t)i0 * out_stride, out_stride);
            cudaError_t e8 = cudaGetLastError();
            if (e8 != cudaSuccess) return e8;
            continue;
        }
        dim3 grid((per + 255) / 256, m, 1);
        yuv420_to_rgb_batch_kernel<<<grid, 256, 0, s>>>(d_pool, slot_stride, off_u, off_v, B, w, h, pw, cpw, d_out + (size_t)i0 * out_stride, out_stride);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_rgb_to_yuv420(const uint8_t *d_rgb, uint32_t w, uint32_t h, uint8_t *d_y, uint8_t *d_u, uint8_t *d_v, cudaStream_t s)
{
    const uint32_t n = (w / 2) * (h / 2);
    rgb_to_yuv420_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_rgb, w, h, d_y, d_u, d_v);
    return cudaGetLastError();
}

}  // namespace pfv
