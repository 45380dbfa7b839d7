// pfv_kernels_fmt.cu — colour / format steps next to the macroblock path (sm_100a).
//
//   rgb_to_yuv420_kernel   packed RGB8 -> tight Y, U, V planes: the test helper load_frame (src/lib.rs:337-363: JPEG
//                          YCbCr in f32, evaluated left to right, `as u8`) followed by VideoFrame::from_planes
//                          (src/frame.rs:51-60), whose `reduce` (src/common.rs:523-536) keeps the chroma sample of the
//                          top-left pixel of every 2x2 block.  One thread per 2x2 block.
//   yuv420_to_rgb_kernel   a frame slot's visible crop (src/dec.rs:195-197) -> packed RGB8: save_frame
//                          (src/lib.rs:365-395) with its `double` (nearest-neighbour chroma, src/common.rs:538-556).
//                          One thread per 4 horizontal pixels (three 32-bit stores).
//
// The reference does this arithmetic in f32 without fused multiply-adds; __fmul_rn / __fadd_rn / __fsub_rn keep nvcc
// from contracting, so the bytes are identical to the CPU's.  Both kernels are pure streaming (HBM bound).
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

__global__ void __launch_bounds__(256)
yuv420_to_rgb_kernel(const uint8_t *__restrict__ yplane, const uint8_t *__restrict__ uplane, const uint8_t *__restrict__ vplane,
                     uint32_t w, uint32_t h, uint32_t pw, uint32_t cpw, uint8_t *__restrict__ rgb)
{
    const uint32_t qw = (w + 3) / 4;                                   // groups of 4 pixels per row
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= qw * h) return;
    const uint32_t y = i / qw, x = (i - y * qw) * 4;
    const uint32_t n = min(4u, w - x);                                 // 4, or 2 at the end of a row whose width is 2 mod 4
    const uint8_t *yr = yplane + (size_t)y * pw + x;
    const uint8_t *ur = uplane + (size_t)(y >> 1) * cpw + (x >> 1);
    const uint8_t *vr = vplane + (size_t)(y >> 1) * cpw + (x >> 1);
    uint8_t out[12];
#pragma unroll
    for (uint32_t k = 0; k < 4; ++k) {
        if (k >= n) break;
        const float fy = (float)yr[k], fu = __fsub_rn((float)ur[k >> 1], 128.0f), fv = __fsub_rn((float)vr[k >> 1], 128.0f);
        const float r = __fadd_rn(fy, __fmul_rn(1.402f, fv));
        const float g = __fsub_rn(__fsub_rn(fy, __fmul_rn(0.344136f, fu)), __fmul_rn(0.714136f, fv));
        const float b = __fadd_rn(fy, __fmul_rn(1.772f, fu));
        out[k * 3] = (uint8_t)f32_as_u8(r);
        out[k * 3 + 1] = (uint8_t)f32_as_u8(g);
        out[k * 3 + 2] = (uint8_t)f32_as_u8(b);
    }
    uint8_t *dst = rgb + ((size_t)y * w + x) * 3;
    if (n == 4 && ((reinterpret_cast<uintptr_t>(dst) & 3u) == 0)) {
        uint32_t *d32 = reinterpret_cast<uint32_t *>(dst);
#pragma unroll
        for (int k = 0; k < 3; ++k)
            d32[k] = (uint32_t)out[4 * k] | ((uint32_t)out[4 * k + 1] << 8) | ((uint32_t)out[4 * k + 2] << 16) | ((uint32_t)out[4 * k + 3] << 24);
    } else {
        for (uint32_t k = 0; k < n * 3; ++k) dst[k] = out[k];
    }
}

// The same for a batch of slots in ONE launch (grid.y = picture): one launch per picture left this step launch bound at
// 1.1 TB/s.  4 pixels per thread: one 32-bit luma load, two 16-bit chroma loads (plane pitches are multiples of 16), three
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

cudaError_t launch_yuv420_to_rgb_batch(const uint8_t *d_pool, size_t slot_stride, uint32_t off_u, uint32_t off_v, const uint32_t *slots,
                                       uint32_t n, uint32_t w, uint32_t h, uint32_t pw, uint32_t cpw, uint8_t *d_out, size_t out_stride,
                                       cudaStream_t s)
{
    const uint32_t per = ((w + 3) / 4) * h;
    for (uint32_t i0 = 0; i0 < n; i0 += 64) {
        RgbBatch B;
        const uint32_t m = n - i0 < 64u ? n - i0 : 64u;
        for (uint32_t i = 0; i < m; i++) B.slot[i] = slots[i0 + i];
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

cudaError_t launch_yuv420_to_rgb(const uint8_t *d_y, const uint8_t *d_u, const uint8_t *d_v, uint32_t w, uint32_t h, uint32_t pw,
                                 uint32_t cpw, uint8_t *d_rgb, cudaStream_t s)
{
    const uint32_t n = ((w + 3) / 4) * h;
    yuv420_to_rgb_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_y, d_u, d_v, w, h, pw, cpw, d_rgb);
    return cudaGetLastError();
}

}  // namespace pfv
