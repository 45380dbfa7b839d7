// pfv_kernels_sparse.cu — device side of the sparse coefficient transport (sm_100a).
//
// The entropy decoder (host) produces, per macroblock, the list of its non-zero coefficients; the reference then
// scatters them into a dense Vec<i16> (src/dec.rs:258-296, :376-417) that is >90 % zeros on real streams.  Only the
// tokens cross PCIe; this kernel rebuilds the dense macroblock (512 B) on the device: one warp per macroblock,
// tile zeroed and filled in shared memory, written out as one coalesced 16-byte store per lane.
#include "pfv_internal.h"

namespace pfv {

constexpr int EXP_WARPS = 8;

__global__ void __launch_bounds__(EXP_WARPS * 32)
expand_tokens_kernel(uint32_t nb, const SparseJob *__restrict__ jobs)
{
    __shared__ __align__(16) int16_t tile[EXP_WARPS][256];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t m = blockIdx.x * EXP_WARPS + warp;
    if (m >= nb) return;
    const SparseJob job = jobs[blockIdx.y];
    if (job.hdr && job.hdr[m].has_coeff == 0) return;        // skipped macroblocks are never read (src/dec.rs:381)
    uint4 *t4 = reinterpret_cast<uint4 *>(tile[warp]);
    t4[lane] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    const uint32_t b = __ldg(job.mb_off + m), e = __ldg(job.mb_off + m + 1);
    for (uint32_t i = b + lane; i < e; i += 32u) {
        const uint32_t t = __ldcs(job.tok + i);
        tile[warp][(t >> 16) & 255u] = (int16_t)(t & 0xffffu);
    }
    __syncwarp();
    __stcg(reinterpret_cast<uint4 *>(job.coeff + (size_t)m * 256) + lane, t4[lane]);
}

cudaError_t launch_expand_tokens(uint32_t nb, const SparseJob *d_jobs, uint32_t njobs, cudaStream_t s)
{
    dim3 grid((nb + EXP_WARPS - 1) / EXP_WARPS, njobs, 1), block(EXP_WARPS * 32, 1, 1);
    expand_tokens_kernel<<<grid, block, 0, s>>>(nb, d_jobs);
    return cudaGetLastError();
}

}  // namespace pfv
