// hgi_reduce_kernels.cu -- residual histogram and `hgi test` error metrics (sm_100a).
//
// Histogram: north_star's "archive.rs residue histogram / frequency-table construction".  The
// reference has no such code (src/archive.rs:34-38 hands the bytes to flate2); the definition is
// hist[v] = #{grid bytes == v}.  Error metrics: src/main.rs:84-92 (sum of squared differences).
#include "hgi_device.cuh"
#include "hgi_kernels.h"

namespace hgi {

namespace {

constexpr int HT = 256;
constexpr int HWARPS = HT / 32;

// Each block owns a contiguous slice of one image; warps count into private shared-memory bins
// (no inter-warp contention), then one global atomic per non-empty bin per block.
__global__ void __launch_bounds__(HT)
hgi_hist_kernel(const uint8_t* __restrict__ data, size_t n_per_image, uint32_t blocks_per_image,
                uint32_t* __restrict__ hist)
{
    __shared__ uint32_t whist[HWARPS * 256];
    const int tid = threadIdx.x;
    for (int i = tid; i < HWARPS * 256; i += HT) whist[i] = 0u;
    __syncthreads();
    const uint32_t img = blockIdx.x / blocks_per_image;
    const uint32_t b = blockIdx.x - img * blocks_per_image;
    const uint8_t* base = data + (size_t)img * n_per_image;
    const size_t per_block = ((n_per_image + blocks_per_image - 1) / blocks_per_image + 15) & ~(size_t)15;
    size_t lo = (size_t)b * per_block, hi = lo + per_block;
    if (hi > n_per_image) hi = n_per_image;
    uint32_t* mine = &whist[(tid >> 5) * 256];
    if (lo < hi) {
        // head up to 16 B alignment, 128-bit body, byte tail
        size_t head = ((16 - ((uintptr_t)(base + lo) & 15)) & 15);
        if (head > hi - lo) head = hi - lo;
        for (size_t i = lo + tid; i < lo + head; i += HT) atomicAdd(&mine[base[i]], 1u);
        const size_t vlo = lo + head;
        const size_t nvec = (hi - vlo) / 16;
        const uint4* v4 = reinterpret_cast<const uint4*>(base + vlo);
        for (size_t i = tid; i < nvec; i += HT) {
            const uint4 v = __ldg(v4 + i);
            const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                atomicAdd(&mine[wds[k] & 0xFFu], 1u);
                atomicAdd(&mine[(wds[k] >> 8) & 0xFFu], 1u);
                atomicAdd(&mine[(wds[k] >> 16) & 0xFFu], 1u);
                atomicAdd(&mine[wds[k] >> 24], 1u);
            }
        }
        for (size_t i = vlo + nvec * 16 + tid; i < hi; i += HT) atomicAdd(&mine[base[i]], 1u);
    }
    __syncthreads();
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < HWARPS; ++w) total += whist[w * 256 + tid];
    if (total) atomicAdd(&hist[(size_t)img * 256 + tid], total);
}

__global__ void __launch_bounds__(256)
hgi_error_kernel(const uint8_t* __restrict__ before, const uint8_t* __restrict__ after, size_t n,
                 unsigned long long* __restrict__ out2)
{
    unsigned long long sum = 0;
    uint32_t mx = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int d = (int)before[i] - (int)after[i];        // src/main.rs:89
        const uint32_t a = (uint32_t)(d < 0 ? -d : d);
        mx = max(mx, a);
        sum += (unsigned long long)(a * a);                  // src/main.rs:91
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
        mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    }
    __shared__ unsigned long long ssum[8];
    __shared__ uint32_t smx[8];
    if ((threadIdx.x & 31) == 0) { ssum[threadIdx.x >> 5] = sum; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { sum += ssum[w]; mx = max(mx, smx[w]); }
        atomicAdd(&out2[0], sum);
        atomicMax(&out2[1], (unsigned long long)mx);
    }
}

}  // namespace

cudaError_t launch_histogram(const uint8_t* data, size_t n_per_image, uint32_t n_images,
                             uint32_t* hist_out, cudaStream_t stream)
{
    if (n_images == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(hist_out, 0, (size_t)n_images * 256 * sizeof(uint32_t), stream);
    if (e != cudaSuccess || n_per_image == 0) return e;
    // ~64 KiB per block, but at least enough blocks to fill the chip for a single big plane
    uint64_t bpi = (n_per_image + 65535) / 65536;
    if (bpi < 1) bpi = 1;
    if (bpi * n_images > 0x7FFFFFFFull) bpi = 0x7FFFFFFFull / n_images;
    hgi_hist_kernel<<<(uint32_t)(bpi * n_images), HT, 0, stream>>>(data, n_per_image, (uint32_t)bpi, hist_out);
    return cudaGetLastError();
}

cudaError_t launch_error_metrics(const uint8_t* before, const uint8_t* after, size_t n,
                                 unsigned long long* out2, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(unsigned long long), stream);
    if (e != cudaSuccess || n == 0) return e;
    uint64_t blocks = (n + 256 * 64 - 1) / (256 * 64);
    if (blocks > 148 * 16) blocks = 148 * 16;
    hgi_error_kernel<<<(uint32_t)blocks, 256, 0, stream>>>(before, after, n, out2);
    return cudaGetLastError();
}

}  // namespace hgi
