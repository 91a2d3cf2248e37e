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

// Each block owns a contiguous slice of one image.  Bins live in shared memory as lane-private columns,
// cell(bin, lane) = bins[bin * 32 + lane], so the 32 lanes of a warp always hit 32 different banks: every
// shared-memory atomic is one conflict-free wavefront whatever the symbol statistics (a residual plane is mostly
// zeros, the worst case for per-warp bins).  The eight warps of the block share the columns; a final pass sums
// each bin over its 32 lanes with warp shuffles and issues one global atomic per non-empty bin per block.
__device__ __forceinline__ void hist_count_word(uint32_t* col, uint32_t w)
{
    atomicAdd(col + ((w & 0xFFu) << 5), 1u);
    atomicAdd(col + ((w >> 3) & 0x1FE0u), 1u);      // ((w >> 8) & 0xFF) << 5
    atomicAdd(col + ((w >> 11) & 0x1FE0u), 1u);     // ((w >> 16) & 0xFF) << 5
    atomicAdd(col + ((w >> 19) & 0x1FE0u), 1u);     // (w >> 24) << 5
}

__global__ void __launch_bounds__(HT)
hgi_hist_kernel(const uint8_t* __restrict__ data, size_t n_per_image, uint32_t blocks_per_image,
                uint32_t* __restrict__ hist)
{
    __shared__ uint32_t bins[256 * 32];             // 32 KB
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 256 * 32; i += HT) bins[i] = 0u;
    __syncthreads();
    const uint32_t img = blockIdx.x / blocks_per_image;
    const uint32_t b = blockIdx.x - img * blocks_per_image;
    const uint8_t* base = data + (size_t)img * n_per_image;
    const size_t per_block = ((n_per_image + blocks_per_image - 1) / blocks_per_image + 15) & ~(size_t)15;
    size_t lo = (size_t)b * per_block, hi = lo + per_block;
    if (hi > n_per_image) hi = n_per_image;
    uint32_t* col = bins + lane;
    if (lo < hi) {
        // head up to 16 B alignment, 128-bit body, byte tail
        size_t head = ((16 - ((uintptr_t)(base + lo) & 15)) & 15);
        if (head > hi - lo) head = hi - lo;
        for (size_t i = lo + tid; i < lo + head; i += HT) atomicAdd(col + ((uint32_t)base[i] << 5), 1u);
        const size_t vlo = lo + head;
        const size_t nvec = (hi - vlo) / 16;
        const uint4* v4 = reinterpret_cast<const uint4*>(base + vlo);
        size_t i = tid;
        for (; i + HT < nvec; i += 2 * HT) {        // two loads in flight per thread
            const uint4 v0 = __ldg(v4 + i), v1 = __ldg(v4 + i + HT);
            hist_count_word(col, v0.x); hist_count_word(col, v0.y); hist_count_word(col, v0.z); hist_count_word(col, v0.w);
            hist_count_word(col, v1.x); hist_count_word(col, v1.y); hist_count_word(col, v1.z); hist_count_word(col, v1.w);
        }
        for (; i < nvec; i += HT) {
            const uint4 v = __ldg(v4 + i);
            hist_count_word(col, v.x); hist_count_word(col, v.y); hist_count_word(col, v.z); hist_count_word(col, v.w);
        }
        for (size_t j = vlo + nvec * 16 + tid; j < hi; j += HT) atomicAdd(col + ((uint32_t)base[j] << 5), 1u);
    }
    __syncthreads();
    // warp w reduces bins [32w, 32w+32): lane-parallel reads of one bin row, shuffle tree
    for (int bin = (tid >> 5) * 32; bin < (tid >> 5) * 32 + 32; ++bin) {
        uint32_t v = bins[bin * 32 + lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0 && v) atomicAdd(&hist[(size_t)img * 256 + bin], v);
    }
}

// 16 pixels per thread and iteration when both planes are 16-byte aligned: |a - b| per byte, squares summed with a
// 4-way dot product (exact: 4 * 255^2 per word fits 32 bits, widened to 64 bits per 16 bytes), running byte-wise max.
__global__ void __launch_bounds__(256)
hgi_error_kernel(const uint8_t* __restrict__ before, const uint8_t* __restrict__ after, size_t n,
                 unsigned long long* __restrict__ out2, int vec_ok)
{
    unsigned long long sum = 0;
    uint32_t mx = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t groups = vec_ok ? n / 16 : 0;
    uint32_t mx4 = 0;
    for (size_t g = tid; g < groups; g += stride) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(before) + g);
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(after) + g);
        const uint32_t d0 = __vabsdiffu4(a.x, b.x), d1 = __vabsdiffu4(a.y, b.y);   // src/main.rs:89
        const uint32_t d2 = __vabsdiffu4(a.z, b.z), d3 = __vabsdiffu4(a.w, b.w);
        mx4 = __vmaxu4(__vmaxu4(mx4, d0), __vmaxu4(d1, __vmaxu4(d2, d3)));
        sum += __dp4a(d0, d0, __dp4a(d1, d1, __dp4a(d2, d2, __dp4a(d3, d3, 0u))));   // src/main.rs:91
    }
    mx = max(max(mx4 & 255u, (mx4 >> 8) & 255u), max((mx4 >> 16) & 255u, mx4 >> 24));
    for (size_t i = groups * 16 + tid; i < n; i += stride) {
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

// RGB -> luma, the `image` 0.19 `to_luma()` that src/main.rs:42,74 calls before encoding:
// l = 0.2126f*r + 0.7152f*g + 0.0722f*b in f32, evaluated left to right with separately rounded
// multiplies and adds (Rust never contracts to FMA), truncated to u8.  __fmul_rn/__fadd_rn are never
// fused by nvcc, independent of -fmad.
__device__ __forceinline__ uint32_t luma_of(uint32_t r, uint32_t g, uint32_t b)
{
    float l = __fmul_rn(0.2126f, (float)r);
    l = __fadd_rn(l, __fmul_rn(0.7152f, (float)g));
    l = __fadd_rn(l, __fmul_rn(0.0722f, (float)b));
    return (uint32_t)l;   // truncation; l < 256
}

// The same arithmetic without the conversion unit: byte `k` of word `w` becomes the float 2^23 + byte by a single
// PRMT into the mantissa of 0x4B000000, minus 2^23 gives (float)byte exactly; the truncation is an add of 2^23
// rounded toward zero, whose low mantissa byte is floor(l) (0 <= l < 256).  Returns that float's bits.
__device__ __forceinline__ float byte_to_f32(uint32_t w, uint32_t k)
{
    return __fadd_rn(__uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + k)), -8388608.0f);
}
__device__ __forceinline__ uint32_t luma_bits(float r, float g, float b)
{
    float l = __fmul_rn(0.2126f, r);
    l = __fadd_rn(l, __fmul_rn(0.7152f, g));
    l = __fadd_rn(l, __fmul_rn(0.0722f, b));
    return __float_as_uint(__fadd_rz(l, 8388608.0f));   // low byte = (u8)l
}

// 16 pixels per thread: three 128-bit loads (48 RGB bytes), one 128-bit store.
__global__ void __launch_bounds__(256)
hgi_luma_kernel(const uint8_t* __restrict__ rgb, size_t n_pixels, uint8_t* __restrict__ luma, int vec_ok)
{
    const size_t groups = n_pixels / 16;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec_ok) {
        for (size_t gidx = tid; gidx < groups; gidx += stride) {
            const uint4* src = reinterpret_cast<const uint4*>(rgb + gidx * 48);
            const uint4 v0 = __ldg(src), v1 = __ldg(src + 1), v2 = __ldg(src + 2);
            const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
            uint32_t out[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {   // 4 pixels = 12 bytes = words 3q..3q+2
                const uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
                const uint32_t l0 = luma_bits(byte_to_f32(a, 0), byte_to_f32(a, 1), byte_to_f32(a, 2));
                const uint32_t l1 = luma_bits(byte_to_f32(a, 3), byte_to_f32(b, 0), byte_to_f32(b, 1));
                const uint32_t l2 = luma_bits(byte_to_f32(b, 2), byte_to_f32(b, 3), byte_to_f32(c, 0));
                const uint32_t l3 = luma_bits(byte_to_f32(c, 1), byte_to_f32(c, 2), byte_to_f32(c, 3));
                out[q] = __byte_perm(__byte_perm(l0, l1, 0x0040u), __byte_perm(l2, l3, 0x0040u), 0x5410u);
            }
            *reinterpret_cast<uint4*>(luma + gidx * 16) = make_uint4(out[0], out[1], out[2], out[3]);
        }
    }
    const size_t first = vec_ok ? groups * 16 : 0;
    for (size_t i = first + tid; i < n_pixels; i += stride)
        luma[i] = (uint8_t)luma_of(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
}

}  // namespace

cudaError_t launch_rgb_to_luma(const uint8_t* rgb, size_t n_pixels, uint8_t* luma, cudaStream_t stream)
{
    if (n_pixels == 0) return cudaSuccess;
    const int vec_ok = (((uintptr_t)rgb | (uintptr_t)luma) & 15u) == 0;
    uint64_t blocks = (n_pixels / 16 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 32) blocks = 148 * 32;
    hgi_luma_kernel<<<(uint32_t)blocks, 256, 0, stream>>>(rgb, n_pixels, luma, vec_ok);
    return cudaGetLastError();
}

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
    const int vec_ok = (((uintptr_t)before | (uintptr_t)after) & 15u) == 0;
    uint64_t blocks = (n + 256 * 64 - 1) / (256 * 64);
    if (blocks > 148 * 16) blocks = 148 * 16;
    hgi_error_kernel<<<(uint32_t)blocks, 256, 0, stream>>>(before, after, n, out2, vec_ok);
    return cudaGetLastError();
}

}  // namespace hgi
