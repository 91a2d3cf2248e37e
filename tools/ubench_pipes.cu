// ubench_pipes.cu -- issue-rate microbenchmark of the instructions the HGI tile kernel is built from (sm_100a).
// Each variant is a loop of independent dependency chains (8 per thread) of one instruction or a mix; every SM runs
// 32 warps (8 per scheduler).  Prints warp-instructions per clock per SM sub-partition (1.0 = the issue limit) so
// that DESIGN.md can say which pipe an instruction loads and at what rate.  Also checks, on the device, that the
// fp16-lane arithmetic the kernel relies on is exact on integer bit patterns (denormals are not flushed).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_pipes tools/ubench_pipes.cu && build/ubench_pipes
#include <cstdint>
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

#define CHAIN8(OP)  OP(0) OP(1) OP(2) OP(3) OP(4) OP(5) OP(6) OP(7)

// x[i] are the chain registers, c0/c1/c2 loop-invariant operands the compiler cannot see through
#define LOP3_(i)   asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(c0), "r"(c1));
#define PRMT_(i)   asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c0), "r"(c1));
#define IADD3_(i)  asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(c0));
#define SHF_(i)    asm volatile("shr.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(c2));
#define IMAD_(i)   asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c0), "r"(c1));
#define IMADHI_(i) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(c0));
#define HFMA2_(i)  asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(c0), "r"(c1));
#define HADD2_(i)  asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(c0));
#define HMUL2_(i)  asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(c0));
#define HSET2_(i)  asm volatile("set.ge.u32.f16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(c0));
#define HMNMX2_(i) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(c0));
#define VIADD_(i)  asm volatile("add.u16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(c0));
#define VIMNMX_(i) asm volatile("min.u16x2 %0, %0, %1;" : "+r"(x[i]) : "r"(c0));
#define FFMA_(i)   asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fc0), "f"(fc1));
// mixes: two or three different instructions per chain step
#define MIX_LOP_IMAD_(i)   LOP3_(i) IMAD_(i)
#define MIX_LOP_HFMA_(i)   LOP3_(i) HFMA2_(i)
#define MIX_IMAD_HFMA_(i)  IMAD_(i) HFMA2_(i)
#define MIX_LOP_IMAD_HFMA_(i) LOP3_(i) IMAD_(i) HFMA2_(i)
#define MIX_LOP_HSET_(i)   LOP3_(i) HSET2_(i)
#define MIX_IMAD_HSET_(i)  IMAD_(i) HSET2_(i)
#define MIX_LOP_PRMT_(i)   LOP3_(i) PRMT_(i)
#define MIX_LOP_VIMNMX_(i) LOP3_(i) VIMNMX_(i)
#define MIX_LOP_FFMA_(i)   LOP3_(i) FFMA_(i)
#define MIX_IMAD_FFMA_(i)  IMAD_(i) FFMA_(i)
#define MIX_HFMA_FFMA_(i)  HFMA2_(i) FFMA_(i)
#define MIX_2LOP_IMAD_HFMA_(i) LOP3_(i) IMAD_(i) PRMT_(i) HFMA2_(i)

#define KERNEL(NAME, OP)                                                                                   \
    __global__ void __launch_bounds__(256) k_##NAME(uint32_t* out, long long* cyc, uint32_t c0, uint32_t c1, \
                                                    uint32_t c2, float fc0, float fc1)                     \
    {                                                                                                      \
        uint32_t x[8];                                                                                     \
        float f[8];                                                                                        \
        for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 0x01010101u + i * 0x00020003u + c0; f[i] = (float)i + fc0; } \
        __syncthreads();                                                                                   \
        const long long t0 = clock64();                                                                    \
        _Pragma("unroll 4") for (int it = 0; it < ITERS; ++it) { CHAIN8(OP) }                                                  \
        const long long t1 = clock64();                                                                    \
        uint32_t s = 0;                                                                                    \
        for (int i = 0; i < 8; ++i) s ^= x[i] ^ __float_as_uint(f[i]);                                     \
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                                    \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                                   \
    }

KERNEL(lop3, LOP3_)
KERNEL(prmt, PRMT_)
KERNEL(iadd, IADD3_)
KERNEL(shf, SHF_)
KERNEL(imad, IMAD_)
KERNEL(imadhi, IMADHI_)
KERNEL(hfma2, HFMA2_)
KERNEL(hadd2, HADD2_)
KERNEL(hmul2, HMUL2_)
KERNEL(hset2, HSET2_)
KERNEL(hmnmx2, HMNMX2_)
KERNEL(viadd16x2, VIADD_)
KERNEL(vimnmx16x2, VIMNMX_)
KERNEL(ffma, FFMA_)
KERNEL(mix_lop_imad, MIX_LOP_IMAD_)
KERNEL(mix_lop_hfma, MIX_LOP_HFMA_)
KERNEL(mix_imad_hfma, MIX_IMAD_HFMA_)
KERNEL(mix_lop_imad_hfma, MIX_LOP_IMAD_HFMA_)
KERNEL(mix_lop_hset, MIX_LOP_HSET_)
KERNEL(mix_imad_hset, MIX_IMAD_HSET_)
KERNEL(mix_lop_prmt, MIX_LOP_PRMT_)
KERNEL(mix_lop_vimnmx, MIX_LOP_VIMNMX_)
KERNEL(mix_lop_ffma, MIX_LOP_FFMA_)
KERNEL(mix_imad_ffma, MIX_IMAD_FFMA_)
KERNEL(mix_hfma_ffma, MIX_HFMA_FFMA_)
KERNEL(mix_lop_imad_prmt_hfma, MIX_2LOP_IMAD_HFMA_)

// ---- exactness of the fp16-lane tricks on integer bit patterns -----------------------------------------------
// quantizer: r = fma(d, K, c1) -> 128 + floor((d+e)/scale)/8 ; q = fma(r, S, c2) -> the integer q in the lane bits
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t hge2(uint32_t a, uint32_t b)
{
    uint32_t r;
    asm("set.ge.u32.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__global__ void k_check(uint32_t K, uint32_t c1, uint32_t S, uint32_t c2, uint32_t error, uint32_t* bad)
{
    // all pairs (d0, d1) in lanes; thread t handles d0 = t & 255, d1 = t >> 8
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t d0 = t & 255u, d1 = (t >> 8) & 255u;
    const uint32_t d = d0 | (d1 << 16);
    const uint32_t r = hfma2(d, K, c1);
    const uint32_t q = hfma2(r, S, c2);
    const uint32_t scale = 2 * error + 1;
    const uint32_t w0 = ((d0 + error) / scale) * scale, w1 = ((d1 + error) / scale) * scale;
    if (q != (w0 | (w1 << 16))) atomicAdd(bad, 1u);
    // unsigned compare of integer lanes through the fp16 comparator (values < 0x7C00)
    const uint32_t a = (d0 * 5u + 3u) | ((d1 * 4u + 1023u) << 16), b = (d1 * 5u) | ((d0 * 4u + 1024u) << 16);
    const uint32_t m = hge2(a, b);
    const uint32_t want = ((a & 0xFFFFu) >= (b & 0xFFFFu) ? 0xFFFFu : 0u) | ((a >> 16) >= (b >> 16) ? 0xFFFF0000u : 0u);
    if (m != want) atomicAdd(bad + 1, 1u);
}

// The predictor's divide-and-floor (pred_pk2 / pred2 of hgi_tile_swar.cuh): for every m = T + 2w in 0..1022, in both lanes
// (lane 1 carries 1022 - m), RN(m b + 512) == 512 + (m+1)/4, RN(256 - m b) == 256 - (m+1)/4 and RN(m b) == (m+1)/4 with
// b = 1/4 - 2^-13 (0x33FF); and the fix-up mask 0x0100 * 255/256 == 0x00FF.
__global__ void k_check_pred(uint32_t* bad)
{
    const uint32_t m0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (m0 > 1022u) return;
    const uint32_t m1 = 1022u - m0, m = m0 | (m1 << 16);
    const uint32_t p = hfma2(m, 0x33FF33FFu, 0x02000200u), pk = hfma2(m, 0xB3FFB3FFu, 0x01000100u), pd = hfma2(m, 0x33FF33FFu, 0u);
    const uint32_t w0 = (m0 + 1u) / 4u, w1 = (m1 + 1u) / 4u;
    if (p != ((512u + w0) | ((512u + w1) << 16)) || pk != ((256u - w0) | ((256u - w1) << 16)) || pd != (w0 | (w1 << 16))) atomicAdd(bad, 1u);
    uint32_t x = (m0 & 1u ? 0x0100u : 0u) | (m0 & 2u ? 0x01000000u : 0u), mk;
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(mk) : "r"(x), "r"(0x3BF83BF8u));
    if (mk != ((m0 & 1u ? 0x00FFu : 0u) | (m0 & 2u ? 0x00FF0000u : 0u))) atomicAdd(bad + 1, 1u);
}

static uint16_t f2h(double v)   // round-to-nearest-even double -> fp16 bits (normal and subnormal), v finite
{
    uint16_t best = 0;
    double bestd = 1e300;
    for (uint32_t b = 0; b < 0x7C00u; ++b) {   // brute force over the positive halves: exactness matters more than speed
        const int e = (b >> 10) & 31, m = b & 1023;
        const double x = e ? ldexp(1.0 + m / 1024.0, e - 15) : ldexp(m / 1024.0, -14);
        const double dd = fabs(fabs(v) - x);
        if (dd < bestd || (dd == bestd && !(b & 1))) { bestd = dd; best = (uint16_t)b; }
    }
    return v < 0 ? (uint16_t)(best | 0x8000u) : best;
}

int main()
{
    int dev = 0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    const int nsm = prop.multiProcessorCount;
    printf("device: %s, %d SMs, clock %d kHz\n", prop.name, nsm, prop.clockRate);
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, (size_t)nsm * 4 * 256 * sizeof(uint32_t));
    cudaMalloc(&cyc, (size_t)nsm * 4 * sizeof(long long));
    long long* h = new long long[nsm * 4];
#define RUN(NAME, PER_STEP)                                                                                  \
    {                                                                                                        \
        for (int rep = 0; rep < 2; ++rep) k_##NAME<<<nsm * 4, 256>>>(out, cyc, 0x3c003c00u, 0x00010001u, 1u, 1.0f, 0.5f); \
        cudaDeviceSynchronize();                                                                             \
        cudaMemcpy(h, cyc, nsm * 4 * sizeof(long long), cudaMemcpyDeviceToHost);                             \
        double s = 0;                                                                                        \
        for (int i = 0; i < nsm * 4; ++i) s += (double)h[i];                                                 \
        s /= nsm * 4;                                                                                        \
        /* 32 warps per SM = 8 per sub-partition, each issuing ITERS * 8 * PER_STEP instructions */          \
        printf("%-24s %6.3f warp-instr/clk/SMSP  (%d instr per step)\n", #NAME, 8.0 * ITERS * 8 * (PER_STEP) / s, PER_STEP); \
    }
    RUN(lop3, 1) RUN(prmt, 1) RUN(iadd, 1) RUN(shf, 1) RUN(imad, 1) RUN(imadhi, 1) RUN(hfma2, 1) RUN(hadd2, 1)
    RUN(hmul2, 1) RUN(hset2, 1) RUN(hmnmx2, 1) RUN(viadd16x2, 1) RUN(vimnmx16x2, 1) RUN(ffma, 1)
    RUN(mix_lop_imad, 2) RUN(mix_lop_hfma, 2) RUN(mix_imad_hfma, 2) RUN(mix_lop_imad_hfma, 3) RUN(mix_lop_hset, 2)
    RUN(mix_imad_hset, 2) RUN(mix_lop_prmt, 2) RUN(mix_lop_vimnmx, 2) RUN(mix_lop_ffma, 2) RUN(mix_imad_ffma, 2)
    RUN(mix_hfma_ffma, 2) RUN(mix_lop_imad_prmt_hfma, 4)

    uint32_t* bad;
    cudaMalloc(&bad, 2 * sizeof(uint32_t));
    for (uint32_t e = 10; e <= 30; e += 10) {
        const uint32_t scale = 2 * e + 1;
        // r = base + floor((d+e)/scale) * ulp(base): base = 128 (ulp 1/8) unless 2^24 / (8 scale) overflows fp16 (e = 10: base 64)
        const double base = (e == 10) ? 64.0 : 128.0, inv_ulp = 1024.0 / base;
        const uint16_t K = f2h(16777216.0 / (inv_ulp * scale));
        const uint16_t c1 = f2h(base + ((double)e / scale - 0.5 + 0.5 / scale) / inv_ulp);
        const uint16_t S = f2h(ldexp(inv_ulp * scale, -24));
        const uint16_t c2 = f2h(-ldexp(1024.0 * scale, -24));
        cudaMemset(bad, 0, 2 * sizeof(uint32_t));
        k_check<<<256, 256>>>(K * 0x10001u, c1 * 0x10001u, S * 0x10001u, c2 * 0x10001u, e, bad);
        uint32_t hb[2];
        cudaMemcpy(hb, bad, sizeof(hb), cudaMemcpyDeviceToHost);
        printf("fp16 quantizer e=%u: K=%04x c1=%04x S=%04x c2=%04x  mismatching lane pairs: %u   fp16-compare mismatches: %u\n",
               e, K, c1, S, c2, hb[0], hb[1]);
    }
    cudaMemset(bad, 0, 2 * sizeof(uint32_t));
    k_check_pred<<<4, 256>>>(bad);
    uint32_t pb[2];
    cudaMemcpy(pb, bad, sizeof(pb), cudaMemcpyDeviceToHost);
    printf("fp16 predictor division by 0x33FF, m = 0..1022 in both lanes: mismatches: %u   fix-up mask mismatches: %u\n", pb[0], pb[1]);
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
