// ubench_hist.cu -- throughput of the primitives a residual histogram can be built from (sm_100a): shared-memory
// atomics under different lane / address patterns, MATCH.ANY, VOTE (ballot), REDUX and non-atomic shared read-modify-write.
// Prints warp-instructions per clock per SM (all four sub-partitions together) measured with clock64 on 32 resident
// warps per SM, scaled so that LOP3 = 2.0 per SM (0.5 per sub-partition; the clock64 tick is not the SM clock).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_hist tools/ubench_hist.cu && build/ubench_hist
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t* out, long long* cyc, uint32_t seed, uint32_t one)
{
    __shared__ uint32_t bins[256 * 32];
    for (int i = threadIdx.x; i < 256 * 32; i += 256) bins[i] = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31;
    uint32_t x = threadIdx.x * 2654435761u + seed, acc = 0;
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        x = x * 1664525u + 1013904223u;                  // cheap per-iteration byte source (2 IMAD-class ops)
        const uint32_t b = x >> 24;
        if (KIND == 0) acc += b;                                                          // baseline: the loop itself
        if (KIND == 1) atomicAdd(&bins[b * 32 + lane], 1u);                               // lane-private column, 32 lanes
        if (KIND == 2) { if (lane < 8) atomicAdd(&bins[b * 32 + lane], 1u); }             // 8 active lanes
        if (KIND == 3) { if (lane < 2) atomicAdd(&bins[b * 32 + lane], 1u); }             // 2 active lanes
        if (KIND == 4) atomicAdd(&bins[(b & 7) * 32 + lane], 1u);                         // few bins (hot lines), lane-private
        if (KIND == 5) atomicAdd(&bins[b], 1u);                                           // warp-shared bins, random conflicts
        if (KIND == 6) atomicAdd(&bins[(it & 255)], 1u);                                  // all lanes one address
        if (KIND == 7) acc += __match_any_sync(0xFFFFFFFFu, b);                           // MATCH.ANY
        if (KIND == 8) acc += __ballot_sync(0xFFFFFFFFu, b == (uint32_t)(it & 255));      // VOTE
        if (KIND == 9) acc += __reduce_add_sync(0xFFFFFFFFu, b);                          // REDUX
        if (KIND == 10) { uint32_t* p = &bins[b * 32 + lane]; *p = *p + 1u; }             // non-atomic RMW, lane-private
        if (KIND == 11) { uint16_t* p = reinterpret_cast<uint16_t*>(bins) + (b * 32 + lane); *p = (uint16_t)(*p + 1u); }
        if (KIND == 12) acc += __popc(x);                                                 // POPC
        if (KIND >= 13 && KIND <= 16) {                                                   // RED, predicated without a branch
            const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bins) + b * 128u + lane * 4u;
            const uint32_t thr = KIND == 13 ? 256u : (KIND == 14 ? 64u : (KIND == 15 ? 16u : 0u));   // P(active) = thr / 256
            asm volatile("{ .reg .pred p; setp.lt.u32 p, %2, %3; @p red.shared.add.u32 [%0], %1; }" ::"r"(addr), "r"(1u), "r"((x >> 8) & 255u), "r"(thr) : "memory");
        }
        if (KIND >= 18 && KIND <= 23) {
            const uint32_t sk = ((x >> 8) & 255u) < 192u ? 0u : b;                         // 75 % zeros, the rest uniform
            const uint32_t v = (KIND == 18 || KIND == 21) ? b : sk;                        // 18/21: uniform bytes; 19/20/22/23: skewed
            uint32_t* cell = (KIND <= 20) ? &bins[v * 32 + lane] : &bins[v];               // 18-20 lane-private, 21-23 warp-shared
            if (KIND == 18 || KIND == 19 || KIND == 21 || KIND == 22) atomicAdd(cell, 1u);  // ATOMS.POPC.INC
            else atomicAdd(cell, one);                                                     // ATOMS.ADD (the addend is opaque)
        }
        if (KIND == 17) {                                                                 // 64 KB layout: address by one PRMT
            const uint32_t addr = __byte_perm(x, lane * 4u, 0x6534u) + (uint32_t)__cvta_generic_to_shared(bins);
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr & 0x7FFFu), "r"(1u) : "memory");
        }
    }
    const long long t1 = clock64();
    uint32_t s = acc;
    for (int i = threadIdx.x; i < 256 * 32; i += 256) s += bins[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(256) k_lop(uint32_t* out, long long* cyc, uint32_t c0, uint32_t c1)
{
    uint32_t x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i + c0;
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it)
        for (int i = 0; i < 8; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(c0), "r"(c1));
    const long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int nsm = prop.multiProcessorCount;
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, (size_t)nsm * 4 * 256 * 4);
    cudaMalloc(&cyc, (size_t)nsm * 4 * 8);
    long long* h = new long long[nsm * 4];
    auto avg = [&] {
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, nsm * 4 * 8, cudaMemcpyDeviceToHost);
        double s = 0;
        for (int i = 0; i < nsm * 4; ++i) s += (double)h[i];
        return s / (nsm * 4);
    };
    k_lop<<<nsm * 4, 256>>>(out, cyc, 1, 2);
    k_lop<<<nsm * 4, 256>>>(out, cyc, 1, 2);
    const double lop = avg();                                    // 32 warps * ITERS * 8 LOP3 at 2.0 per clock per SM
    const double tick = lop / (32.0 * ITERS * 8 / 2.0);          // clock64 ticks per SM clock
    printf("clock64 ticks per SM clock (from LOP3 = 0.5/clk/SMSP): %.3f\n", tick);
    const char* names[] = {"loop only (LCG byte source)", "ATOMS lane-private column, 32 lanes", "ATOMS lane-private, 8 lanes active",
                           "ATOMS lane-private, 2 lanes active", "ATOMS lane-private, 8 hot bins", "ATOMS warp-shared bins, random",
                           "ATOMS one address per warp", "MATCH.ANY", "VOTE.ballot", "REDUX.add", "LDS+STS read-modify-write u32",
                           "LDS+STS read-modify-write u16", "POPC", "RED predicated, all lanes active", "RED predicated, 25% of lanes active",
                           "RED predicated, 6% of lanes active", "RED predicated, no lane active", "RED, address by one PRMT",
                           "lane-private POPC.INC, uniform bytes", "lane-private POPC.INC, 75% zeros", "lane-private ATOMS.ADD, 75% zeros",
                           "warp-shared POPC.INC, uniform bytes", "warp-shared POPC.INC, 75% zeros", "warp-shared ATOMS.ADD, 75% zeros"};
    double base = 0;
#define RUN(K)                                                                                               \
    {                                                                                                        \
        k<K><<<nsm * 4, 256>>>(out, cyc, 1, 1);                                                                 \
        k<K><<<nsm * 4, 256>>>(out, cyc, 2, 1);                                                                 \
        const double c = avg() / tick;                                                                       \
        if (K == 0) base = c;                                                                                \
        printf("%-42s %8.0f SM clocks for %d iterations x 32 warps: %6.2f clk per warp-iteration, %6.2f beyond the loop\n", \
               names[K], c, ITERS, c / (32.0 * ITERS), (c - base) / (32.0 * ITERS));                         \
    }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14) RUN(15) RUN(16) RUN(17) RUN(18) RUN(19) RUN(20) RUN(21) RUN(22) RUN(23)
    printf("cuda status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
