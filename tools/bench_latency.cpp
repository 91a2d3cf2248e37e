// bench_latency.cpp -- per-call latency of the device-pointer C-ABI entry points (include/hgi.h) without any
// interpreter in the way: the reference's own benchmark is one 1080p plane per call (benches/bench.rs:24-28,54-110).
// For each plane size: N back-to-back hgi_encode_dev / hgi_decode_dev calls on one stream, CUDA events around the
// batch (us per call = GPU-side cadence), and the same with a host synchronise after every call (round-trip latency).
//   nvcc -O2 -std=c++17 -o build/bench_latency tools/bench_latency.cpp -Iinclude -Lrustyhgi_b200 -lhgi_b200 \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../rustyhgi_b200'
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "hgi.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)
#define HK(x) do { int r_ = (x); if (r_ != HGI_OK) { fprintf(stderr, "%s: %s\n", #x, hgi_strerror(r_)); exit(1); } } while (0)

struct Case { const char* name; uint32_t w, h, levels; int qlevel; int reps; };

int main(int argc, char** argv)
{
    const double peak = argc > 1 ? atof(argv[1]) : 6454.6;
    const int only = argc > 2 ? atoi(argv[2]) : -1;          // run one case only (index), e.g. under ncu
    const int reps_override = argc > 3 ? atoi(argv[3]) : 0;
    hgi_ctx_t* ctx = nullptr;
    HK(hgi_ctx_create(0, &ctx));
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    const Case cases[] = {{"C1 256x256 L4 Medium", 256, 256, 4, 2, 2000},
                          {"C2 1920x1080 L4 Lossless", 1920, 1080, 4, 0, 2000},
                          {"C2 1920x1080 L4 Medium", 1920, 1080, 4, 2, 2000},
                          {"C3 2368x2614 L6 High", 2368, 2614, 6, 3, 1000},
                          {"C4 16384x16384 L8 Medium", 16384, 16384, 8, 2, 100}};
    printf("| plane | encode us/call (stream) | decode us/call (stream) | encode us (call + sync) | decode us (call + sync) | "
           "encode GB/s (2 B/px) | %% of %.1f |\n|---|---|---|---|---|---|---|\n", peak);
    int idx = -1;
    for (Case c : cases) {
        if (++idx != only && only >= 0) continue;
        if (reps_override > 0) c.reps = reps_override;
        const size_t n = (size_t)c.w * c.h;
        std::vector<uint8_t> host(n);
        for (uint32_t y = 0; y < c.h; ++y)
            for (uint32_t x = 0; x < c.w; ++x) host[(size_t)y * c.w + x] = (uint8_t)((x * y) & 255u);   // benches/bench.rs:26-28
        uint8_t *img, *grid, *out;
        CK(cudaMalloc(&img, n)); CK(cudaMalloc(&grid, n)); CK(cudaMalloc(&out, n));
        CK(cudaMemcpy(img, host.data(), n, cudaMemcpyHostToDevice));
        hgi_params_t p{c.levels, HGI_INTERP_CROSSED, HGI_QUANT_LINEAR, c.qlevel};
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        double res[4];
        for (int mode = 0; mode < 2; ++mode) {
            auto call = [&] {
                if (mode == 0) HK(hgi_encode_dev(ctx, img, 1, c.w, c.h, &p, grid, nullptr, nullptr, st));
                else HK(hgi_decode_dev(ctx, grid, 1, c.w, c.h, &p, out, st));
            };
            for (int i = 0; i < 20; ++i) call();
            CK(cudaStreamSynchronize(st));
            std::vector<float> best;
            for (int s = 0; s < 5; ++s) {
                CK(cudaEventRecord(a, st));
                for (int i = 0; i < c.reps; ++i) call();
                CK(cudaEventRecord(b, st));
                CK(cudaEventSynchronize(b));
                float ms;
                CK(cudaEventElapsedTime(&ms, a, b));
                best.push_back(ms * 1e3f / c.reps);
            }
            std::sort(best.begin(), best.end());
            res[mode] = best[best.size() / 2];
            const auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < c.reps; ++i) { call(); CK(cudaStreamSynchronize(st)); }
            res[2 + mode] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / c.reps;
        }
        const double gbs = 2.0 * n / (res[0] * 1e-6) / 1e9;
        printf("| %s | %.2f | %.2f | %.2f | %.2f | %.0f | %.1f |\n", c.name, res[0], res[1], res[2], res[3], gbs, 100.0 * gbs / peak);
        fflush(stdout);
        CK(cudaFree(img)); CK(cudaFree(grid)); CK(cudaFree(out));
    }
    hgi_ctx_destroy(ctx);
    return 0;
}
