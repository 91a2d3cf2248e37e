// bench_criterion.cpp -- the reference's criterion benchmarks (benches/bench.rs:33-152) run against the C++ host
// mirror (include/hgi.hpp) of this library: same image (1920x1080, (x*y) as u8, benches/bench.rs:24-28), same
// levels (4), same eight benchmarks, same accounting (Throughput::Bytes(width*height); image clone / buffer
// allocation untimed where the reference uses iter_with_large_setup, :61).  Every iteration goes through the
// host-pointer C ABI, i.e. includes H2D + kernel + D2H, because the reference benchmarks operate on host memory.
//
//   g++ -std=c++17 -O2 tools/bench_criterion.cpp -Lrustyhgi_b200 -l:libhgi_b200.so -Wl,-rpath,$PWD/rustyhgi_b200 -o build/bench_criterion
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <sstream>
#include <vector>

#include "../include/hgi.hpp"

using namespace hgi;
using clk = std::chrono::steady_clock;

static GrayImage get_test_image(uint32_t w, uint32_t h)            // benches/bench.rs:15-31
{
    GrayImage img(w, h);
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) img.at(x, y) = (uint8_t)(x * y);
    return img;
}

template <class F>
static void bench(const char* name, size_t bytes, F&& body, int samples = 25)   // sample_size(25), :156
{
    for (int i = 0; i < 3; ++i) body();                            // warm-up
    std::vector<double> t;
    for (int i = 0; i < samples; ++i) {
        const auto a = clk::now();
        body();
        t.push_back(std::chrono::duration<double>(clk::now() - a).count());
    }
    std::sort(t.begin(), t.end());
    const double med = t[t.size() / 2];
    if (bytes)
        std::printf("| %-24s | %10.1f us | %10.1f us | %9.2f GB/s |\n", name, med * 1e6, t[0] * 1e6, bytes / med / 1e9);
    else
        std::printf("| %-24s | %10.1f us | %10.1f us | %14s |\n", name, med * 1e6, t[0] * 1e6, "-");
}

int main()
{
    const uint32_t width = 1920, height = 1080;
    const size_t size = (size_t)width * height, levels = 4;
    const GrayImage image = get_test_image(width, height);
    std::printf("| benchmark (benches/bench.rs) | median | best | throughput |\n|---|---|---|---|\n");

    {   // "memory": memcpy of the plane (:38-52)
        std::vector<uint8_t> v(size), mem(size);
        for (size_t i = 0; i < size; ++i) v[i] = (uint8_t)i;
        bench("memory", size, [&] { std::memcpy(mem.data(), v.data(), size); });
    }
    {   // :54-96, the four encoder instantiations
        Encoder<LeftTop, NoOp> e1(LeftTop{}, NoOp{}, levels);
        bench("left_top_nop_encode", size, [&] { e1.encode(image); });
        Encoder<LeftTop, Linear> e2(LeftTop{}, Linear::from(QuantizationLevel::Lossless), levels);
        bench("left_top_quanted_encode", size, [&] { e2.encode(image); });
        Encoder<Crossed, NoOp> e3(Crossed{}, NoOp{}, levels);
        bench("crossed_nop_encode", size, [&] { e3.encode(image); });
        Encoder<Crossed, Linear> e4(Crossed{}, Linear::from(QuantizationLevel::Lossless), levels);
        bench("crossed_quanted_encode", size, [&] { e4.encode(image); });
    }
    Encoder<Crossed, Linear> encoder(Crossed{}, Linear::from(QuantizationLevel::Lossless), levels);
    const Grid grid = encoder.encode(image);
    {   // "decode" (:98-110)
        Decoder<Crossed> decoder(Crossed{});
        bench("decode", size, [&] { decoder.decode({width, height}, levels, grid); });
    }
    const Metadata metadata{QuantizationLevel::Medium, InterpolationType::Crossed, width, height, levels};   // :16-22
    {   // "serialization" (:112-127): Archive::serialize_to_writer = zlib level 9 here, flate2 best() there
        const Archive archive{metadata, grid};
        bench("serialization", 0, [&] { std::ostringstream os; archive.serialize_to_writer(os); }, 5);
    }
    {   // the same container through the run-length entropy stage (GPU-built token tables, host bit-packing)
        const Archive archive{metadata, grid};
        bench("serialization (rle)", 0, [&] { std::ostringstream os; archive.serialize_to_writer_rle(os); }, 5);
        bench("compression (rle)", 0, [&] {
            const Archive a2{metadata, encoder.encode(image)};
            std::ostringstream os;
            a2.serialize_to_writer_rle(os);
        }, 5);
    }
    {   // "compression" = encode + serialise (:129-151)
        bench("compression", 0, [&] {
            const Archive archive{metadata, encoder.encode(image)};
            std::ostringstream os;
            archive.serialize_to_writer(os);
        }, 5);
    }
    std::printf("\nkernel launches: %llu\n", (unsigned long long)Context::shared()->kernel_launches());
    return 0;
}
