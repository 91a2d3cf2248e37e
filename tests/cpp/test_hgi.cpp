// C++ transcription of the reference's own unit tests (src/lib.rs:25-126) against the C++ host
// mirror (include/hgi.hpp) -- with the one change that makes `test_error` non-vacuous: the
// reference shadows the source image with the decoded one (src/lib.rs:61), so its loop compares the
// decoded image with itself; here the decoded image is compared with the source.
#include <cstdio>
#include <cstdlib>
#include <sstream>

#include "../../include/hgi.hpp"

using namespace hgi;

#define ASSERT(c) do { if (!(c)) { std::fprintf(stderr, "ASSERT FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); std::exit(1); } } while (0)

static GrayImage get_test_image(uint32_t width, uint32_t height)   // src/lib.rs:36-43
{
    GrayImage image(width, height);
    for (uint32_t y = 0; y < height; ++y)
        for (uint32_t x = 0; x < width; ++x) image.at(x, y) = (uint8_t)(x * y);
    return image;
}

static void test_error(QuantizationLevel quantization_level)       // src/lib.rs:45-77
{
    const size_t levels = 3;
    const uint32_t width = 12, height = 8;
    const GrayImage image = get_test_image(width, height);
    Linear quantizator = Linear::from(quantization_level);
    const size_t max_error = quantizator.error();
    Encoder<Crossed, Linear> encoder(Crossed{}, quantizator, levels);
    GrayImage recon;
    const Grid grid = encoder.encode(image, &recon);
    Decoder<Crossed> decoder(Crossed{});
    const GrayImage decoded = decoder.decode({width, height}, levels, grid);
    for (uint32_t y = 0; y < height; ++y)
        for (uint32_t x = 0; x < width; ++x) {
            const int before = image.at(x, y), after = decoded.at(x, y);
            ASSERT((size_t)std::abs(before - after) <= max_error);
            ASSERT(decoded.at(x, y) == recon.at(x, y));     // decoder sees what the encoder predicted from
        }
}

static void serde()                                                 // src/lib.rs:99-125
{
    const size_t levels = 3;
    const uint32_t width = 8, height = 8;
    const GrayImage image = get_test_image(8, 8);
    const QuantizationLevel quantization_level = QuantizationLevel::Lossless;
    Encoder<Crossed, Linear> encoder(Crossed{}, Linear::from(quantization_level), levels);
    const Grid grid = encoder.encode(image);
    const Metadata metadata{quantization_level, InterpolationType::Crossed, width, height, levels};
    const Archive archive{metadata, grid};
    std::stringstream buffer;
    archive.serialize_to_writer(buffer);
    const Archive back = Archive::deserialize_from_reader(buffer);
    ASSERT(back == archive);
    std::stringstream rle;                                          // the fast entropy stage writes the same container
    archive.serialize_to_writer_rle(rle);
    ASSERT(Archive::deserialize_from_reader(rle) == archive);
    std::stringstream bad("\x01\x02\x03\x04 not an archive at all ........");
    bool threw = false;
    try { Archive::deserialize_from_reader(bad); } catch (const Error& e) { threw = e.status() == HGI_ERR_BAD_MAGIC; }
    ASSERT(threw);                                                  // src/archive.rs:47-50
}

static void bench_variants()   // the four encoder instantiations of benches/bench.rs:54-96, on a small plane
{
    const GrayImage image = get_test_image(192, 108);
    const Grid a = Encoder<LeftTop, NoOp>(LeftTop{}, NoOp{}, 4).encode(image);
    const Grid b = Encoder<LeftTop, Linear>(LeftTop{}, Linear::from(QuantizationLevel::Lossless), 4).encode(image);
    const Grid c = Encoder<Crossed, NoOp>(Crossed{}, NoOp{}, 4).encode(image);
    const Grid d = Encoder<Crossed, Linear>(Crossed{}, Linear::from(QuantizationLevel::Lossless), 4).encode(image);
    ASSERT(a == b && c == d && !(a == c));
    ASSERT(Decoder<Crossed>(Crossed{}).decode({192, 108}, 4, d).data == image.data);
    ASSERT(Decoder<LeftTop>(LeftTop{}).decode({192, 108}, 4, a).data == image.data);
}

static void pool_equals_single_context()   // by image and by row band, byte-identical to one context
{
    Pool pool({0, 0});                      // two contexts on device 0 (a box with several GPUs would list them)
    ASSERT(pool.size() == 2);
    const uint32_t w = 640, h = 700;
    const GrayImage image = get_test_image(w, h);
    const Linear q = Linear::from(QuantizationLevel::Medium);
    const Grid one = Encoder<Crossed, Linear>(Crossed{}, q, 6).encode(image);
    const Grid banded = pool.encode_plane(Crossed{}, q, 6, image);
    ASSERT(banded == one);
    ASSERT(pool.decode_plane(Crossed{}, {w, h}, 6, banded).data == Decoder<Crossed>(Crossed{}).decode({w, h}, 6, one).data);
    Bytes batch((size_t)5 * w * h);
    for (int k = 0; k < 5; ++k)
        for (size_t i = 0; i < (size_t)w * h; ++i) batch[(size_t)k * w * h + i] = (uint8_t)(image.data[i] + 31 * k);
    const Bytes grids = pool.encode_batch(Crossed{}, q, 4, batch, 5, w, h);
    for (int k = 0; k < 5; ++k) {
        GrayImage f(w, h);
        std::memcpy(f.data.data(), batch.data() + (size_t)k * w * h, (size_t)w * h);
        const Grid g = Encoder<Crossed, Linear>(Crossed{}, q, 4).encode(f);
        ASSERT(std::memcmp(g.buffer.data(), grids.data() + (size_t)k * w * h, (size_t)w * h) == 0);
    }
    const Bytes back = pool.decode_batch(Crossed{}, 4, grids, 5, w, h);
    int worst = 0;
    for (size_t i = 0; i < back.size(); ++i) { const int d = (int)back[i] - (int)batch[i]; worst = std::max(worst, d < 0 ? -d : d); }
    ASSERT(worst <= q.error());
}

int main()
{
    test_error(QuantizationLevel::Lossless);   // lossless_compression  src/lib.rs:79-82
    test_error(QuantizationLevel::Low);        // low_compression       :84-87
    test_error(QuantizationLevel::Medium);     // medium_compression    :89-92
    test_error(QuantizationLevel::High);       // high_compression      :94-97
    serde();
    bench_variants();
    pool_equals_single_context();
    std::printf("test_hgi: all reference unit tests passed on the GPU path (%llu kernel launches)\n",
                (unsigned long long)Context::shared()->kernel_launches());
    return 0;
}
