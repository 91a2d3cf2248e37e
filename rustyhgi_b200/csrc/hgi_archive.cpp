// hgi_archive.cpp -- the `.hgi` container of src/archive.rs, host side (no CUDA).
//
// Layout (src/archive.rs:31-41, bincode 1.x default config = little-endian fixed-width ints,
// enum = u32 variant index, usize = u64, Vec<u8> = u64 length + bytes):
//   u32 MAGIC 0xBAADA555 | u32 quantization_level | u32 interpolation | u32 width | u32 height |
//   u64 scale_level | raw-DEFLATE( u64 grid_len | grid bytes | u64 grid_width )
// The DEFLATE bitstream itself comes from the un-vendored `flate2` crate in the reference
// (Compression::best()); here zlib level 9 raw deflate produces it.  Byte identity of that
// stream with flate2's miniz backend is NOT claimed (DESIGN.md "parity unpinned: archive bytes");
// header bytes and inflate(payload) are exact, and either side can read the other's files.
#include <cstdint>
#include <cstring>
#include <vector>
#include <zlib.h>

#include "../../include/hgi.h"

namespace {

void put_u32(uint8_t* p, uint32_t v) { for (int i = 0; i < 4; ++i) p[i] = (uint8_t)(v >> (8 * i)); }
void put_u64(uint8_t* p, uint64_t v) { for (int i = 0; i < 8; ++i) p[i] = (uint8_t)(v >> (8 * i)); }
uint32_t get_u32(const uint8_t* p) { uint32_t v = 0; for (int i = 0; i < 4; ++i) v |= (uint32_t)p[i] << (8 * i); return v; }
uint64_t get_u64(const uint8_t* p) { uint64_t v = 0; for (int i = 0; i < 8; ++i) v |= (uint64_t)p[i] << (8 * i); return v; }

}  // namespace

extern "C" {

size_t hgi_archive_bound(size_t n)
{
    // deflateBound for raw streams is n + n/1000-ish + small; be generous and simple.
    return HGI_ARCHIVE_HEADER_BYTES + (n + 16) + ((n + 16) >> 8) + 64 + 5 * (((n + 16) >> 14) + 1);
}

int hgi_archive_serialize(const hgi_metadata_t* m, const uint8_t* grid, size_t grid_len, uint64_t grid_width,
                          uint8_t* out, size_t out_capacity, size_t* out_len)
{
    if (!m || !out || !out_len || (grid_len && !grid)) return HGI_ERR_INVALID_ARG;
    if (m->quantization_level > 3 || m->interpolation > 2) return HGI_ERR_INVALID_ARG;
    if (out_capacity < HGI_ARCHIVE_HEADER_BYTES) return HGI_ERR_BUFFER_TOO_SMALL;
    put_u32(out + 0, HGI_ARCHIVE_MAGIC);               // src/archive.rs:32
    put_u32(out + 4, m->quantization_level);           // :33 bincode(Metadata), field order :16-22
    put_u32(out + 8, m->interpolation);
    put_u32(out + 12, m->width);
    put_u32(out + 16, m->height);
    put_u64(out + 20, m->scale_level);

    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    // raw deflate (no zlib header), best compression: DeflateEncoder::new(_, Compression::best()) (:36)
    if (deflateInit2(&zs, 9, Z_DEFLATED, -15, 9, Z_DEFAULT_STRATEGY) != Z_OK) return HGI_ERR_ALLOC;
    uint8_t lenb[8], widthb[8];
    put_u64(lenb, (uint64_t)grid_len);                 // :35 bincode(Grid): Vec<u8> length prefix
    put_u64(widthb, grid_width);                       //     then `width: usize` (src/grid.rs:2-5)
    zs.next_out = out + HGI_ARCHIVE_HEADER_BYTES;
    size_t out_left = out_capacity - HGI_ARCHIVE_HEADER_BYTES;
    auto feed = [&](const uint8_t* p, size_t n, int flush) -> int {
        size_t done = 0;
        do {
            const size_t in_now = (n - done) > 0x40000000u ? 0x40000000u : (n - done);
            zs.next_in = const_cast<Bytef*>(p + done);
            zs.avail_in = (uInt)in_now;
            const int fl = (done + in_now == n) ? flush : Z_NO_FLUSH;
            int zr;
            do {
                const size_t out_now = out_left > 0x40000000u ? 0x40000000u : out_left;
                if (out_now == 0) return HGI_ERR_BUFFER_TOO_SMALL;
                zs.avail_out = (uInt)out_now;
                zr = deflate(&zs, fl);
                if (zr == Z_STREAM_ERROR) return HGI_ERR_INVALID_ARG;
                out_left -= out_now - zs.avail_out;
            } while (zs.avail_out == 0 || (fl == Z_FINISH && zr != Z_STREAM_END));
            done += in_now;
        } while (done < n);
        return HGI_OK;
    };
    int rc = feed(lenb, 8, Z_NO_FLUSH);
    if (rc == HGI_OK && grid_len) rc = feed(grid, grid_len, Z_NO_FLUSH);
    if (rc == HGI_OK) rc = feed(widthb, 8, Z_FINISH);
    deflateEnd(&zs);
    if (rc != HGI_OK) return rc;
    *out_len = out_capacity - out_left;
    return HGI_OK;
}

int hgi_archive_read_header(const uint8_t* data, size_t len, hgi_metadata_t* m)
{
    if (!data || !m) return HGI_ERR_INVALID_ARG;
    if (len < 4) return HGI_ERR_TRUNCATED;
    if (get_u32(data) != HGI_ARCHIVE_MAGIC) return HGI_ERR_BAD_MAGIC;   // src/archive.rs:47-50
    if (len < HGI_ARCHIVE_HEADER_BYTES) return HGI_ERR_TRUNCATED;
    m->quantization_level = get_u32(data + 4);                          // :51
    m->interpolation = get_u32(data + 8);
    m->width = get_u32(data + 12);
    m->height = get_u32(data + 16);
    m->scale_level = get_u64(data + 20);
    // bincode rejects unknown enum variants
    if (m->quantization_level > 3 || m->interpolation > 2) return HGI_ERR_TRUNCATED;
    return HGI_OK;
}

int hgi_archive_read_grid(const uint8_t* data, size_t len, uint8_t* grid_out, size_t grid_capacity,
                          size_t* grid_len_out, uint64_t* grid_width_out)
{
    if (!data || !grid_len_out) return HGI_ERR_INVALID_ARG;
    hgi_metadata_t m;
    int rc = hgi_archive_read_header(data, len, &m);
    if (rc) return rc;
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return HGI_ERR_ALLOC;            // src/archive.rs:52 DeflateDecoder
    const uint8_t* in = data + HGI_ARCHIVE_HEADER_BYTES;
    size_t in_left = len - HGI_ARCHIVE_HEADER_BYTES;
    bool ended = false;
    auto pull = [&](uint8_t* dst, size_t n) -> int {
        size_t got = 0;
        while (got < n) {
            if (ended) return HGI_ERR_TRUNCATED;
            const size_t in_now = in_left > 0x40000000u ? 0x40000000u : in_left;
            const size_t out_now = (n - got) > 0x40000000u ? 0x40000000u : (n - got);
            zs.next_in = const_cast<Bytef*>(in);
            zs.avail_in = (uInt)in_now;
            zs.next_out = dst + got;
            zs.avail_out = (uInt)out_now;
            const int zr = inflate(&zs, Z_NO_FLUSH);
            const size_t used = in_now - zs.avail_in, made = out_now - zs.avail_out;
            in += used;
            in_left -= used;
            got += made;
            if (zr == Z_STREAM_END) ended = true;
            else if (zr != Z_OK && zr != Z_BUF_ERROR) return HGI_ERR_TRUNCATED;
            else if (used == 0 && made == 0) return HGI_ERR_TRUNCATED;   // no progress: input exhausted
        }
        return HGI_OK;
    };
    uint8_t b8[8];
    rc = pull(b8, 8);                                                     // :53 bincode Vec<u8> length
    uint64_t glen = 0;
    if (rc == HGI_OK) {
        glen = get_u64(b8);
        *grid_len_out = (size_t)glen;
        if (!grid_out || glen > grid_capacity) rc = HGI_ERR_BUFFER_TOO_SMALL;
    }
    if (rc == HGI_OK) rc = pull(grid_out, (size_t)glen);
    if (rc == HGI_OK) rc = pull(b8, 8);
    if (rc == HGI_OK && grid_width_out) *grid_width_out = get_u64(b8);
    inflateEnd(&zs);
    return rc;
}

}  // extern "C"
