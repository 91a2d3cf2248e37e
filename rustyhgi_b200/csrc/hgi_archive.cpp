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
#include <algorithm>
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

// ---- Huffman-only DEFLATE driven by GPU-built frequency tables --------------------------------------------
// north_star: "archive.rs residue histogram / frequency-table construction [on the GPU] ... the entropy
// bitstream stays [on the host] after the GPU-built tables".  Each block of residual bytes becomes one
// dynamic-Huffman DEFLATE block (RFC 1951 3.2.7) whose literal code is built from that block's 256-bin
// histogram -- no LZ77 matching, so the host only bit-packs.  Any inflate (zlib, flate2/miniz) reads it.
class BitWriter {
public:
    BitWriter(uint8_t* out, size_t cap) : p_(out), end_(out + cap) {}
    inline void put(uint32_t bits, int n)   // LSB-first, n <= 32
    {
        acc_ |= (uint64_t)bits << nbits_;
        nbits_ += n;
        while (nbits_ >= 8) {
            if (p_ < end_) *p_ = (uint8_t)acc_; else overflow_ = true;
            ++p_;
            acc_ >>= 8;
            nbits_ -= 8;
        }
    }
    void finish() { if (nbits_ > 0) put(0, 8 - nbits_); }
    bool overflow() const { return overflow_; }
    uint8_t* pos() const { return p_; }
private:
    uint8_t *p_, *end_;
    uint64_t acc_ = 0;
    int nbits_ = 0;
    bool overflow_ = false;
};

// Huffman code lengths (<= max_len) for `n` symbols; zero-frequency symbols get length 0.  Plain heap
// construction; if the tree is too deep the small counts are flattened and it is rebuilt (rarely needed).
void huffman_lengths(const uint64_t* freq, int n, int max_len, uint8_t* len_out)
{
    std::vector<uint64_t> f(freq, freq + n);
    for (;;) {
        std::vector<int> parent(2 * n, -1);
        std::vector<uint64_t> w(2 * n, 0);
        std::vector<int> heap;
        auto less = [&](int a, int b) { return w[a] > w[b] || (w[a] == w[b] && a > b); };
        for (int i = 0; i < n; ++i)
            if (f[i]) { w[i] = f[i]; heap.push_back(i); }
        std::fill(len_out, len_out + n, 0);
        if (heap.empty()) return;
        if (heap.size() == 1) { len_out[heap[0]] = 1; return; }
        std::make_heap(heap.begin(), heap.end(), less);
        int next = n;
        while (heap.size() > 1) {
            std::pop_heap(heap.begin(), heap.end(), less); const int a = heap.back(); heap.pop_back();
            std::pop_heap(heap.begin(), heap.end(), less); const int b = heap.back(); heap.pop_back();
            w[next] = w[a] + w[b];
            parent[a] = parent[b] = next;
            heap.push_back(next);
            std::push_heap(heap.begin(), heap.end(), less);
            ++next;
        }
        int deepest = 0;
        for (int i = 0; i < n; ++i) {
            if (!f[i]) continue;
            int d = 0;
            for (int v = i; parent[v] >= 0; v = parent[v]) ++d;
            len_out[i] = (uint8_t)d;
            if (d > deepest) deepest = d;
        }
        if (deepest <= max_len) return;
        for (int i = 0; i < n; ++i)
            if (f[i]) f[i] = (f[i] >> 2) + 1;   // flatten and retry
    }
}

// Canonical codes (RFC 1951 3.2.2), bit-reversed for the LSB-first stream.
void canonical_codes(const uint8_t* len, int n, uint16_t* code_out)
{
    int bl_count[16] = {0};
    for (int i = 0; i < n; ++i) bl_count[len[i]]++;
    bl_count[0] = 0;
    int next_code[16] = {0}, code = 0;
    for (int b = 1; b <= 15; ++b) { code = (code + bl_count[b - 1]) << 1; next_code[b] = code; }
    for (int i = 0; i < n; ++i) {
        if (!len[i]) { code_out[i] = 0; continue; }
        int c = next_code[len[i]]++, r = 0;
        for (int b = 0; b < len[i]; ++b) { r = (r << 1) | (c & 1); c >>= 1; }
        code_out[i] = (uint16_t)r;
    }
}

// One dynamic-Huffman block holding `pre` + `data` + `post` as literals, code built from `freq[257]`.
void write_huffman_block(BitWriter& bw, const uint64_t freq[257], const uint8_t* pre, size_t npre, const uint8_t* data,
                         size_t n, const uint8_t* post, size_t npost, bool final_block)
{
    uint8_t lit_len[257];
    uint16_t lit_code[257];
    huffman_lengths(freq, 257, 15, lit_len);
    canonical_codes(lit_len, 257, lit_code);
    // lengths to transmit: 257 literal/length codes + 1 distance code (length 1: "one distance code ... using one bit")
    uint8_t seq[258];
    std::memcpy(seq, lit_len, 257);
    seq[257] = 1;
    // run-length encode with symbols 16/17/18 (3.2.7)
    struct Tok { uint8_t sym, extra_bits; uint16_t extra; };
    std::vector<Tok> toks;
    for (int i = 0; i < 258;) {
        const uint8_t v = seq[i];
        int run = 1;
        while (i + run < 258 && seq[i + run] == v) ++run;
        int left = run;
        if (v == 0) {
            while (left >= 11) { const int r = left > 138 ? 138 : left; toks.push_back({18, 7, (uint16_t)(r - 11)}); left -= r; }
            if (left >= 3) { toks.push_back({17, 3, (uint16_t)(left - 3)}); left = 0; }
            while (left-- > 0) toks.push_back({0, 0, 0});
        } else {
            toks.push_back({v, 0, 0});
            --left;
            while (left >= 3) { const int r = left > 6 ? 6 : left; toks.push_back({16, 2, (uint16_t)(r - 3)}); left -= r; }
            while (left-- > 0) toks.push_back({v, 0, 0});
        }
        i += run;
    }
    uint64_t cl_freq[19] = {0};
    for (const Tok& t : toks) cl_freq[t.sym]++;
    uint8_t cl_len[19];
    uint16_t cl_code[19];
    huffman_lengths(cl_freq, 19, 7, cl_len);
    canonical_codes(cl_len, 19, cl_code);
    static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19;
    while (hclen > 4 && cl_len[order[hclen - 1]] == 0) --hclen;
    bw.put(final_block ? 1u : 0u, 1);   // BFINAL
    bw.put(2u, 2);                      // BTYPE = 10 (dynamic Huffman)
    bw.put(257 - 257, 5);               // HLIT
    bw.put(1 - 1, 5);                   // HDIST
    bw.put((uint32_t)(hclen - 4), 4);   // HCLEN
    for (int i = 0; i < hclen; ++i) bw.put(cl_len[order[i]], 3);
    for (const Tok& t : toks) {
        bw.put(cl_code[t.sym], cl_len[t.sym]);
        if (t.extra_bits) bw.put(t.extra, t.extra_bits);
    }
    auto emit = [&](const uint8_t* p, size_t m) { for (size_t i = 0; i < m; ++i) bw.put(lit_code[p[i]], lit_len[p[i]]); };
    emit(pre, npre);
    emit(data, n);
    emit(post, npost);
    bw.put(lit_code[256], lit_len[256]);   // end of block
}

// ---- run-length DEFLATE: literals + distance-1 matches, tables from the GPU ---------------------------------------
// The parse (which bytes become literals, which runs become matches) is a pure function of the data and MUST equal
// the one hgi_rle_hist_kernel counted (hgi_reduce_kernels.cu): 512-byte segments relative to the block start; in a
// segment every maximal run of a byte b: literal b, then matches of min(258, rest) while rest >= 3, then `rest`
// literals b.  RFC 1951 3.2.5 length symbols; the only distance is 1 (distance symbol 0, a one-bit code).
constexpr size_t kRleSeg = HGI_RLE_SEGMENT_BYTES;
inline uint32_t rle_len_sym(uint32_t len, uint32_t* extra_bits, uint32_t* extra)
{
    if (len == 258u) { *extra_bits = 0; *extra = 0; return 285u; }
    const uint32_t l = len - 3u;
    uint32_t e = 0;
    if (l >= 8u) { e = 0; while ((l >> (e + 3)) != 0) ++e; }   // floor(log2(l)) - 2
    *extra_bits = e;
    *extra = l & ((1u << e) - 1u);
    return 257u + 4u * e + (l >> e);
}

// One dynamic-Huffman block: `pre` literals, the RLE parse of data[0, n), `post` literals; code from freq[286]
// (the GPU table of this block plus the pre/post bytes and the end-of-block symbol).  Returns false when the table
// does not cover a symbol the parse needs (i.e. it was not built from this data).
bool write_rle_block(BitWriter& bw, const uint64_t freq[286], const uint8_t* pre, size_t npre, const uint8_t* data, size_t n,
                     const uint8_t* post, size_t npost, bool final_block)
{
    uint8_t lit_len[286];
    uint16_t lit_code[286];
    huffman_lengths(freq, 286, 15, lit_len);
    canonical_codes(lit_len, 286, lit_code);
    int hlit = 286;
    while (hlit > 257 && lit_len[hlit - 1] == 0) --hlit;
    uint8_t seq[287];
    std::memcpy(seq, lit_len, (size_t)hlit);
    seq[hlit] = 1;   // one distance code (distance 1), "encoded using one bit" (RFC 1951 3.2.7)
    const int nseq = hlit + 1;
    struct Tok { uint8_t sym, extra_bits; uint16_t extra; };
    std::vector<Tok> toks;
    for (int i = 0; i < nseq;) {
        const uint8_t v = seq[i];
        int run = 1;
        while (i + run < nseq && seq[i + run] == v) ++run;
        int left = run;
        if (v == 0) {
            while (left >= 11) { const int r = left > 138 ? 138 : left; toks.push_back({18, 7, (uint16_t)(r - 11)}); left -= r; }
            if (left >= 3) { toks.push_back({17, 3, (uint16_t)(left - 3)}); left = 0; }
            while (left-- > 0) toks.push_back({0, 0, 0});
        } else {
            toks.push_back({v, 0, 0});
            --left;
            while (left >= 3) { const int r = left > 6 ? 6 : left; toks.push_back({16, 2, (uint16_t)(r - 3)}); left -= r; }
            while (left-- > 0) toks.push_back({v, 0, 0});
        }
        i += run;
    }
    uint64_t cl_freq[19] = {0};
    for (const Tok& t : toks) cl_freq[t.sym]++;
    uint8_t cl_len[19];
    uint16_t cl_code[19];
    huffman_lengths(cl_freq, 19, 7, cl_len);
    canonical_codes(cl_len, 19, cl_code);
    static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19;
    while (hclen > 4 && cl_len[order[hclen - 1]] == 0) --hclen;
    bw.put(final_block ? 1u : 0u, 1);   // BFINAL
    bw.put(2u, 2);                      // BTYPE = 10 (dynamic Huffman)
    bw.put((uint32_t)(hlit - 257), 5);  // HLIT
    bw.put(1 - 1, 5);                   // HDIST
    bw.put((uint32_t)(hclen - 4), 4);   // HCLEN
    for (int i = 0; i < hclen; ++i) bw.put(cl_len[order[i]], 3);
    for (const Tok& t : toks) {
        bw.put(cl_code[t.sym], cl_len[t.sym]);
        if (t.extra_bits) bw.put(t.extra, t.extra_bits);
    }
    bool ok = true;
    auto lit = [&](uint32_t b) { ok &= lit_len[b] != 0; bw.put(lit_code[b], lit_len[b]); };
    for (size_t i = 0; i < npre; ++i) lit(pre[i]);
    for (size_t s0 = 0; s0 < n; s0 += kRleSeg) {
        const size_t s1 = s0 + kRleSeg < n ? s0 + kRleSeg : n;
        size_t i = s0;
        while (i < s1) {
            const uint8_t b = data[i];
            size_t j = i + 1;
            while (j < s1 && data[j] == b) ++j;
            lit(b);
            uint32_t rem = (uint32_t)(j - i - 1);
            while (rem >= 3u) {
                const uint32_t m = rem < 258u ? rem : 258u;
                uint32_t eb, ev;
                const uint32_t sym = rle_len_sym(m, &eb, &ev);
                ok &= lit_len[sym] != 0;
                bw.put(lit_code[sym], lit_len[sym]);
                if (eb) bw.put(ev, (int)eb);
                bw.put(0u, 1);          // distance symbol 0 (distance 1), the one-bit code
                rem -= m;
            }
            for (; rem; --rem) lit(b);
            i = j;
        }
    }
    for (size_t i = 0; i < npost; ++i) lit(post[i]);
    bw.put(lit_code[256], lit_len[256]);   // end of block
    return ok;
}

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

size_t hgi_archive_huffman_bound(size_t n, size_t n_blocks)
{
    // <= 15 bits per literal, <= ~330 bytes of table per block, 16 bincode bytes, header
    return HGI_ARCHIVE_HEADER_BYTES + 2 * (n + 16) + 512 * (n_blocks ? n_blocks : 1) + 64;
}

int hgi_archive_serialize_huffman(const hgi_metadata_t* m, const uint8_t* grid, size_t grid_len, uint64_t grid_width,
                                  const uint32_t* hist, size_t n_blocks, size_t block_bytes, uint8_t* out,
                                  size_t out_capacity, size_t* out_len)
{
    if (!m || !out || !out_len || !hist || n_blocks == 0 || (grid_len && !grid)) return HGI_ERR_INVALID_ARG;
    if (m->quantization_level > 3 || m->interpolation > 2) return HGI_ERR_INVALID_ARG;
    if (n_blocks > 1 && (block_bytes == 0 || (n_blocks - 1) * block_bytes >= grid_len || n_blocks * block_bytes < grid_len))
        return HGI_ERR_INVALID_ARG;
    if (out_capacity < HGI_ARCHIVE_HEADER_BYTES) return HGI_ERR_BUFFER_TOO_SMALL;
    put_u32(out + 0, HGI_ARCHIVE_MAGIC);
    put_u32(out + 4, m->quantization_level);
    put_u32(out + 8, m->interpolation);
    put_u32(out + 12, m->width);
    put_u32(out + 16, m->height);
    put_u64(out + 20, m->scale_level);
    uint8_t lenb[8], widthb[8];
    put_u64(lenb, (uint64_t)grid_len);       // bincode(Grid): Vec<u8> length prefix, bytes, then `width: usize`
    put_u64(widthb, grid_width);
    BitWriter bw(out + HGI_ARCHIVE_HEADER_BYTES, out_capacity - HGI_ARCHIVE_HEADER_BYTES);
    for (size_t b = 0; b < n_blocks; ++b) {
        const size_t lo = n_blocks == 1 ? 0 : b * block_bytes;
        const size_t hi = (n_blocks == 1 || b + 1 == n_blocks) ? grid_len : lo + block_bytes;
        uint64_t freq[257];
        uint64_t total = 0;
        for (int i = 0; i < 256; ++i) { freq[i] = hist[b * 256 + i]; total += freq[i]; }
        if (total != hi - lo) return HGI_ERR_INVALID_ARG;   // the table must describe exactly this block
        freq[256] = 1;                                       // end-of-block
        const bool first = (b == 0), last = (b + 1 == n_blocks);
        if (first) for (int i = 0; i < 8; ++i) freq[lenb[i]]++;
        if (last) for (int i = 0; i < 8; ++i) freq[widthb[i]]++;
        write_huffman_block(bw, freq, lenb, first ? 8 : 0, grid + lo, hi - lo, widthb, last ? 8 : 0, last);
    }
    bw.finish();
    if (bw.overflow()) return HGI_ERR_BUFFER_TOO_SMALL;
    *out_len = (size_t)(bw.pos() - out);
    return HGI_OK;
}

int hgi_archive_serialize_rle(const hgi_metadata_t* m, const uint8_t* grid, size_t grid_len, uint64_t grid_width,
                              const uint32_t* hist, size_t n_blocks, size_t block_bytes, uint8_t* out,
                              size_t out_capacity, size_t* out_len)
{
    if (!m || !out || !out_len || !hist || n_blocks == 0 || (grid_len && !grid)) return HGI_ERR_INVALID_ARG;
    if (m->quantization_level > 3 || m->interpolation > 2) return HGI_ERR_INVALID_ARG;
    if (n_blocks > 1 && (block_bytes == 0 || block_bytes % kRleSeg != 0 || (n_blocks - 1) * block_bytes >= grid_len ||
                         n_blocks * block_bytes < grid_len))
        return HGI_ERR_INVALID_ARG;
    if (out_capacity < HGI_ARCHIVE_HEADER_BYTES) return HGI_ERR_BUFFER_TOO_SMALL;
    put_u32(out + 0, HGI_ARCHIVE_MAGIC);
    put_u32(out + 4, m->quantization_level);
    put_u32(out + 8, m->interpolation);
    put_u32(out + 12, m->width);
    put_u32(out + 16, m->height);
    put_u64(out + 20, m->scale_level);
    uint8_t lenb[8], widthb[8];
    put_u64(lenb, (uint64_t)grid_len);       // bincode(Grid): Vec<u8> length prefix, bytes, then `width: usize`
    put_u64(widthb, grid_width);
    BitWriter bw(out + HGI_ARCHIVE_HEADER_BYTES, out_capacity - HGI_ARCHIVE_HEADER_BYTES);
    for (size_t b = 0; b < n_blocks; ++b) {
        const size_t lo = n_blocks == 1 ? 0 : b * block_bytes;
        const size_t hi = (n_blocks == 1 || b + 1 == n_blocks) ? grid_len : lo + block_bytes;
        uint64_t freq[286];
        for (int i = 0; i < 286; ++i) freq[i] = hist[b * HGI_RLE_TABLE_SYMBOLS + i];
        freq[256] = 1;                                       // end-of-block
        const bool first = (b == 0), last = (b + 1 == n_blocks);
        if (first) for (int i = 0; i < 8; ++i) freq[lenb[i]]++;
        if (last) for (int i = 0; i < 8; ++i) freq[widthb[i]]++;
        if (!write_rle_block(bw, freq, lenb, first ? 8 : 0, grid + lo, hi - lo, widthb, last ? 8 : 0, last))
            return HGI_ERR_INVALID_ARG;                      // the table was not built from this block
    }
    bw.finish();
    if (bw.overflow()) return HGI_ERR_BUFFER_TOO_SMALL;
    *out_len = (size_t)(bw.pos() - out);
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
        // a deflate stream expands at most 1032x: a length prefix beyond that (or beyond width * height of the
        // header when that is known) cannot be honest -- refuse before anybody allocates for it
        const uint64_t by_ratio = (uint64_t)(len - HGI_ARCHIVE_HEADER_BYTES) * 1032u + 1032u;
        const uint64_t by_dims = (uint64_t)m.width * m.height;
        if (glen > by_ratio || (by_dims != 0 && glen > by_dims)) rc = HGI_ERR_TRUNCATED;
        else if (glen > 0 && (!grid_out || glen > grid_capacity)) rc = HGI_ERR_BUFFER_TOO_SMALL;
    }
    if (rc == HGI_OK && glen) rc = pull(grid_out, (size_t)glen);
    if (rc == HGI_OK) rc = pull(b8, 8);
    if (rc == HGI_OK && grid_width_out) *grid_width_out = get_u64(b8);
    inflateEnd(&zs);
    return rc;
}

}  // extern "C"
