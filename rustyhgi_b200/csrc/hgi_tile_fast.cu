// hgi_tile_fast.cu -- the fast fused HGI tile kernel (sm_100a): D == 1 passes on planes whose rows
// are 16-byte aligned.  Same pass/tile/halo decomposition as hgi_tile_kernels.cu (the generic
// kernel, which still serves D > 1 passes and unaligned planes), restructured so the ALU work per
// pixel drops ~4x:
//
//  * every level s in {2,4,8,16} of the tile lives in its own DENSE shared-memory plane P_s
//    (lattice-s points only, one byte each), so each level is "the finest level of a half-size
//    image": even rows read [R a R a ...], odd rows [a a a a ...] -- natural 16-bit-lane SWAR;
//  * a level reads its corners (and the coarser symbols) from the coarser plane P_2s / Q_2s and
//    writes complete words into P_s / Q_s, so no separate "insert" step exists;
//  * the finest level never touches shared memory for pixels: each thread loads its own 16x2
//    pixels with two 128-bit global loads at kernel start (a software prefetch that is in flight
//    while the coarse levels run), and stores finished grid / image words straight to HBM;
//  * two pixels per 32-bit register: averages, residuals, the Linear quantizer (an exact
//    multiply-shift, checked on the host against src/quantizator.rs:50-60 for all 256 inputs) and
//    the overflow fix-up (src/encoder.rs:56-60) are all 16-bit-lane SWAR.
//
// Reference semantics: src/encoder.rs:39-71, src/decoder.rs:18-46, src/utils.rs:11-41,
// src/interpolator.rs:15-28,41-91, src/quantizator.rs:41-74.
#include "hgi_device.cuh"
#include "hgi_kernels.h"

namespace hgi {

namespace {

#ifndef HGI_FAST_TILE_H
#define HGI_FAST_TILE_H 128
#endif
constexpr int TW = 128;                  // tile width  (lattice points)
constexpr int TH = HGI_FAST_TILE_H;      // tile height: 64 or 128 (NU = TH/64 16x2 units per thread)
constexpr int NT = 256;
constexpr int NU = TH / 64;
constexpr int NWARPS = NT / 32;
constexpr int FMAX = 1 << kMaxPassLevels;
constexpr uint32_t M16 = 0x00FF00FFu;

static_assert(TW == 128 && (TH == 64 || TH == 128) && NT == 256, "thread mapping assumes 128-wide tiles, 256 threads");

// Dense level planes.  P_s holds lattice-s points of the tile + halo: columns 0..TW/s+1,
// rows 0..TH/s+1 (the last ones are only partially needed, see need_limit).
__host__ __device__ constexpr int plane_pitch(int s) { return s == 2 ? 96 : (s == 4 ? 48 : (s == 8 ? 32 : 16)); }
__host__ __device__ constexpr int plane_rows(int s) { return TH / s + 2; }
__host__ __device__ constexpr int plane_bytes(int s) { return plane_rows(s) * plane_pitch(s); }
__host__ __device__ constexpr int plane_off(int s)
{
    return s == 2 ? 0 : (s == 4 ? plane_bytes(2) : (s == 8 ? plane_bytes(2) + plane_bytes(4)
                                                               : plane_bytes(2) + plane_bytes(4) + plane_bytes(8)));
}
constexpr int PLANE_BYTES = plane_bytes(2) + plane_bytes(4) + plane_bytes(8) + plane_bytes(16);  // 4544 (TH=64) / 8704 (TH=128)

__device__ __forceinline__ int need_limit(int tile_extent, int s)
{
    return s == 1 ? tile_extent - 1 : (s == 2 ? tile_extent : tile_extent + s);
}

// ---- 16-bit-lane SWAR primitives (two pixels per register) -----------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
// bytes (i, i+1) of w -> lanes
__device__ __forceinline__ uint32_t lanes01(uint32_t w) { return prmt(w, 0u, 0x4140u); }
__device__ __forceinline__ uint32_t lanes12(uint32_t w) { return prmt(w, 0u, 0x4241u); }
__device__ __forceinline__ uint32_t lanes23(uint32_t w) { return prmt(w, 0u, 0x4342u); }
// bytes (0,2) / (1,3) of w -> lanes
__device__ __forceinline__ uint32_t lanes_even(uint32_t w) { return w & M16; }
__device__ __forceinline__ uint32_t lanes_odd(uint32_t w) { return prmt(w, 0u, 0x4341u); }
// lanes (lo0, lo1), (hi0, hi1) -> bytes lo0 hi0 lo1 hi1
__device__ __forceinline__ uint32_t interleave(uint32_t even_lanes, uint32_t odd_lanes)
{
    return even_lanes + (odd_lanes << 8);
}
__device__ __forceinline__ uint32_t avg2(uint32_t x, uint32_t y)   // src/interpolator.rs:44 per lane
{
    return ((x + y + 0x00010001u) >> 1) & M16;
}
// src/interpolator.rs:43-54 per lane.  With avg(x,y) = (x+y+1)>>1 = (x + y + ((x^y)&1)) / 2, the sum of the
// four edge averages is T + E/2, T = A+B+C+D, where E counts the edges of the cycle A-B-D-C-A whose
// endpoints differ in parity (0, 2 or 4).  Working through floor((T + E/2)/4) by the parity of T gives
//     pred = (((T + 1) >> 1) + w) >> 1,    w = (A^B) & (C^D) & (A^C) & 1
// (E/2 only matters when it makes T+E/2 cross a multiple of 4: T odd -> the +1; T = 2 mod 4 with all
// three parity tests true -> the w).  Checked against the four-average form in tests/test_swar_model.py.
// 8 ALU-pipe operations per register (two cells) instead of 16.
template <int INTERP>
__device__ __forceinline__ uint32_t pred2(uint32_t A, uint32_t B, uint32_t C, uint32_t D)
{
    if (INTERP == kInterpLeftTop) return A;                                    // src/interpolator.rs:26
    const uint32_t x1 = (A ^ B) & 0x00010001u;
    const uint32_t w = x1 & (C ^ D) & (A ^ C);
    const uint32_t h = ((A + B + C) + (D + 0x00010001u)) >> 1;                 // lanes <= 510 (+ stray bit 15)
    return ((h + w) >> 1) & M16;                                               // :51
}

// Linear quantizer as an exact per-lane multiply-shift: ((d + e) / scale) * scale for d in 0..255.
struct QuantSwar {
    uint32_t mul, add, shift, scale;
    uint32_t rmask, qmul;   // q = umulhi(t & rmask, qmul): rmask = 0xF << shift per lane, qmul = scale << (32 - shift)
};
__host__ __device__ inline QuantSwar quant_swar(uint32_t error)
{
    // (k, c, n) with ((x*k + c) >> n) == x / (2e+1) for all x in [e, 255+e] and x*k + c < 2^16
    uint32_t k = 0, c = 0, n = 0;
    if (error == 10) { k = 195; c = 195; n = 12; }
    else if (error == 20) { k = 25; c = 0; n = 10; }
    else if (error == 30) { k = 67; c = 67; n = 12; }
    QuantSwar q;
    q.mul = k;
    q.add = (error * k + c) * 0x00010001u;
    q.shift = n;
    q.scale = 2 * error + 1;
    q.rmask = (0xFu << n) * 0x00010001u;
    q.qmul = n ? (q.scale << (32 - n)) : 0u;
    return q;
}

// src/encoder.rs:52-64 for two pixels.  Returns the symbols; `recon` = what the decoder rebuilds.
template <bool IDENTITY>
__device__ __forceinline__ uint32_t encode2(uint32_t a, uint32_t p, uint32_t pk, const QuantSwar& qc, uint32_t& recon)
{
    const uint32_t dd = a + pk;                       // pk = 0x01000100 - p: per lane a + 256 - p, bit 8 = [a >= p]
    const uint32_t d = dd & M16;                      // :53 wrapping_sub
    if (IDENTITY) {
        recon = a;                                    // p + (a - p) == a
        return d;
    }
    const uint32_t t = d * qc.mul + qc.add;           // lanes: (d + e) * k + c  < 2^16
    // r = t >> shift per lane, q = r * scale -- done as one high multiply on the masked quotient bits, which
    // moves the shift off the ALU pipe: ((r << n) * (scale << (32 - n))) >> 32 == r * scale in both lanes
    uint32_t q = __umulhi(t & qc.rmask, qc.qmul);     // :54 table[d]
    const uint32_t ov = q + p;                        // bit 8 = overflow                      (:56)
    // overflow_is_expected = [a < p] = !bit8(dd)  =>  mismatch iff bit8(ov) == bit8(dd)       (:57-58)
    const uint32_t x = ~(ov ^ dd) & 0x01000100u;
    const uint32_t m = __umulhi(x, 0xFF000000u);      // (x * 255) >> 8: 0x00FF in every mismatching lane
    q = (q & ~m) | (d & m);                           // :59
    recon = ((ov & ~m) | (a & m)) & M16;              // :63 (p + q) mod 256, == a after a fix-up
    return q;
}

__device__ __forceinline__ uint32_t decode2(uint32_t g, uint32_t p) { return (p + g) & M16; }  // src/decoder.rs:39

__device__ __forceinline__ uint32_t valid_mask(int col0, int row, int xin_s, int yin_s)
{
    const int n = xin_s - col0;
    if (row >= yin_s || n <= 0) return 0u;
    return n >= 4 ? 0xFFFFFFFFu : ((1u << (8 * n)) - 1u);
}

struct FastSmem {
    alignas(16) uint8_t P[PLANE_BYTES];   // encode: pixels -> reconstruction; decode: residuals -> pixels
    alignas(16) uint8_t Q[PLANE_BYTES];   // encode: residual symbols
};

// Stage one 16-byte chunk (columns 16c..16c+15 of tile row y, y even) into the dense planes of the
// levels that are computed in this pass (s < F).  c == 8 is the right-halo chunk (x = TW..TW+15).
template <int F>
__device__ __forceinline__ void stage_chunk(uint8_t* P, const uint4 v, int y, int c)
{
    if (F > 2 && y <= TH) {
        uint8_t* row = P + plane_off(2) + (y >> 1) * plane_pitch(2);
        if (c < 8)
            *reinterpret_cast<uint2*>(row + 8 * c) = make_uint2(prmt(v.x, v.y, 0x6420u), prmt(v.z, v.w, 0x6420u));
        else
            row[64] = (uint8_t)v.x;
    }
    if (F > 4 && (y & 3) == 0 && y <= TH + 4) {
        uint8_t* row = P + plane_off(4) + (y >> 2) * plane_pitch(4);
        if (c < 8)
            *reinterpret_cast<uint32_t*>(row + 4 * c) = prmt(prmt(v.x, v.y, 0x0040u), prmt(v.z, v.w, 0x0040u), 0x5410u);
        else
            *reinterpret_cast<uint16_t*>(row + 32) = (uint16_t)prmt(v.x, v.y, 0x0040u);
    }
    if (F > 8 && (y & 7) == 0 && y <= TH + 8) {
        uint8_t* row = P + plane_off(8) + (y >> 3) * plane_pitch(8);
        *reinterpret_cast<uint16_t*>(row + 2 * c) = (uint16_t)prmt(v.x, v.z, 0x0040u);
    }
}

// One word (two cells, four plane columns) of a coarse level s >= 2, SWAR.
template <int MODE, int INTERP, bool IDENTITY, int S>
__device__ __forceinline__ void level_word(FastSmem& sm, int g, int cy, const QuantSwar& qc, bool edge,
                                           int xin_s, int yin_s)
{
    constexpr int ps = plane_pitch(S), pc = plane_pitch(2 * S);
    uint8_t* Ps = sm.P + plane_off(S);
    const uint8_t* Pc = sm.P + plane_off(2 * S);
    const uint8_t* ct = Pc + cy * pc + 2 * g;
    const uint32_t cwt = (uint32_t)*reinterpret_cast<const uint16_t*>(ct) | ((uint32_t)ct[2] << 16);
    const uint32_t cwb = (uint32_t)*reinterpret_cast<const uint16_t*>(ct + pc) | ((uint32_t)ct[pc + 2] << 16);
    const uint32_t A = lanes01(cwt), C = lanes12(cwt), B = lanes01(cwb), D = lanes12(cwb);
    const uint32_t p = pred2<INTERP>(A, B, C, D);
    uint32_t* pev = reinterpret_cast<uint32_t*>(Ps + (2 * cy) * ps + 4 * g);
    uint32_t* pod = reinterpret_cast<uint32_t*>(Ps + (2 * cy + 1) * ps + 4 * g);
    const uint32_t ev = *pev, od = *pod;
    const uint32_t a1 = lanes_odd(ev), a2 = lanes_even(od), a3 = lanes_odd(od);
    uint32_t r1, r2, r3;
    if (MODE == kModeEncode) {
        const uint32_t pk = 0x01000100u - p;
        const uint32_t q1 = encode2<IDENTITY>(a1, p, pk, qc, r1);
        const uint32_t q2 = encode2<IDENTITY>(a2, p, pk, qc, r2);
        const uint32_t q3 = encode2<IDENTITY>(a3, p, pk, qc, r3);
        uint8_t* Qs = sm.Q + plane_off(S);
        const uint8_t* Qc = sm.Q + plane_off(2 * S);
        const uint32_t QA = lanes01((uint32_t)*reinterpret_cast<const uint16_t*>(Qc + cy * pc + 2 * g));
        *reinterpret_cast<uint32_t*>(Qs + (2 * cy) * ps + 4 * g) = interleave(QA, q1);
        *reinterpret_cast<uint32_t*>(Qs + (2 * cy + 1) * ps + 4 * g) = interleave(q2, q3);
    } else {
        r1 = decode2(a1, p);
        r2 = decode2(a2, p);
        r3 = decode2(a3, p);
    }
    uint32_t wev = interleave(A, r1), wod = interleave(r2, r3);
    if (edge) {   // out-of-image reconstruction must read as 0 (src/interpolator.rs:75-82)
        wev &= valid_mask(4 * g, 2 * cy, xin_s, yin_s);
        wod &= valid_mask(4 * g, 2 * cy + 1, xin_s, yin_s);
    }
    *pev = wev;
    *pod = wod;
}

// One fringe cell (extra cell column / row right of and below the tile) of a coarse level, scalar.
template <int MODE, int INTERP, bool IDENTITY, int S>
__device__ __forceinline__ void fringe_cell(FastSmem& sm, int cx, int cy, const QuantSwar& qc, int xin_s, int yin_s)
{
    constexpr int ps = plane_pitch(S), pc = plane_pitch(2 * S);
    constexpr int xlim = (S == 2 ? TW : TW + S) / S, ylim = (S == 2 ? TH : TH + S) / S;   // need_limit / S
    uint8_t* Ps = sm.P + plane_off(S);
    const uint8_t* Pc = sm.P + plane_off(2 * S);
    const uint32_t A = Pc[cy * pc + cx], C = Pc[cy * pc + cx + 1];
    const uint32_t B = Pc[(cy + 1) * pc + cx], D = Pc[(cy + 1) * pc + cx + 1];
    const uint32_t pred = predict<INTERP>(A, B, C, D);
    const int x0 = 2 * cx, y0 = 2 * cy;
    Ps[y0 * ps + x0] = (uint8_t)A;   // the coarser lattice point itself (already 0 when out of image)
    const int px[3] = {x0 + 1, x0, x0 + 1};
    const int py[3] = {y0, y0 + 1, y0 + 1};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int x = px[k], y = py[k];
        if (x > xlim || y > ylim || x >= xin_s || y >= yin_s) continue;
        uint8_t* r = &Ps[y * ps + x];
        if (MODE == kModeEncode) {
            const uint32_t a = *r;
            const uint32_t diff = (a - pred) & 0xFFu;
            uint32_t q = diff;
            if (!IDENTITY) {
                q = (((diff * qc.mul + (qc.add & 0xFFFFu)) >> qc.shift) & 0xFu) * qc.scale;
                if (((pred + q) > 255u) != ((pred + diff) > 255u)) q = diff;
            }
            *r = (uint8_t)((pred + q) & 0xFFu);       // symbols of fringe points are never output
        } else {
            *r = (uint8_t)((pred + *r) & 0xFFu);
        }
    }
}

// One coarse level (sub-step S >= 2) of the tile: SWAR words over the tile's own cells on the low
// threads, the fringe cells (cell column TW/(2S), cell row TH/(2S)) on the high threads.
template <int MODE, int INTERP, bool IDENTITY, int S>
__device__ __forceinline__ void coarse_level(FastSmem& sm, int tid, const QuantSwar& qc, bool edge, int xin, int yin)
{
    constexpr int wpr = TW / (4 * S);               // SWAR words per cell row (2 cells each)
    constexpr int ncy = TH / (2 * S), ncx = TW / (2 * S);
    constexpr int nfr = (ncy + 1) + ncx;
    const int xin_s = (xin + S - 1) / S, yin_s = (yin + S - 1) / S;
    for (int it = tid; it < wpr * ncy; it += NT)
        level_word<MODE, INTERP, IDENTITY, S>(sm, it % wpr, it / wpr, qc, edge, xin_s, yin_s);
    for (int it = NT - 1 - tid; it < nfr; it += NT) {
        const int cx = it <= ncy ? ncx : it - (ncy + 1);
        const int cy = it <= ncy ? it : ncy;
        if (2 * cx < xin_s && 2 * cy < yin_s)
            fringe_cell<MODE, INTERP, IDENTITY, S>(sm, cx, cy, qc, xin_s, yin_s);
        else
            (sm.P + plane_off(S))[(2 * cy) * plane_pitch(S) + 2 * cx] = 0;
    }
    __syncthreads();
}

#ifndef HGI_FAST_MIN_BLOCKS
#define HGI_FAST_MIN_BLOCKS 6
#endif

template <int MODE, int INTERP, bool IDENTITY, bool EXTRA, int NLEV>
__global__ void __launch_bounds__(NT, HGI_FAST_MIN_BLOCKS)
hgi_tile_fast_kernel(const PassArgs p)
{
    __shared__ FastSmem sm;
    __shared__ uint32_t whist[(MODE == kModeEncode && EXTRA) ? NWARPS * 256 : 1];
    constexpr int F = 1 << NLEV;

    const int tid = threadIdx.x;
    const uint32_t img = blockIdx.z;
    const uint32_t X0 = blockIdx.x * TW, Y0 = blockIdx.y * TH;
    const int xin = (int)min((uint32_t)(TW + FMAX + 1), p.w - X0);   // in-image extent of tile + halo
    const int yin = (int)min((uint32_t)(TH + FMAX + 1), p.h - Y0);
    const bool edge = (xin < TW + FMAX + 1) || (yin < TH + FMAX + 1);
    const size_t tile_off = ((size_t)img * p.h + Y0) * p.w + X0;     // CTA-uniform
    const uint8_t* __restrict__ tile = p.src + tile_off;
    const bool top = (p.c_recon == nullptr);
    const QuantSwar qc = quant_swar(p.quant_error);

    // ---- 1. global loads: this thread's NU 16x2-pixel units (kept in registers for the finest level) ----
    const int sx = tid & 7, ry = tid >> 3;          // 8 column strips x 32 row pairs (x NU units, 64 rows apart)
    const bool col_ok = 16 * sx < xin;
    const uint32_t toff = (uint32_t)(2 * ry) * p.w + (uint32_t)(16 * sx);   // tile-relative, fits 32 bits
    uint4 ev[NU], od[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        ev[u] = make_uint4(0u, 0u, 0u, 0u);
        od[u] = make_uint4(0u, 0u, 0u, 0u);
        const int y = 2 * ry + 64 * u;
        if (col_ok && y < yin) ev[u] = __ldg(reinterpret_cast<const uint4*>(tile + toff + (uint32_t)(64 * u) * p.w));
        if (col_ok && y + 1 < yin) od[u] = __ldg(reinterpret_cast<const uint4*>(tile + toff + (uint32_t)(64 * u + 1) * p.w));
    }

    // halo chunks (right of / below the tile) feed only the coarse planes; the upper half of the CTA
    // fetches them: TH/2 right-halo chunks (rows 0,2,..,TH-2; column TW) + rows TH, TH+4, TH+8 (chunks 0..8)
    constexpr int NRIGHT = TH / 2, NHALO = NRIGHT + 27;
    const int hj = tid - (NT - 128);
    int hy = 2 * hj, hc = 8;
    if (hj >= NRIGHT) {
        const int r = (hj - NRIGHT) / 9;
        hc = (hj - NRIGHT) - 9 * r;
        hy = TH + 4 * r;
    }
    const bool halo = NLEV > 1 && hj >= 0 && hj < NHALO;
    uint4 hv = make_uint4(0u, 0u, 0u, 0u);
    if (halo && hy < yin && 16 * hc < xin)
        hv = __ldg(reinterpret_cast<const uint4*>(tile + (uint32_t)hy * p.w + (uint32_t)(16 * hc)));

    // ---- 2. stage the dense coarse planes + the coarse lattice of this pass ---------------------
#pragma unroll
    for (int u = 0; u < NU; ++u) stage_chunk<F>(sm.P, ev[u], 2 * ry + 64 * u, sx);
    if (halo) stage_chunk<F>(sm.P, hv, hy, hc);
    {
        constexpr int ncx = TW / F + 2, ncy = TH / F + 2;
        constexpr int pf = plane_pitch(F);
        uint8_t* Pf = sm.P + plane_off(F);
        uint8_t* Qf = sm.Q + plane_off(F);
        for (int it = tid; it < ncx * ncy; it += NT) {
            const int cj = it / ncx, ci = it - cj * ncx;
            const int x = ci * F, y = cj * F;
            uint8_t rv = 0, qv = 0;
            if (x < xin && y < yin) {
                if (top) {   // src/encoder.rs:26-37 / src/decoder.rs:22-28: the seed is the source byte
                    rv = __ldg(tile + (uint32_t)y * p.w + (uint32_t)x);
                    qv = rv;
                } else {
                    const size_t co = (size_t)img * p.cw * p.ch + (size_t)((Y0 + y) >> NLEV) * p.cw + ((X0 + x) >> NLEV);
                    rv = __ldg(p.c_recon + co);
                    if (MODE == kModeEncode) qv = __ldg(p.c_q + co);
                }
            }
            Pf[cj * pf + ci] = rv;
            if (MODE == kModeEncode) Qf[cj * pf + ci] = qv;
        }
    }
    __syncthreads();

    // ---- 3. coarse levels of the pass, s = F/2 .. 2 ---------------------------------------------
    if (F >= 16) coarse_level<MODE, INTERP, IDENTITY, 8>(sm, tid, qc, edge, xin, yin);
    if (F >= 8) coarse_level<MODE, INTERP, IDENTITY, 4>(sm, tid, qc, edge, xin, yin);
    if (F >= 4) coarse_level<MODE, INTERP, IDENTITY, 2>(sm, tid, qc, edge, xin, yin);

    // ---- 4. finest level: registers + P_2 / Q_2 -> HBM -------------------------------------------
    uint8_t* __restrict__ out = (MODE == kModeEncode ? p.grid_out : p.recon_out) + tile_off;
    if (MODE == kModeEncode && EXTRA && p.hist != nullptr) {
        for (int i = tid; i < NWARPS * 256; i += NT) whist[i] = 0u;
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int rp = ry + 32 * u;                 // row pair (cell row) of this unit
        const bool row0_ok = 2 * rp < yin, row1_ok = 2 * rp + 1 < yin;
        const uint32_t uoff = toff + (uint32_t)(64 * u) * p.w;
        const uint8_t* P2r = sm.P + plane_off(2) + rp * plane_pitch(2) + 8 * sx;
        const uint2 ctw = *reinterpret_cast<const uint2*>(P2r);
        const uint2 cbw = *reinterpret_cast<const uint2*>(P2r + plane_pitch(2));
        const uint32_t cte = P2r[8], cbe = P2r[plane_pitch(2) + 8];
        uint32_t A[4], B[4], C[4], D[4];
        A[0] = lanes01(ctw.x); A[1] = lanes23(ctw.x); A[2] = lanes01(ctw.y); A[3] = lanes23(ctw.y);
        B[0] = lanes01(cbw.x); B[1] = lanes23(cbw.x); B[2] = lanes01(cbw.y); B[3] = lanes23(cbw.y);
        C[0] = lanes12(ctw.x); C[1] = __funnelshift_r(A[1], A[2], 16); C[2] = lanes12(ctw.y); C[3] = __funnelshift_r(A[3], cte, 16);
        D[0] = lanes12(cbw.x); D[1] = __funnelshift_r(B[1], B[2], 16); D[2] = lanes12(cbw.y); D[3] = __funnelshift_r(B[3], cbe, 16);
        const uint32_t evw[4] = {ev[u].x, ev[u].y, ev[u].z, ev[u].w};
        const uint32_t odw[4] = {od[u].x, od[u].y, od[u].z, od[u].w};
        uint32_t out_ev[4], out_od[4], rec_ev[4], rec_od[4];
        uint2 qcw = make_uint2(0u, 0u);
        if (MODE == kModeEncode) qcw = *reinterpret_cast<const uint2*>(sm.Q + plane_off(2) + rp * plane_pitch(2) + 8 * sx);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t pr = pred2<INTERP>(A[k], B[k], C[k], D[k]);
            const uint32_t a1 = lanes_odd(evw[k]), a2 = lanes_even(odw[k]), a3 = lanes_odd(odw[k]);
            if (MODE == kModeEncode) {
                uint32_t r1, r2, r3;
                const uint32_t pk = 0x01000100u - pr;
                const uint32_t q1 = encode2<IDENTITY>(a1, pr, pk, qc, r1);
                const uint32_t q2 = encode2<IDENTITY>(a2, pr, pk, qc, r2);
                const uint32_t q3 = encode2<IDENTITY>(a3, pr, pk, qc, r3);
                const uint32_t qw = (k < 2) ? qcw.x : qcw.y;
                const uint32_t QA = (k & 1) ? lanes23(qw) : lanes01(qw);
                out_ev[k] = interleave(QA, q1);
                out_od[k] = interleave(q2, q3);
                if (EXTRA) {
                    rec_ev[k] = interleave(A[k], r1);
                    rec_od[k] = interleave(r2, r3);
                }
            } else {
                out_ev[k] = interleave(A[k], decode2(a1, pr));
                out_od[k] = interleave(decode2(a2, pr), decode2(a3, pr));
            }
        }
        if (col_ok && row0_ok) *reinterpret_cast<uint4*>(out + uoff) = make_uint4(out_ev[0], out_ev[1], out_ev[2], out_ev[3]);
        if (col_ok && row1_ok) *reinterpret_cast<uint4*>(out + uoff + p.w) = make_uint4(out_od[0], out_od[1], out_od[2], out_od[3]);
        if (MODE == kModeEncode && EXTRA) {
            if (p.recon_out != nullptr) {
                uint8_t* __restrict__ rout = p.recon_out + tile_off;
                if (col_ok && row0_ok) *reinterpret_cast<uint4*>(rout + uoff) = make_uint4(rec_ev[0], rec_ev[1], rec_ev[2], rec_ev[3]);
                if (col_ok && row1_ok) *reinterpret_cast<uint4*>(rout + uoff + p.w) = make_uint4(rec_od[0], rec_od[1], rec_od[2], rec_od[3]);
            }
            // residual histogram (north_star's archive.rs stage): warp-private shared-memory bins
            if (p.hist != nullptr && col_ok) {
                uint32_t* mine = &whist[(tid >> 5) * 256];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (row0_ok) atomicAdd(&mine[(out_ev[k] >> (8 * b)) & 0xFFu], 1u);
                        if (row1_ok) atomicAdd(&mine[(out_od[k] >> (8 * b)) & 0xFFu], 1u);
                    }
                }
            }
        }
    }
    if (MODE == kModeEncode && EXTRA && p.hist != nullptr) {   // one global atomic per non-empty bin per tile
        __syncthreads();
        uint32_t total = 0;
#pragma unroll
        for (int wv = 0; wv < NWARPS; ++wv) total += whist[wv * 256 + tid];
        if (total) atomicAdd(&p.hist[(size_t)img * 256 + tid], total);
    }
}

template <int MODE, int INTERP, int NLEV>
cudaError_t launch_fast_n(const PassArgs& args, cudaStream_t stream)
{
    const uint32_t tiles_x = (args.w + TW - 1) / TW, tiles_y = (args.h + TH - 1) / TH;
    if (tiles_x == 0 || tiles_y == 0 || args.n_images == 0) return cudaSuccess;
    if (tiles_y > 65535u) return cudaErrorInvalidConfiguration;
    const size_t plane = (size_t)args.w * args.h;
    for (uint32_t first = 0; first < args.n_images; first += 65535u) {   // gridDim.z limit
        PassArgs a = args;
        a.n_images = args.n_images - first < 65535u ? args.n_images - first : 65535u;
        a.src = args.src + (size_t)first * plane;
        if (args.grid_out) a.grid_out = args.grid_out + (size_t)first * plane;
        if (args.recon_out) a.recon_out = args.recon_out + (size_t)first * plane;
        if (args.hist) a.hist = args.hist + (size_t)first * 256;
        if (args.c_recon) a.c_recon = args.c_recon + (size_t)first * args.cw * args.ch;
        if (args.c_q) a.c_q = args.c_q + (size_t)first * args.cw * args.ch;
        const dim3 nb(tiles_x, tiles_y, a.n_images);
        if (MODE == kModeDecode) {
            hgi_tile_fast_kernel<kModeDecode, INTERP, true, false, NLEV><<<nb, NT, 0, stream>>>(a);
        } else {
            const bool extra = (a.recon_out != nullptr) || (a.hist != nullptr);
            const bool ident = (a.quant_error == 0);
            if (ident && !extra) hgi_tile_fast_kernel<kModeEncode, INTERP, true, false, NLEV><<<nb, NT, 0, stream>>>(a);
            else if (ident) hgi_tile_fast_kernel<kModeEncode, INTERP, true, true, NLEV><<<nb, NT, 0, stream>>>(a);
            else if (!extra) hgi_tile_fast_kernel<kModeEncode, INTERP, false, false, NLEV><<<nb, NT, 0, stream>>>(a);
            else hgi_tile_fast_kernel<kModeEncode, INTERP, false, true, NLEV><<<nb, NT, 0, stream>>>(a);
        }
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <int MODE, int INTERP>
cudaError_t launch_fast_t(const PassArgs& a, cudaStream_t stream)
{
    switch (a.nlev) {
        case 1: return launch_fast_n<MODE, INTERP, 1>(a, stream);
        case 2: return launch_fast_n<MODE, INTERP, 2>(a, stream);
        case 3: return launch_fast_n<MODE, INTERP, 3>(a, stream);
        case 4: return launch_fast_n<MODE, INTERP, 4>(a, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace

// Host-side proof obligation for the SWAR quantizer: it must equal the reference table for all
// 256 residuals at every QuantizationLevel (src/quantizator.rs:50-60).
bool quant_swar_self_check()
{
    for (uint32_t e = 10; e <= 30; e += 10) {
        const QuantSwar q = quant_swar(e);
        for (uint32_t d0 = 0; d0 < 256; ++d0) {
            const uint32_t d1 = 255u - d0;
            const uint32_t d = d0 | (d1 << 16);
            const uint32_t t = d * q.mul + q.add;
            const uint32_t v = (uint32_t)(((unsigned long long)(t & q.rmask) * q.qmul) >> 32);
            if ((v & 0xFFFFu) != quant_entry(d0, e) || (v >> 16) != quant_entry(d1, e)) return false;
        }
    }
    return true;
}

cudaError_t launch_tile_pass_fast(int mode, int interp, const PassArgs& a, cudaStream_t stream)
{
    if (mode == kModeEncode)
        return interp == kInterpLeftTop ? launch_fast_t<kModeEncode, kInterpLeftTop>(a, stream)
                                        : launch_fast_t<kModeEncode, kInterpCrossed>(a, stream);
    return interp == kInterpLeftTop ? launch_fast_t<kModeDecode, kInterpLeftTop>(a, stream)
                                    : launch_fast_t<kModeDecode, kInterpCrossed>(a, stream);
}

}  // namespace hgi
