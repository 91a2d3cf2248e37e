// hgi_tile_tma.cu -- persistent, TMA-pipelined variant of the fast fused HGI tile kernel (sm_100a).
//
// Same arithmetic and plane geometry as hgi_tile_fast.cu (shared code: hgi_tile_swar.cuh); what changes
// is how pixels reach the SM:
//
//  * the planes are described once per launch by two 3-D tensor maps (x, y, image) and every 128x64 tile
//    plus its right/bottom halo arrives with four `cp.async.bulk.tensor.3d` (TMA) copies issued by ONE
//    thread: a 160x65 box (rows 0..TH, columns 0..TW+31) and three 160x1 boxes for rows TH+4, TH+8, TH+16.
//    TMA's out-of-bounds zero fill is exactly the reference's `get_pixel` rule (out-of-image reads as 0,
//    src/interpolator.rs:75-82) and applies only at true image edges, because the z coordinate keeps the
//    images of a batch apart;
//  * CTAs are persistent (grid = resident CTAs, static round-robin over tiles) with a two-stage
//    shared-memory ring guarded by mbarriers: the copy of tile i+1 is in flight while the CTA computes
//    tile i, so the DRAM latency that the register-prefetch kernel exposes at the start of every CTA is
//    hidden, no registers are spent on prefetched pixels, and no thread computes load addresses or
//    edge predicates for loads;
//  * the finest level reads its pixels from the staged tile with 128-bit shared-memory loads and still
//    stores finished grid / image words straight to HBM.
//
// Reference semantics: src/encoder.rs:39-71, src/decoder.rs:18-46, src/utils.rs:11-41,
// src/interpolator.rs:15-28,41-91, src/quantizator.rs:41-74.
#include <cuda.h>

#ifndef HGI_TMA_TILE_H
#define HGI_TMA_TILE_H 128
#endif
#define HGI_TILE_H HGI_TMA_TILE_H
#define HGI_TILE_NT (HGI_TMA_TILE_H * 4)     // one 16x2 unit per thread: 256 threads for 128x64, 512 for 128x128
#include "hgi_tile_swar.cuh"

namespace hgi {

namespace {

constexpr int RAW_PITCH = TW + 32;                                  // 160 staged columns per row
constexpr int RAW_MAIN_ROWS = TH + 1;                               // rows 0..TH
constexpr int RAW_MAIN_BYTES = RAW_MAIN_ROWS * RAW_PITCH;           // 10400
constexpr int RAW_ROW4_OFF = (RAW_MAIN_BYTES + 127) & ~127;         // row TH+4
constexpr int RAW_ROW8_OFF = RAW_ROW4_OFF + 256;                    // row TH+8
constexpr int RAW_ROW16_OFF = RAW_ROW8_OFF + 256;                   // row TH+16
constexpr int RAW_BYTES = RAW_ROW16_OFF + 256;                      // 11264
constexpr uint32_t TX_BYTES = RAW_MAIN_BYTES + 3 * RAW_PITCH;       // bytes one tile's four copies deliver
static_assert(RAW_BYTES % 128 == 0 && RAW_ROW4_OFF % 128 == 0, "TMA destinations must be 128-byte aligned");

#ifndef HGI_TMA_MIN_BLOCKS
#define HGI_TMA_MIN_BLOCKS (HGI_TMA_TILE_H == 128 ? 3 : 6)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost copy must end in a trap (reported as a CUDA error), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int x, int y, int z, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

template <int MODE, int INTERP, bool IDENTITY, bool EXTRA, int NLEV>
__global__ void __launch_bounds__(NT, HGI_TMA_MIN_BLOCKS)
hgi_tile_tma_kernel(const __grid_constant__ CUtensorMap tm_main, const __grid_constant__ CUtensorMap tm_row,
                    const PassArgs p, const uint32_t tiles_x, const uint32_t tiles_per_image, const uint32_t total_tiles)
{
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    uint8_t (*raw)[RAW_BYTES] = reinterpret_cast<uint8_t (*)[RAW_BYTES]>(dyn_smem);
    FastSmem& sm = *reinterpret_cast<FastSmem*>(dyn_smem + 2 * RAW_BYTES);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(dyn_smem + 2 * RAW_BYTES + sizeof(FastSmem));
    constexpr int F = 1 << NLEV;

    const int tid = threadIdx.x;
    const int sx = tid & 7, ry = tid >> 3;          // 8 column strips x TH/2 row pairs
    const bool top = (p.c_recon == nullptr);
    const QuantSwar qc = {p.q_one, p.q_mul, p.q_add, p.q_shift, p.q_scale, p.q_rmask, p.q_qmul, p.q_hK, p.q_hc1, p.q_hS, p.q_hc2};   // filled by the launcher

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // one thread feeds the ring: four TMA copies per tile, completion counted in bytes on the stage's mbarrier
    auto issue = [&](uint32_t t, int buf) {
        const uint32_t im = t / tiles_per_image, r = t - im * tiles_per_image;
        const uint32_t tyy = r / tiles_x, txx = r - tyy * tiles_x;
        const int x = (int)(txx * TW), y = (int)(tyy * TH), z = (int)im;
        mbar_arrive_expect_tx(&mbar[buf], TX_BYTES);
        tma_load_3d(&raw[buf][0], &tm_main, x, y, z, &mbar[buf]);
        tma_load_3d(&raw[buf][RAW_ROW4_OFF], &tm_row, x, y + TH + 4, z, &mbar[buf]);
        tma_load_3d(&raw[buf][RAW_ROW8_OFF], &tm_row, x, y + TH + 8, z, &mbar[buf]);
        tma_load_3d(&raw[buf][RAW_ROW16_OFF], &tm_row, x, y + TH + 16, z, &mbar[buf]);
    };

    uint32_t t = blockIdx.x;
    if (t >= total_tiles) return;
    if (tid == 0) issue(t, 0);
    // tile coordinates advance by gridDim.x tiles per iteration: keep (img, ty, tx) incrementally
    const uint32_t tiles_y = tiles_per_image / tiles_x;
    uint32_t img = t / tiles_per_image, tyy = (t - img * tiles_per_image) / tiles_x, txx = t - img * tiles_per_image - tyy * tiles_x;
    const uint32_t d_img = gridDim.x / tiles_per_image, d_rem = gridDim.x - d_img * tiles_per_image;
    const uint32_t d_ty = d_rem / tiles_x, d_tx = d_rem - d_ty * tiles_x;

    for (uint32_t iter = 0; t < total_tiles; t += gridDim.x, ++iter) {
        const int buf = (int)(iter & 1u);
        // prefetch the next tile into the other stage (its last readers passed the barrier that ended the previous iteration)
        if (tid == 0 && t + gridDim.x < total_tiles) issue(t + gridDim.x, buf ^ 1);

        const uint32_t X0 = txx * TW, Y0 = tyy * TH;
        const int xin = (int)min((uint32_t)(TW + FMAX + 1), p.w - X0);   // in-image extent of tile + halo
        const int yin = (int)min((uint32_t)(TH + FMAX + 1), p.h - Y0);
        const bool edge = (xin < TW + FMAX + 1) || (yin < TH + FMAX + 1);
        const size_t tile_off = ((size_t)img * p.h + Y0) * p.w + X0;
        const uint8_t* rw = raw[buf];

        mbar_wait(&mbar[buf], (iter >> 1) & 1u);

        // ---- stage the dense coarse planes + the coarse lattice of this pass ------------------------
        if (NLEV > 1) {
            stage_chunk<F>(sm.P, *reinterpret_cast<const uint4*>(rw + (2 * ry) * RAW_PITCH + 16 * sx), 2 * ry, sx);
            // halo chunks: column TW on even rows, rows TH / TH+4 / TH+8 (chunks 0..8)
            constexpr int NRIGHT = TH / 2, NHALO = NRIGHT + 27;
            const int hj = tid - (NT - 128);
            if (hj >= 0 && hj < NHALO) {
                int hy = 2 * hj, hc = 8;
                const uint8_t* src = rw + hy * RAW_PITCH + 128;
                if (hj >= NRIGHT) {
                    const int r = (hj - NRIGHT) / 9;
                    hc = (hj - NRIGHT) - 9 * r;
                    hy = TH + 4 * r;
                    src = rw + (r == 0 ? TH * RAW_PITCH : (r == 1 ? RAW_ROW4_OFF : RAW_ROW8_OFF)) + 16 * hc;
                }
                stage_chunk<F>(sm.P, *reinterpret_cast<const uint4*>(src), hy, hc);
            }
        }
        {
            constexpr int ncx = TW / F + 2, ncy = TH / F + 2;
            constexpr int pf = plane_pitch(F);
            uint8_t* Pf = sm.P + plane_off(F);
            uint8_t* Qf = sm.Q + plane_off(F);
            for (int it = tid; it < ncx * ncy; it += NT) {
                const int cj = it / ncx, ci = it - cj * ncx;
                const int x = ci * F, y = cj * F;
                uint8_t rv = 0, qv = 0;
                if (top) {   // src/encoder.rs:26-37 / src/decoder.rs:22-28: the seed is the source byte (0 outside)
                    const int off = y <= TH ? y * RAW_PITCH
                                            : (y == TH + 4 ? RAW_ROW4_OFF : (y == TH + 8 ? RAW_ROW8_OFF : (y == TH + 16 ? RAW_ROW16_OFF : -1)));
                    if (off >= 0) rv = rw[off + x];   // other rows of the over-covering fill are never read
                    qv = rv;
                } else if (x < xin && y < yin) {
                    const size_t co = ((size_t)img * p.ch + ((Y0 + y) >> NLEV)) * p.cpitch + ((X0 + x) >> NLEV);
                    rv = __ldg(p.c_recon + co);
                    if (MODE == kModeEncode) qv = __ldg(p.c_q + co);
                }
                Pf[cj * pf + ci] = rv;
                if (MODE == kModeEncode) Qf[cj * pf + ci] = qv;
            }
        }
        __syncthreads();

        // ---- coarse levels of the pass, s = F/2 .. 2 ------------------------------------------------
        if (F >= 16) coarse_level<MODE, INTERP, IDENTITY, 8>(sm, tid, qc, edge, xin, yin);
        if (F >= 8) coarse_level<MODE, INTERP, IDENTITY, 4>(sm, tid, qc, edge, xin, yin);
        if (F >= 4) coarse_level<MODE, INTERP, IDENTITY, 2>(sm, tid, qc, edge, xin, yin);

        // ---- finest level: staged pixels + P_2 / Q_2 -> HBM -----------------------------------------
        {
            const bool col_ok = 16 * sx < xin;
            const bool row0_ok = 2 * ry < yin, row1_ok = 2 * ry + 1 < yin;
            const uint32_t toff = (uint32_t)(2 * ry) * p.w + (uint32_t)(16 * sx);
            const uint4 ev = *reinterpret_cast<const uint4*>(rw + (2 * ry) * RAW_PITCH + 16 * sx);
            const uint4 od = *reinterpret_cast<const uint4*>(rw + (2 * ry + 1) * RAW_PITCH + 16 * sx);
            const uint8_t* P2r = sm.P + plane_off(2) + ry * plane_pitch(2) + 8 * sx;
            const uint2 ctw = *reinterpret_cast<const uint2*>(P2r);
            const uint2 cbw = *reinterpret_cast<const uint2*>(P2r + plane_pitch(2));
            const uint32_t cte = P2r[8], cbe = P2r[plane_pitch(2) + 8];
            uint32_t A[4], B[4], C[4], D[4];
            A[0] = lanes01(ctw.x); A[1] = lanes23(ctw.x); A[2] = lanes01(ctw.y); A[3] = lanes23(ctw.y);
            B[0] = lanes01(cbw.x); B[1] = lanes23(cbw.x); B[2] = lanes01(cbw.y); B[3] = lanes23(cbw.y);
            C[0] = lanes12(ctw.x); C[1] = __funnelshift_r(A[1], A[2], 16); C[2] = lanes12(ctw.y); C[3] = __funnelshift_r(A[3], cte, 16);
            D[0] = lanes12(cbw.x); D[1] = __funnelshift_r(B[1], B[2], 16); D[2] = lanes12(cbw.y); D[3] = __funnelshift_r(B[3], cbe, 16);
            const uint32_t evw[4] = {ev.x, ev.y, ev.z, ev.w};
            const uint32_t odw[4] = {od.x, od.y, od.z, od.w};
            uint32_t out_ev[4], out_od[4], rec_ev[4], rec_od[4];
            uint2 qcw = make_uint2(0u, 0u);
            if (MODE == kModeEncode) qcw = *reinterpret_cast<const uint2*>(sm.Q + plane_off(2) + ry * plane_pitch(2) + 8 * sx);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t a1 = lanes_odd(evw[k]), a2 = lanes_even(odw[k]), a3 = lanes_odd(odw[k]);
                if (MODE == kModeEncode) {
                    uint32_t q[3], r[3];
                    encode_cells<INTERP, IDENTITY>(A[k], B[k], C[k], D[k], a1, a2, a3, qc, q, r);
                    const uint32_t qw = (k < 2) ? qcw.x : qcw.y;
                    out_ev[k] = pack_even_row(qw, q[0], (k & 1) != 0);
                    out_od[k] = pack_sym<IDENTITY>(q[1], q[2]);
                    if (EXTRA) {
                        rec_ev[k] = interleave(A[k], r[0]);
                        rec_od[k] = interleave(r[1], r[2]);
                    }
                } else {
                    const uint32_t pr = pred2<INTERP, true>(A[k], B[k], C[k], D[k], qc.one);
                    out_ev[k] = pack_lo(A[k], decode2(a1, pr, qc.one));
                    out_od[k] = pack_lo(decode2(a2, pr, qc.one), decode2(a3, pr, qc.one));
                }
            }
            uint8_t* __restrict__ out = (MODE == kModeEncode ? p.grid_out : p.recon_out) + tile_off;
            if (col_ok && row0_ok) *reinterpret_cast<uint4*>(out + toff) = make_uint4(out_ev[0], out_ev[1], out_ev[2], out_ev[3]);
            if (col_ok && row1_ok) *reinterpret_cast<uint4*>(out + toff + p.w) = make_uint4(out_od[0], out_od[1], out_od[2], out_od[3]);
            if (MODE == kModeEncode && EXTRA) {
                if (p.recon_out != nullptr) {
                    uint8_t* __restrict__ rout = p.recon_out + tile_off;
                    if (col_ok && row0_ok) *reinterpret_cast<uint4*>(rout + toff) = make_uint4(rec_ev[0], rec_ev[1], rec_ev[2], rec_ev[3]);
                    if (col_ok && row1_ok) *reinterpret_cast<uint4*>(rout + toff + p.w) = make_uint4(rec_od[0], rec_od[1], rec_od[2], rec_od[3]);
                }
            }
        }
        __syncthreads();   // the tile is done: planes, this stage of the ring may be reused
        {   // advance (img, ty, tx) by gridDim.x tiles (mixed-radix add with carries)
            txx += d_tx;
            uint32_t cy = 0;
            if (txx >= tiles_x) { txx -= tiles_x; cy = 1; }
            tyy += d_ty + cy;
            uint32_t ci = 0;
            if (tyy >= tiles_y) { tyy -= tiles_y; ci = 1; }
            img += d_img + ci;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        (void)cudaGetLastError();
        return (EncodeTiledFn)f;
    }();
    return fn;
}

bool make_maps(const uint8_t* base, uint32_t w, uint32_t h, uint32_t n, CUtensorMap* main_map, CUtensorMap* row_map)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {w, h, n};
    const cuuint64_t strides[2] = {(cuuint64_t)w, (cuuint64_t)w * h};   // bytes, dims 1 and 2
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint32_t box_main[3] = {(cuuint32_t)RAW_PITCH, (cuuint32_t)RAW_MAIN_ROWS, 1};
    const cuuint32_t box_row[3] = {(cuuint32_t)RAW_PITCH, 1, 1};
    if (strides[1] >= (1ull << 40)) return false;
    void* g = const_cast<uint8_t*>(base);
    // OOB_FILL_NONE = out-of-bounds elements are written as zeros: the reference's get_pixel rule
    if (enc(main_map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, g, dims, strides, box_main, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (enc(row_map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, g, dims, strides, box_row, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    return true;
}

int resident_ctas()
{
    static int n = [] {
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        return sms * HGI_TMA_MIN_BLOCKS;
    }();
    return n;
}

constexpr size_t smem_bytes(bool) { return 2 * RAW_BYTES + sizeof(FastSmem) + 16; }

template <class K>
cudaError_t launch_one(K kernel, bool hist, uint32_t grid, cudaStream_t stream, const CUtensorMap& tm_main,
                       const CUtensorMap& tm_row, const PassArgs& a, uint32_t tiles_x, uint32_t per_img, uint32_t total)
{
    const size_t smem = smem_bytes(hist);
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // pixels arrive through TMA, L1 is almost unused: give the whole array to shared memory so that
    // HGI_TMA_MIN_BLOCKS CTAs really are resident per SM
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    kernel<<<grid, NT, smem, stream>>>(tm_main, tm_row, a, tiles_x, per_img, total); ++launch_count();
    return cudaGetLastError();
}

template <int MODE, int INTERP, int NLEV>
cudaError_t launch_tma_n(const PassArgs& a, cudaStream_t stream, bool* used)
{
    const uint32_t tiles_x = (a.w + TW - 1) / TW, tiles_y = (a.h + TH - 1) / TH;
    const uint64_t total64 = (uint64_t)tiles_x * tiles_y * a.n_images;
    *used = false;
    if (total64 == 0) { *used = true; return cudaSuccess; }
    if (total64 > 0x7FFFFFFFull) return cudaSuccess;   // fall back to the register-prefetch kernel
    CUtensorMap tm_main, tm_row;
    if (!make_maps(a.src, a.w, a.h, a.n_images, &tm_main, &tm_row)) return cudaSuccess;
    *used = true;
    const uint32_t total = (uint32_t)total64, per_img = tiles_x * tiles_y;
    const uint32_t grid = total < (uint32_t)resident_ctas() ? total : (uint32_t)resident_ctas();
    if (MODE == kModeDecode)
        return launch_one(hgi_tile_tma_kernel<kModeDecode, INTERP, true, false, NLEV>, false, grid, stream, tm_main, tm_row, a, tiles_x, per_img, total);
    const bool extra = (a.recon_out != nullptr);
    const bool ident = (a.quant_error == 0);
    if (ident && !extra)
        return launch_one(hgi_tile_tma_kernel<kModeEncode, INTERP, true, false, NLEV>, false, grid, stream, tm_main, tm_row, a, tiles_x, per_img, total);
    if (ident)
        return launch_one(hgi_tile_tma_kernel<kModeEncode, INTERP, true, true, NLEV>, true, grid, stream, tm_main, tm_row, a, tiles_x, per_img, total);
    if (!extra)
        return launch_one(hgi_tile_tma_kernel<kModeEncode, INTERP, false, false, NLEV>, false, grid, stream, tm_main, tm_row, a, tiles_x, per_img, total);
    return launch_one(hgi_tile_tma_kernel<kModeEncode, INTERP, false, true, NLEV>, true, grid, stream, tm_main, tm_row, a, tiles_x, per_img, total);
}

template <int MODE, int INTERP>
cudaError_t launch_tma_t(const PassArgs& a, cudaStream_t stream, bool* used)
{
    switch (a.nlev) {
        case 1: return launch_tma_n<MODE, INTERP, 1>(a, stream, used);
        case 2: return launch_tma_n<MODE, INTERP, 2>(a, stream, used);
        case 3: return launch_tma_n<MODE, INTERP, 3>(a, stream, used);
        case 4: return launch_tma_n<MODE, INTERP, 4>(a, stream, used);
        default: *used = false; return cudaSuccess;
    }
}

}  // namespace

// Returns cudaSuccess with *used == false when the TMA path cannot serve this launch (no driver entry point,
// tensor-map limits); the caller then uses the register-prefetch kernel.
cudaError_t launch_tile_pass_tma(int mode, int interp, const PassArgs& args, cudaStream_t stream, bool* used)
{
    PassArgs a = args;
    fill_quant_args(a);
    if (mode == kModeEncode)
        return interp == kInterpLeftTop ? launch_tma_t<kModeEncode, kInterpLeftTop>(a, stream, used)
                                        : launch_tma_t<kModeEncode, kInterpCrossed>(a, stream, used);
    return interp == kInterpLeftTop ? launch_tma_t<kModeDecode, kInterpLeftTop>(a, stream, used)
                                    : launch_tma_t<kModeDecode, kInterpCrossed>(a, stream, used);
}

}  // namespace hgi
