// B3 -- depthwise 3x3 + bias + ReLU6 over NHWC activations (fp32 | bf16 storage, fp32 math).
// Replaces SeperableConv.depthwise + relu6 (posenet/models/mobilenet_v1.py:60-62,66) with the
// stride / dilation / padding rule of _to_output_strided_layers and _get_padding (:8-44):
// pad = ((s-1) + 2d) / 2, symmetric zero padding, out = (in + 2p - 2d - 1)/s + 1.
//
// HBM-bound (9 MAC per element moved), so the kernel is a TMA-fed streaming pipeline:
//   * the input is described by ONE 4-D tensor map (C, W, H, N).  A work item is an output tile of
//     TH x TW pixels x CB channels of one image; its input patch (with the 3x3 halo) is one
//     cp.async.bulk.tensor.4d box.  Out-of-bounds box elements are zero-filled by the TMA unit, which
//     IS the convolution's zero padding, so there is no border code at all.
//   * persistent CTAs (1 per SM): one producer warp keeps a 2-4 deep mbarrier ring of patches in
//     flight (~100+ KB per SM outstanding with no register cost), fifteen consumer warps compute.
//   * a group of CB/8 (bf16) lanes owns one pixel's channel block, 16 B per lane; a group computes a
//     strip of 4 (stride 1) or 2 (stride 2) adjacent output pixels so that each shared-memory vector
//     is loaded once per strip and tap row, math is packed FFMA2 (fma.rn.f32x2), and the result
//     leaves as 16 B stores that cover whole 32 B sectors.
// The tile shape is chosen per layer on the host (odd map sizes 257, 129, 65, 33, 17 ...).
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"

namespace pn {

constexpr int DW_CONSUMER_WARPS = 15;           // + 1 producer = 512 threads -> 128 registers each
constexpr int DW_CONSUMERS = DW_CONSUMER_WARPS * 32;
constexpr int DW_THREADS = DW_CONSUMERS + 32;        // + the producer warp
constexpr int DW_MAX_STAGES = 4;
constexpr int DW_SMEM_PER_CTA = 200 * 1024;          // one CTA per SM

template <typename T> struct Vec16;                  // one 16-byte channel vector
template <> struct Vec16<float> {
    static constexpr int VE = 4;
    static __device__ __forceinline__ void unpack(const uint4 &r, float2 (&v)[2]) {
        v[0] = make_float2(__uint_as_float(r.x), __uint_as_float(r.y));
        v[1] = make_float2(__uint_as_float(r.z), __uint_as_float(r.w));
    }
    static __device__ __forceinline__ void store_relu6(float *p, const float2 (&v)[2]) {
        *reinterpret_cast<float4 *>(p) = make_float4(relu6f(v[0].x), relu6f(v[0].y), relu6f(v[1].x), relu6f(v[1].y));
    }
};
template <> struct Vec16<__nv_bfloat16> {
    static constexpr int VE = 8;
    static __device__ __forceinline__ void unpack(const uint4 &r, float2 (&v)[4]) {   // bf16 -> f32 is a 16-bit shift
        v[0] = make_float2(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u));
        v[1] = make_float2(__uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
        v[2] = make_float2(__uint_as_float(r.z << 16), __uint_as_float(r.z & 0xffff0000u));
        v[3] = make_float2(__uint_as_float(r.w << 16), __uint_as_float(r.w & 0xffff0000u));
    }
    static __device__ __forceinline__ void store_relu6(__nv_bfloat16 *p, const float2 (&v)[4]) {
        // round first, clamp after: 0 and 6 are exact in bf16 and rounding is monotone, so this equals
        // clamp-then-round while the clamp runs on packed pairs
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = relu6_bf16x2(v[j]);
        *reinterpret_cast<uint4 *>(p) = make_uint4(o[0], o[1], o[2], o[3]);
    }
};

struct DwGeom {
    int n, h, w, c, ho, wo, pad;
    int cb, lpp;               // channels per block, lanes per pixel (cb / VE)
    int th, tw;                // output tile
    int thi, twi;              // input box (twi possibly padded to stagger shared-memory banks)
    int spr, nstrips;          // strips per tile row, strips per tile
    int tiles_x, tiles_y, cblocks;
    int stages;
    unsigned stage_bytes;      // distance between stages (128 B aligned)
    unsigned box_bytes;        // bytes one TMA box delivers
    long long items;
};

// item -> (channel block, tile x, tile y, image); channel blocks fastest so that neighbouring CTAs share halos in L2
__device__ __forceinline__ void dw_item(const DwGeom &g, long long it, int &cbk, int &tx, int &ty, int &img) {
    cbk = (int)(it % g.cblocks);
    long long r = it / g.cblocks;
    tx = (int)(r % g.tiles_x);
    r /= g.tiles_x;
    ty = (int)(r % g.tiles_y);
    img = (int)(r / g.tiles_y);
}

template <typename T, int S, int D>
__global__ void __launch_bounds__(DW_THREADS, 1)
dwconv_tma_kernel(const __grid_constant__ CUtensorMap tmap, const float *__restrict__ w, const float *__restrict__ bias,
                  T *__restrict__ y, const DwGeom g) {
    constexpr int VE = Vec16<T>::VE;
    constexpr int VP = VE / 2;                            // float2 pairs per vector
    constexpr int PXT = (S == 1) ? 4 : 2;                 // output pixels per strip
    constexpr int NCOLS = (PXT - 1) * S + 2 * D + 1;      // input columns a strip touches

    extern __shared__ uint8_t dw_smem_raw[];
    const uint32_t base = (smem_u32(dw_smem_raw) + 127u) & ~127u;
    auto full_bar = [&](int s) { return base + 8u * s; };
    auto empty_bar = [&](int s) { return base + 8u * (DW_MAX_STAGES + s); };
    auto stage_addr = [&](int s) { return base + 128u + (uint32_t)s * g.stage_bytes; };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        for (int s = 0; s < g.stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), DW_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    pdl_launch_dependents();
    __syncthreads();
    pdl_wait();                                           // (ptx.cuh) the previous layer has completed

    if (warp == DW_CONSUMER_WARPS) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long long it = blockIdx.x; it < g.items; it += gridDim.x) {
                int cbk, tx, ty, img;
                dw_item(g, it, cbk, tx, ty, img);
                mbar_wait(empty_bar(s), ph ^ 1);
                mbar_expect_tx(full_bar(s), g.box_bytes);
                tma_load_4d(stage_addr(s), &tmap, full_bar(s), cbk * g.cb, tx * g.tw * S - g.pad, ty * g.th * S - g.pad, img);
                if (++s == g.stages) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ===================== consumers =====================
    const int cl = threadIdx.x % g.lpp;                   // which 16 B of the pixel's channel block
    const int group = threadIdx.x / g.lpp;
    const int ngroups = DW_CONSUMERS / g.lpp;
    const uint32_t pixb = (uint32_t)g.cb * sizeof(T);     // bytes per pixel in the patch
    const uint32_t rowb = (uint32_t)g.twi * pixb;

    int s = 0;
    uint32_t ph = 0;
    for (long long it = blockIdx.x; it < g.items; it += gridDim.x) {
        int cbk, tx, ty, img;
        dw_item(g, it, cbk, tx, ty, img);
        const int c0 = cbk * g.cb + cl * VE;
        float2 bs[VP];
#pragma unroll
        for (int j = 0; j < VP; j += 2) {
            const float4 b4 = __ldg(reinterpret_cast<const float4 *>(bias + c0 + 2 * j));
            bs[j] = make_float2(b4.x, b4.y);
            bs[j + 1] = make_float2(b4.z, b4.w);
        }
        mbar_wait(full_bar(s), ph);
        const uint32_t patch = stage_addr(s) + (uint32_t)cl * 16u;

        for (int q = group; q < g.nstrips; q += ngroups) {
            const int r = q % g.th, sg = q / g.th;        // row-fastest: neighbouring groups sit on neighbouring rows
            const int oy = ty * g.th + r, ox0 = tx * g.tw + sg * PXT;
            if (oy >= g.ho || ox0 >= g.wo) continue;
            float2 acc[PXT][VP];
#pragma unroll
            for (int p = 0; p < PXT; ++p)
#pragma unroll
                for (int j = 0; j < VP; ++j) acc[p][j] = bs[j];
            const uint32_t strip = patch + (uint32_t)(r * S) * rowb + (uint32_t)(sg * PXT * S) * pixb;
#pragma unroll 1
            for (int ky = 0; ky < 3; ++ky) {                  // not unrolled: keeps one tap row of weights live
                float2 wv[3][VP];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                    for (int j = 0; j < VP; j += 2) {
                        const float4 w4 = __ldg(reinterpret_cast<const float4 *>(w + (size_t)(ky * 3 + kx) * g.c + c0 + 2 * j));
                        wv[kx][j] = make_float2(w4.x, w4.y);
                        wv[kx][j + 1] = make_float2(w4.z, w4.w);
                    }
                const uint32_t rowp = strip + (uint32_t)(ky * D) * rowb;
#pragma unroll
                for (int col = 0; col < NCOLS; ++col) {
                    bool used = false;
#pragma unroll
                    for (int p = 0; p < PXT; ++p)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) used |= (p * S + kx * D == col);
                    if (!used) continue;
                    float2 v[VP];
                    Vec16<T>::unpack(ld_shared_v4(rowp + (uint32_t)col * pixb), v);
#pragma unroll
                    for (int p = 0; p < PXT; ++p)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
                            if (p * S + kx * D == col) {
#pragma unroll
                                for (int j = 0; j < VP; ++j) acc[p][j] = ffma2(v[j], wv[kx][j], acc[p][j]);
                            }
                }
            }
            T *op = y + (((size_t)img * g.ho + oy) * g.wo + ox0) * g.c + c0;
#pragma unroll
            for (int p = 0; p < PXT; ++p)
                if (ox0 + p < g.wo) Vec16<T>::store_relu6(op + (size_t)p * g.c, acc[p]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));         // this warp no longer reads the patch
        if (++s == g.stages) { s = 0; ph ^= 1; }
    }
}

// ---- fallback for (stride, dilation) pairs the network tables never produce: direct global gathers ----------
template <typename T>
__global__ void __launch_bounds__(128) dwconv_direct_kernel(const T *__restrict__ x, const float *__restrict__ w,
                                                            const float *__restrict__ bias, T *__restrict__ y, int n, int h,
                                                            int wd, int c, int ho, int wo, int stride, int dil, int pad) {
    constexpr int VE = Vec16<T>::VE, VP = VE / 2;
    const int cg = c / VE;
    const long long total = (long long)n * ho * wo * cg;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int c0 = (int)(t % cg) * VE;
    long long r = t / cg;
    const int ox = (int)(r % wo);
    r /= wo;
    const int oy = (int)(r % ho), img = (int)(r / ho);
    float2 acc[VP];
#pragma unroll
    for (int j = 0; j < VP; ++j) acc[j] = make_float2(bias[c0 + 2 * j], bias[c0 + 2 * j + 1]);
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * stride - pad + ky * dil;
        if (iy < 0 || iy >= h) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * stride - pad + kx * dil;
            if (ix < 0 || ix >= wd) continue;
            float2 v[VP];
            Vec16<T>::unpack(__ldg(reinterpret_cast<const uint4 *>(x + (((size_t)img * h + iy) * wd + ix) * c + c0)), v);
            const float *wp = w + (size_t)(ky * 3 + kx) * c + c0;
#pragma unroll
            for (int j = 0; j < VP; ++j) {
                acc[j].x = fmaf(v[j].x, wp[2 * j], acc[j].x);
                acc[j].y = fmaf(v[j].y, wp[2 * j + 1], acc[j].y);
            }
        }
    }
    Vec16<T>::store_relu6(y + (((size_t)img * ho + oy) * wo + ox) * c + c0, acc);
}

// ---- host side: tile selection ---------------------------------------------------------------------------
static bool dw_tma_supported(int stride, int dil) {
    return (stride == 1 && (dil == 1 || dil == 2 || dil == 4)) || (stride == 2 && dil == 1);
}

int dw_prepare(DwOp *op, const void *x, int n, int h, int wd, int c, int stride, int dil, int dtype) {
    PN_CHECK_ARG(x && n > 0 && h > 0 && wd > 0, "pn_dwconv3x3: bad argument");
    PN_CHECK_ARG(dtype == PN_BF16 || dtype == PN_F32, "pn_dwconv3x3: bad dtype %d", dtype);
    PN_CHECK_ARG(c > 0 && c % 8 == 0, "pn_dwconv3x3: channels must be a multiple of 8 (got %d)", c);
    PN_CHECK_ARG(stride == 1 || stride == 2, "pn_dwconv3x3: stride must be 1 or 2 (got %d)", stride);
    PN_CHECK_ARG(dil >= 1, "pn_dwconv3x3: dilation must be >= 1");
    PN_CHECK_ARG(((uintptr_t)x & 15) == 0, "pn_dwconv3x3: input must be 16-byte aligned");
    memset(op, 0, sizeof(*op));
    op->x = x; op->n = n; op->h = h; op->w = wd; op->c = c; op->stride = stride; op->dil = dil; op->dtype = dtype;
    DwGeom g;
    memset(&g, 0, sizeof(g));
    g.n = n; g.h = h; g.w = wd; g.c = c;
    g.pad = ((stride - 1) + dil * 2) / 2;
    g.ho = (h + 2 * g.pad - 2 * dil - 1) / stride + 1;
    g.wo = (wd + 2 * g.pad - 2 * dil - 1) / stride + 1;
    PN_CHECK_ARG(g.ho > 0 && g.wo > 0, "pn_dwconv3x3: empty output");
    op->ho = g.ho; op->wo = g.wo;
    if (dwwarp_supported(c, stride, dil, dtype)) {
        op->warp_kind = true;
        return dwwarp_prepare(&op->warp, x, n, h, wd, c, dil);
    }
    op->use_tma = dw_tma_supported(stride, dil);
    if (!op->use_tma) return PN_OK;

    const int esize = dtype == PN_BF16 ? 2 : 4, ve = 16 / esize;
    g.lpp = 8;
    while (c % (g.lpp * ve) != 0) g.lpp >>= 1;            // c % 8 == 0 guarantees lpp >= 1 (bf16) / 2 (fp32)
    g.cb = g.lpp * ve;
    g.cblocks = c / g.cb;
    const int pixb = g.cb * esize;
    const int ngroups = DW_CONSUMERS / g.lpp;
    const int pxt = stride == 1 ? 4 : 2;
    // Search the tile shape: fewest strip slots (tiles x passes x groups), then fewest staged bytes.
    const int stage_cap = 64 * 1024;                    // at least three stages per CTA
    double best = 1e300;
    for (int th = 1; th <= 64 && th <= g.ho + 7; ++th) {
        for (int tw = pxt; tw <= 128; tw += pxt) {
            if (tw - pxt >= g.wo) break;
            const int thi = (th - 1) * stride + 2 * dil + 1;
            int twi = (tw - 1) * stride + 2 * dil + 1;
            if (pixb < 128) {                             // stagger consecutive rows by one pixel's worth of banks
                const int m = 128 / pixb;
                while (twi % m != 1 % m) ++twi;
            }
            if (thi > 256 || twi > 256) continue;
            const long long bytes = (long long)thi * twi * pixb;
            if (bytes > stage_cap) continue;
            const int spr = tw / pxt, nstrips = th * spr;
            const int passes = ceil_div(nstrips, ngroups);
            if (passes > 4) continue;
            const long long tiles = (long long)ceil_div(g.ho, th) * ceil_div(g.wo, tw);
            // cost model in SM cycles per tile: a pass issues ~660 cycles of instructions, the patch arrives at
            // ~40 B/clk (L2-assisted: halos are shared with neighbouring tiles), ~100 cycles of hand-off
            const double issue = passes * 660.0, fill = bytes / 40.0;
            const double cost = (double)tiles * ((issue > fill ? issue : fill) + 100.0);
            if (cost < best) {
                best = cost;
                g.th = th; g.tw = tw; g.thi = thi; g.twi = twi; g.spr = spr; g.nstrips = nstrips;
                g.stage_bytes = (unsigned)bytes;
            }
        }
    }
    PN_CHECK_ARG(best < 1e300, "pn_dwconv3x3: no tile shape fits (c %d stride %d dilation %d)", c, stride, dil);
    g.tiles_x = ceil_div(g.wo, g.tw);
    g.tiles_y = ceil_div(g.ho, g.th);
    g.items = (long long)n * g.tiles_y * g.tiles_x * g.cblocks;
    g.box_bytes = g.stage_bytes;
    g.stage_bytes = (g.stage_bytes + 127u) & ~127u;       // keep every stage 128-byte aligned
    g.stages = (DW_SMEM_PER_CTA - 256) / (int)g.stage_bytes;
    if (g.stages > DW_MAX_STAGES) g.stages = DW_MAX_STAGES;
    PN_CHECK_ARG(g.stages >= 2, "pn_dwconv3x3: stage too large");
    static_assert(sizeof(DwGeom) <= sizeof(op->geom), "DwOp::geom too small");
    memcpy(op->geom, &g, sizeof(g));
    const uint64_t dims[4] = {(uint64_t)c, (uint64_t)wd, (uint64_t)h, (uint64_t)n};
    const uint64_t strides[3] = {(uint64_t)c * esize, (uint64_t)wd * c * esize, (uint64_t)h * wd * c * esize};
    const uint32_t box[4] = {(uint32_t)g.cb, (uint32_t)g.twi, (uint32_t)g.thi, 1u};
    return encode_tmap(op->tmap, x, esize, 4, dims, strides, box, 0);
}

template <typename T, int S, int D>
static int launch_tma(const DwOp *op, const DwGeom &g, const float *w, const float *b, void *y, cudaStream_t st) {
    static DeviceOnce once;
    const int dev = current_device();
    auto kern = dwconv_tma_kernel<T, S, D>;
    if (!once.get(dev)) {
        PN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, DW_SMEM_PER_CTA));
        once.set(dev, 1);
    }
    const int smem = 256 + g.stages * (int)g.stage_bytes;
    const long long max_ctas = num_sms();
    const int grid = (int)(g.items < max_ctas ? g.items : max_ctas);
    PN_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(DW_THREADS), (size_t)smem, st, *reinterpret_cast<const CUtensorMap *>(op->tmap), w, b, (T *)y, g));
    return PN_OK;
}

template <typename T>
static int launch_typed(const DwOp *op, const float *w, const float *b, void *y, cudaStream_t st) {
    if (!op->use_tma) {
        const int ve = Vec16<T>::VE;
        const int pad = ((op->stride - 1) + op->dil * 2) / 2;
        const long long total = (long long)op->n * op->ho * op->wo * (op->c / ve);
        const long long blocks = (total + 127) / 128;
        PN_CHECK_ARG(blocks < (1ll << 31), "pn_dwconv3x3: problem too large");
        dwconv_direct_kernel<T><<<(unsigned)blocks, 128, 0, st>>>((const T *)op->x, w, b, (T *)y, op->n, op->h, op->w, op->c,
                                                                 op->ho, op->wo, op->stride, op->dil, pad);
        PN_CHECK_LAUNCH();
        return PN_OK;
    }
    DwGeom g;
    memcpy(&g, op->geom, sizeof(g));
    if (op->stride == 2) return launch_tma<T, 2, 1>(op, g, w, b, y, st);
    if (op->dil == 1) return launch_tma<T, 1, 1>(op, g, w, b, y, st);
    if (op->dil == 2) return launch_tma<T, 1, 2>(op, g, w, b, y, st);
    return launch_tma<T, 1, 4>(op, g, w, b, y, st);
}

int dw_launch(const DwOp *op, const float *w, const float *b, void *y, cudaStream_t st) {
    PN_CHECK_ARG(op && w && b && y, "pn_dwconv3x3: null pointer");
    PN_CHECK_ARG(((uintptr_t)y & 15) == 0, "pn_dwconv3x3: output must be 16-byte aligned");
    if (op->warp_kind) return dwwarp_launch(&op->warp, w, b, y, st);
    return op->dtype == PN_BF16 ? launch_typed<__nv_bfloat16>(op, w, b, y, st) : launch_typed<float>(op, w, b, y, st);
}

int launch_dwconv(const void *x, const float *w, const float *b, void *y, int n, int h, int wd, int c,
                  int stride, int dilation, int dtype, cudaStream_t st) {
    DwOp op;
    int rc = dw_prepare(&op, x, n, h, wd, c, stride, dilation, dtype);
    if (rc != PN_OK) return rc;
    return dw_launch(&op, w, b, y, st);
}

}  // namespace pn
