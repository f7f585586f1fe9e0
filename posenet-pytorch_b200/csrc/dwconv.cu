// B3 -- depthwise 3x3 + bias + ReLU6 over NHWC activations (fp32 | bf16 storage, fp32 math).
// Replaces SeperableConv.depthwise + relu6 (posenet/models/mobilenet_v1.py:60-62,66) with the
// stride / dilation / padding rule of _to_output_strided_layers and _get_padding (:8-44):
// pad = ((s-1) + 2d) / 2, symmetric zero padding, out = (in + 2p - 2d - 1)/s + 1.
//
// HBM-bound (9 MAC per element moved), so the kernel is organised around instruction count:
// a thread owns CH = 4 consecutive channels and a PX = 4 pixel wide column strip of TH output rows.
//   * stride 1, dilation 1 (most layers): the thread marches DOWN its strip.  Each input row is
//     loaded once (6 vectors), converted to fp32 once, and scattered into three rotating
//     accumulator rows (ky = 0, 1, 2), so every input element is fetched once per thread instead of
//     9 times, and the next row is prefetched while the current one is being consumed.
//   * strided / atrous layers: same strip ownership (weights stay in registers), three rows gathered
//     per output row.
// Consecutive lanes take consecutive channel groups, so every warp-level load / store is a run of
// contiguous NHWC bytes; the work list is flattened over (image, row block, strip, channel group),
// so the odd map sizes (257, 129, 65, 33 ...) cost at most one partial strip per row.
#include <type_traits>

#include "common.cuh"

namespace pn {

constexpr int DW_PX = 4;
constexpr int DW_CH = 4;

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    typedef float4 raw;
    static __device__ __forceinline__ raw zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ raw load(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
    static __device__ __forceinline__ void unpack(const raw &r, float (&v)[4]) { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
    static __device__ __forceinline__ void store_relu6(float *p, const float (&v)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(relu6f(v[0]), relu6f(v[1]), relu6f(v[2]), relu6f(v[3]));
    }
};
template <> struct Vec4<__nv_bfloat16> {
    typedef uint2 raw;
    static __device__ __forceinline__ raw zero() { return make_uint2(0u, 0u); }
    static __device__ __forceinline__ raw load(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
    static __device__ __forceinline__ void unpack(const raw &r, float (&v)[4]) {   // bf16 -> f32 is a 16-bit shift
        v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
        v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void store_relu6(__nv_bfloat16 *p, const float (&v)[4]) {
        // round first, clamp after: 0 and 6 are exact in bf16 and rounding is monotone, so this equals
        // clamp-then-round while the clamp runs on packed pairs
        const __nv_bfloat162 lo = __floats2bfloat162_rn(0.f, 0.f), hi = __floats2bfloat162_rn(6.f, 6.f);
        __nv_bfloat162 a = __hmin2(__hmax2(__floats2bfloat162_rn(v[0], v[1]), lo), hi);
        __nv_bfloat162 b = __hmin2(__hmax2(__floats2bfloat162_rn(v[2], v[3]), lo), hi);
        *reinterpret_cast<uint2 *>(p) = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
    }
};

struct DwGeom {
    int n, h, w, c, ho, wo, dil, pad, th, strips, yblocks;
};

// flat thread id -> (channel group, strip, row block, image)
__device__ __forceinline__ bool dw_decompose(const DwGeom &g, int &c0, int &ox0, int &oy0, int &img) {
    const int cg = g.c / DW_CH;
    const long long total = (long long)g.n * g.yblocks * g.strips * cg;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return false;
    c0 = (int)(t % cg) * DW_CH;
    long long r = t / cg;
    ox0 = (int)(r % g.strips) * DW_PX;
    r /= g.strips;
    oy0 = (int)(r % g.yblocks) * g.th;
    img = (int)(r / g.yblocks);
    return true;
}

__device__ __forceinline__ void dw_load_weights(const float *__restrict__ w, const float *__restrict__ bias, int c, int c0,
                                                float (&wt)[9][DW_CH], float (&bs)[DW_CH]) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(w + (size_t)k * c + c0));
        wt[k][0] = a.x; wt[k][1] = a.y; wt[k][2] = a.z; wt[k][3] = a.w;
    }
    const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + c0));
    bs[0] = b.x; bs[1] = b.y; bs[2] = b.z; bs[3] = b.w;
}

// ---- stride 1, dilation 1: marching column strip with rotating accumulator rows ------------------------
template <typename T>
__global__ void __launch_bounds__(128) dwconv_s1_kernel(const T *__restrict__ x, const float *__restrict__ w,
                                                        const float *__restrict__ bias, T *__restrict__ y, DwGeom g) {
    typedef typename Vec4<T>::raw Raw;
    constexpr int NC = DW_PX + 2;                 // input columns ox0-1 .. ox0+PX
    int c0, ox0, oy0, img;
    if (!dw_decompose(g, c0, ox0, oy0, img)) return;
    float wt[9][DW_CH], bs[DW_CH];
    dw_load_weights(w, bias, g.c, c0, wt, bs);

    const T *ximg = x + (size_t)img * g.h * g.w * g.c + c0;
    T *yimg = y + (size_t)img * g.ho * g.wo * g.c + c0;
    bool colok[NC];
#pragma unroll
    for (int cc = 0; cc < NC; ++cc) colok[cc] = (ox0 - 1 + cc) >= 0 && (ox0 - 1 + cc) < g.w;
    const int ylast = min(oy0 + g.th, g.ho) - 1;  // last output row of this strip
    const T *colbase = ximg + (ptrdiff_t)(ox0 - 1) * g.c;

    auto load_row = [&](int iy, Raw (&r)[NC]) {
        const bool rowok = iy >= 0 && iy < g.h;
        const T *rp = colbase + (ptrdiff_t)iy * g.w * g.c;
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) r[cc] = (rowok && colok[cc]) ? Vec4<T>::load(rp + (ptrdiff_t)cc * g.c) : Vec4<T>::zero();
    };

    float acc[3][DW_PX][DW_CH];
    Raw cur[NC], nxt[NC];
    load_row(oy0 - 1, cur);

    // One step consumes input row iy: it finishes output row iy-1 (ky = 2), continues row iy (ky = 1) and
    // opens row iy+1 (ky = 0, initialised with the bias).  PH rotates which accumulator plays which role.
    auto step = [&](auto ph, int iy) {
        constexpr int PH = decltype(ph)::value;
        constexpr int A = PH % 3, B = (PH + 1) % 3, C = (PH + 2) % 3;
        load_row(iy + 1, nxt);                                        // prefetch while computing
#pragma unroll
        for (int p = 0; p < DW_PX; ++p)
#pragma unroll
            for (int j = 0; j < DW_CH; ++j) acc[C][p][j] = bs[j];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
            float v[DW_CH];
            Vec4<T>::unpack(cur[cc], v);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int p = cc - kx;
                if (p < 0 || p >= DW_PX) continue;
#pragma unroll
                for (int j = 0; j < DW_CH; ++j) {
                    acc[A][p][j] = fmaf(v[j], wt[6 + kx][j], acc[A][p][j]);
                    acc[B][p][j] = fmaf(v[j], wt[3 + kx][j], acc[B][p][j]);
                    acc[C][p][j] = fmaf(v[j], wt[kx][j], acc[C][p][j]);
                }
            }
        }
        const int oy = iy - 1;
        if (oy >= oy0) {
            T *op = yimg + ((size_t)oy * g.wo + ox0) * g.c;
#pragma unroll
            for (int p = 0; p < DW_PX; ++p)
                if (ox0 + p < g.wo) Vec4<T>::store_relu6(op + (size_t)p * g.c, acc[A][p]);
        }
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) cur[cc] = nxt[cc];
    };

    // rows oy0-1 .. ylast+1; accumulators that receive contributions before being opened are harmless
    // (row oy0-2 / oy0-1 garbage is never stored), but keep them finite:
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int p = 0; p < DW_PX; ++p)
#pragma unroll
            for (int j = 0; j < DW_CH; ++j) acc[a][p][j] = 0.f;

    int iy = oy0 - 1;
#pragma unroll 1
    for (;;) {
        step(std::integral_constant<int, 0>(), iy); if (++iy > ylast + 1) break;
        step(std::integral_constant<int, 1>(), iy); if (++iy > ylast + 1) break;
        step(std::integral_constant<int, 2>(), iy); if (++iy > ylast + 1) break;
    }
}

// ---- strided / atrous layers: strip ownership without vertical reuse ------------------------------------
template <typename T, int STRIDE>
__global__ void __launch_bounds__(128) dwconv_gen_kernel(const T *__restrict__ x, const float *__restrict__ w,
                                                         const float *__restrict__ bias, T *__restrict__ y, DwGeom g) {
    int c0, ox0, oy0, img;
    if (!dw_decompose(g, c0, ox0, oy0, img)) return;
    float wt[9][DW_CH], bs[DW_CH];
    dw_load_weights(w, bias, g.c, c0, wt, bs);
    const T *ximg = x + (size_t)img * g.h * g.w * g.c + c0;
    T *yimg = y + (size_t)img * g.ho * g.wo * g.c + c0;
    const int ylast = min(oy0 + g.th, g.ho) - 1;
#pragma unroll 1
    for (int oy = oy0; oy <= ylast; ++oy) {
        float acc[DW_PX][DW_CH];
#pragma unroll
        for (int p = 0; p < DW_PX; ++p)
#pragma unroll
            for (int j = 0; j < DW_CH; ++j) acc[p][j] = bs[j];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * STRIDE - g.pad + ky * g.dil;
            if (iy < 0 || iy >= g.h) continue;
            const T *rp = ximg + (size_t)iy * g.w * g.c;
#pragma unroll
            for (int p = 0; p < DW_PX; ++p) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int ix = (ox0 + p) * STRIDE - g.pad + kx * g.dil;
                    if (ix < 0 || ix >= g.w) continue;
                    float v[DW_CH];
                    Vec4<T>::unpack(Vec4<T>::load(rp + (size_t)ix * g.c), v);
#pragma unroll
                    for (int j = 0; j < DW_CH; ++j) acc[p][j] = fmaf(v[j], wt[ky * 3 + kx][j], acc[p][j]);
                }
            }
        }
        T *op = yimg + ((size_t)oy * g.wo + ox0) * g.c;
#pragma unroll
        for (int p = 0; p < DW_PX; ++p)
            if (ox0 + p < g.wo) Vec4<T>::store_relu6(op + (size_t)p * g.c, acc[p]);
    }
}

// rows per strip: the divisor-ish of `ho` in [6, 13] that wastes the fewest rows
static int pick_th(int ho) {
    int best = 8, best_waste = 1 << 30;
    for (int th = 13; th >= 6; --th) {
        const int waste = ceil_div(ho, th) * th - ho;
        if (waste < best_waste) { best_waste = waste; best = th; }
    }
    return ho < 6 ? ho : best;
}

template <typename T>
static int launch_t(const T *x, const float *w, const float *b, T *y, int n, int h, int wd, int c, int stride,
                    int dil, cudaStream_t st) {
    DwGeom g;
    g.n = n; g.h = h; g.w = wd; g.c = c; g.dil = dil;
    g.pad = ((stride - 1) + dil * 2) / 2;
    g.ho = (h + 2 * g.pad - 2 * dil - 1) / stride + 1;
    g.wo = (wd + 2 * g.pad - 2 * dil - 1) / stride + 1;
    g.th = pick_th(g.ho);
    g.strips = ceil_div(g.wo, DW_PX);
    g.yblocks = ceil_div(g.ho, g.th);
    const long long total = (long long)n * g.yblocks * g.strips * (c / DW_CH);
    const long long blocks = (total + 127) / 128;
    PN_CHECK_ARG(blocks < (1ll << 31), "pn_dwconv3x3: problem too large");
    if (stride == 1 && dil == 1)
        dwconv_s1_kernel<T><<<(unsigned)blocks, 128, 0, st>>>(x, w, b, y, g);
    else if (stride == 1)
        dwconv_gen_kernel<T, 1><<<(unsigned)blocks, 128, 0, st>>>(x, w, b, y, g);
    else
        dwconv_gen_kernel<T, 2><<<(unsigned)blocks, 128, 0, st>>>(x, w, b, y, g);
    PN_CHECK_LAUNCH();
    return PN_OK;
}

int launch_dwconv(const void *x, const float *w, const float *b, void *y, int n, int h, int wd, int c,
                  int stride, int dilation, int dtype, cudaStream_t st) {
    PN_CHECK_ARG(x && w && b && y && n > 0 && h > 0 && wd > 0, "pn_dwconv3x3: bad argument");
    PN_CHECK_ARG(c > 0 && c % 8 == 0, "pn_dwconv3x3: channels must be a multiple of 8 (got %d)", c);
    PN_CHECK_ARG(stride == 1 || stride == 2, "pn_dwconv3x3: stride must be 1 or 2 (got %d)", stride);
    PN_CHECK_ARG(dilation >= 1, "pn_dwconv3x3: dilation must be >= 1");
    if (dtype == PN_BF16)
        return launch_t<__nv_bfloat16>((const __nv_bfloat16 *)x, w, b, (__nv_bfloat16 *)y, n, h, wd, c, stride, dilation, st);
    if (dtype == PN_F32) return launch_t<float>((const float *)x, w, b, (float *)y, n, h, wd, c, stride, dilation, st);
    set_error("pn_dwconv3x3: bad dtype %d", dtype);
    return PN_ERR_ARG;
}

}  // namespace pn
