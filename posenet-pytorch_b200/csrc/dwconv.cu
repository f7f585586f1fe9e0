// B3 -- depthwise 3x3 + bias + ReLU6 over NHWC activations (fp32 | bf16 storage, fp32 math).
// Replaces SeperableConv.depthwise + relu6 (posenet/models/mobilenet_v1.py:60-62,66) with the
// stride / dilation / padding rule of _to_output_strided_layers and _get_padding (:8-44):
// pad = ((s-1) + 2d) / 2, symmetric zero padding, out = (in + 2p - 2d - 1)/s + 1.
//
// HBM-bound (9 MAC per element moved).  Each thread owns 8 consecutive channels (one 16 B bf16
// vector) and PX horizontally adjacent output pixels, so the 3 x (PX-1)*s+2d+1 input window is
// loaded once per row and the 9 weight vectors stay in registers.  Consecutive threads take
// consecutive channel groups -> every warp-level load/store is a contiguous run of the NHWC row.
#include "common.cuh"

namespace pn {

template <typename T, int STRIDE, int PX>
__global__ void __launch_bounds__(256) dwconv_kernel(const T *__restrict__ x, const float *__restrict__ w,
                                                      const float *__restrict__ bias, T *__restrict__ y, int n,
                                                      int h, int wd, int c, int ho, int wo, int dil, int pad) {
    const int cg = c >> 3;                       // channel groups of 8
    const int wo_t = (wo + PX - 1) / PX;          // x tiles per row
    const long long total = (long long)n * ho * wo_t * cg;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int g = (int)(t % cg);
    long long r = t / cg;
    const int xt = (int)(r % wo_t);
    r /= wo_t;
    const int oy = (int)(r % ho);
    const int img = (int)(r / ho);
    const int c0 = g * 8;
    const int ox0 = xt * PX;

    float wt[9][8];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(w + (size_t)k * c + c0));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(w + (size_t)k * c + c0 + 4));
        wt[k][0] = a.x; wt[k][1] = a.y; wt[k][2] = a.z; wt[k][3] = a.w;
        wt[k][4] = b.x; wt[k][5] = b.y; wt[k][6] = b.z; wt[k][7] = b.w;
    }
    float acc[PX][8];
    {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(bias + c0));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + c0 + 4));
#pragma unroll
        for (int p = 0; p < PX; ++p) {
            acc[p][0] = a.x; acc[p][1] = a.y; acc[p][2] = a.z; acc[p][3] = a.w;
            acc[p][4] = b.x; acc[p][5] = b.y; acc[p][6] = b.z; acc[p][7] = b.w;
        }
    }
    const T *ximg = x + (size_t)img * h * wd * c;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * STRIDE - pad + ky * dil;
        if (iy < 0 || iy >= h) continue;
        const T *row = ximg + (size_t)iy * wd * c + c0;
#pragma unroll
        for (int p = 0; p < PX; ++p) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = (ox0 + p) * STRIDE - pad + kx * dil;
                if (ix < 0 || ix >= wd || ox0 + p >= wo) continue;
                float v[8];
                Vec8<T>::load(row + (size_t)ix * c, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(v[j], wt[ky * 3 + kx][j], acc[p][j]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < PX; ++p) {
        if (ox0 + p >= wo) break;
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = relu6f(acc[p][j]);
        Vec8<T>::store(y + (((size_t)img * ho + oy) * wo + ox0 + p) * c + c0, o);
    }
}

template <typename T>
static int launch_t(const T *x, const float *w, const float *b, T *y, int n, int h, int wd, int c, int stride,
                    int dil, cudaStream_t st) {
    const int pad = ((stride - 1) + dil * 2) / 2;
    const int ho = (h + 2 * pad - 2 * dil - 1) / stride + 1;
    const int wo = (wd + 2 * pad - 2 * dil - 1) / stride + 1;
    constexpr int PX = 2;
    const long long total = (long long)n * ho * ((wo + PX - 1) / PX) * (c / 8);
    const long long blocks = (total + 255) / 256;
    PN_CHECK_ARG(blocks < (1ll << 31), "pn_dwconv3x3: problem too large");
    if (stride == 1)
        dwconv_kernel<T, 1, PX><<<(unsigned)blocks, 256, 0, st>>>(x, w, b, y, n, h, wd, c, ho, wo, dil, pad);
    else
        dwconv_kernel<T, 2, PX><<<(unsigned)blocks, 256, 0, st>>>(x, w, b, y, n, h, wd, c, ho, wo, dil, pad);
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
