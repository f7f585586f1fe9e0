// B2 -- stem: relu6(conv3x3(x, stride s, pad 1) + b), 3 -> COUT (16 | 24 | 32), NHWC out.
// Replaces InputConv (posenet/models/mobilenet_v1.py:47-54).  K = 27 is far too small for the tensor
// cores; the layer is bounded by its output write (COUT*2 B per pixel) and FFMA issue.
// One thread per output pixel, all COUT channels in registers, weights broadcast from shared memory.
// Input is either the reference's f32 NCHW tensor, or (fused P1, identity resize) the uint8 BGR
// HWC image itself, normalised on the fly with the same two rounded fp32 ops as utils.py:23.
#include "common.cuh"

namespace pn {

template <int COUT, bool IN_U8, typename TOut>
__global__ void __launch_bounds__(128) stem_kernel(const void *__restrict__ xin, const float *__restrict__ w,
                                                    const float *__restrict__ bias, TOut *__restrict__ y,
                                                    int n, int h, int wd, int ho, int wo, int stride) {
    __shared__ float sw[27 * COUT];
    __shared__ float sb[COUT];
    for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) sw[i] = w[i];
    for (int i = threadIdx.x; i < COUT; i += blockDim.x) sb[i] = bias[i];
    __syncthreads();

    const long long total = (long long)n * ho * wo;
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= total) return;
    const int ox = (int)(pix % wo);
    const int oy = (int)((pix / wo) % ho);
    const int img = (int)(pix / ((long long)wo * ho));

    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = sb[c];

    const size_t plane = (size_t)h * wd;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * stride - 1 + ky;
        if (iy < 0 || iy >= h) continue;               // zero padding of the (normalised) input
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * stride - 1 + kx;
            if (ix < 0 || ix >= wd) continue;
            float v[3];
            if (IN_U8) {
                const uint8_t *p = reinterpret_cast<const uint8_t *>(xin) + ((size_t)img * plane + (size_t)iy * wd + ix) * 3;
                // BGR -> RGB, x*(2/255)-1 exactly as P1 does
                v[0] = __fsub_rn(__fmul_rn((float)p[2], (float)(2.0 / 255.0)), 1.0f);
                v[1] = __fsub_rn(__fmul_rn((float)p[1], (float)(2.0 / 255.0)), 1.0f);
                v[2] = __fsub_rn(__fmul_rn((float)p[0], (float)(2.0 / 255.0)), 1.0f);
            } else {
                const float *p = reinterpret_cast<const float *>(xin) + (size_t)img * 3 * plane + (size_t)iy * wd + ix;
                v[0] = __ldg(p); v[1] = __ldg(p + plane); v[2] = __ldg(p + 2 * plane);
            }
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float *wr = &sw[((ky * 3 + kx) * 3 + ci) * COUT];
#pragma unroll
                for (int c = 0; c < COUT; ++c) acc[c] = fmaf(v[ci], wr[c], acc[c]);
            }
        }
    }
    TOut *o = y + (size_t)pix * COUT;
#pragma unroll
    for (int c0 = 0; c0 < COUT; c0 += 8) {
        float v8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v8[j] = relu6f(acc[c0 + j]);
        Vec8<TOut>::store(o + c0, v8);
    }
}

template <int COUT, bool IN_U8>
static int launch_t(const void *x, const float *w, const float *b, void *y, int n, int h, int wd, int ho, int wo,
                    int stride, int out_dtype, cudaStream_t st) {
    const long long total = (long long)n * ho * wo;
    const int blocks = (int)((total + 127) / 128);
    if (out_dtype == PN_BF16)
        stem_kernel<COUT, IN_U8, __nv_bfloat16><<<blocks, 128, 0, st>>>(x, w, b, (__nv_bfloat16 *)y, n, h, wd, ho, wo, stride);
    else
        stem_kernel<COUT, IN_U8, float><<<blocks, 128, 0, st>>>(x, w, b, (float *)y, n, h, wd, ho, wo, stride);
    PN_CHECK_LAUNCH();
    return PN_OK;
}

int launch_stem(const void *x, bool x_is_u8, const float *w, const float *b, void *y, int n, int h, int wd,
                int cout, int stride, int out_dtype, cudaStream_t st) {
    PN_CHECK_ARG(x && w && b && y && n > 0 && h > 0 && wd > 0, "pn_stem_conv: bad argument");
    PN_CHECK_ARG(stride == 1 || stride == 2, "pn_stem_conv: stride must be 1 or 2 (got %d)", stride);
    PN_CHECK_ARG(out_dtype == PN_F32 || out_dtype == PN_BF16, "pn_stem_conv: bad dtype %d", out_dtype);
    const int ho = (h + 2 - 3) / stride + 1, wo = (wd + 2 - 3) / stride + 1;
#define PN_STEM_CASE(C)                                                                                  \
    if (cout == C)                                                                                       \
        return x_is_u8 ? launch_t<C, true>(x, w, b, y, n, h, wd, ho, wo, stride, out_dtype, st)          \
                       : launch_t<C, false>(x, w, b, y, n, h, wd, ho, wo, stride, out_dtype, st);
    PN_STEM_CASE(16)
    PN_STEM_CASE(24)
    PN_STEM_CASE(32)
#undef PN_STEM_CASE
    set_error("pn_stem_conv: unsupported cout %d (16, 24, 32)", cout);
    return PN_ERR_UNSUPPORTED;
}

}  // namespace pn
