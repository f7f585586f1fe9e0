// B4 / H1, fp32 parity path -- Y = epilogue(A[M,K] . W[N,K]^T + bias) on the FFMA pipe.
// The reference's fp32 oracle (oneDNN, true fp32) must be matched to 1e-3 end to end; single-pass
// TF32 tensor-core math leaves only ~25% margin over 27 layers, so the fp32
// mode stays on CUDA cores with fp32 accumulation.  The bf16 production path is gemm_tc.cu.
// Classic register-blocked tiling: 128x64 block tile, 16-deep k slices, 8x4 outputs per thread.
#include "common.cuh"

namespace pn {

constexpr int SBM = 128, SBN = 64, SBK = 16;

template <int EPI>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float *__restrict__ A, const float *__restrict__ W,
                                                         int M, int K, int N, EpiParams ep) {
    __shared__ __align__(16) float As[SBK][SBM + 4];
    __shared__ __align__(16) float Ws[SBK][SBN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15;        // 16 column groups of 4
    const int ty = tid >> 4;        // 16 row groups of 8
    const int m0 = blockIdx.x * SBM, n0 = blockIdx.y * SBN;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += SBK) {
        // A tile 128x16: 512 float4, two per thread.  K % 4 == 0 is guaranteed by the caller.
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int f = tid + it * 256;
            const int row = f >> 2, kq = (f & 3) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + row < M && k0 + kq < K) v = __ldg(reinterpret_cast<const float4 *>(A + (size_t)(m0 + row) * K + k0 + kq));
            As[kq + 0][row] = v.x; As[kq + 1][row] = v.y; As[kq + 2][row] = v.z; As[kq + 3][row] = v.w;
        }
        {
            const int row = tid >> 2, kq = (tid & 3) * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + row < N && k0 + kq < K) v = __ldg(reinterpret_cast<const float4 *>(W + (size_t)(n0 + row) * K + k0 + kq));
            Ws[kq + 0][row] = v.x; Ws[kq + 1][row] = v.y; Ws[kq + 2][row] = v.z; Ws[kq + 3][row] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < SBK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Ws[kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    const int nb = n0 + tx * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
        if (EPI == EPI_RELU6) {
            if (nb + 3 < N) {
                const float4 bb = __ldg(reinterpret_cast<const float4 *>(ep.bias + nb));
                float4 o = make_float4(relu6f(acc[i][0] + bb.x), relu6f(acc[i][1] + bb.y), relu6f(acc[i][2] + bb.z),
                                       relu6f(acc[i][3] + bb.w));
                *reinterpret_cast<float4 *>(reinterpret_cast<float *>(ep.y) + (size_t)m * N + nb) = o;
            } else {
                for (int j = 0; j < 4; ++j)
                    if (nb + j < N) reinterpret_cast<float *>(ep.y)[(size_t)m * N + nb + j] = relu6f(acc[i][j] + ep.bias[nb + j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (nb + j < N) store_head(ep, m, nb + j, acc[i][j] + ep.bias[nb + j]);
        }
    }
}

int launch_gemm_simt(const float *a, const float *w, int m, int k, int n, int epi, const EpiParams &ep,
                     cudaStream_t st) {
    PN_CHECK_ARG(a && w && m > 0 && k > 0 && n > 0, "gemm(fp32): bad argument");
    PN_CHECK_ARG(k % 4 == 0, "gemm(fp32): K must be a multiple of 4 (got %d)", k);
    PN_CHECK_ARG(epi == EPI_HEADS || n % 4 == 0, "gemm(fp32): N must be a multiple of 4 (got %d)", n);
    dim3 grid(ceil_div(m, SBM), ceil_div(n, SBN));
    if (epi == EPI_RELU6)
        gemm_simt_kernel<EPI_RELU6><<<grid, 256, 0, st>>>(a, w, m, k, n, ep);
    else
        gemm_simt_kernel<EPI_HEADS><<<grid, 256, 0, st>>>(a, w, m, k, n, ep);
    PN_CHECK_LAUNCH();
    return PN_OK;
}

}  // namespace pn
