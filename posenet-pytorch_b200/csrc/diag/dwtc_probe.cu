// Diagnostic kernel for the tensor-pipe depthwise (sepconv_tc.cu): ONE 128-position chunk of one 64-channel image, linear
// control flow, every intermediate written out.  It pins down three hardware behaviours the production kernel relies on:
//   (1) a K-major SWIZZLE_128B UMMA descriptor whose start address is a whole number of 128-byte rows (not a multiple of
//       1024 bytes) into a TMA-written patch reads rows [start, start + 128) with the swizzle phase of the ABSOLUTE address
//       (flag bit 0 additionally sets the descriptor's base-offset field to (addr >> 7) & 7 for comparison);
//   (2) tcgen05.mma with N = 16 against block-diagonal weight tiles accumulates the 9 taps per 16-channel group;
//   (3) tcgen05.st of packed bf16 pairs + tcgen05.mma with the A operand in TMEM (the pointwise conv on the depthwise result).
#include <cuda.h>

#include "../common.cuh"
#include "../ptx.cuh"
#include "../../../include/posenet_b200_diag.h"

namespace pn {

struct DwtcProbeArgs {
    int wp, dil, qoff, rows_box, x_org, y_org, flags;
};

__device__ __forceinline__ uint64_t probe_desc(uint32_t saddr, int with_base_offset) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    if (with_base_offset) d |= (uint64_t)((saddr >> 7) & 7) << 49;
    return d;
}
__device__ __forceinline__ uint32_t probe_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void tc_mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tc_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(128, 1)
dwtc_probe_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_diag,
                  const __grid_constant__ CUtensorMap tm_w, const float *dw_bias, float *out_dw, float *out_pw, DwtcProbeArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t s_diag = base;                       // 9 x [16 x 64] bf16 = 18432 B
    const uint32_t s_w = base + 18432;                  // [64 x 64] bf16 = 8192 B
    const uint32_t s_patch = base + 18432 + 8192;       // rows_box * wp rows of 128 B (+ slack)
    __shared__ __align__(8) uint64_t bars[3];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar_load = smem_u32(&bars[0]), bar_dw = smem_u32(&bars[1]), bar_pw = smem_u32(&bars[2]);
    if (tid == 0) {
        mbar_init(bar_load, 1);
        mbar_init(bar_dw, 1);
        mbar_init(bar_pw, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // zero the patch slack so rows past the box are finite
    for (uint32_t o = tid * 16; o < (uint32_t)(a.rows_box * a.wp + 160) * 128u; o += 128 * 16) st_shared_v4(s_patch + o, 0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)a.rows_box * a.wp * 128u + 18432u + 8192u;
        mbar_expect_tx(bar_load, bytes);
        tma_load_4d(s_patch, &tm_x, bar_load, 0, a.x_org, a.y_org, 0);
        tma_load_2d(s_diag, &tm_diag, bar_load, 0, 0);
        tma_load_2d(s_w, &tm_w, bar_load, 0, 0);
        mbar_wait(bar_load, 0);
        tc_fence_after();
        const uint32_t idesc16 = probe_idesc(16);
        for (int g = 0; g < 4; ++g)
            for (int t = 0; t < 9; ++t) {
                const int shift = (t / 3) * a.dil * a.wp + (t % 3) * a.dil + a.qoff;
                const uint64_t ad = probe_desc(s_patch + (uint32_t)shift * 128u + g * 32, a.flags & 1);
                const uint64_t bd = probe_desc(s_diag + t * 2048 + g * 32, 0);
                tc_mma_bf16(tmem + g * 16, ad, bd, idesc16, t > 0);
            }
        tc_commit(bar_dw);
    }
    __syncwarp();
    mbar_wait(bar_dw, 0);
    tc_fence_after();
    {   // depthwise accumulator -> + bias, ReLU6 -> bf16 pairs: written out and stored to TMEM as the A operand
        uint32_t v[64], pk[32];
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        tc_ld32(lane_addr, v);
        tc_ld32(lane_addr + 32, v + 32);
        tc_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float lo = __uint_as_float(v[2 * j]) + dw_bias[2 * j], hi = __uint_as_float(v[2 * j + 1]) + dw_bias[2 * j + 1];
            pk[j] = relu6_bf16x2(lo, hi);
            out_dw[tid * 64 + 2 * j] = __uint_as_float(pk[j] << 16);
            out_dw[tid * 64 + 2 * j + 1] = __uint_as_float(pk[j] & 0xffff0000u);
        }
        tc_st32(lane_addr + 64, pk);
        tc_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t idesc64 = probe_idesc(64);
        for (int k = 0; k < 4; ++k) tc_mma_bf16_ts(tmem + 128, tmem + 64 + k * 8, probe_desc(s_w + k * 32, 0), idesc64, k > 0);
        tc_commit(bar_pw);
    }
    __syncwarp();
    mbar_wait(bar_pw, 0);
    tc_fence_after();
    {
        uint32_t v[64];
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16) + 128;
        tc_ld32(lane_addr, v);
        tc_ld32(lane_addr + 32, v + 32);
        tc_ld_wait();
#pragma unroll
        for (int j = 0; j < 64; ++j) out_pw[tid * 64 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

}  // namespace pn

// x: bf16 [h, wd, 64]; diag: bf16 [9*16, 64] (tap-major tiles: row n of tile t holds w[t][16 g + n] at column 16 g + n);
// pw_w: bf16 [64, 64] (row = output channel); dw_bias: f32 [64]; out_dw / out_pw: f32 [128, 64].
extern "C" int pn_dwtc_probe(const void *x, int h, int wd, const void *diag, const void *pw_w, const float *dw_bias, float *out_dw,
                             float *out_pw, int wp, int dil, int qoff, int rows_box, int x_org, int y_org, int flags,
                             pn_stream_t stream) {
    using namespace pn;
    alignas(64) CUtensorMap tx, td, tw;
    {
        const uint64_t dims[4] = {64, (uint64_t)wd, (uint64_t)h, 1};
        const uint64_t str[3] = {128, (uint64_t)wd * 128, (uint64_t)wd * h * 128};
        const uint32_t box[4] = {64, (uint32_t)wp, (uint32_t)rows_box, 1};
        int r = encode_tmap(&tx, x, 2, 4, dims, str, box, 3);
        if (r != PN_OK) return r;
    }
    {
        const uint64_t dims[2] = {64, 144};
        const uint64_t str[1] = {128};
        const uint32_t box[2] = {64, 144};
        int r = encode_tmap(&td, diag, 2, 2, dims, str, box, 3);
        if (r != PN_OK) return r;
    }
    {
        const uint64_t dims[2] = {64, 64};
        const uint64_t str[1] = {128};
        const uint32_t box[2] = {64, 64};
        int r = encode_tmap(&tw, pw_w, 2, 2, dims, str, box, 3);
        if (r != PN_OK) return r;
    }
    const size_t smem = 1024 + 18432 + 8192 + (size_t)(rows_box * wp + 160) * 128;
    PN_CHECK_ARG(smem <= 227 * 1024, "pn_dwtc_probe: patch too large");
    PN_CHECK_CUDA(cudaFuncSetAttribute(dwtc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DwtcProbeArgs a{wp, dil, qoff, rows_box, x_org, y_org, flags};
    dwtc_probe_kernel<<<1, 128, smem, as_stream(stream)>>>(tx, td, tw, dw_bias, out_dw, out_pw, a);
    PN_CHECK_LAUNCH();
    return PN_OK;
}

// ---- UMMA cost microbenchmark (diagnostics): cycles per tcgen05.mma for small N and different operand layouts -----------------
namespace pn {
__global__ void __launch_bounds__(128, 1) umma_cost_kernel(int n, int layout, int reps, int a_step16, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bars[1];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar0 = smem_u32(&bars[0]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (uint32_t o = tid * 16; o < 96u * 1024u; o += 128 * 16) st_shared_v4(base + o, 0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        // layout 0: K-major SWIZZLE_128B (128-byte rows, SBO 1024); layout 1: K-major SWIZZLE_32B (32-byte rows, SBO 256);
        // layout 2: K-major SWIZZLE_64B (64-byte rows, SBO 512)
        const uint64_t sbo = layout == 0 ? 1024 : layout == 1 ? 256 : 512;
        const uint64_t lt = layout == 0 ? 2 : layout == 1 ? 6 : 4;
        const uint64_t hi = (1ull << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (lt << 61);
        const uint32_t a0 = (base & 0x3FFFF) >> 4, b0 = ((base + 65536u) & 0x3FFFF) >> 4;
        const uint32_t idesc = probe_idesc(n);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int u = 0; u < 9; ++u) tc_mma_bf16(tmem, hi | (uint64_t)(a0 + u * a_step16), hi | (uint64_t)(b0 + u * 128), idesc, 1);
        }
        const long long t1 = clock64();
        tc_commit(bar0);
        mbar_wait(bar0, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0;
        out[1] = t2 - t0;
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// Same measurement with the issue loop executed by the WHOLE warp (uniform control flow, operands derived from kernel
// parameters only) and the MMA itself predicated on elect.sync -- lets ptxas keep the descriptors in uniform registers.
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}
__global__ void __launch_bounds__(128, 1) umma_cost_uniform_kernel(int n, int reps, int a_step16, long long *out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bars[1];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t bar0 = smem_u32(&bars[0]);
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (uint32_t o = tid * 16; o < 96u * 1024u; o += 128 * 16) st_shared_v4(base + o, 0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const uint64_t hi = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint32_t a0 = (base & 0x3FFFF) >> 4, b0 = ((base + 65536u) & 0x3FFFF) >> 4;
        const uint32_t idesc = probe_idesc(n);
        const uint32_t leader = elect_one();
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int u = 0; u < 9; ++u)
                if (leader) tc_mma_bf16(tmem, hi | (uint64_t)(a0 + u * a_step16), hi | (uint64_t)(b0 + u * 128), idesc, 1);
        }
        const long long t1 = clock64();
        if (leader) tc_commit(bar0);
        __syncwarp();
        mbar_wait(bar0, 0);
        const long long t2 = clock64();
        if (leader) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
}  // namespace pn

// out_host[0] = cycles until the last MMA was issued, out_host[1] = cycles until all completed; 9 * reps MMAs of M128 x N x K16
extern "C" int pn_debug_umma_cost(int n, int layout, int reps, int a_step16, long long *out_host) {
    using namespace pn;
    long long *d = nullptr;
    PN_CHECK_CUDA(cudaMalloc(&d, 16));
    const size_t smem = 1024 + 96 * 1024;
    PN_CHECK_CUDA(cudaFuncSetAttribute(umma_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (layout == 3) {
        PN_CHECK_CUDA(cudaFuncSetAttribute(umma_cost_uniform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        umma_cost_uniform_kernel<<<1, 128, smem>>>(n, reps, a_step16, d);
    } else
        umma_cost_kernel<<<1, 128, smem>>>(n, layout, reps, a_step16, d);
    PN_CHECK_LAUNCH();
    PN_CHECK_CUDA(cudaMemcpy(out_host, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return PN_OK;
}
