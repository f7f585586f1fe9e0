// B4 / H1, bf16 production path -- pointwise 1x1 convolutions and the fused 4-head projection as
// tcgen05 tensor-core GEMMs:  D[M,N] = A[M,K] . W[N,K]^T  (bf16 in, fp32 accumulate in TMEM).
// Replaces SeperableConv.pointwise + relu6 (posenet/models/mobilenet_v1.py:63,67) and the four head
// convs + sigmoid (:151-154,158-161) of the reference.
//
// Shape of the kernel (persistent, warp-specialised, one CTA per SM):
//   warp 0      TMA producer   cp.async.bulk.tensor 2D loads of a 128 x 64 A tile (activations, pixel-
//                              major == K-major) and a BLOCK_N x 64 W tile (OIHW 1x1 weight == K-major)
//                              into a STAGES-deep 128B-swizzled shared-memory ring; OOB rows / K tail
//                              are zero-filled by TMA, so ragged M (n*h*w) and K in {16,24,...} are free.
//   warp 1      MMA issuer     one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128,
//                              N=BLOCK_N, K=16) x 4 per stage, tcgen05.commit frees the stage / publishes
//                              the accumulator.
//   warp 2      TMEM allocator 2 x BLOCK_N fp32 columns: the accumulator is double-buffered so the
//                              epilogue of tile i overlaps the mainloop of tile i+1.
//   warps 4..7  epilogue       tcgen05.ld 32x32b (one accumulator row per thread), + bias, ReLU6 -> bf16,
//                              staged through 128B-swizzled shared memory in 128 x 64 panels and written
//                              with TMA stores (full-line writes; the M tail is clipped by the tensor map);
//                              the TMEM accumulator is released as soon as its last column is in registers.
//                              Heads: scatter (sigmoid on the 17 heatmap columns) into four fp32 NCHW
//                              tensors, coalesced across the warp's 32 pixels.
// All mbarrier waits are bounded: a protocol error traps instead of hanging the GPU.
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "ptx.cuh"

namespace pn {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;      // 64 bf16 = 128 B = one SWIZZLE_128B span
constexpr int TC_UMMA_K = 16;
constexpr int TC_THREADS = 256;
constexpr int TC_SMEM_MAX = 232448;           // 227 KB dynamic shared memory per CTA
constexpr int TC_STAGING_BYTES = 2 * 16384;   // two 128 x 64 bf16 output panels (ping-pong)
constexpr int TC_SMEM_BUDGET = TC_SMEM_MAX - 1024 /*alignment slack*/ - 256 /*barriers*/ - TC_STAGING_BYTES;

__host__ __device__ constexpr int tc_stage_bytes(int block_n) { return (TC_BLOCK_M + block_n) * TC_BLOCK_K * 2; }
__host__ __device__ constexpr int tc_stages(int block_n) {
    return TC_SMEM_BUDGET / tc_stage_bytes(block_n) > 8 ? 8 : TC_SMEM_BUDGET / tc_stage_bytes(block_n);
}
__host__ __device__ constexpr int tc_tmem_cols(int block_n) {
    return 2 * block_n <= 32 ? 32 : 2 * block_n <= 64 ? 64 : 2 * block_n <= 128 ? 128 : 2 * block_n <= 256 ? 256 : 512;
}
__host__ __device__ constexpr int tc_smem_bytes(int block_n) {
    return tc_stages(block_n) * tc_stage_bytes(block_n) + TC_STAGING_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1" format):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1)
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D fp32 (bits 4-5 = 1), A/B bf16 (bits 7-9, 10-12 = 1), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- the kernel ----------------------------------------------------------------------------------
template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, EpiParams ep, int M, int N, int K) {
    constexpr int STAGES = tc_stages(BLOCK_N);
    constexpr int A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;
    constexpr int STAGE_BYTES = tc_stage_bytes(BLOCK_N);
    constexpr uint32_t IDESC = make_idesc(TC_BLOCK_M, BLOCK_N);
    static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA N for M=128");
    constexpr bool TMA_STORE = (EPI == EPI_RELU6) && (BLOCK_N % 64 == 0);

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B needs 1024 B alignment
    const uint32_t staging = smem_base + STAGES * STAGE_BYTES;              // 2 x 16 KB, 1024 B aligned
    const uint32_t bars = staging + TC_STAGING_BYTES;                        // 8 B each
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bars + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
    uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t *>(smem_gen + STAGES * STAGE_BYTES + TC_STAGING_BYTES + 8 * (2 * STAGES + 4));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m_tiles = (M + TC_BLOCK_M - 1) / TC_BLOCK_M;
    const int num_n_tiles = N / BLOCK_N;
    const int num_tiles = num_m_tiles * num_n_tiles;
    const int num_k_blocks = (K + TC_BLOCK_K - 1) / TC_BLOCK_K;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        if (TMA_STORE) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_c) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 128);      // every epilogue thread arrives
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                     "r"((uint32_t)tc_tmem_cols(BLOCK_N))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (EPI == EPI_HEADS && threadIdx.x < 128) {                             // head bias (128 padded entries) -> the staging area
        reinterpret_cast<float *>(smem_gen + STAGES * STAGE_BYTES)[threadIdx.x] = threadIdx.x < (unsigned)N ? __ldg(ep.bias + threadIdx.x) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_launch_dependents();                                  // after the TMEM allocation is made (see stem.cu): a dependent CTA must not allocate first
    const uint32_t tmem_base = *tmem_slot_ptr;
    pdl_wait();                                                              // (ptx.cuh) the previous layer has completed

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_tile = tile / num_n_tiles, n_tile = tile % num_n_tiles;
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    mbar_expect_tx(full_bar(stage), STAGE_BYTES);
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    tma_load_2d(sa, &tmap_a, full_bar(stage), kb * TC_BLOCK_K, m_tile * TC_BLOCK_M);
                    tma_load_2d(sa + A_BYTES, &tmap_b, full_bar(stage), kb * TC_BLOCK_K, n_tile * BLOCK_N);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1);          // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);               // TMA bytes have landed
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * STAGE_BYTES;
                    const uint32_t sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < TC_BLOCK_K / TC_UMMA_K; ++k) {
                        // advancing 16 bf16 = 32 B along K inside the 128 B swizzle span
                        const uint64_t adesc = make_smem_desc(sa + k * TC_UMMA_K * 2);
                        const uint64_t bdesc = make_smem_desc(sb + k * TC_UMMA_K * 2);
                        tc_mma_bf16(d_tmem, adesc, bdesc, IDESC, (uint32_t)((kb | k) != 0));
                    }
                    tc_commit(empty_bar(stage));                     // smem stage reusable once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(tfull_bar(acc));                           // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int q = warp & 3;                                      // TMEM lane quarter this warp may read
        const int row_in_tile = q * 32 + lane;
        const bool issuer = (threadIdx.x == 128);                    // one thread owns the bulk-store groups
        int acc = 0, buf = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_tile = tile / num_n_tiles, n_tile = tile % num_n_tiles;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const int row = m_tile * TC_BLOCK_M + row_in_tile;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            if constexpr (TMA_STORE) {
#pragma unroll 1
                for (int p = 0; p < BLOCK_N / 64; ++p) {
                    uint32_t v[64];
                    tc_ld32(taddr + (uint32_t)(p * 64), v);
                    tc_ld32(taddr + (uint32_t)(p * 64 + 32), v + 32);
                    tc_ld_wait();
                    if (p == BLOCK_N / 64 - 1) {                     // accumulator fully in registers: hand it back
                        tc_fence_before();
                        mbar_arrive(tempty_bar(acc));
                    }
                    if (issuer) bulk_wait_read<1>();                 // the store that last read this buffer is done
                    epi_bar_sync();
                    const int col0 = n_tile * BLOCK_N + p * 64;
                    const float4 *bp = reinterpret_cast<const float4 *>(ep.bias + col0);
                    const uint32_t srow = staging + (uint32_t)buf * 16384u + (uint32_t)row_in_tile * 128u;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {                    // 8 columns -> one 16 B chunk, 128B-swizzled
                        const float4 b0 = __ldg(bp + 2 * c), b1 = __ldg(bp + 2 * c + 1);
                        const float2 s0 = fadd2(make_float2(__uint_as_float(v[8 * c + 0]), __uint_as_float(v[8 * c + 1])), make_float2(b0.x, b0.y));
                        const float2 s1 = fadd2(make_float2(__uint_as_float(v[8 * c + 2]), __uint_as_float(v[8 * c + 3])), make_float2(b0.z, b0.w));
                        const float2 s2 = fadd2(make_float2(__uint_as_float(v[8 * c + 4]), __uint_as_float(v[8 * c + 5])), make_float2(b1.x, b1.y));
                        const float2 s3 = fadd2(make_float2(__uint_as_float(v[8 * c + 6]), __uint_as_float(v[8 * c + 7])), make_float2(b1.z, b1.w));
                        st_shared_v4(srow + (uint32_t)((c ^ (row_in_tile & 7)) << 4), relu6_bf16x2(s0), relu6_bf16x2(s1), relu6_bf16x2(s2),
                                     relu6_bf16x2(s3));
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to TMA
                    epi_bar_sync();
                    if (issuer) {
                        tma_store_2d(&tmap_c, staging + (uint32_t)buf * 16384u, col0, m_tile * TC_BLOCK_M);
                        bulk_commit();
                    }
                    buf ^= 1;
                }
            } else if constexpr (EPI == EPI_HEADS) {
                // Scatter to four fp32 NCHW tensors.  Fully unrolled over the 115 columns, so the column -> (tensor, channel)
                // mapping, the bias offsets and the sigmoid split are compile-time; per element: one shared-memory bias read
                // (4 per LDS.128), one add, one pointer add, one 4-byte store that the warp's 32 pixels coalesce into a
                // 128 B line.  The accumulator is handed back as soon as its last column is in registers.
                static_assert(BLOCK_N == 128, "heads epilogue expects one 128-column tile (115 head channels + padding)");
                const bool valid = row < M;
                const int img = valid ? row / ep.hw : 0, pix = valid ? row - img * ep.hw : 0;
                const size_t hw = (size_t)ep.hw;
                float *ph = ep.heat + (size_t)img * 17 * hw + pix, *po = ep.off + (size_t)img * 34 * hw + pix;
                float *pf = ep.fwd + (size_t)img * 32 * hw + pix, *pb = ep.bwd + (size_t)img * 32 * hw + pix;
#pragma unroll
                for (int c = 0; c < BLOCK_N; c += 16) {
                    uint32_t v[16];
                    tc_ld16(taddr + (uint32_t)c, v);
                    tc_ld_wait();
                    if (c + 16 == BLOCK_N) {
                        tc_fence_before();
                        mbar_arrive(tempty_bar(acc));
                    }
                    float bs[16];
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const uint4 b4 = ld_shared_v4(staging + (uint32_t)(c + j) * 4u);
                        bs[j] = __uint_as_float(b4.x); bs[j + 1] = __uint_as_float(b4.y);
                        bs[j + 2] = __uint_as_float(b4.z); bs[j + 3] = __uint_as_float(b4.w);
                    }
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int col = c + j;                       // compile-time
                            const float x = __uint_as_float(v[j]) + bs[j];
                            if (col < 17) ph[(size_t)col * hw] = 1.f / (1.f + expf(-x));          // mobilenet_v1.py:158
                            else if (col < 51) po[(size_t)(col - 17) * hw] = x;
                            else if (col < 83) pf[(size_t)(col - 51) * hw] = x;
                            else if (col < 115) pb[(size_t)(col - 83) * hw] = x;
                        }
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 16) {
                    uint32_t v[16];
                    tc_ld16(taddr + (uint32_t)c, v);
                    tc_ld_wait();
                    const int col0 = n_tile * BLOCK_N + c;
                    if (row < M) {
                        const float4 *bp = reinterpret_cast<const float4 *>(ep.bias + col0);
                        uint32_t packed[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 b = __ldg(bp + j);
                            packed[2 * j] = relu6_bf16x2(__uint_as_float(v[4 * j + 0]) + b.x, __uint_as_float(v[4 * j + 1]) + b.y);
                            packed[2 * j + 1] = relu6_bf16x2(__uint_as_float(v[4 * j + 2]) + b.z, __uint_as_float(v[4 * j + 3]) + b.w);
                        }
                        uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(ep.y) + (size_t)row * N + col0);
                        dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                        dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                    }
                }
                tc_fence_before();
                mbar_arrive(tempty_bar(acc));
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (TMA_STORE && issuer) bulk_wait_all();                    // smem must outlive the last bulk store
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)tc_tmem_cols(BLOCK_N))
                     : "memory");
    }
}

// ---- host side -------------------------------------------------------------------------------------
// 2D bf16 row-major [rows, cols] tensor, box = [box_rows, 64 cols], 128B swizzle, zero OOB fill.
static int encode_2d(void *out, const void *base, int rows, int cols, int box_rows) {
    const uint64_t dims[2] = {(uint64_t)cols, (uint64_t)rows};
    const uint64_t strides[1] = {(uint64_t)cols * 2};
    const uint32_t box[2] = {(uint32_t)TC_BLOCK_K, (uint32_t)box_rows};
    return encode_tmap(out, base, 2, 2, dims, strides, box, 3);
}

// Largest supported tile width that divides N.
static int pick_block_n(int n) {
    static const int cand[] = {256, 192, 128, 96, 64, 48, 32, 16};
    for (int c : cand)
        if (n % c == 0) return c;
    return 0;
}

int gemm_tc_prepare(GemmTc *g, const void *a, const void *w, void *y, int m, int k, int n, int epi) {
    PN_CHECK_ARG(a && w && m > 0 && k > 0 && n > 0, "gemm(bf16): bad argument");
    PN_CHECK_ARG(k % 8 == 0, "gemm(bf16): K must be a multiple of 8 (TMA 16-byte row pitch), got %d", k);
    PN_CHECK_ARG(((uintptr_t)a & 15) == 0 && ((uintptr_t)w & 15) == 0, "gemm(bf16): operands must be 16-byte aligned");
    const int bn = pick_block_n(n);
    PN_CHECK_ARG(bn != 0, "gemm(bf16): N must be a multiple of 16 (got %d)", n);
    g->m = m; g->k = k; g->n = n; g->block_n = bn; g->epi = epi;
    int rc = encode_2d(g->tmap_a, a, m, k, TC_BLOCK_M);
    if (rc != PN_OK) return rc;
    rc = encode_2d(g->tmap_b, w, n, k, bn);
    if (rc != PN_OK) return rc;
    if (epi == EPI_RELU6 && bn % 64 == 0) {          // output written with TMA stores in 128 x 64 panels
        PN_CHECK_ARG(y && ((uintptr_t)y & 15) == 0, "gemm(bf16): output must be 16-byte aligned");
        return encode_2d(g->tmap_c, y, m, n, TC_BLOCK_M);
    }
    memcpy(g->tmap_c, g->tmap_a, sizeof(g->tmap_c));  // unused by this instantiation; keep it a valid map
    return PN_OK;
}

template <int BLOCK_N, int EPI>
static int launch_tc(const GemmTc *g, const EpiParams &ep, cudaStream_t st) {
    static DeviceOnce once;
    const int dev = current_device();
    auto kern = gemm_tc_kernel<BLOCK_N, EPI>;
    constexpr int smem = tc_smem_bytes(BLOCK_N);
    if (!once.get(dev)) {
        PN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        once.set(dev, 1);
    }
    const int tiles = ceil_div(g->m, TC_BLOCK_M) * (g->n / BLOCK_N);
    const int grid = tiles < num_sms() ? tiles : num_sms();
    PN_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(TC_THREADS), (size_t)smem, st, *reinterpret_cast<const CUtensorMap *>(g->tmap_a),
                             *reinterpret_cast<const CUtensorMap *>(g->tmap_b), *reinterpret_cast<const CUtensorMap *>(g->tmap_c), ep, g->m,
                             g->n, g->k));
    return PN_OK;
}

int gemm_tc_launch(const GemmTc *g, const EpiParams &ep, cudaStream_t st) {
    if (g->epi == EPI_HEADS) {                                               // the heads are one 128-column tile (115 + padding)
        PN_CHECK_ARG(g->block_n == 128 && g->n == 128, "pn_heads_gemm: expected %d packed head rows", PN_HEAD_ROWS);
        return launch_tc<128, EPI_HEADS>(g, ep, st);
    }
#define PN_TC_CASE(BN)                                                       \
    if (g->block_n == BN) return launch_tc<BN, EPI_RELU6>(g, ep, st);
    PN_TC_CASE(256)
    PN_TC_CASE(192)
    PN_TC_CASE(128)
    PN_TC_CASE(96)
    PN_TC_CASE(64)
    PN_TC_CASE(48)
    PN_TC_CASE(32)
    PN_TC_CASE(16)
#undef PN_TC_CASE
    set_error("gemm(bf16): no kernel for block_n %d", g->block_n);
    return PN_ERR_UNSUPPORTED;
}

}  // namespace pn
