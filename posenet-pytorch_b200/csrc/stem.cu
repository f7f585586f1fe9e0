// B2 -- stem: relu6(conv3x3(x, stride s, pad 1) + b), 3 -> COUT (16 | 24 | 32), NHWC out.
// Replaces InputConv (posenet/models/mobilenet_v1.py:47-54).  K = 27 is far too small for the tensor
// cores; the layer is bounded by its output write (COUT*2 B per pixel) and FFMA issue.
// One thread per output pixel, all COUT channels in registers, weights broadcast from shared memory.
// Input is either the reference's f32 NCHW tensor, or (fused P1, identity resize) the uint8 BGR
// HWC image itself, normalised on the fly with the same two rounded fp32 ops as utils.py:23.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

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

// ---- tensor-core stem for the production path (uint8 image in, bf16 NHWC out) ----------------------------------
// The SIMT kernel above is bound by FFMA issue (27 x COUT FMAs per pixel, 6x the HBM time).  Here the layer is an
// im2col GEMM on tcgen05: per tile of 128 output pixels, A[128, 32] holds the 27 taps (+5 zero columns) as bf16 and
// W[32, 32] the weights; two UMMAs (K = 16) replace 864 FMAs per pixel.  Pixel bytes are exact in bf16, so the
// normalisation x*(2/255)-1 of utils.py:23 is folded into the operands without losing the pixel value:
//     x*(2/255) - 1 = (x - 128)*(2/255) + 1/255,   A = x - 128 (exact, in [-128, 127]),   W' = bf16(w * 2/255),
//     bias' = b + sum_k w_k / 255,   and a padded tap holds -0.5, which normalises to exactly 0 (the reference's zero
// padding of the normalised image).  Centring on 128 avoids the cancellation a raw-pixel formulation would have.
// One CTA = 128 pixel threads + a control warp; several CTAs per SM.

// K-major SWIZZLE_64B: the stem's K is 32 (27 taps + 5 zero columns) = 64-byte rows, 8-row groups 512 B apart.  Half the shared
// memory of the 128-byte-row layout for A and W (10 KB instead of 20 KB per CTA), i.e. 7 instead of 5 CTAs per SM.
__device__ __forceinline__ uint64_t stc_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
constexpr int STC_A_BYTES = 128 * 64, STC_W_BYTES = 32 * 64;

__device__ __forceinline__ uint32_t lds_u32s(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}

struct StemTcArgs {
    const uint8_t *img;
    const float *w27, *bias;
    __nv_bfloat16 *y;
    int n, h, w, ho, wo, stride, cout;
    long long total_px, total_bytes, total_lo16;   // pixels; bytes of the image batch; the same rounded DOWN to 16
    int span_cap;                          // bytes per span buffer (multiple of 16)
    int seg, piece_cap;                    // segmented spans (wide images): 6 pieces of piece_cap bytes per buffer, see the kernel
    int nbuf;                              // span ring slots (stem_tc_kernel)
};
constexpr int STC_PIECES = 6;              // 2 output-row segments x 3 input rows

// The CTA is a software pipeline (round 2; the first version ran a tile as one chain -- im2col, CTA barrier, MMA, commit wait,
// TMEM read, stores -- in which every warp sat out the barrier + MMA round trip, warp 0 additionally walked the refill's span
// arithmetic with the other three waiting for it, and ~200 of a warp's ~580 instructions per tile were the (image, row, column)
// divisions of its pixel index: C2 85.9 -> 80.1 us, C3 169.7 -> 126.9 us, C4 202.7 -> 159.0 us, same bits):
//   * a fifth warp owns the control work: it waits for "A complete" (an mbarrier the 128 pixel threads arrive on), issues the
//     tile's two MMAs and refills the span slot the tile has just released;
//   * A and the accumulator are double-buffered, so a pixel warp goes from the im2col of tile i straight to the epilogue of
//     tile i - 1 (whose MMAs retired long ago) and on to the im2col of tile i + 1: it never waits for a round trip;
//   * the pixel index advances incrementally (conditional subtractions instead of divisions).
// Input staging, per tile, into a ring of a.nbuf slots:
//   * whole rows (default): the tile's input bytes are one contiguous span of the image batch [lo16, lo16 + size), one bulk copy;
//     the <= 15 bytes between total_lo16 and the real end of the caller's buffer are fetched with ordinary byte loads by the
//     issuing thread BEFORE its arrive.expect_tx (release) on the slot's barrier, so the consumers' wait (acquire) covers them;
//   * segments (a.seg: images at least 128 output pixels wide whose whole rows would leave too few CTAs per SM, or do not fit
//     at all -- 19 KB instead of 2.3 KB per tile at 1281 x 721): a 128-pixel tile lies in at most two output rows; per
//     output-row segment and window row ky ONE bulk copy of just the columns the segment reads (16-byte aligned outwards),
//     issued by lane p = segment * 3 + ky with its own expect_tx, so the slot's barrier counts STC_PIECES arrivals.
constexpr int STP_THREADS = 160;
constexpr int STP_MAX_NBUF = 4;
constexpr int STP_CTRL_BYTES = 384;         // bias 128 | span_full[4] a_full[2] mma_done[2] 64 | tmem slot 16 | span_lo[4] 32 | piece_adj[24] 96

__global__ void __launch_bounds__(STP_THREADS) stem_tc_kernel(const StemTcArgs a) {
    extern __shared__ uint8_t stc_raw[];
    const uint32_t base = (smem_u32(stc_raw) + 1023u) & ~1023u;
    uint8_t *gen = stc_raw + (base - smem_u32(stc_raw));
    const uint32_t sA = base;                                     // 2 x (128 rows x 64 B), 64B swizzle
    const uint32_t sW = base + 2 * STC_A_BYTES;                   // 32 rows x 64 B
    const uint32_t sSpan = sW + STC_W_BYTES;                      // a.nbuf x span_cap
    const uint32_t ctrl_off = 2 * STC_A_BYTES + STC_W_BYTES + (uint32_t)a.nbuf * (uint32_t)a.span_cap;
    float *sBias = reinterpret_cast<float *>(gen + ctrl_off);
    const uint32_t span_full = base + ctrl_off + 128, a_full = span_full + 32, mma_done = a_full + 16, tmem_slot_addr = mma_done + 16;
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen + ctrl_off + 192);
    volatile long long *span_lo = reinterpret_cast<volatile long long *>(gen + ctrl_off + 208);
    volatile int *piece_adj = reinterpret_cast<volatile int *>(gen + ctrl_off + 240);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long row_bytes = (long long)a.w * 3;
    const long long num_tiles = (a.total_px + 127) / 128;

    if (tid == 0) {
        for (int i = 0; i < a.nbuf; ++i) mbar_init(span_full + 8u * i, a.seg ? STC_PIECES : 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(a_full + 8u * i, 128);
            mbar_init(mma_done + 8u * i, 1);
        }
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot_addr), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // W'[n][k] = bf16(w[k][n] * 2/255) in the swizzled K-major layout; bias' = b + sum_k w_k / 255
    for (int i = tid; i < 32 * 4; i += STP_THREADS) {               // (row n, 16-byte chunk c) -> 8 k values
        const int nrow = i >> 2, c = i & 3;
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k0 = c * 8 + 2 * j;
            const float w0 = (nrow < a.cout && k0 < 27) ? a.w27[k0 * a.cout + nrow] * (float)(2.0 / 255.0) : 0.f;
            const float w1 = (nrow < a.cout && k0 + 1 < 27) ? a.w27[(k0 + 1) * a.cout + nrow] * (float)(2.0 / 255.0) : 0.f;
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(w0, w1);
            pk[j] = *reinterpret_cast<const uint32_t *>(&h2);
        }
        st_shared_v4(sW + (uint32_t)nrow * 64u + (uint32_t)((c ^ ((nrow >> 1) & 3)) << 4), pk[0], pk[1], pk[2], pk[3]);
    }
    if (tid < 32) {
        float sum = 0.f;
        if (tid < a.cout)
            for (int k = 0; k < 27; ++k) sum += a.w27[k * a.cout + tid];
        sBias[tid] = tid < a.cout ? a.bias[tid] + sum * (float)(1.0 / 255.0) : 0.f;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // only now may the next kernel's CTAs start their prologue on this SM: the TMEM allocation above has been made (a dependent
    // that allocated first could leave this CTA blocked in tcgen05.alloc while it waits in griddepcontrol.wait for this grid)
    pdl_launch_dependents();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();                                                   // (ptx.cuh) whatever produced the image / used the output buffer is done
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    if (warp == 4) {
        // ===================== control warp: span loads, MMA issue =====================
        auto issue_span = [&](int slot, long long t) {
            const uint32_t m0 = (uint32_t)t * 128u, m1 = min(m0 + 127u, (uint32_t)a.total_px - 1u);
            const uint32_t r0 = m0 / (uint32_t)a.wo, r1 = m1 / (uint32_t)a.wo;
            const int i0 = (int)(r0 / (uint32_t)a.ho), i1 = (int)(r1 / (uint32_t)a.ho);
            const int oy0 = (int)(r0 - (uint32_t)i0 * (uint32_t)a.ho), oy1 = (int)(r1 - (uint32_t)i1 * (uint32_t)a.ho);
            const int iy0 = max(oy0 * a.stride - 1, 0), iy1 = min(oy1 * a.stride + 1, a.h - 1);
            const long long lo = ((long long)i0 * a.h + iy0) * row_bytes, hi = ((long long)i1 * a.h + iy1 + 1) * row_bytes;
            const long long lo16 = lo & ~15ll;
            long long end = (hi + 15) & ~15ll;
            if (end > a.total_lo16) end = a.total_lo16;
            const uint32_t size = (uint32_t)(end - lo16);
            span_lo[slot] = lo16;
            const uint32_t dst = sSpan + (uint32_t)slot * (uint32_t)a.span_cap;
            if (lo16 + size == a.total_lo16)
                for (long long b = a.total_lo16; b < a.total_bytes; ++b)
                    asm volatile("st.shared.u8 [%0], %1;" ::"r"(dst + (uint32_t)(b - lo16)), "r"((uint32_t)a.img[b]) : "memory");
            mbar_expect_tx(span_full + 8u * slot, size);
            if (size) bulk_load_1d(dst, a.img + lo16, size, span_full + 8u * slot);
        };
        auto issue_piece = [&](int slot, long long t, int p) {
            const uint32_t m0 = (uint32_t)t * 128u, m1 = min(m0 + 127u, (uint32_t)a.total_px - 1u);
            const uint32_t r0 = m0 / (uint32_t)a.wo, r1 = m1 / (uint32_t)a.wo;
            const int sg = p / 3, ky = p - sg * 3;
            const uint32_t r = r0 + (uint32_t)sg;
            const uint32_t bar = span_full + 8u * slot;
            if (r <= r1) {
                const int img_i = (int)(r / (uint32_t)a.ho), oy = (int)(r - (uint32_t)img_i * (uint32_t)a.ho);
                const int xa = sg == 0 ? (int)(m0 - r0 * (uint32_t)a.wo) : 0, xb = r == r1 ? (int)(m1 - r1 * (uint32_t)a.wo) : a.wo - 1;
                const int iy = oy * a.stride - 1 + ky;
                if (iy >= 0 && iy < a.h) {
                    const long long rowstart = ((long long)img_i * a.h + iy) * row_bytes;
                    const long long lo = rowstart + (long long)max(xa * a.stride - 1, 0) * 3, hi = rowstart + (long long)min(xb * a.stride + 1, a.w - 1) * 3 + 3;
                    const long long lo16 = lo & ~15ll;
                    long long end = (hi + 15) & ~15ll;
                    if (end > a.total_lo16) end = a.total_lo16;
                    const uint32_t size = end > lo16 ? (uint32_t)(end - lo16) : 0u;
                    const uint32_t dst = sSpan + (uint32_t)slot * (uint32_t)a.span_cap + (uint32_t)p * (uint32_t)a.piece_cap;
                    piece_adj[slot * STC_PIECES + p] = p * a.piece_cap - (int)(lo16 - rowstart);
                    for (long long b = max(a.total_lo16, lo16); b < hi; ++b)
                        asm volatile("st.shared.u8 [%0], %1;" ::"r"(dst + (uint32_t)(b - lo16)), "r"((uint32_t)a.img[b]) : "memory");
                    mbar_expect_tx(bar, size);
                    if (size) bulk_load_1d(dst, a.img + lo16, size, bar);
                    return;
                }
            }
            mbar_arrive(bar);
        };
        const int npiece = a.seg ? STC_PIECES : 1;
        auto issue = [&](int slot, long long t) {
            if (lane < npiece) {
                if (a.seg) issue_piece(slot, t, lane); else issue_span(slot, t);
            }
        };
        {
            long long t = blockIdx.x;
            for (int i = 0; i < a.nbuf && t < num_tiles; ++i, t += gridDim.x) issue(i, t);
        }
        int slot = 0;
        uint32_t it = 0;
        for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const uint32_t s = it & 1u;
            mbar_wait(a_full + 8u * s, (it >> 1) & 1u);              // all 128 rows of A[s] written; the tile's span slot is free;
            tc_fence_after();                                       // accumulator s was read out two tiles ago
            if (lane == 0) {
                const uint32_t sa = sA + s * (uint32_t)STC_A_BYTES, d = tmem + s * 32u;
                tc_mma_bf16(d, stc_smem_desc(sa), stc_smem_desc(sW), IDESC, 0u);
                tc_mma_bf16(d, stc_smem_desc(sa + 32), stc_smem_desc(sW + 32), IDESC, 1u);
                tc_commit(mma_done + 8u * s);
            }
            const long long next = tile + (long long)a.nbuf * gridDim.x;
            if (next < num_tiles) issue(slot, next);
            if (++slot == a.nbuf) slot = 0;
            __syncwarp();
        }
    } else {
        // ===================== pixel warps: im2col of tile i, epilogue of tile i - 1 =====================
        const uint32_t step = 128u * gridDim.x;                      // pixels between a thread's consecutive tiles
        uint32_t m = blockIdx.x * 128u + (uint32_t)tid;
        int ox, oy, img_i;
        {
            const uint32_t r = m / (uint32_t)a.wo;
            ox = (int)(m - r * (uint32_t)a.wo);
            img_i = (int)(r / (uint32_t)a.ho);
            oy = (int)(r - (uint32_t)img_i * (uint32_t)a.ho);
        }
        const uint32_t drs = step / (uint32_t)a.wo;
        const int dxs = (int)(step - drs * (uint32_t)a.wo), dimg = (int)(drs / (uint32_t)a.ho), dys = (int)(drs - (uint32_t)dimg * (uint32_t)a.ho);
        const uint32_t a_row = (uint32_t)tid * 64u, a_sw = (uint32_t)((tid >> 1) & 3);
        auto epilogue = [&](uint32_t j, uint32_t mp, bool livep) {   // tile j of this CTA: TMEM -> bias, ReLU6, bf16 -> global
            const uint32_t s = j & 1u;
            mbar_wait(mma_done + 8u * s, (j >> 1) & 1u);
            tc_fence_after();
            uint32_t v[32];
            tc_ld32(tmem + ((uint32_t)(warp * 32) << 16) + s * 32u, v);
            tc_ld_wait();
            if (livep) {
                // a pixel's cout bf16 values are contiguous (32 / 48 / 64 bytes): 256-bit stores where the row is 32-byte aligned
                // (cout 16, 32) -- a warp store then fills whole 32-byte sectors -- 128-bit stores otherwise (cout 24)
                __nv_bfloat16 *dst = a.y + (size_t)mp * a.cout;
                const bool wide = (a.cout & 15) == 0;
#pragma unroll
                for (int c2 = 0; c2 < 2; ++c2) {
                    if (c2 * 16 < a.cout) {
                        uint32_t o[8];
#pragma unroll
                        for (int j2 = 0; j2 < 8; ++j2) {
                            const int col = c2 * 16 + 2 * j2;
                            o[j2] = relu6_bf16x2(fadd2(make_float2(__uint_as_float(v[col]), __uint_as_float(v[col + 1])),
                                                       *reinterpret_cast<const float2 *>(sBias + col)));
                        }
                        if (wide) {
                            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + c2 * 16), "r"(o[0]), "r"(o[1]),
                                         "r"(o[2]), "r"(o[3]), "r"(o[4]), "r"(o[5]), "r"(o[6]), "r"(o[7])
                                         : "memory");
                        } else {
                            *reinterpret_cast<uint4 *>(dst + c2 * 16) = make_uint4(o[0], o[1], o[2], o[3]);
                            if (c2 * 16 + 8 < a.cout) *reinterpret_cast<uint4 *>(dst + c2 * 16 + 8) = make_uint4(o[4], o[5], o[6], o[7]);
                        }
                    }
                }
            }
        };
        int slot = 0;
        uint32_t span_phase = 0, it = 0, m_prev = 0;
        bool live_prev = false;
        for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const bool live = (long long)m < a.total_px;
            mbar_wait(span_full + 8u * slot, (span_phase >> slot) & 1u);
            span_phase ^= 1u << slot;
            const uint32_t sp = sSpan + (uint32_t)slot * (uint32_t)a.span_cap;
            const int ix0 = ox * a.stride - 1, iy0 = oy * a.stride - 1;
            const int lead = ix0 < 0 ? 1 : 0;                                  // window column 0 is padding (pad = 1)
            bool okx[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) okx[kx] = live && ix0 + kx >= 0 && ix0 + kx < a.w;
            const bool interior_px = live && ix0 >= 0 && ix0 + 2 < a.w && iy0 >= 0 && iy0 + 2 < a.h;
            int rowoff[3];
            if (a.seg) {
                const int sg = (live && ox < tid) ? 1 : 0;                     // the tile's second output row (wo >= 128: at most two)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) rowoff[ky] = piece_adj[slot * STC_PIECES + sg * 3 + ky] + (ix0 + lead) * 3;
            } else {
                const long long lo16 = span_lo[slot];                          // (published before the load was issued)
                const int rowoff0 = (int)(((long long)img_i * a.h + iy0) * row_bytes - lo16) + (ix0 + lead) * 3;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) rowoff[ky] = rowoff0 + ky * (int)row_bytes;
            }
            // ---- im2col: this thread's pixel -> row `tid` of A.  The 9 bytes of a window row (3 pixels x BGR) are contiguous: three
            // aligned 32-bit loads + funnel shifts bring them to byte 0 (one pixel later when the window starts left of the image);
            // interior pixels (all 27 taps inside the image) take a copy of the loop without the padding selects
            float f[32];
#pragma unroll
            for (int k = 27; k < 32; ++k) f[k] = 0.f;
            auto im2col = [&](const bool interior) {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int iy = iy0 + ky;
                    const bool row_ok = live && iy >= 0 && iy < a.h;
                    const uint32_t b0 = row_ok ? (uint32_t)rowoff[ky] : 0u;
                    const uint32_t a0 = sp + (b0 & ~3u), sh = (b0 & 3u) * 8u;
                    const uint32_t w0 = lds_u32s(a0), w1 = lds_u32s(a0 + 4), w2 = lds_u32s(a0 + 8);
                    uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = w2 >> sh;
                    const uint32_t ls = (uint32_t)lead * 24u;                  // shift the 9 bytes up by one pixel
                    v2 = __funnelshift_l(v1, v2, ls);
                    v1 = __funnelshift_l(v0, v1, ls);
                    v0 = v0 << ls;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const bool ok = row_ok && okx[kx];
#pragma unroll
                        for (int ci = 0; ci < 3; ++ci) {                       // BGR bytes -> RGB taps; one PRMT builds 2^23 + x
                            const int i = kx * 3 + (2 - ci);
                            const uint32_t src = i < 4 ? v0 : i < 8 ? v1 : v2;
                            const uint32_t bits = __byte_perm(src, 0x4B000000u, 0x7540u + (uint32_t)(i & 3));
                            const float val = __uint_as_float(bits) - 8388736.0f;
                            f[(ky * 3 + kx) * 3 + ci] = (interior || ok) ? val : -0.5f;
                        }
                    }
                }
            };
            if (interior_px) im2col(true); else im2col(false);
            const uint32_t s = it & 1u;
            // (this thread last wrote A[s] two tiles ago and has since seen that tile's MMAs retire, in its epilogue)
            const uint32_t arow = sA + s * (uint32_t)STC_A_BYTES + a_row;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(f[c * 8 + 2 * j], f[c * 8 + 2 * j + 1]);
                    pk[j] = *reinterpret_cast<const uint32_t *>(&h2);
                }
                st_shared_v4(arow + (((uint32_t)c ^ a_sw) << 4), pk[0], pk[1], pk[2], pk[3]);
            }
            fence_async_smem();                                                // generic-proxy writes of A -> visible to the MMA
            tc_fence_before();                                                 // (and the TMEM reads of the last epilogue are ordered)
            mbar_arrive(a_full + 8u * s);
            if (it) epilogue(it - 1u, m_prev, live_prev);
            m_prev = m;
            live_prev = live;
            // next tile of this thread: + step pixels
            m += step;
            ox += dxs;
            const int cx = ox >= a.wo ? 1 : 0;
            ox -= cx ? a.wo : 0;
            oy += dys + cx;
            const int cy = oy >= a.ho ? 1 : 0;
            oy -= cy ? a.ho : 0;
            img_i += dimg + cy;
            if (++slot == a.nbuf) slot = 0;
        }
        if (it) epilogue(it - 1u, m_prev, live_prev);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
    }
}

static bool stem_tc_usable(const void *x, const void *y, int cout, int h, int wd) {
    return ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && cout % 8 == 0 && cout <= 32 && h >= 1 && wd >= 1 &&
           getenv("PN_STEM_SIMT") == nullptr;
}

static int launch_stem_tc(const uint8_t *img, const float *w, const float *b, void *y, int n, int h, int wd, int ho, int wo,
                          int stride, int cout, cudaStream_t st) {
    StemTcArgs a;
    a.img = img; a.w27 = w; a.bias = b; a.y = (__nv_bfloat16 *)y;
    a.n = n; a.h = h; a.w = wd; a.ho = ho; a.wo = wo; a.stride = stride; a.cout = cout;
    a.total_px = (long long)n * ho * wo;
    if (a.total_px >= (1ll << 31) - 256) return 1;                // 32-bit pixel indexing inside the kernel
    a.total_bytes = (long long)n * h * wd * 3;
    a.total_lo16 = a.total_bytes & ~15ll;
    // input rows a 128-pixel tile can touch: it covers at most 128/wo + 2 output rows (across an image boundary too), i.e.
    // (rows_out - 1) * stride + 3 contiguous input rows (tight: checked by brute force over every tile of 28 geometries)
    const long long rows_out = 128 / wo + 2;
    long long span = ((rows_out - 1) * stride + 3) * (long long)wd * 3 + 32;
    span = (span + 127) & ~127ll;
    // Whole rows or segments (see the kernel).  Measured with the pipelined kernel (us, whole rows with 2 slots / segments with 3):
    // 513 x 513 x 64 -> 32: 79.9 / 83.5; 257 x 257 x 512 -> 24: 158.9 / 170.4; 1281 x 721 x 32 -> 16: 137.1 / 126.9 -- one bulk copy
    // per tile beats six until the whole-row spans (19 KB there) leave only three CTAs per SM.  PN_STEM_SEGMENTS=1 forces
    // segments (tests), =0 forbids them; PN_STEM_NBUF / PN_STEM_CTAS override the ring depth / CTAs per SM.
    const long long piece = (((127ll * stride + 3) * 3 + 30) + 15) & ~15ll;
    auto smem_of = [](int nbuf, long long sp) { return 2ll * STC_A_BYTES + STC_W_BYTES + nbuf * sp + STP_CTRL_BYTES + 1024; };
    const char *e_seg = getenv("PN_STEM_SEGMENTS");
    const bool seg_ok = wo >= 128 && !(e_seg && e_seg[0] == '0');
    a.seg = (wo >= 128 && e_seg && e_seg[0] == '1') || (seg_ok && smem_of(2, span) > 44 * 1024) ? 1 : 0;
    a.piece_cap = (int)piece;
    if (a.seg) span = (STC_PIECES * piece + 127) & ~127ll;
    int nbuf = a.seg ? 3 : 2;
    if (const char *e = getenv("PN_STEM_NBUF")) nbuf = atoi(e);
    if (nbuf < 2) nbuf = 2;
    if (nbuf > STP_MAX_NBUF) nbuf = STP_MAX_NBUF;
    while (nbuf > 2 && smem_of(nbuf, span) > 200 * 1024) --nbuf;
    const long long smem = smem_of(nbuf, span);
    if (smem > 200 * 1024) return 1;                              // does not fit: caller falls back to the SIMT kernel
    a.span_cap = (int)span;
    a.nbuf = nbuf;
    static DeviceOnce once;                                       // largest dynamic shared memory size configured, per device
    const int dev = current_device();
    if (once.get(dev) < (int)smem) {
        PN_CHECK_CUDA(cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        once.set(dev, (int)smem);
    }
    const long long tiles = (a.total_px + 127) / 128;
    int per_sm = (int)((220 * 1024) / smem);
    if (per_sm > 8) per_sm = 8;                                   // 8 x 64 accumulator columns = all of TMEM
    if (per_sm < 1) per_sm = 1;
    if (const char *e = getenv("PN_STEM_CTAS")) per_sm = atoi(e) > 0 && atoi(e) < per_sm ? atoi(e) : per_sm;
    const long long max_ctas = (long long)num_sms() * per_sm;
    const int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
    PN_CHECK_CUDA(launch_pdl(stem_tc_kernel, dim3(grid), dim3(STP_THREADS), (size_t)smem, st, a));
    return PN_OK;
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
    if (x_is_u8 && out_dtype == PN_BF16 && stem_tc_usable(x, y, cout, h, wd)) {   // production path: tcgen05 im2col GEMM
        const int rc = launch_stem_tc((const uint8_t *)x, w, b, y, n, h, wd, ho, wo, stride, cout, st);
        if (rc <= 0) return rc;                                                  // rc > 0: span too large, use the SIMT kernel
    }
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
