// C1 + D1..D4 -- part candidates and the greedy multi-pose decoder, on the device.
// Replaces posenet/decode_multi.py:27-34 (build_part_with_score_torch), :61-148 (decode_multiple_poses),
// :8-24 (NMS / instance score) and posenet/decode.py:9-63,131-182 (traverse_to_targ_keypoint,
// decode_pose) of the reference, including the six .cpu().numpy() round trips they imply.
//
// Bit-exactness: the reference decodes in numpy float64 from fp32 maps.  Everything below uses the
// same IEEE operations in the same order (explicit _rn intrinsics so nothing is contracted into an
// FMA), round-half-even (rint), numpy's pairwise summation order for the instance score, and the
// "score == 0.0 means not yet decoded" flag.  The one defined difference: the reference's candidate
// sort is unstable (torch.argsort); here the order is (score desc, flat (part,y,x) index asc).
#include "common.cuh"

namespace pn {

__constant__ int c_parent[PN_NUM_EDGES] = {0, 1, 0, 2, 0, 5, 7, 5, 11, 13, 0, 6, 8, 6, 12, 14};
__constant__ int c_child[PN_NUM_EDGES] = {1, 3, 2, 4, 5, 7, 9, 11, 13, 15, 6, 8, 10, 12, 14, 16};

__device__ __forceinline__ float map_at(const pn_map &m, int img, int ch, int y, int x) {
    return __ldg(m.ptr + img * m.s_img + ch * m.s_ch + y * m.s_y + x * m.s_x);
}

// ---------------------------------------------------------------------------------- C1: candidates
// Sort key: high word = ~orderable(score) so that ascending key == descending score; low word = flat
// index, which both breaks ties the way the oracle defines and identifies the cell.
__device__ __forceinline__ uint64_t make_key(float score, uint32_t flat) {
    uint32_t u = (score == 0.0f) ? 0u : __float_as_uint(score);   // -0.0 and +0.0 tie, as they do for argsort
    u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;       // monotone float -> uint
    return ((uint64_t)(~u) << 32) | flat;
}

__global__ void __launch_bounds__(256) candidates_kernel(pn_map heat, int h, int w, float thr,
                                                          uint64_t *__restrict__ keys, int capacity,
                                                          int *__restrict__ counts) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int part = blockIdx.y;
    const int img = blockIdx.z;
    bool keep = false;
    float v = 0.f;
    if (p < h * w) {
        const int y = p / w, x = p - y * w;
        v = map_at(heat, img, part, y, x);
        if (v >= thr) {                                // decode_multi.py:30 (threshold already fp32)
            keep = true;
            // 3x3 window, -inf padding (decode_multi.py:29): a cell survives iff no in-bounds neighbour
            // is larger.  !(nb <= v) also rejects NaN neighbours like max_pool2d's NaN propagation.
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int yy = y + dy;
                if (yy < 0 || yy >= h) continue;
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int xx = x + dx;
                    if (xx < 0 || xx >= w || (dx == 0 && dy == 0)) continue;
                    if (!(map_at(heat, img, part, yy, xx) <= v)) keep = false;
                }
            }
        }
    }
    // warp-aggregated compaction: one atomic per warp, ballot/popc prefix inside it
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, keep);
    if (mask) {
        const int lane = threadIdx.x & 31;
        int base = 0;
        if (lane == __ffs(mask) - 1) base = atomicAdd(&counts[img], __popc(mask));
        base = __shfl_sync(0xFFFFFFFFu, base, __ffs(mask) - 1);
        if (keep) {
            const int pos = base + __popc(mask & ((1u << lane) - 1));
            if (pos < capacity) keys[(size_t)img * capacity + pos] = make_key(v, (uint32_t)(part * h * w + p));
        }
    }
}

// ---------------------------------------------------------------------------------- D: greedy decode
// The reference walks the sorted candidates one by one; per accepted pose it pays 16 dependent displacement hops
// (2 dependent gathers each), which is pure latency.  decode_pose (decode.py:131-182) does not depend on the poses
// accepted so far -- only the root NMS test and the instance score do -- so a block decodes in ROUNDS:
//   (a) screen: all threads test the next candidates against the accepted poses and compact the survivors, in
//       order, into up to DEC_SLOTS slots (candidates suppressed now stay suppressed: poses are only ever added);
//   (b) speculate: 17 threads per slot, one per part.  decode_pose reaches every part along the unique tree path from the
//       root part (child -> parent hops of the backward pass first, then parent -> child hops of the forward pass), and a hop
//       is a pure function of its source coordinates, so the thread of part k walks the path root -> k on its own: at most 8
//       dependent hops instead of the 16 a single thread pays, prefixes shared with other parts are recomputed (same
//       addresses, same warp-level requests), and a path stops where the reference's `score[source] > 0.0` gate stops it;
//   (c) commit: warp 0 replays the reference's greedy loop over the slots in order -- root NMS against every accepted
//       pose (including those of this round), instance score, acceptance -- from shared memory only.
// The sequence of accepted poses and every float64 operation is the reference's; only the latency is shared.
#ifndef PN_DEC_THREADS
#define PN_DEC_THREADS 512
#endif
constexpr int DEC_THREADS = PN_DEC_THREADS;
constexpr int DEC_KBUF = 4096;          // capacity of the sorted chunk in shared memory (32 KB)
constexpr int DEC_CHUNK = 1024;         // keys taken per chunk when an image has more: the greedy loop rarely needs more
constexpr int DEC_BATCH = 128;          // slot capacity of the speculative pose records
constexpr int DEC_SLOTS = DEC_THREADS / PN_NUM_PARTS < DEC_BATCH ? DEC_THREADS / PN_NUM_PARTS : DEC_BATCH;   // slots per round
constexpr int DEC_ACC = 64;             // accepted poses whose coordinates are cached in shared memory

struct DecodeArgs {
    pn_map heat, off, fwd, bwd;
    int h, w;
    const uint64_t *keys;
    int capacity;
    const int *counts;
    pn_decode_params prm;
    double *pose_scores, *kp_scores, *kp_coords, *kp_offsets;
    int *pose_counts;
};

struct DecodeShared {
    uint64_t keys[DEC_KBUF];
    // speculative pose records, [part][slot]: a warp's accesses to one part are conflict-free
    double kc[PN_NUM_PARTS][2][DEC_BATCH];
    float ks[PN_NUM_PARTS][DEC_BATCH];
    float ko[PN_NUM_PARTS][2][DEC_BATCH];
    uint32_t cand[DEC_BATCH];           // flat (part, y, x) index of each slot's root
    double full[DEC_BATCH];             // instance score of a slot when no part is masked: np.sum(all 17 scores) / 17
    int up[PN_NUM_PARTS], up_edge[PN_NUM_PARTS];
    unsigned anc[PN_NUM_PARTS];         // bit a set: part a is part k's ancestor or k itself
    double acc[DEC_ACC][PN_NUM_PARTS][2];   // keypoint coordinates of the accepted poses (the first DEC_ACC)
    double vals[PN_NUM_PARTS];
    unsigned hist[256];
    int warp_cnt[DEC_THREADS / 32];
    uint64_t pivot, kmax;
    int cnt, npose, done, next_ci;
};

// decode.py:15-16 / :50-51 -- np.clip(np.round(p / stride), 0, hi).astype(int32): f64 divide, half-even.
// inv > 0: the stride is a power of two and p * (1 / stride) is the same double as p / stride (an exact scaling), which
// saves the ~30-instruction division sequence four times per hop; otherwise (inv == 0) the division is done as written.
__device__ __forceinline__ int to_cell(double p, double stride, double inv, int hi) {
    double r = rint(inv > 0.0 ? __dmul_rn(p, inv) : __ddiv_rn(p, stride));
    r = fmin(fmax(r, 0.0), (double)hi);
    return (int)r;
}

// numpy's pairwise float64 sum for n <= 128 contiguous values (verified against np.sum for n = 0..17).
__device__ double np_sum(const double *a, int n) {
    if (n < 8) {
        double r = 0.0;
        for (int i = 0; i < n; ++i) r = __dadd_rn(r, a[i]);
        return r;
    }
    double r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

// decode.py:9-63: one displacement hop along edge e from the source at (cy, cx) to part `tgt`; on return (cy, cx) are the
// target's coordinates, sc its score, (oy, ox) its offset vector and (dy, dx) the displacement vector that was followed.
__device__ __forceinline__ void hop_full(const DecodeArgs &a, int img, const pn_map &disp, int e, int tgt, double &cy, double &cx,
                                         float &sc, float &oy, float &ox, float &dy, float &dx) {
    const int os = a.prm.output_stride;
    const double stride = (double)os, inv = (os & (os - 1)) == 0 ? 1.0 / stride : 0.0;
    const int iy = to_cell(cy, stride, inv, a.h - 1), ix = to_cell(cx, stride, inv, a.w - 1);
    dy = map_at(disp, img, e, iy, ix);                                                        // decode.py:39
    dx = map_at(disp, img, PN_NUM_EDGES + e, iy, ix);
    const double py = __dadd_rn(cy, (double)dy);                                              // decode.py:40
    const double px = __dadd_rn(cx, (double)dx);
    const int ty = to_cell(py, stride, inv, a.h - 1), tx = to_cell(px, stride, inv, a.w - 1);
    sc = map_at(a.heat, img, tgt, ty, tx);                                                    // decode.py:53
    oy = map_at(a.off, img, tgt, ty, tx);
    ox = map_at(a.off, img, PN_NUM_PARTS + tgt, ty, tx);
    cy = __dadd_rn((double)(ty * os), (double)oy);                                            // decode.py:55-56
    cx = __dadd_rn((double)(tx * os), (double)ox);
}
__device__ __forceinline__ void hop(const DecodeArgs &a, int img, const pn_map &disp, int e, int tgt, double &cy, double &cx,
                                    float &sc, float &oy, float &ox) {
    float dy, dx;
    hop_full(a, img, disp, e, tgt, cy, cx, sc, oy, ox, dy, dx);
}

__device__ __forceinline__ double sqdist(double ay, double ax, double by, double bx) {
    const double dy = __dsub_rn(ay, by), dx = __dsub_rn(ax, bx);
    return __dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx));
}

// keypoint coordinate c of part `part` of accepted pose p (shared-memory cache, or the output array beyond it)
__device__ __forceinline__ double acc_coord(const DecodeShared &S, const double *out_kc, int p, int part, int c) {
    return p < DEC_ACC ? S.acc[p][part][c] : __ldcg(out_kc + ((size_t)p * PN_NUM_PARTS + part) * 2 + c);
}

__global__ void __launch_bounds__(DEC_THREADS, DEC_THREADS <= 512 ? 2 : 1) decode_kernel(DecodeArgs a) {
    extern __shared__ __align__(16) unsigned char dec_smem_raw[];
    DecodeShared &S = *reinterpret_cast<DecodeShared *>(dec_smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int img = blockIdx.x;
    const int P = a.prm.max_pose_detections;
    const int n = min(a.counts[img], a.capacity);
    const uint64_t *keys = a.keys + (size_t)img * a.capacity;
    const int hw = a.h * a.w;
    const int os = a.prm.output_stride;
    const double r2 = a.prm.squared_nms_radius;
    double *out_ps = a.pose_scores + (size_t)img * P;
    double *out_ks = a.kp_scores + (size_t)img * P * PN_NUM_PARTS;
    double *out_kc = a.kp_coords + (size_t)img * P * PN_NUM_PARTS * 2;
    double *out_ko = a.kp_offsets + (size_t)img * P * PN_NUM_PARTS * 2;

#ifdef PN_DEC_TRACE
    long long tr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tr_t = clock64();
    int tr_rounds = 0, tr_slots = 0;
#define DEC_STAMP(i) do { const long long now_ = clock64(); tr[i] += now_ - tr_t; tr_t = now_; } while (0)
#else
#define DEC_STAMP(i) do { } while (0)
#endif
    // the skeleton as a tree rooted at the nose (part 0): parent part, the edge whose child a part is, ancestors-or-self
    if (tid < PN_NUM_PARTS) {
        int up = 0, up_edge = -1;
        for (int e = 0; e < PN_NUM_EDGES; ++e)
            if (c_child[e] == tid) { up = c_parent[e]; up_edge = e; }
        S.up[tid] = up;
        S.up_edge[tid] = up_edge;
    }
    __syncthreads();
    if (tid < PN_NUM_PARTS) {
        unsigned m = 1u << tid;
        for (int k = tid; k != 0; k = S.up[k]) m |= 1u << S.up[k];
        S.anc[tid] = m;
    }
    __syncthreads();
    uint64_t lo = 0;            // every key consumed so far is <= lo (real keys are never 0)
    int remaining = n;
    int npose = 0;

    while (remaining > 0 && npose < P) {
        // ---- 1. pick a pivot so that (lo, pivot] holds between 1 and DEC_CHUNK of the best keys
        uint64_t pivot = ~0ull;
        if (remaining > DEC_CHUNK) {
            // MSB-first radix descent, 8 bits per level, among keys > lo.  Keys are unique (low word is
            // the cell index), so the descent always terminates with a non-empty prefix set.  The descent starts below the
            // bits all remaining keys share (one min / max pass): scores of one image often agree in their top 2-3 bytes,
            // and every such byte would cost a histogram pass that ends in a single bucket.
            if (tid == 0) { S.pivot = ~0ull; S.kmax = 0ull; }
            __syncthreads();
            {
                uint64_t mn = ~0ull, mx = 0ull;
                for (int i0 = tid; i0 < n; i0 += 4 * DEC_THREADS) {
                    uint64_t k4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) k4[j] = i0 + j * DEC_THREADS < n ? __ldg(keys + i0 + j * DEC_THREADS) : 0ull;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (k4[j] > lo) { mn = k4[j] < mn ? k4[j] : mn; mx = k4[j] > mx ? k4[j] : mx; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const uint64_t a_ = __shfl_xor_sync(0xFFFFFFFFu, mn, o), b_ = __shfl_xor_sync(0xFFFFFFFFu, mx, o);
                    mn = a_ < mn ? a_ : mn;
                    mx = b_ > mx ? b_ : mx;
                }
                if (lane == 0) {
                    atomicMin(reinterpret_cast<unsigned long long *>(&S.pivot), (unsigned long long)mn);
                    atomicMax(reinterpret_cast<unsigned long long *>(&S.kmax), (unsigned long long)mx);
                }
            }
            __syncthreads();
            int bits = __clzll((long long)(S.pivot ^ S.kmax));           // > DEC_CHUNK distinct keys remain, so min != max
            uint64_t prefix = bits ? S.pivot >> (64 - bits) : 0ull;
            __syncthreads();
            for (;;) {
                for (int i = tid; i < 256; i += DEC_THREADS) S.hist[i] = 0;
                __syncthreads();
                const int width = 64 - bits < 8 ? 64 - bits : 8;
                const int shift = 64 - bits - width;
                // four keys in flight per thread: the pass is bound by the latency of the key loads, not by the atomics
                for (int i0 = tid; i0 < n; i0 += 4 * DEC_THREADS) {
                    uint64_t k4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) k4[j] = i0 + j * DEC_THREADS < n ? __ldg(keys + i0 + j * DEC_THREADS) : 0ull;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint64_t k = k4[j];                       // 0 is never a key and never > lo
                        if (k > lo && (bits == 0 || (k >> (64 - bits)) == prefix)) atomicAdd(&S.hist[(k >> shift) & ((1u << width) - 1u)], 1u);
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    unsigned cum = 0;
                    int sel = -1, first = -1;
                    for (int d = 0; d < 256; ++d) {
                        const unsigned c = S.hist[d];
                        if (c && first < 0) first = d;
                        if (cum + c > (unsigned)DEC_CHUNK) break;
                        cum += c;
                        if (c) sel = d;
                    }
                    if (sel >= 0) {
                        const uint64_t low_ones = shift ? ((1ull << shift) - 1) : 0ull;
                        S.pivot = (((prefix << width) | (uint64_t)sel) << shift) | low_ones;
                        S.done = 1;
                    } else {
                        S.pivot = (prefix << width) | (uint64_t)first;   // descend into the first non-empty bucket
                        S.done = 0;
                    }
                }
                __syncthreads();
                const int done = S.done;
                const uint64_t pv = S.pivot;
                __syncthreads();
                if (done) { pivot = pv; break; }
                prefix = pv;
                bits += width;
            }
        }
        DEC_STAMP(0);
        // ---- 2. gather (lo, pivot] into shared memory and sort ascending (bitonic)
        if (tid == 0) S.cnt = 0;
        __syncthreads();
        for (int i0 = tid; i0 < n; i0 += 4 * DEC_THREADS) {
            uint64_t k4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) k4[j] = i0 + j * DEC_THREADS < n ? __ldg(keys + i0 + j * DEC_THREADS) : 0ull;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k4[j] > lo && k4[j] <= pivot) S.keys[atomicAdd(&S.cnt, 1)] = k4[j];
        }
        __syncthreads();
        const int cnt = S.cnt;
        DEC_STAMP(1);
        if (cnt <= DEC_THREADS) {
            // Small chunk (the usual case: a few hundred candidates, one per thread): rank sort.  Keys are unique, so the
            // number of smaller keys IS the sorted position; every thread reads the same key at a time (a shared-memory
            // broadcast), and the whole sort is two barriers instead of ~50 bitonic passes.
            // The list is padded to a multiple of 8 with all-ones keys (never smaller than a key), so the count runs over
            // 128-bit shared-memory loads, eight independent comparisons at a time.
            uint64_t *scratch = S.keys + DEC_KBUF / 2;
            const int cnt8 = (cnt + 7) & ~7;
            if (tid < cnt8 - cnt) S.keys[cnt + tid] = ~0ull;
            __syncthreads();
            for (int i = tid; i < cnt; i += DEC_THREADS) {
                const uint64_t k = S.keys[i];
                int rank = 0;
                for (int j = 0; j < cnt8; j += 8) {
                    ulonglong2 q[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) q[u] = *reinterpret_cast<const ulonglong2 *>(&S.keys[j + 2 * u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) rank += (q[u].x < k ? 1 : 0) + (q[u].y < k ? 1 : 0);
                }
                scratch[rank] = k;
            }
            __syncthreads();
            for (int i = tid; i < cnt; i += DEC_THREADS) S.keys[i] = scratch[i];
            __syncthreads();
        } else {
            int m = 2;
            while (m < cnt) m <<= 1;
            for (int i = cnt + tid; i < m; i += DEC_THREADS) S.keys[i] = ~0ull;
            __syncthreads();
            for (int k = 2; k <= m; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = tid; i < m; i += DEC_THREADS) {
                        const int ixj = i ^ j;
                        if (ixj > i) {
                            const uint64_t x = S.keys[i], y = S.keys[ixj];
                            const bool up = ((i & k) == 0);
                            if ((x > y) == up) { S.keys[i] = y; S.keys[ixj] = x; }
                        }
                    }
                    __syncthreads();
                }
            }
        }
        DEC_STAMP(2);
        // ---- 3. greedy pass over this chunk, in rounds of up to DEC_SLOTS speculatively decoded candidates
        int ci = 0;
        while (ci < cnt && npose < P) {
            // (a) screen candidates ci.. against the npose accepted poses; survivors -> slots, in order
            constexpr int slots = DEC_SLOTS;
            int nb = 0, next = ci;
            while (next < cnt && nb < slots) {
                const int base = next, i = base + tid;
                if (tid == 0) S.next_ci = 0x7fffffff;
                bool keep = false;
                uint32_t flat = 0;
                if (i < cnt) {
                    flat = (uint32_t)S.keys[i];
                    const int part = flat / hw;
                    const int rem = flat - part * hw;
                    const int y = rem / a.w, x = rem - y * a.w;
                    // decode_multi.py:106-109: int64 cell * stride + fp32 offset -> float64
                    const double ry = __dadd_rn((double)(y * os), (double)map_at(a.off, img, part, y, x));
                    const double rx = __dadd_rn((double)(x * os), (double)map_at(a.off, img, PN_NUM_PARTS + part, y, x));
                    // decode_multi.py:8-11,111-113: same part of any accepted pose within the radius (<=)
                    keep = true;
                    for (int p = 0; p < npose; ++p)
                        if (sqdist(acc_coord(S, out_kc, p, part, 0), acc_coord(S, out_kc, p, part, 1), ry, rx) <= r2) { keep = false; break; }
                }
                const unsigned bal = __ballot_sync(0xFFFFFFFFu, keep);
                if (lane == 0) S.warp_cnt[warp] = __popc(bal);
                __syncthreads();
                int pos = nb, total = nb;
#pragma unroll
                for (int wi = 0; wi < DEC_THREADS / 32; ++wi) {
                    const int c = S.warp_cnt[wi];
                    if (wi < warp) pos += c;
                    total += c;
                }
                pos += __popc(bal & ((1u << lane) - 1));
                if (keep) {
                    if (pos < slots) S.cand[pos] = flat;
                    else atomicMin(&S.next_ci, i);                  // first survivor that did not fit: resume there
                }
                __syncthreads();
                next = total > slots ? S.next_ci : min(base + DEC_THREADS, cnt);
                nb = min(total, slots);
                __syncthreads();
            }
            DEC_STAMP(3);
            // (b) one thread per (slot, part): decode.py:131-182, the part's own path from the root
            if (tid < nb * PN_NUM_PARTS) {
                const int s = tid / PN_NUM_PARTS, k = tid - s * PN_NUM_PARTS;
                const uint32_t flat = S.cand[s];
                const int root = flat / hw;
                const int rem = flat - root * hw;
                const int y = rem / a.w, x = rem - y * a.w;
                float sc = map_at(a.heat, img, root, y, x), oy = 0.f, ox = 0.f;       // the root's offset row stays 0 (decode.py:150)
                double cy = __dadd_rn((double)(y * os), (double)map_at(a.off, img, root, y, x));
                double cx = __dadd_rn((double)(x * os), (double)map_at(a.off, img, PN_NUM_PARTS + root, y, x));
                const unsigned anc_k = S.anc[k];
                int cur = root;
                bool reached = true;
                while (cur != k) {
                    // decode.py:152-153,169-170: a hop needs `score[source] > 0.0`; its target is still 0.0 (every part has
                    // exactly one path from the root), so a non-positive score on the path leaves the rest of it undecoded
                    if (!(sc > 0.f)) { reached = false; break; }
                    const bool climb = ((anc_k >> cur) & 1u) == 0;                      // cur is not above k yet: child -> parent
                    int nxt;
                    if (climb) nxt = S.up[cur];
                    else { nxt = k; while (S.up[nxt] != cur) nxt = S.up[nxt]; }         // the child of cur on the way down to k
                    const int e = climb ? S.up_edge[cur] : S.up_edge[nxt];
                    hop(a, img, climb ? a.bwd : a.fwd, e, nxt, cy, cx, sc, oy, ox);
                    cur = nxt;
                }
                if (!reached) { sc = 0.f; oy = ox = 0.f; cy = cx = 0.0; }
                S.ks[k][s] = sc;
                S.kc[k][0][s] = cy; S.kc[k][1][s] = cx;
                S.ko[k][0][s] = oy; S.ko[k][1][s] = ox;
            }
            __syncthreads();
            // the instance score of a slot none of whose parts is masked (decode_multi.py:14-24 with an all-true mask)
            if (tid < nb) {
                double v[PN_NUM_PARTS];
#pragma unroll
                for (int k = 0; k < PN_NUM_PARTS; ++k) v[k] = (double)S.ks[k][tid];
                S.full[tid] = __ddiv_rn(np_sum(v, PN_NUM_PARTS), (double)PN_NUM_PARTS);
            }
            __syncthreads();
            DEC_STAMP(4);
#ifdef PN_DEC_TRACE
            ++tr_rounds; tr_slots += nb;
#endif
            // (c) warp 0: the reference's greedy loop over the slots, in candidate order.  The screen left no slot within the
            // radius of a pose accepted in an EARLIER round, so the root NMS test (decode_multi.py:8-11,111-113) only has to look
            // at the poses accepted in THIS round: lane l owns slot l, and after every acceptance each lane tests its own pending
            // slot against the new pose -- a suppressed slot costs nothing instead of a pass over all accepted poses.
            static_assert(DEC_SLOTS <= 32, "one lane per slot in the commit loop");
            if (warp == 0) {
                int my_part = 0;
                double my_ry = 0.0, my_rx = 0.0;
                if (lane < nb) {
                    my_part = S.cand[lane] / hw;
                    my_ry = S.kc[my_part][0][lane]; my_rx = S.kc[my_part][1][lane];
                }
                unsigned pending = nb >= 32 ? 0xFFFFFFFFu : (1u << nb) - 1u;     // slots neither visited nor suppressed yet
                while (pending && npose < P) {
                    const int s = __ffs(pending) - 1;
                    pending &= pending - 1;
                    // the mask of decode_multi.py:14-24: parts strictly farther than the radius from that part of EVERY accepted pose
                    bool far = lane < PN_NUM_PARTS;
                    double ky = 0.0, kx = 0.0;
                    if (lane < PN_NUM_PARTS) {
                        ky = S.kc[lane][0][s]; kx = S.kc[lane][1][s];
                        for (int p = 0; p < npose; ++p)
                            if (!(sqdist(acc_coord(S, out_kc, p, lane, 0), acc_coord(S, out_kc, p, lane, 1), ky, kx) > r2)) far = false;
                    }
                    const unsigned fmask = __ballot_sync(0xFFFFFFFFu, far);
                    double score;
                    if (fmask == (1u << PN_NUM_PARTS) - 1u) {
                        score = S.full[s];                       // nothing masked: computed with the slot, in parallel
                    } else {
                        // sum the kept scores in numpy's order; always divide by 17
                        if (far) S.vals[__popc(fmask & ((1u << lane) - 1))] = (double)S.ks[lane][s];
                        __syncwarp();
                        score = 0.0;
                        if (lane == 0) score = __ddiv_rn(np_sum(S.vals, __popc(fmask)), (double)PN_NUM_PARTS);
                        score = __shfl_sync(0xFFFFFFFFu, score, 0);
                    }
                    if (a.prm.min_pose_score == 0.0 || score >= a.prm.min_pose_score) {    // decode_multi.py:128
                        if (lane == 0) out_ps[npose] = score;
                        if (lane < PN_NUM_PARTS) {
                            const size_t o = (size_t)npose * PN_NUM_PARTS + lane;
                            out_ks[o] = (double)S.ks[lane][s];
                            out_kc[o * 2] = ky;
                            out_kc[o * 2 + 1] = kx;
                            out_ko[o * 2] = (double)S.ko[lane][0][s];
                            out_ko[o * 2 + 1] = (double)S.ko[lane][1][s];
                            if (npose < DEC_ACC) { S.acc[npose][lane][0] = ky; S.acc[npose][lane][1] = kx; }
                        }
                        ++npose;
                        if (npose > DEC_ACC) __threadfence_block();   // poses beyond the shared-memory cache are re-read from global
                        // pending slots whose root lies within the radius (<=) of the same part of the new pose are suppressed
                        bool hit = false;
                        if ((pending >> lane) & 1u) hit = sqdist(S.kc[my_part][0][s], S.kc[my_part][1][s], my_ry, my_rx) <= r2;
                        pending &= ~__ballot_sync(0xFFFFFFFFu, hit);
                    }
                    __syncwarp();
                }
                if (lane == 0) S.npose = npose;
            }
            __syncthreads();
            DEC_STAMP(5);
            npose = S.npose;
            ci = next;
        }
        lo = pivot;
        remaining -= cnt;
        __syncthreads();
    }
    if (tid == 0) a.pose_counts[img] = npose;
    // decode_multi.py:94-100: the outputs are np.zeros; rows past the last accepted pose are zero-padded here, so the
    // caller hands over uninitialised buffers and the step needs no separate fill launch
    for (int i = npose + tid; i < P; i += DEC_THREADS) out_ps[i] = 0.0;
    for (int i = npose * PN_NUM_PARTS + tid; i < P * PN_NUM_PARTS; i += DEC_THREADS) out_ks[i] = 0.0;
    for (int i = npose * PN_NUM_PARTS * 2 + tid; i < P * PN_NUM_PARTS * 2; i += DEC_THREADS) { out_kc[i] = 0.0; out_ko[i] = 0.0; }
#ifdef PN_DEC_TRACE
    if (tid == 0 && (img == 0 || img == 7))
        printf("decode img %d: n %d poses %d rounds %d slots %d | cycles: pivot %lld gather %lld sort %lld screen %lld spec %lld commit %lld\n", img, n,
               npose, tr_rounds, tr_slots, tr[0], tr[1], tr[2], tr[3], tr[4], tr[5]);
#endif
}

// ---------------------------------------------------------------------------------- D2 / D3 as stand-alone calls
// posenet/decode.py:9-63 -- traverse_to_targ_keypoint: one hop, one thread.  out[7] = score, image_coord (y, x),
// displacement_vector (y, x), offset (y, x); the fp32 values are stored as doubles (exact).
__global__ void traverse_kernel(DecodeArgs a, pn_map disp, int edge, int target, double sy, double sx, double *__restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float sc, oy, ox, dy, dx;
    double cy = sy, cx = sx;
    hop_full(a, 0, disp, edge, target, cy, cx, sc, oy, ox, dy, dx);
    out[0] = (double)sc; out[1] = cy; out[2] = cx; out[3] = (double)dy; out[4] = (double)dx; out[5] = (double)oy; out[6] = (double)ox;
}

// posenet/decode.py:131-182 -- decode_pose for ONE root: the thread of part k walks the tree path root -> k exactly like a slot of
// decode_kernel's speculation phase (climbing hops use displacements_bwd, descending hops displacements_fwd; a hop needs
// `score[source] > 0.0`).  out = keypoint_scores[17] | keypoint_coords[17][2] | offsets[17][2], float64.
__global__ void decode_pose_kernel(DecodeArgs a, double root_score, int root, double ry, double rx, double *__restrict__ out) {
    const int k = threadIdx.x;
    if (k >= PN_NUM_PARTS) return;
    int up[PN_NUM_PARTS], up_edge[PN_NUM_PARTS];
#pragma unroll
    for (int p = 0; p < PN_NUM_PARTS; ++p) { up[p] = 0; up_edge[p] = -1; }
#pragma unroll
    for (int e = 0; e < PN_NUM_EDGES; ++e) { up[c_child[e]] = c_parent[e]; up_edge[c_child[e]] = e; }
    unsigned anc_k = 1u << k;
    for (int p = k; p != 0; p = up[p]) anc_k |= 1u << up[p];
    double ks = 0.0, cy = 0.0, cx = 0.0;
    float oy = 0.f, ox = 0.f;
    if (k == root) {                                   // decode.py:142-143: the root keeps the caller's score and coordinates
        ks = root_score; cy = ry; cx = rx;
    } else if (root_score > 0.0) {
        float sc = 1.f;                                // the root passed the `> 0.0` gate; every hop overwrites sc
        cy = ry; cx = rx;
        int cur = root;
        bool reached = true;
        while (cur != k) {
            if (!(sc > 0.f)) { reached = false; break; }
            const bool climb = ((anc_k >> cur) & 1u) == 0;
            int nxt;
            if (climb) nxt = up[cur];
            else { nxt = k; while (up[nxt] != cur) nxt = up[nxt]; }
            const int e = climb ? up_edge[cur] : up_edge[nxt];
            hop(a, 0, climb ? a.bwd : a.fwd, e, nxt, cy, cx, sc, oy, ox);
            cur = nxt;
        }
        if (reached) ks = (double)sc;
        else { cy = cx = 0.0; oy = ox = 0.f; }
    }
    out[k] = ks;
    out[PN_NUM_PARTS + 2 * k] = cy;
    out[PN_NUM_PARTS + 2 * k + 1] = cx;
    out[3 * PN_NUM_PARTS + 2 * k] = (double)oy;
    out[3 * PN_NUM_PARTS + 2 * k + 1] = (double)ox;
}

// ---------------------------------------------------------------------------------- N4: coordinates back to the source frame
// image_demo.py:50 -- `keypoint_coords *= output_scale` (float64 multiply by (src_h / target_h, src_w / target_w), utils.py:19)
// for a whole batch of pose records on the device, so the scaled records leave with the same single D2H copy.
__global__ void scale_coords_kernel(double *__restrict__ kc, long long pairs, int pairs_per_img, const double *__restrict__ scales,
                                    double sy, double sx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pairs) return;
    double fy = sy, fx = sx;
    if (scales) {
        const long long img = i / pairs_per_img;
        fy = scales[2 * img]; fx = scales[2 * img + 1];
    }
    double2 v = reinterpret_cast<double2 *>(kc)[i];
    v.x = __dmul_rn(v.x, fy);
    v.y = __dmul_rn(v.y, fx);
    reinterpret_cast<double2 *>(kc)[i] = v;
}

}  // namespace pn

// ---- C ABI ---------------------------------------------------------------------------------------
using namespace pn;

extern "C" int pn_candidates(const pn_map *heat, int n_img, int h, int wd, float score_threshold, uint64_t *keys,
                             int capacity, int *counts, pn_stream_t stream) {
    PN_CHECK_ARG(heat && heat->ptr && keys && counts, "pn_candidates: null pointer");
    PN_CHECK_ARG(n_img > 0 && h > 0 && wd > 0 && capacity > 0, "pn_candidates: bad shape");
    PN_CHECK_ARG((long long)PN_NUM_PARTS * h * wd < (1ll << 31), "pn_candidates: map too large");
    PN_CHECK_ARG(n_img <= 65535, "pn_candidates: at most 65535 images per call");
    cudaStream_t st = as_stream(stream);
    PN_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * n_img, st));
    dim3 grid(ceil_div(h * wd, 256), PN_NUM_PARTS, n_img);
    candidates_kernel<<<grid, 256, 0, st>>>(*heat, h, wd, score_threshold, keys, capacity, counts);
    PN_CHECK_LAUNCH();
    return PN_OK;
}

extern "C" int pn_decode_greedy(const pn_map *heat, const pn_map *off, const pn_map *fwd, const pn_map *bwd, int n_img,
                                int h, int wd, const uint64_t *keys, int capacity, const int *counts,
                                const pn_decode_params *params, double *pose_scores, double *kp_scores,
                                double *kp_coords, double *kp_offsets, int *pose_counts, pn_stream_t stream) {
    PN_CHECK_ARG(heat && off && fwd && bwd && heat->ptr && off->ptr && fwd->ptr && bwd->ptr, "pn_decode_greedy: null map");
    PN_CHECK_ARG(keys && counts && params && pose_scores && kp_scores && kp_coords && kp_offsets && pose_counts,
                 "pn_decode_greedy: null pointer");
    PN_CHECK_ARG(n_img > 0 && h > 0 && wd > 0 && capacity > 0, "pn_decode_greedy: bad shape");
    PN_CHECK_ARG(params->max_pose_detections >= 0 && params->output_stride > 0, "pn_decode_greedy: bad params");
    DecodeArgs a;
    a.heat = *heat; a.off = *off; a.fwd = *fwd; a.bwd = *bwd;
    a.h = h; a.w = wd; a.keys = keys; a.capacity = capacity; a.counts = counts; a.prm = *params;
    a.pose_scores = pose_scores; a.kp_scores = kp_scores; a.kp_coords = kp_coords; a.kp_offsets = kp_offsets;
    a.pose_counts = pose_counts;
    static DeviceOnce once;
    const int dev = current_device();
    if (!once.get(dev)) {
        PN_CHECK_CUDA(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecodeShared)));
        once.set(dev, 1);
    }
    decode_kernel<<<n_img, DEC_THREADS, sizeof(DecodeShared), as_stream(stream)>>>(a);
    PN_CHECK_LAUNCH();
    return PN_OK;
}

extern "C" int pn_traverse_to_targ_keypoint(int edge_id, const double *source_keypoint_host, int target_keypoint_id,
                                            const pn_map *heat, const pn_map *off, const pn_map *disp, int h, int wd,
                                            int output_stride, double *out7, pn_stream_t stream) {
    PN_CHECK_ARG(source_keypoint_host && heat && off && disp && heat->ptr && off->ptr && disp->ptr && out7,
                 "pn_traverse_to_targ_keypoint: null pointer");
    PN_CHECK_ARG(edge_id >= 0 && edge_id < PN_NUM_EDGES && target_keypoint_id >= 0 && target_keypoint_id < PN_NUM_PARTS,
                 "pn_traverse_to_targ_keypoint: edge %d / part %d out of range", edge_id, target_keypoint_id);
    PN_CHECK_ARG(h > 0 && wd > 0 && output_stride > 0, "pn_traverse_to_targ_keypoint: bad shape");
    DecodeArgs a = {};
    a.heat = *heat; a.off = *off; a.h = h; a.w = wd; a.prm.output_stride = output_stride;
    traverse_kernel<<<1, 32, 0, as_stream(stream)>>>(a, *disp, edge_id, target_keypoint_id, source_keypoint_host[0],
                                                     source_keypoint_host[1], out7);
    PN_CHECK_LAUNCH();
    return PN_OK;
}

extern "C" int pn_decode_pose(double root_score, int root_id, const double *root_image_coord_host, const pn_map *heat,
                              const pn_map *off, const pn_map *fwd, const pn_map *bwd, int h, int wd, int output_stride,
                              double *out85, pn_stream_t stream) {
    PN_CHECK_ARG(root_image_coord_host && heat && off && fwd && bwd && heat->ptr && off->ptr && fwd->ptr && bwd->ptr && out85,
                 "pn_decode_pose: null pointer");
    PN_CHECK_ARG(root_id >= 0 && root_id < PN_NUM_PARTS, "pn_decode_pose: root part %d out of range", root_id);
    PN_CHECK_ARG(h > 0 && wd > 0 && output_stride > 0, "pn_decode_pose: bad shape");
    DecodeArgs a = {};
    a.heat = *heat; a.off = *off; a.fwd = *fwd; a.bwd = *bwd; a.h = h; a.w = wd; a.prm.output_stride = output_stride;
    decode_pose_kernel<<<1, 32, 0, as_stream(stream)>>>(a, root_score, root_id, root_image_coord_host[0], root_image_coord_host[1],
                                                        out85);
    PN_CHECK_LAUNCH();
    return PN_OK;
}

extern "C" int pn_scale_keypoint_coords(double *kp_coords, int n_img, int points_per_img, const double *scales, double scale_y,
                                        double scale_x, pn_stream_t stream) {
    PN_CHECK_ARG(kp_coords && n_img > 0 && points_per_img > 0, "pn_scale_keypoint_coords: bad argument");
    PN_CHECK_ARG(((uintptr_t)kp_coords & 15) == 0, "pn_scale_keypoint_coords: kp_coords must be 16-byte aligned");
    const long long pairs = (long long)n_img * points_per_img;
    scale_coords_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, as_stream(stream)>>>(kp_coords, pairs, points_per_img, scales, scale_y,
                                                                                        scale_x);
    PN_CHECK_LAUNCH();
    return PN_OK;
}
