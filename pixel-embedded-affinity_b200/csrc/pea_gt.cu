// Ground-truth generation on device: instance labels -> affinities / validity masks / class-balance weights.
// Integer compares and exact counts: results are bit-identical to the reference's NumPy/SciPy code.
//   gen_affs_ours        scripts_cvppp/utils/affinity_ours.py:17-39
//   seg_to_aff           scripts_cvppp/data/data_affinity.py:53-102
//   weight_binary_ratio  scripts_cvppp/data/data_segmentation.py:205-228
#include <cstring>
#include "pea_common.cuh"

namespace pea {

#ifndef PEA_GT_UNROLL
#define PEA_GT_UNROLL 4
#endif
constexpr int kGtUnroll = PEA_GT_UNROLL;  // offsets whose neighbour loads are in flight together in the fused passes

struct GtTable {
    int K;
    int dz[PEA_MAX_OFFSETS];
    int dy[PEA_MAX_OFFSETS];
    int dx[PEA_MAX_OFFSETS];
};

// 2-D: aff = 1 where label == neighbour label (an out-of-image neighbour reads as 0, as scipy.ndimage.shift
// with cval=0 does), optional `ignore` of pairs touching label 0, then the padding rule on out-of-image pairs.
template <typename L>
__global__ void __launch_bounds__(256)
gt_affs2d_kernel(const L* __restrict__ labels, int B, int H, int W, GtTable tab, int ignore, int padding,
                 float* __restrict__ affs, uint8_t* __restrict__ mask) {
    const int64_t N = (int64_t)H * W, total = (int64_t)B * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / N, p = i - b * N;
        const int x = (int)(p % W), y = (int)(p / W);
        const L* lab = labels + b * N;
        const L here = lab[p];
        for (int k = 0; k < tab.K; ++k) {
            const int qy = y + tab.dy[k], qx = x + tab.dx[k];
            const bool inb = (unsigned)qy < (unsigned)H && (unsigned)qx < (unsigned)W;
            float a;
            if (inb) {
                const L there = __ldg(lab + (int64_t)qy * W + qx);
                bool same = (here == there);
                if (ignore) same = same && (here != (L)0) && (there != (L)0);
                a = same ? 1.f : 0.f;
            } else {
                a = padding ? 1.f : 0.f;
            }
            const int64_t o = (b * tab.K + k) * N + p;
            affs[o] = a;
            if (mask != nullptr) mask[o] = inb ? 1 : 0;
        }
    }
}

// 3-D (or 2-D with D = 1): aff = (seg[p] == seg[p+n]) & seg[p] > 0 & seg[p+n] > 0 on the in-volume pairs, 0 elsewhere;
// replicate_pad overwrites the first slice along axis e of channel e with seg > 0 (channels 0..2 only).
// Signed label types compare `> 0` as the reference's NumPy does.
template <typename L>
__global__ void __launch_bounds__(256)
gt_affs3d_kernel(const L* __restrict__ seg, int B, int D, int H, int W, GtTable tab, int replicate_pad, int nd,
                 float* __restrict__ affs) {
    const int64_t N = (int64_t)D * H * W, total = (int64_t)B * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / N, p = i - b * N;
        const int x = (int)(p % W), y = (int)((p / W) % H), z = (int)(p / ((int64_t)W * H));
        const L* s = seg + b * N;
        const L here = s[p];
        const bool fg = here > (L)0;
        for (int k = 0; k < tab.K; ++k) {
            const int qz = z + tab.dz[k], qy = y + tab.dy[k], qx = x + tab.dx[k];
            float a = 0.f;
            if ((unsigned)qz < (unsigned)D && (unsigned)qy < (unsigned)H && (unsigned)qx < (unsigned)W) {
                const L there = __ldg(s + ((int64_t)qz * H + qy) * W + qx);
                a = (fg && there > (L)0 && here == there) ? 1.f : 0.f;
            }
            if (replicate_pad) {
                // nd == 3: channel 0/1/2 <-> first slice along z/y/x;  nd == 2 (D == 1): channel 0/1 <-> y/x
                const int first = (nd == 3) ? (k == 0 ? z : (k == 1 ? y : x)) : (k == 0 ? y : x);
                if (first == 0) a = fg ? 1.f : 0.f;
            }
            affs[(b * tab.K + k) * N + p] = a;
        }
    }
}

// ---- weight_binary_ratio ------------------------------------------------------------------------
// stats per channel: [0] count of non-zeros (u64), [1] ordered-int max, [2] ordered-int min (stored as u64)
__device__ __forceinline__ int ordered(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}

__global__ void wbr_init_kernel(unsigned long long* stats, int n_channels) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_channels) {
        stats[3 * c + 0] = 0ull;
        reinterpret_cast<long long*>(stats)[3 * c + 1] = (long long)INT32_MIN;
        reinterpret_cast<long long*>(stats)[3 * c + 2] = (long long)INT32_MAX;
    }
}

__global__ void __launch_bounds__(256)
wbr_count_kernel(const float* __restrict__ label, int64_t n, int chunks, unsigned long long* __restrict__ stats) {
    const int ch = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    const float* src = label + (int64_t)ch * n;
    const int64_t per = (n + chunks - 1) / chunks;
    const int64_t lo = (int64_t)chunk * per, hi = (lo + per < n) ? lo + per : n;
    unsigned int cnt = 0;
    int mx = INT32_MIN, mn = INT32_MAX;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float v = __ldg(src + i);
        cnt += (v != 0.f) ? 1u : 0u;
        const int o = ordered(v + 0.f);  // +0.f folds -0 onto +0 so that max == min matches float equality
        mx = max(mx, o);
        mn = min(mn, o);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(stats + 3 * ch + 0, (unsigned long long)cnt);
        atomicMax(reinterpret_cast<long long*>(stats) + 3 * ch + 1, (long long)mx);
        atomicMin(reinterpret_cast<long long*>(stats) + 3 * ch + 2, (long long)mn);
    }
}

__global__ void __launch_bounds__(256)
wbr_write_kernel(const float* __restrict__ label, int64_t n, int chunks, const unsigned long long* __restrict__ stats,
                 double alpha, float* __restrict__ weight) {
    const int ch = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    const unsigned long long cnt = stats[3 * ch + 0];
    const bool uniform = reinterpret_cast<const long long*>(stats)[3 * ch + 1] ==
                         reinterpret_cast<const long long*>(stats)[3 * ch + 2];
    float w_pos = 1.f, w_neg = 1.f;
    if (!uniform) {
        // Python float64 arithmetic of the reference, evaluated left to right
        double wf = (double)cnt / (double)n;
        wf = fmin(fmax(wf, 5e-2), 0.99);
        if (wf > 0.5) {
            w_pos = 1.f;
            w_neg = (float)(alpha * wf / (1.0 - wf));
        } else {
            w_pos = (float)(alpha * (1.0 - wf) / wf);
            w_neg = 1.f;
        }
    }
    const float* src = label + (int64_t)ch * n;
    float* dst = weight + (int64_t)ch * n;
    const int64_t per = (n + chunks - 1) / chunks;
    const int64_t lo = (int64_t)chunk * per, hi = (lo + per < n) ? lo + per : n;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) dst[i] = (__ldg(src + i) != 0.f) ? w_pos : w_neg;
}

// ---- fused batch GT: labels -> target (+ mask) with the per-(sample, channel) non-zero counts accumulated in the
//      same pass, then weights recomputed from the labels (the target is never re-read): 2 label reads + 4K (+K)
//      + 4K bytes per voxel, the algorithmic minimum of SURVEY.md 8(d).
// MODE 0: gen_affs_ours semantics (2-D, validity mask, padding/ignore);  MODE 1: seg_to_aff semantics (background-aware).
template <typename L, int MODE>
__device__ __forceinline__ float gt_value(const L* __restrict__ s, L here, int z, int y, int x, int dz, int dy, int dx,
                                          int D, int H, int W, int ignore, int padding, bool& inb) {
    const int qz = z + dz, qy = y + dy, qx = x + dx;
    inb = (unsigned)qz < (unsigned)D && (unsigned)qy < (unsigned)H && (unsigned)qx < (unsigned)W;
    if (!inb) return (MODE == 0 && padding) ? 1.f : 0.f;
    const L there = __ldg(s + ((int64_t)qz * H + qy) * W + qx);
    if (MODE == 0) {
        bool same = (here == there);
        if (ignore) same = same && (here != (L)0) && (there != (L)0);
        return same ? 1.f : 0.f;
    }
    return (here > (L)0 && there > (L)0 && here == there) ? 1.f : 0.f;
}

// The own labels of a thread's NV (4 or 8) consecutive voxels of one row, loaded once and reused for every offset.
template <typename L, int NV>
struct Own {
    L here[NV];
    uint32_t hw[NV / 4];  // 8-bit labels: four labels per word
};

template <typename L, int NV>
__device__ __forceinline__ void load_own(const L* __restrict__ s, bool bytes4, uint32_t p, Own<L, NV>& q) {
#pragma unroll
    for (int i = 0; i < NV / 4; ++i) q.hw[i] = 0;
    if (sizeof(L) == 1 && bytes4) {
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) {
            q.hw[i] = reinterpret_cast<const uint32_t*>(s)[(p >> 2) + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) q.here[4 * i + j] = (L)((q.hw[i] >> (8 * j)) & 0xffu);
        }
    } else if (sizeof(L) == 4 && (reinterpret_cast<uintptr_t>(s) & 15) == 0) {
#pragma unroll
        for (int i = 0; i < NV / 4; ++i) {
            const uint4 h4 = *reinterpret_cast<const uint4*>(s + p + 4 * i);
            memcpy(q.here + 4 * i, &h4, 16);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NV; ++j) q.here[j] = s[p + j];
    }
}

template <typename L, int MODE>
__device__ __forceinline__ bool gt_same(L here, L there, int ignore) {
    bool same = (here == there);
    if (MODE == 0) { if (ignore) same = same && (here != (L)0) && (there != (L)0); }
    else same = same && (here > (L)0) && (there > (L)0);
    return same;
}

// NV consecutive voxels p .. p + NV - 1 of one row against one offset: bit j of the result = "the target of voxel p + j
// is 1", bit j of `inb` = "the pair of voxel p + j lies inside the volume".  The validity of the NV pairs is ONE bit mask
// computed arithmetically; almost every group has all its partners inside: that path carries no per-element predicate.
// 8-bit labels: the neighbours are NV / 4 (+ 1) aligned words funnel-shifted into place, every four comparisons one SIMD
// instruction, and a multiply gathers four result bits.  Other widths: NV loads (128-bit loads for 4-byte labels when the
// offset keeps the alignment) and NV compares.  Partial groups (row ends) load only their valid elements.
// W % NV == 0 and an aligned base make every vector access of a full group lie inside the array.
template <typename L, int MODE, int NV>
__device__ __forceinline__ uint32_t gt_group(const L* __restrict__ s, bool bytes4, const Own<L, NV>& q, uint32_t p, int z, int y,
                                             int x, int dz, int dy, int dx, int32_t delta, int D, int H, int W, int ignore,
                                             int padding, uint32_t& inb) {
    constexpr uint32_t FULL = (1u << NV) - 1u;
    const bool row_in = (unsigned)(z + dz) < (unsigned)D && (unsigned)(y + dy) < (unsigned)H;
    const int xs = x + dx;  // column of the first partner
    const int lo = min(max(-xs, 0), NV), hi = min(max(xs + NV - W, 0), NV);
    const uint32_t m = row_in ? ((FULL << lo) & (FULL >> hi) & FULL) : 0u;
    inb = m;
    const int32_t nb = (int32_t)p + delta;
    uint32_t eq = 0;
    if (m == FULL) {
        if (sizeof(L) == 1 && bytes4) {
            const uint32_t* w32 = reinterpret_cast<const uint32_t*>(s) + ((uint32_t)nb >> 2);
            const uint32_t r3 = (uint32_t)nb & 3u;
            uint32_t w[NV / 4 + 1];
#pragma unroll
            for (int i = 0; i < NV / 4; ++i) w[i] = __ldg(w32 + i);
            w[NV / 4] = r3 ? __ldg(w32 + NV / 4) : 0u;
#pragma unroll
            for (int i = 0; i < NV / 4; ++i) {
                const uint32_t tw = __funnelshift_r(w[i], w[i + 1], 8u * r3);
                uint32_t e8 = __vcmpeq4(q.hw[i], tw);  // 0xff where equal
                if (MODE == 0) { if (ignore) e8 &= __vcmpne4(q.hw[i], 0u) & __vcmpne4(tw, 0u); }
                else e8 &= __vcmpne4(q.hw[i], 0u);  // unsigned 8-bit ids: > 0 is != 0 (there == here)
                eq |= (((e8 & 0x01010101u) * 0x10204080u) >> 28) << (4 * i);  // bit 8 j -> bit j
            }
            return eq;
        }
        L there[NV];
        if (sizeof(L) == 4 && (nb & 3) == 0 && (reinterpret_cast<uintptr_t>(s) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < NV / 4; ++i) {
                const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(s + nb + 4 * i));
                memcpy(there + 4 * i, &t4, 16);
            }
        } else {
#pragma unroll
            for (int j = 0; j < NV; ++j) there[j] = __ldg(s + (nb + j));
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) eq |= gt_same<L, MODE>(q.here[j], there[j], ignore) ? (1u << j) : 0u;
        return eq;
    }
    if (m != 0u) {
        // a row end: branch-free -- a pair outside the row reads the voxel itself (always a valid address) and is masked out
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const L there = __ldg(s + (((m >> j) & 1u) ? (uint32_t)(nb + j) : p + (uint32_t)j));
            eq |= gt_same<L, MODE>(q.here[j], there, ignore) ? (1u << j) : 0u;
        }
        eq &= m;
    }
    return eq | ((MODE == 0 && padding) ? (~m & FULL) : 0u);
}

__device__ __forceinline__ float4 quad_values(uint32_t bits, float one, float zero) {
    return make_float4((bits & 1u) ? one : zero, (bits & 2u) ? one : zero, (bits & 4u) ? one : zero, (bits & 8u) ? one : zero);
}
__device__ __forceinline__ uint32_t quad_bytes(uint32_t bits) {  // bit j -> byte j (0 / 1): sixteen distinct partial products, no carry
    return ((bits & 15u) * 0x00204081u) & 0x01010101u;
}

// the two values of weight_binary_ratio for one (sample, channel): Python float64 arithmetic of the reference, left to right
__device__ __forceinline__ void pair_weights(unsigned long long cnt, unsigned long long N, double alpha, float& w_pos, float& w_neg) {
    w_pos = 1.f;
    w_neg = 1.f;
    if (cnt != 0ull && cnt != N) {  // values are 0 / 1: the reference's "all equal -> ones" test is count in {0, N}
        double wf = (double)cnt / (double)N;
        wf = fmin(fmax(wf, 5e-2), 0.99);
        if (wf > 0.5) w_neg = (float)(alpha * wf / (1.0 - wf));
        else w_pos = (float)(alpha * (1.0 - wf) / wf);
    }
}

// Two-launch route, NV (8 when W % 8 == 0, else 4) consecutive x per thread: one coordinate decode per group, 128-bit
// stores, 32-bit index arithmetic inside a sample.  PASS 1 writes target (+ mask) and accumulates the exact
// per-(sample, channel) non-zero counts (warp redux -> per-warp shared counters -> one global atomic per block and
// channel; integer, so the order does not matter).  PASS 2 recomputes the 0/1 values from the labels and writes the
// two-valued weights.  grid = (chunks, B); a block handles a contiguous range of voxel groups of one sample.
template <typename L, int MODE, int PASS, int NV>
__global__ void __launch_bounds__(256)
gt_fused_x4(const L* __restrict__ labels, int D, int H, int W, GtTable tab, int ignore, int padding, double alpha,
            float* __restrict__ target, uint8_t* __restrict__ mask, unsigned long long* __restrict__ counts,
            float* __restrict__ weight) {
    __shared__ unsigned int s_cnt[8][PEA_MAX_OFFSETS];
    __shared__ float s_pos[PEA_MAX_OFFSETS], s_neg[PEA_MAX_OFFSETS];
    const int b = blockIdx.y, K = tab.K;
    const uint32_t HW = (uint32_t)H * W, N = HW * (uint32_t)D, NG = N / NV;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (PASS == 1) {
        for (int i = threadIdx.x; i < 8 * PEA_MAX_OFFSETS; i += blockDim.x) (&s_cnt[0][0])[i] = 0;
    } else {
        for (int k = threadIdx.x; k < K; k += blockDim.x) pair_weights(counts[(int64_t)b * K + k], N, alpha, s_pos[k], s_neg[k]);
    }
    __syncthreads();
    const L* s = labels + (int64_t)b * N;
    // packed-byte fast path: unsigned 8-bit codes only (a signed -1 would be "not > 0" in seg_to_aff semantics)
    const bool bytes4 = sizeof(L) == 1 && (L)(-1) > (L)0 && (reinterpret_cast<uintptr_t>(s) & 3) == 0;
    float* out = (PASS == 1 ? target : weight) + (int64_t)b * K * N;
    uint8_t* mout = (PASS == 1 && MODE == 0 && mask) ? mask + (int64_t)b * K * N : nullptr;
    const uint32_t per = (NG + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = blockIdx.x * per, hi = (lo + per < NG) ? lo + per : NG;
    for (uint32_t base = lo; base < hi; base += blockDim.x) {
        const uint32_t gd = base + threadIdx.x;
        const bool live = gd < hi;
        const uint32_t p = (live ? gd : lo) * NV;
        const int z = (int)(p / HW);
        const uint32_t r = p - (uint32_t)z * HW;
        const int y = (int)(r / (uint32_t)W), x = (int)(r - (uint32_t)y * W);
        Own<L, NV> q;
        load_own(s, bytes4, p, q);
#pragma unroll kGtUnroll
        for (int k = 0; k < K; ++k) {  // unrolled: the neighbour loads of several offsets are in flight together
            const int dz = tab.dz[k], dy = tab.dy[k], dx = tab.dx[k];
            uint32_t inb;
            const uint32_t bits = gt_group<L, MODE, NV>(s, bytes4, q, p, z, y, x, dz, dy, dx, dz * (int32_t)HW + dy * W + dx, D, H,
                                                        W, ignore, padding, inb);
            const uint32_t o = (uint32_t)k * N + p;
            if (PASS == 1) {
                if (live && target) {  // target == nullptr: counts only (pea_gt_pair_weights)
#pragma unroll
                    for (int i = 0; i < NV / 4; ++i) {
                        *reinterpret_cast<float4*>(out + o + 4 * i) = quad_values(bits >> (4 * i), 1.f, 0.f);
                        if (mout) *reinterpret_cast<uint32_t*>(mout + o + 4 * i) = quad_bytes(inb >> (4 * i));
                    }
                }
                const unsigned int tot = __reduce_add_sync(0xffffffffu, live ? (unsigned)__popc(bits) : 0u);
                if (lane == 0) atomicAdd(&s_cnt[warp][k], tot);
            } else if (live) {
#pragma unroll
                for (int i = 0; i < NV / 4; ++i)
                    *reinterpret_cast<float4*>(out + o + 4 * i) = quad_values(bits >> (4 * i), s_pos[k], s_neg[k]);
            }
        }
    }
    if (PASS == 1) {
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            unsigned int c = 0;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) c += s_cnt[w8][k];
            if (c) atomicAdd(counts + (int64_t)b * K + k, (unsigned long long)c);
        }
    }
}

// One-launch route (cooperative launch: every block resident).  Work = slices of blockDim.x groups of NV voxels, dealt
// round-robin to the blocks.  A block
//   1. COUNTS all its slices: compares only, the NV result bits (and the NV validity bits of the mask) of every
//      (slice, offset, thread) parked in shared memory, counts -> global atomics, one completion counter per SAMPLE;
//   2a. streams target (+ mask) of the same slices out of the parked bits -- nothing here depends on the counts, so this
//      phase hides the wait for the other blocks;
//   2b. per slice, waits until every slice of the slice's sample has been counted and streams the weights.
// The store phases re-deal the parked bits so that consecutive lanes own consecutive quads whatever NV is: every 128-bit
// store instruction of a warp covers 512 contiguous bytes.  No grid-wide barrier, no second round of compares, and the
// store phases run at the pace of the memory system (about 20 instructions per offset and quad).
template <typename L, int MODE, int NV>
__global__ void __launch_bounds__(256)
gt_fused_one(const L* __restrict__ labels, int B, int D, int H, int W, GtTable tab, int ignore, int padding, double alpha,
             float* __restrict__ target, uint8_t* __restrict__ mask, unsigned long long* __restrict__ counts,
             float* __restrict__ weight, unsigned int sps, unsigned int n_slices, unsigned int its) {
    constexpr uint32_t Q = NV / 4;  // quads per thread in phase 1
    extern __shared__ uint8_t s_park[];  // [slice of the block][k][thread]: result bits, then (2-D) validity bits
    __shared__ unsigned int s_cnt[8][PEA_MAX_OFFSETS];
    __shared__ float s_pos[PEA_MAX_OFFSETS], s_neg[PEA_MAX_OFFSETS];
    const int K = tab.K;
    const uint32_t HW = (uint32_t)H * W, N = HW * (uint32_t)D, NG = N / NV, T = blockDim.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool with_mask = MODE == 0 && mask != nullptr;
    uint8_t* s_bits = s_park;
    uint8_t* s_inb = s_park + (size_t)its * K * T;
    unsigned int* done = reinterpret_cast<unsigned int*>(counts + (int64_t)B * K);  // [B], zeroed with the counts
    for (int i = threadIdx.x; i < 8 * PEA_MAX_OFFSETS; i += T) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    // ---- 1: count
    uint32_t it = 0;
    for (uint32_t sl = blockIdx.x; sl < n_slices; sl += gridDim.x, ++it) {
        const uint32_t b = sl / sps, gd = (sl - b * sps) * T + threadIdx.x;
        const bool live = gd < NG;
        const uint32_t p = (live ? gd : 0u) * NV;
        const int z = (int)(p / HW);
        const uint32_t r = p - (uint32_t)z * HW;
        const int y = (int)(r / (uint32_t)W), x = (int)(r - (uint32_t)y * W);
        const L* s = labels + (int64_t)b * N;
        const bool bytes4 = sizeof(L) == 1 && (L)(-1) > (L)0 && (reinterpret_cast<uintptr_t>(s) & 3) == 0;
        Own<L, NV> q;
        load_own(s, bytes4, p, q);
#pragma unroll kGtUnroll
        for (int k = 0; k < K; ++k) {
            const int dz = tab.dz[k], dy = tab.dy[k], dx = tab.dx[k];
            uint32_t inb;
            const uint32_t bits = gt_group<L, MODE, NV>(s, bytes4, q, p, z, y, x, dz, dy, dx, dz * (int32_t)HW + dy * W + dx, D, H,
                                                        W, ignore, padding, inb);
            const unsigned int tot = __reduce_add_sync(0xffffffffu, live ? (unsigned)__popc(bits) : 0u);
            if (lane == 0) atomicAdd(&s_cnt[warp][k], tot);
            const uint32_t slot = (it * (uint32_t)K + (uint32_t)k) * T + threadIdx.x;
            s_bits[slot] = (uint8_t)bits;
            if (with_mask) s_inb[slot] = (uint8_t)inb;
        }
        __syncthreads();
        if ((int)threadIdx.x < K) {
            const int k = threadIdx.x;
            unsigned int c = 0;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) {
                c += s_cnt[w8][k];
                s_cnt[w8][k] = 0;
            }
            if (c) atomicAdd(counts + (int64_t)b * K + k, (unsigned long long)c);
            __threadfence();  // the count is visible before the slice is reported
        }
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(done + b, 1u);
    }
    // the parked byte of thread g / Q holds quad g % Q of its group: lane -> consecutive quads
    const uint32_t sub = threadIdx.x / Q, qsel = 4u * (threadIdx.x % Q);
    // ---- 2a: target (+ mask)
    it = 0;
    for (uint32_t sl = blockIdx.x; sl < n_slices; sl += gridDim.x, ++it) {
        const uint32_t b = sl / sps, g0 = (sl - b * sps) * T;
        float* tout = target + (int64_t)b * K * N;
        uint8_t* mout = with_mask ? mask + (int64_t)b * K * N : nullptr;
#pragma unroll
        for (uint32_t h = 0; h < Q; ++h) {
            const uint32_t g = h * (T / Q) + sub;  // thread whose byte this lane expands
            const uint32_t gd = g0 + g;
            if (gd < NG) {
                const uint32_t px = gd * NV + qsel;
#pragma unroll 4
                for (int k = 0; k < K; ++k) {
                    const uint32_t slot = (it * (uint32_t)K + (uint32_t)k) * T + g;
                    *reinterpret_cast<float4*>(tout + ((uint32_t)k * N + px)) = quad_values((uint32_t)s_bits[slot] >> qsel, 1.f, 0.f);
                    if (with_mask) *reinterpret_cast<uint32_t*>(mout + ((uint32_t)k * N + px)) = quad_bytes((uint32_t)s_inb[slot] >> qsel);
                }
            }
        }
    }
    // ---- 2b: weights, sample by sample
    it = 0;
    for (uint32_t sl = blockIdx.x; sl < n_slices; sl += gridDim.x, ++it) {
        const uint32_t b = sl / sps, g0 = (sl - b * sps) * T;
        if (threadIdx.x == 0) {
            unsigned int seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(done + b) : "memory");
            } while (seen < sps);
        }
        __syncthreads();
        if ((int)threadIdx.x < K) pair_weights(__ldcg(counts + (int64_t)b * K + threadIdx.x), N, alpha, s_pos[threadIdx.x], s_neg[threadIdx.x]);
        __syncthreads();
        float* wout = weight + (int64_t)b * K * N;
#pragma unroll
        for (uint32_t h = 0; h < Q; ++h) {
            const uint32_t g = h * (T / Q) + sub;
            const uint32_t gd = g0 + g;
            if (gd < NG) {
                const uint32_t px = gd * NV + qsel;
#pragma unroll 4
                for (int k = 0; k < K; ++k) {
                    const uint32_t bits = (uint32_t)s_bits[(it * (uint32_t)K + (uint32_t)k) * T + g] >> qsel;
                    *reinterpret_cast<float4*>(wout + ((uint32_t)k * N + px)) = quad_values(bits, s_pos[k], s_neg[k]);
                }
            }
        }
        __syncthreads();  // s_pos / s_neg are rewritten for the next slice
    }
}

constexpr int kGtNotLaunched = -1000;  // run_gt_fused_nv: the one-launch route is not available here (not an error)
constexpr int kGtOneMaxBytes = 24 * 1024;  // parked bits of the one-launch route: (slices per block) * K * threads (* 2 with a mask)

template <typename L, int MODE, int NV>
static int run_gt_fused_nv(const void* labels, int B, int D, int H, int W, const GtTable& tab, int ignore, int padding,
                           double alpha, float* target, uint8_t* mask, float* weight, unsigned long long* counts,
                           cudaStream_t st) {
    const int64_t N = (int64_t)D * H * W;
    // counts + the per-sample completion counters of the one-launch route
    cudaError_t e = cudaMemsetAsync(counts, 0, pea_gt_counts_bytes(B, tab.K), st);
    if (e != cudaSuccess) return (int)e;
    const L* lab = static_cast<const L*>(labels);
    // ---- one cooperative launch when every block can be resident and a block's results fit in shared memory
    if (debug_knob(PEA_KNOB_GT_TWO_PASS) == 0) {
        auto kern = gt_fused_one<L, MODE, NV>;
        int dev = 0, sms = 0, coop = 0;
        const int64_t NG = N / NV;
        const int planes = (MODE == 0 && mask) ? 2 : 1;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coop && sms > 0) {
            // every block gets the SAME number of slices (a block that finishes early only waits for the others at its
            // sample's completion counter); among block sizes and slices per block pick the most even load per SM
            int best_thr = 0, best_its = 0;
            unsigned int best_grid = 0;
            double best_eff = 0.0;
            for (int thr = 256; thr >= 128; thr >>= 1) {
                const int64_t sps = (NG + thr - 1) / thr, n_slices = sps * B;
                if (n_slices >= ((int64_t)1 << 31)) continue;
                for (int its = 1; its * tab.K * thr * planes <= kGtOneMaxBytes; ++its) {
                    int per_sm = 0;
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, thr, (size_t)its * tab.K * thr * planes) != cudaSuccess ||
                        per_sm < 1)
                        break;
                    const int64_t grid = (n_slices + its - 1) / its;
                    if (grid > (int64_t)per_sm * sms) continue;
                    const int64_t worst = (grid + sms - 1) / sms;  // blocks on the fullest SM
                    const double eff = (double)n_slices / ((double)worst * its * sms);
                    if (eff > best_eff + 0.02) {
                        best_eff = eff; best_thr = thr; best_its = its; best_grid = (unsigned)grid;
                    }
                }
            }
            if (best_thr) {
                const int64_t sps = (NG + best_thr - 1) / best_thr;
                int B_ = B, D_ = D, H_ = H, W_ = W, ig = ignore, pad = padding;
                unsigned int sps_ = (unsigned)sps, ns_ = (unsigned)(sps * B), its_ = (unsigned)best_its;
                GtTable t2 = tab;
                void* args[] = {(void*)&lab, &B_, &D_, &H_, &W_, &t2, &ig, &pad, &alpha, &target, &mask, &counts, &weight, &sps_, &ns_, &its_};
                e = cudaLaunchCooperativeKernel((const void*)kern, dim3(best_grid), dim3(best_thr), args,
                                                (size_t)best_its * tab.K * best_thr * planes, st);
                if (e == cudaSuccess) return 0;
                (void)cudaGetLastError();  // not launchable as a cooperative grid here: the two-launch route below
            }
        }
    }
    return kGtNotLaunched;  // the caller takes the two-launch route
}

// One launch: eight voxels per thread in the counting phase when the rows allow it (the store phases are re-dealt quad
// by quad either way).  Two launches (volumes too large to be resident at once, or no cooperative launch): four voxels
// per thread -- eight would leave every 128-bit store instruction with half-filled 32-byte sectors (measured 77.6 us
// against 61.6 us at B = 8, 520 x 696, K = 10 when the stores sat in the counting loop).
template <typename L, int MODE>
static int run_gt_fused(const void* labels, int B, int D, int H, int W, const GtTable& tab, int ignore, int padding,
                        double alpha, float* target, uint8_t* mask, float* weight, unsigned long long* counts,
                        cudaStream_t st) {
    int rc;
    if (W % 8 == 0 && sizeof(L) <= 4)
        rc = run_gt_fused_nv<L, MODE, 8>(labels, B, D, H, W, tab, ignore, padding, alpha, target, mask, weight, counts, st);
    else
        rc = run_gt_fused_nv<L, MODE, 4>(labels, B, D, H, W, tab, ignore, padding, alpha, target, mask, weight, counts, st);
    if (rc != kGtNotLaunched) return rc;
    const int64_t N = (int64_t)D * H * W;
    int64_t chunks = (148 * 8 + B - 1) / B;
    const int64_t max_chunks = (N / 4 + 1023) / 1024;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    const L* lab = static_cast<const L*>(labels);
    dim3 grid((unsigned)chunks, (unsigned)B);
    gt_fused_x4<L, MODE, 1, 4><<<grid, 256, 0, st>>>(lab, D, H, W, tab, ignore, padding, alpha, target, mask, counts, weight);
    gt_fused_x4<L, MODE, 2, 4><<<grid, 256, 0, st>>>(lab, D, H, W, tab, ignore, padding, alpha, target, mask, counts, weight);
    return (int)cudaGetLastError();
}

// per-(sample, channel) counts -> the two values of the class-balance weight map: wtab[2 i] where the target is 1,
// wtab[2 i + 1] where it is 0 (weight_binary_ratio, scripts_cvppp/data/data_segmentation.py:205-228; same float64
// arithmetic as PASS 2 above, so a loss kernel that derives its weights from this table sees bit-identical values)
__global__ void gt_weight_table_kernel(const unsigned long long* __restrict__ counts, int n, unsigned long long N, double alpha,
                                       float* __restrict__ wtab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float w_pos, w_neg;
    pair_weights(counts[i], N, alpha, w_pos, w_neg);
    wtab[2 * i] = w_pos;
    wtab[2 * i + 1] = w_neg;
}

template <typename L, int MODE, int NV>
static int run_gt_weights_nv(const void* labels, int B, int D, int H, int W, const GtTable& tab, int ignore, int padding,
                             double alpha, float* wtab, unsigned long long* counts, cudaStream_t st) {
    const int64_t N = (int64_t)D * H * W;
    // counting only: no stores to amortise, so one group per thread and every SM full (8 blocks of 256 threads each)
    int64_t chunks = (148 * 8 + B - 1) / B;
    const int64_t max_chunks = (N / NV + 255) / 256;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)B * tab.K * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)chunks, (unsigned)B);
    gt_fused_x4<L, MODE, 1, NV><<<grid, 256, 0, st>>>(static_cast<const L*>(labels), D, H, W, tab, ignore, padding, alpha, nullptr,
                                                     nullptr, counts, nullptr);
    const int n = B * tab.K;
    gt_weight_table_kernel<<<(n + 127) / 128, 128, 0, st>>>(counts, n, (unsigned long long)N, alpha, wtab);
    return (int)cudaGetLastError();
}

template <typename L, int MODE>
static int run_gt_weights(const void* labels, int B, int D, int H, int W, const GtTable& tab, int ignore, int padding,
                          double alpha, float* wtab, unsigned long long* counts, cudaStream_t st) {
    // counting only (no stores): eight voxels per thread when the rows allow it
    if (W % 8 == 0 && sizeof(L) <= 4) return run_gt_weights_nv<L, MODE, 8>(labels, B, D, H, W, tab, ignore, padding, alpha, wtab, counts, st);
    return run_gt_weights_nv<L, MODE, 4>(labels, B, D, H, W, tab, ignore, padding, alpha, wtab, counts, st);
}

// ---- launchers -----------------------------------------------------------------------------------
static int grid1d(int64_t total) {
    int64_t g = (total + 255) / 256;
    return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

template <typename L>
static int run_gt2d(const void* labels, int B, int H, int W, const GtTable& tab, int ignore, int padding, float* affs,
                    uint8_t* mask, cudaStream_t st) {
    gt_affs2d_kernel<L><<<grid1d((int64_t)B * H * W), 256, 0, st>>>(static_cast<const L*>(labels), B, H, W, tab, ignore,
                                                                   padding, affs, mask);
    return (int)cudaGetLastError();
}

template <typename L>
static int run_gt3d(const void* seg, int B, int D, int H, int W, const GtTable& tab, int replicate_pad, int nd,
                    float* affs, cudaStream_t st) {
    gt_affs3d_kernel<L><<<grid1d((int64_t)B * D * H * W), 256, 0, st>>>(static_cast<const L*>(seg), B, D, H, W, tab,
                                                                       replicate_pad, nd, affs);
    return (int)cudaGetLastError();
}

}  // namespace pea

using namespace pea;

extern "C" int pea_gt_affs2d(const void* labels, int label_dtype, int B, int H, int W, const int32_t* offsets, int K,
                             int ignore, int padding, float* affs, uint8_t* mask, void* stream) {
    if (!labels || !offsets || !affs) return PEA_E_NULL;
    if (B <= 0 || H <= 0 || W <= 0) return PEA_E_SHAPE;
    if (K < 1 || K > PEA_MAX_OFFSETS) return PEA_E_OFFSETS;
    GtTable tab;
    tab.K = K;
    for (int k = 0; k < K; ++k) { tab.dz[k] = 0; tab.dy[k] = offsets[2 * k]; tab.dx[k] = offsets[2 * k + 1]; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (label_dtype) {
        // equality and `!= 0` do not depend on signedness: same-width codes share a kernel
        case PEA_LABEL_U8: case PEA_LABEL_S8: return run_gt2d<uint8_t>(labels, B, H, W, tab, ignore, padding, affs, mask, st);
        case PEA_LABEL_U16: case PEA_LABEL_S16: return run_gt2d<uint16_t>(labels, B, H, W, tab, ignore, padding, affs, mask, st);
        case PEA_LABEL_U32: case PEA_LABEL_S32: return run_gt2d<uint32_t>(labels, B, H, W, tab, ignore, padding, affs, mask, st);
        case PEA_LABEL_U64: case PEA_LABEL_S64: return run_gt2d<uint64_t>(labels, B, H, W, tab, ignore, padding, affs, mask, st);
        case PEA_LABEL_F32: return run_gt2d<float>(labels, B, H, W, tab, ignore, padding, affs, mask, st);
        default: return PEA_E_DTYPE;
    }
}

extern "C" int pea_gt_affs3d(const void* seg, int label_dtype, int B, int D, int H, int W, const int32_t* nhood, int K,
                             int replicate_pad, float* affs, void* stream) {
    if (!seg || !nhood || !affs) return PEA_E_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return PEA_E_SHAPE;
    if (K < 1 || K > PEA_MAX_OFFSETS) return PEA_E_OFFSETS;
    GtTable tab;
    tab.K = K;
    for (int k = 0; k < K; ++k) { tab.dz[k] = nhood[3 * k]; tab.dy[k] = nhood[3 * k + 1]; tab.dx[k] = nhood[3 * k + 2]; }
    // replicate_pad: 3 -> three edges of a 3-D volume, 2 -> two edges of a 2-D image passed as D = 1
    const int nd = (replicate_pad == 2) ? 2 : 3;
    if (replicate_pad && K != nd) return PEA_E_MODE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the `seg > 0` test is signed exactly when the caller's dtype is
    switch (label_dtype) {
        case PEA_LABEL_U8: return run_gt3d<uint8_t>(seg, B, D, H, W, tab, replicate_pad, nd, affs, st);
        case PEA_LABEL_U16: return run_gt3d<uint16_t>(seg, B, D, H, W, tab, replicate_pad, nd, affs, st);
        case PEA_LABEL_U32: return run_gt3d<uint32_t>(seg, B, D, H, W, tab, replicate_pad, nd, affs, st);
        case PEA_LABEL_U64: return run_gt3d<uint64_t>(seg, B, D, H, W, tab, replicate_pad, nd, affs, st);
        case PEA_LABEL_S8: return run_gt3d<int8_t>(seg, B, D, H, W, tab, replicate_pad, nd, affs, st);
        case PEA_LABEL_S16: return run_gt3d<int16_t>(seg, B, D, H, W, tab, replicate_pad, nd, affs, st);
        case PEA_LABEL_S32: return run_gt3d<int32_t>(seg, B, D, H, W, tab, replicate_pad, nd, affs, st);
        case PEA_LABEL_S64: return run_gt3d<int64_t>(seg, B, D, H, W, tab, replicate_pad, nd, affs, st);
        case PEA_LABEL_F32: return run_gt3d<float>(seg, B, D, H, W, tab, replicate_pad, nd, affs, st);
        default: return PEA_E_DTYPE;
    }
}

extern "C" size_t pea_wbr_stats_bytes(int n_channels) {
    return n_channels > 0 ? (size_t)n_channels * 3 * sizeof(unsigned long long) : 0;
}

extern "C" int pea_weight_binary_ratio(const float* label, int n_channels, int64_t n_per_channel, double alpha,
                                       float* weight, void* stats, size_t stats_bytes, void* stream) {
    if (!label || !weight || !stats) return PEA_E_NULL;
    if (n_channels <= 0 || n_per_channel <= 0) return PEA_E_SHAPE;
    if (stats_bytes < pea_wbr_stats_bytes(n_channels)) return PEA_E_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(stats) % 8) return PEA_E_ALIGN;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto* s = static_cast<unsigned long long*>(stats);
    // enough blocks per channel to fill the machine, at least ~16 K elements each
    int64_t chunks = (148 * 8 + n_channels - 1) / n_channels;
    const int64_t max_chunks = (n_per_channel + 16383) / 16384;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks < 1) chunks = 1;
    if ((int64_t)n_channels * chunks > 0x7fffffffLL) return PEA_E_SHAPE;
    wbr_init_kernel<<<(n_channels + 127) / 128, 128, 0, st>>>(s, n_channels);
    wbr_count_kernel<<<(unsigned)(n_channels * chunks), 256, 0, st>>>(label, n_per_channel, (int)chunks, s);
    wbr_write_kernel<<<(unsigned)(n_channels * chunks), 256, 0, st>>>(label, n_per_channel, (int)chunks, s, alpha,
                                                                      weight);
    return (int)cudaGetLastError();
}

extern "C" size_t pea_gt_counts_bytes(int B, int K) {
    // per-(sample, channel) counts + one completion counter per sample (one-launch route), rounded up to 16 bytes
    return (B > 0 && K > 0) ? ((size_t)B * K * sizeof(unsigned long long) + (size_t)B * sizeof(unsigned int) + 15) / 16 * 16 : 0;
}

extern "C" int pea_gt_targets(const void* labels, int label_dtype, int B, int D, int H, int W, const int32_t* offsets,
                              int K, int semantics, int ignore, int padding, double alpha, float* target, uint8_t* mask,
                              float* weight, void* counts, size_t counts_bytes, void* stream) {
    if (!labels || !offsets || !target || !weight || !counts) return PEA_E_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || B > 65535) return PEA_E_SHAPE;
    if (K < 1 || K > PEA_MAX_OFFSETS) return PEA_E_OFFSETS;
    if (semantics != 0 && semantics != 1) return PEA_E_MODE;
    if (counts_bytes < pea_gt_counts_bytes(B, K)) return PEA_E_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(counts) % 8) return PEA_E_ALIGN;
    // four voxels per thread with 128-bit stores and 32-bit in-sample indices
    if (W % 4 != 0 || (int64_t)K * D * H * W >= ((int64_t)1 << 31)) return PEA_E_SHAPE;
    if ((reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(weight)) & 15) return PEA_E_ALIGN;
    if (mask && (reinterpret_cast<uintptr_t>(mask) & 3)) return PEA_E_ALIGN;
    GtTable tab;
    tab.K = K;
    for (int k = 0; k < K; ++k) { tab.dz[k] = offsets[3 * k]; tab.dy[k] = offsets[3 * k + 1]; tab.dx[k] = offsets[3 * k + 2]; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto* c = static_cast<unsigned long long*>(counts);
#define PEA_GT_CASE(CODE, TYPE)                                                                                        \
    case CODE:                                                                                                         \
        return semantics == 0 ? run_gt_fused<TYPE, 0>(labels, B, D, H, W, tab, ignore, padding, alpha, target, mask,  \
                                                      weight, c, st)                                                   \
                              : run_gt_fused<TYPE, 1>(labels, B, D, H, W, tab, ignore, padding, alpha, target, mask,  \
                                                      weight, c, st);
    switch (label_dtype) {
        PEA_GT_CASE(PEA_LABEL_U8, uint8_t)
        PEA_GT_CASE(PEA_LABEL_U16, uint16_t)
        PEA_GT_CASE(PEA_LABEL_U32, uint32_t)
        PEA_GT_CASE(PEA_LABEL_U64, uint64_t)
        PEA_GT_CASE(PEA_LABEL_S8, int8_t)
        PEA_GT_CASE(PEA_LABEL_S16, int16_t)
        PEA_GT_CASE(PEA_LABEL_S32, int32_t)
        PEA_GT_CASE(PEA_LABEL_S64, int64_t)
        PEA_GT_CASE(PEA_LABEL_F32, float)
        default: return PEA_E_DTYPE;
    }
#undef PEA_GT_CASE
}

extern "C" int pea_gt_pair_weights(const void* labels, int label_dtype, int B, int D, int H, int W, const int32_t* offsets,
                                   int K, int semantics, int ignore, int padding, double alpha, float* wtab, void* counts,
                                   size_t counts_bytes, void* stream) {
    if (!labels || !offsets || !wtab || !counts) return PEA_E_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || B > 65535) return PEA_E_SHAPE;
    if (K < 1 || K > PEA_MAX_OFFSETS) return PEA_E_OFFSETS;
    if (semantics != 0 && semantics != 1) return PEA_E_MODE;
    if (counts_bytes < pea_gt_counts_bytes(B, K)) return PEA_E_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(counts) % 8) return PEA_E_ALIGN;
    if (W % 4 != 0 || (int64_t)K * D * H * W >= ((int64_t)1 << 31)) return PEA_E_SHAPE;
    GtTable tab;
    tab.K = K;
    for (int k = 0; k < K; ++k) { tab.dz[k] = offsets[3 * k]; tab.dy[k] = offsets[3 * k + 1]; tab.dx[k] = offsets[3 * k + 2]; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto* c = static_cast<unsigned long long*>(counts);
#define PEA_GW_CASE(CODE, TYPE)                                                                                        \
    case CODE:                                                                                                         \
        return semantics == 0 ? run_gt_weights<TYPE, 0>(labels, B, D, H, W, tab, ignore, padding, alpha, wtab, c, st)  \
                              : run_gt_weights<TYPE, 1>(labels, B, D, H, W, tab, ignore, padding, alpha, wtab, c, st);
    switch (label_dtype) {
        PEA_GW_CASE(PEA_LABEL_U8, uint8_t)
        PEA_GW_CASE(PEA_LABEL_U16, uint16_t)
        PEA_GW_CASE(PEA_LABEL_U32, uint32_t)
        PEA_GW_CASE(PEA_LABEL_U64, uint64_t)
        PEA_GW_CASE(PEA_LABEL_S8, int8_t)
        PEA_GW_CASE(PEA_LABEL_S16, int16_t)
        PEA_GW_CASE(PEA_LABEL_S32, int32_t)
        PEA_GW_CASE(PEA_LABEL_S64, int64_t)
        PEA_GW_CASE(PEA_LABEL_F32, float)
        default: return PEA_E_DTYPE;
    }
#undef PEA_GW_CASE
}
