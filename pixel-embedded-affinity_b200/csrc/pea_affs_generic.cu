// Table-driven gather kernels: any offset list (diagonals, either sign, |o| >= extent), both border rules,
// self and cross (two-view) variants.  One thread per voxel, channel-planar coalesced loads, neighbours
// served from L1/L2.  This is the reference-semantics workhorse that the tiled kernels (pea_affs_tiled.cu)
// are checked against and fall back to for offset sets they do not specialise.
//
// Math (reference: scripts_cvppp/loss/loss_embedding_mse.py:7-95,
// scripts_ac3ac4/loss/loss_embedding_mse.py:143-194):
//   e^ = x / max(|x|, 1e-12)                      a_k[p] = <e^[p], u^[p+o_k]>
//   S_k = sum w (a m - t m)^2                     g_k[p] = gcoef_k * w m^2 (a - t)
//   G[p] = sum_k g_k[p] u^[p+o_k]  (+ g_k[p-o_k] e^[p-o_k] when u^ == e^)
//   dL/dx[p] = (G - e^ <e^, G>) / |x|
#include "pea_common.cuh"

namespace pea {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 148 * 8;

// ------------------------------------------------------------------------------------------------
// 1 / max(|x|_2, eps) per voxel
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads) rinv_kernel(const float* __restrict__ emb, Dims d, float* __restrict__ rinv) {
    const int64_t total = (int64_t)d.B * d.N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / d.N, p = i - b * d.N;
        const float* src = emb + (b * C) * d.N + p;
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float v = __ldg(src + (int64_t)c * d.N);
            ss = fmaf(v, v, ss);
        }
        rinv[i] = 1.0f / fmaxf(sqrtf(ss), kNormEps);
    }
}

template <int C>
__device__ __forceinline__ void load_unit(const float* __restrict__ base, const float* __restrict__ rinv, int64_t b,
                                          int64_t p, int64_t N, float (&v)[C]) {
    const float r = __ldg(rinv + b * N + p);
    const float* src = base + (b * C) * N + p;
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = __ldg(src + (int64_t)c * N) * r;
}

template <int C>
__device__ __forceinline__ float dot(const float (&a)[C], const float (&b)[C]) {
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s = fmaf(a[c], b[c], s);
    return s;
}

// ------------------------------------------------------------------------------------------------
// forward only
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads)
fwd_generic(const float* __restrict__ emb, const float* __restrict__ par, const float* __restrict__ rinvE,
            const float* __restrict__ rinvU, Dims d, OffsetTable tab, int wrap, int z_skip, int64_t affs_plane,
            float* __restrict__ affs, int relu) {
    const int64_t total = (int64_t)d.B * d.N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / d.N, p = i - b * d.N;
        const int x = (int)(p % d.W);
        const int y = (int)((p / d.W) % d.H);
        const int z = (int)(p / ((int64_t)d.W * d.H));
        if (z < z_skip) continue;  // halo slices of the previous z-shard: neighbours only
        float e[C];
        load_unit<C>(emb, rinvE, b, p, d.N, e);
        const int64_t pout = p - (int64_t)z_skip * d.W * d.H;
        for (int k = 0; k < tab.K; ++k) {
            int64_t q;
            float a = 0.f;
            if (neighbour(z, y, x, tab.dz[k], tab.dy[k], tab.dx[k], d, wrap != 0, q)) {
                float un[C];
                load_unit<C>(par, rinvU, b, q, d.N, un);
                a = dot<C>(e, un);
            }
            affs[(b * tab.K + k) * affs_plane + pout] = relu ? fmaxf(a, 0.f) : a;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fused loss forward + backward, and (upstream != NULL) the generic backward
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(kThreads)
fused_generic(const float* __restrict__ emb, const float* __restrict__ par, const float* __restrict__ rinvE,
              const float* __restrict__ rinvU, Dims d, OffsetTable tab, int wrap, int cross,
              const float* __restrict__ target, const float* __restrict__ weight, const void* __restrict__ mask,
              int mask_dtype, const float* __restrict__ upstream, float* __restrict__ affs,
              float* __restrict__ partials, float* __restrict__ grad_emb, float* __restrict__ grad_par) {
    __shared__ float s_part[PEA_MAX_OFFSETS][kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < PEA_MAX_OFFSETS * (kThreads / 32); k += kThreads) (&s_part[0][0])[k] = 0.f;
    __syncthreads();

    const bool want_rev_self = !cross && grad_emb != nullptr;
    const bool want_rev_par = cross && grad_par != nullptr;
    const int64_t total = (int64_t)d.B * d.N;
    const int64_t span = ((total + 31) / 32) * 32;  // keep whole warps in the loop for the shuffles
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < span; i += (int64_t)gridDim.x * blockDim.x) {
        const bool live = i < total;
        const int64_t ii = live ? i : total - 1;
        const int64_t b = ii / d.N, p = ii - b * d.N;
        const int x = (int)(p % d.W);
        const int y = (int)((p / d.W) % d.H);
        const int z = (int)(p / ((int64_t)d.W * d.H));

        float e[C], u[C], G[C], GU[C];
        load_unit<C>(emb, rinvE, b, p, d.N, e);
        if (want_rev_par) load_unit<C>(par, rinvU, b, p, d.N, u);
#pragma unroll
        for (int c = 0; c < C; ++c) G[c] = GU[c] = 0.f;

        for (int k = 0; k < tab.K; ++k) {
            const int64_t idx = (b * tab.K + k) * d.N + p;
            // ---- pair anchored at p, partner at p + o_k
            int64_t q;
            const bool valid = neighbour(z, y, x, tab.dz[k], tab.dy[k], tab.dx[k], d, wrap != 0, q);
            float a = 0.f, g = 0.f, s = 0.f;
            float un[C];
            if (valid) {
                load_unit<C>(par, rinvU, b, q, d.N, un);
                a = dot<C>(e, un);
            }
            if (upstream != nullptr) {
                if (valid) g = __ldg(upstream + idx);
            } else {
                if (affs != nullptr && live) affs[idx] = a;
                if (valid) {
                    const float t = __ldg(target + idx), w = __ldg(weight + idx), m = load_mask(mask, mask_dtype, idx);
                    const float dd = a * m - t * m;
                    s = w * dd * dd;
                    g = tab.gcoef[k] * w * m * m * (a - t);
                }
                s = warp_sum(live ? s : 0.f);
                if (lane == 0) s_part[k][warp] += s;
            }
            if (valid && grad_emb != nullptr) {
#pragma unroll
                for (int c = 0; c < C; ++c) G[c] = fmaf(g, un[c], G[c]);
            }
            // ---- pair anchored at p - o_k, whose partner is p
            if (want_rev_self || want_rev_par) {
                int64_t q2;
                if (neighbour(z, y, x, -tab.dz[k], -tab.dy[k], -tab.dx[k], d, wrap != 0, q2)) {
                    float en[C];
                    load_unit<C>(emb, rinvE, b, q2, d.N, en);
                    const int64_t idx2 = (b * tab.K + k) * d.N + q2;
                    float g2;
                    if (upstream != nullptr) {
                        g2 = __ldg(upstream + idx2);
                    } else {
                        const float a2 = want_rev_par ? dot<C>(en, u) : dot<C>(en, e);
                        const float t2 = __ldg(target + idx2), w2 = __ldg(weight + idx2),
                                    m2 = load_mask(mask, mask_dtype, idx2);
                        g2 = tab.gcoef[k] * w2 * m2 * m2 * (a2 - t2);
                    }
                    if (want_rev_par) {
#pragma unroll
                        for (int c = 0; c < C; ++c) GU[c] = fmaf(g2, en[c], GU[c]);
                    } else {
#pragma unroll
                        for (int c = 0; c < C; ++c) G[c] = fmaf(g2, en[c], G[c]);
                    }
                }
            }
        }
        // ---- back through the normalisation: (G - e^ <e^,G>) / |x|   (eps branch: G / eps)
        if (live && grad_emb != nullptr) {
            const float r = __ldg(rinvE + b * d.N + p);
            const float eg = (r < 1e12f) ? dot<C>(e, G) : 0.f;
            float* dst = grad_emb + (b * C) * d.N + p;
#pragma unroll
            for (int c = 0; c < C; ++c) dst[(int64_t)c * d.N] = (G[c] - e[c] * eg) * r;
        }
        if (live && want_rev_par) {
            const float r = __ldg(rinvU + b * d.N + p);
            const float ug = (r < 1e12f) ? dot<C>(u, GU) : 0.f;
            float* dst = grad_par + (b * C) * d.N + p;
#pragma unroll
            for (int c = 0; c < C; ++c) dst[(int64_t)c * d.N] = (GU[c] - u[c] * ug) * r;
        }
    }
    if (partials != nullptr) {
        __syncthreads();
        for (int k = threadIdx.x; k < tab.K; k += kThreads) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) s += s_part[k][w];
            partials[(int64_t)blockIdx.x * tab.K + k] = s;
        }
    }
}

}  // namespace pea

// ------------------------------------------------------------------------------------------------
// host side launchers (used by pea_abi.cu)
// ------------------------------------------------------------------------------------------------
namespace pea {

static int grid_for(int64_t total) {
    int64_t g = (total + kThreads - 1) / kThreads;
    return (int)(g < 1 ? 1 : (g > kMaxBlocks ? kMaxBlocks : g));
}

int generic_partial_blocks(int64_t total) { return grid_for(((total + 31) / 32) * 32); }

#define PEA_DISPATCH_C(C_, ...)                    \
    switch (C_) {                                  \
        case 4: { constexpr int CC = 4; __VA_ARGS__; } break;   \
        case 8: { constexpr int CC = 8; __VA_ARGS__; } break;   \
        case 16: { constexpr int CC = 16; __VA_ARGS__; } break; \
        case 32: { constexpr int CC = 32; __VA_ARGS__; } break; \
        default: return PEA_E_CHANNELS;            \
    }

int launch_rinv(const float* emb, int C, const Dims& d, float* rinv, cudaStream_t st) {
    const int grid = grid_for((int64_t)d.B * d.N);
    PEA_DISPATCH_C(C, rinv_kernel<CC><<<grid, kThreads, 0, st>>>(emb, d, rinv));
    return (int)cudaGetLastError();
}

int launch_fwd_generic(const float* emb, const float* par, const float* rinvE, const float* rinvU, int C, const Dims& d,
                       const OffsetTable& tab, int wrap, int z_skip, int64_t affs_plane, float* affs, int relu,
                       cudaStream_t st) {
    const int grid = grid_for((int64_t)d.B * d.N);
    PEA_DISPATCH_C(C, fwd_generic<CC><<<grid, kThreads, 0, st>>>(emb, par, rinvE, rinvU, d, tab, wrap, z_skip, affs_plane,
                                                                 affs, relu));
    return (int)cudaGetLastError();
}

int launch_fused_generic(const float* emb, const float* par, const float* rinvE, const float* rinvU, int C,
                         const Dims& d, const OffsetTable& tab, int wrap, int cross, const float* target,
                         const float* weight, const void* mask, int mask_dtype, const float* upstream, float* affs,
                         float* partials, float* grad_emb, float* grad_par, cudaStream_t st) {
    const int grid = generic_partial_blocks((int64_t)d.B * d.N);
    PEA_DISPATCH_C(C, fused_generic<CC><<<grid, kThreads, 0, st>>>(emb, par, rinvE, rinvU, d, tab, wrap, cross, target,
                                                                   weight, mask, mask_dtype, upstream, affs, partials,
                                                                   grad_emb, grad_par));
    return (int)cudaGetLastError();
}

}  // namespace pea
