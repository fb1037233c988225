// One launch for SEVERAL small fused losses: the deep-supervision scales of one training step
//   scripts_ac3ac4/main.py:227-230   embedding_loss_norm1 at 18 x {10, 20, 40, 80}^2
//   scripts_cvppp/main.py:284-287    embedding_loss at 34^2 .. 272^2 with the first 2 / 4 / 6 / 8 offsets
// Each of those calls is a few hundred kB to a few MB -- microseconds of HBM time -- and used to cost three launches
// (1/|x| plane, table-driven fused kernel, loss finalize) plus the gradient-scale launch in backward.  Here a work list
// of up to 8 problems runs through ONE grid: a warp takes a 32-voxel chunk of whichever problem the chunk index falls
// into, evaluates the gather-form fused loss + gradient with the neighbours' 1/|x| recomputed in registers (the
// volumes live in L1 / L2), and the last CTA to arrive (ticket) reduces the per-CTA partials of every problem in a
// fixed order and writes the loss scalars.  Same per-pair arithmetic as pea_affs_generic.cu (reference semantics:
// any offsets, both border rules, optional mask, optional second view whose gradient is not requested).
#include "pea_common.cuh"

namespace pea {
namespace multi {

constexpr int kMaxProblems = PEA_MULTI_MAX_PROBLEMS;
constexpr int kMaxK = PEA_MULTI_MAX_OFFSETS;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct Problem {
    const float* emb;
    const float* par;  // second view (read only) or emb itself
    const float* target;
    const float* weight;
    const void* mask;
    float* affs;   // may be null
    float* grad;   // may be null
    float* loss_out;
    int B, D, H, W, K, wrap, mask_dtype, cross;
    int upstream;     // != 0: `target` holds d loss / d affs (K5, pea_affs_bwd): gradient only, no loss
    int chunk_begin;  // first 32-voxel chunk of this problem in the launch-wide chunk numbering
    signed char dz[kMaxK], dy[kMaxK], dx[kMaxK];
    float gcoef[kMaxK], norm[kMaxK], coef[kMaxK];
};

struct Batch {
    int n, total_chunks;
    float* partials;  // [gridDim.x][n][kMaxK]
    uint32_t* ticket;
    Problem p[kMaxProblems];
};

__device__ __forceinline__ float ld_cg(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// 1 / max(sqrt(ss), 1e-12) == min(rsqrt(ss), 1e12) (one MUFU.RSQ, 2 ulp; ss == 0 -> +inf -> 1e12): the tiled kernels' form
__device__ __forceinline__ float inv_norm(float ss) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(ss));
    return fminf(r, 1e12f);
}

// neighbour under the border rule in 32-bit arithmetic; an invalid neighbour yields the voxel itself (a safe address)
__device__ __forceinline__ bool nbr(int z, int y, int x, int dz, int dy, int dx, int D, int H, int W, bool wrap, int p, int& q) {
    int qz = z + dz, qy = y + dy, qx = x + dx;
    bool ok = true;
    if (wrap) {
        qz = pmod(qz, D);
        qy = pmod(qy, H);
        qx = pmod(qx, W);
    } else {
        ok = (unsigned)qz < (unsigned)D && (unsigned)qy < (unsigned)H && (unsigned)qx < (unsigned)W;
    }
    q = ok ? (qz * H + qy) * W + qx : p;
    return ok;
}

template <int C>
__global__ void __launch_bounds__(kThreads, 2) multi_fused(const __grid_constant__ Batch bt) {
    __shared__ float s_part[kWarps][kMaxProblems][kMaxK];
    __shared__ float s_loss[kMaxProblems][kMaxK];
    __shared__ double s_acc[kWarps][kMaxProblems * kMaxK];
    __shared__ uint32_t s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kWarps * kMaxProblems * kMaxK; i += kThreads) (&s_part[0][0][0])[i] = 0.f;
    __syncthreads();

    for (int chunk = blockIdx.x * kWarps + warp; chunk < bt.total_chunks; chunk += gridDim.x * kWarps) {
        int pi = 0;
#pragma unroll
        for (int i = 1; i < kMaxProblems; ++i)
            if (i < bt.n && chunk >= bt.p[i].chunk_begin) pi = i;
        const Problem& pr = bt.p[pi];
        // a problem holds < 2^31 voxels (checked by the host): 32-bit index arithmetic throughout
        const int D = pr.D, H = pr.H, W = pr.W, K = pr.K;
        const int N = D * H * W, total = pr.B * N;
        const int i0 = (chunk - pr.chunk_begin) * 32 + lane;
        const bool live = i0 < total;
        const int ii = live ? i0 : total - 1;
        const int b = ii / N, p = ii - b * N;
        const int x = p % W, yz = p / W, y = yz % H, z = yz / H;
        const bool wrap = pr.wrap != 0;
        const bool want_grad = pr.grad != nullptr, rev = want_grad && !pr.cross;
        const float* embB = pr.emb + (int64_t)b * C * N;
        const float* parB = pr.par + (int64_t)b * C * N;
        const int64_t kbase = (int64_t)b * K * N;

        float e[C], G[C];
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            e[c] = __ldg(embB + (int64_t)c * N + p);
            ss = fmaf(e[c], e[c], ss);
        }
        const float rp = inv_norm(ss);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            e[c] *= rp;
            G[c] = 0.f;
        }

        for (int k = 0; k < K; ++k) {
            // the pair anchored at p (partner q) and the pair anchored at q2 = p - o_k whose partner is p: both raw
            // vectors and both (target, weight, mask) triples are requested before anything is consumed
            int q, q2;
            const bool vf = nbr(z, y, x, pr.dz[k], pr.dy[k], pr.dx[k], D, H, W, wrap, p, q);
            const bool vr = rev && nbr(z, y, x, -pr.dz[k], -pr.dy[k], -pr.dx[k], D, H, W, wrap, p, q2);
            if (!rev) q2 = p;
            const int64_t idx = kbase + (int64_t)k * N + p, idx2 = kbase + (int64_t)k * N + q2;
            float xq[C], xr[C];
#pragma unroll
            for (int c = 0; c < C; ++c) xq[c] = __ldg(parB + (int64_t)c * N + q);
            if (rev) {
#pragma unroll
                for (int c = 0; c < C; ++c) xr[c] = __ldg(embB + (int64_t)c * N + q2);
            }
            if (pr.upstream) {
                // K5: the gradient arrives through the affinities; a pair the border rule drops has a constant affinity
                const float u = vf ? __ldg(pr.target + idx) : 0.f, u2 = vr ? __ldg(pr.target + idx2) : 0.f;
                float sq = 0.f, sr = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) sq = fmaf(xq[c], xq[c], sq);
                const float g = u * inv_norm(sq);
#pragma unroll
                for (int c = 0; c < C; ++c) G[c] = fmaf(g, xq[c], G[c]);
                if (rev) {
#pragma unroll
                    for (int c = 0; c < C; ++c) sr = fmaf(xr[c], xr[c], sr);
                    const float g2 = u2 * inv_norm(sr);
#pragma unroll
                    for (int c = 0; c < C; ++c) G[c] = fmaf(g2, xr[c], G[c]);
                }
                continue;
            }
            const float t = __ldg(pr.target + idx), w = vf ? __ldg(pr.weight + idx) : 0.f, m = load_mask(pr.mask, pr.mask_dtype, idx);
            float t2 = 0.f, w2 = 0.f, m2 = 0.f;
            if (rev) {
                t2 = __ldg(pr.target + idx2);
                w2 = vr ? __ldg(pr.weight + idx2) : 0.f;
                m2 = load_mask(pr.mask, pr.mask_dtype, idx2);
            }
            float sq = 0.f, dq = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                sq = fmaf(xq[c], xq[c], sq);
                dq = fmaf(e[c], xq[c], dq);
            }
            const float rq = inv_norm(sq);
            const float a = vf ? dq * rq : 0.f;
            const float df = a * m - t * m;
            float s = w * df * df;
            if (pr.affs != nullptr && live) pr.affs[idx] = a;
            s = warp_sum(live ? s : 0.f);
            if (lane == 0) s_part[warp][pi][k] += s;
            if (want_grad) {
                const float g = pr.gcoef[k] * w * m * m * (a - t) * rq;
#pragma unroll
                for (int c = 0; c < C; ++c) G[c] = fmaf(g, xq[c], G[c]);
            }
            if (rev) {
                float sr = 0.f, dr = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    sr = fmaf(xr[c], xr[c], sr);
                    dr = fmaf(e[c], xr[c], dr);
                }
                const float rr = inv_norm(sr);
                const float g2 = pr.gcoef[k] * w2 * m2 * m2 * (dr * rr - t2) * rr;
#pragma unroll
                for (int c = 0; c < C; ++c) G[c] = fmaf(g2, xr[c], G[c]);
            }
        }
        if (live && want_grad) {  // back through the normalisation: (G - e^ <e^, G>) / |x|   (eps branch: G / eps)
            float eg = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) eg = fmaf(e[c], G[c], eg);
            if (!(rp < 1e12f)) eg = 0.f;
            float* dst = pr.grad + (int64_t)b * C * N + p;
#pragma unroll
            for (int c = 0; c < C; ++c) dst[(int64_t)c * N] = (G[c] - e[c] * eg) * rp;
        }
    }

    if (bt.ticket == nullptr) return;  // gradient-only launch (upstream mode): no loss to finalize
    // ---- per-CTA partials (fixed order over warps), then the last CTA to arrive finalizes every problem
    __syncthreads();
    const int slots = bt.n * kMaxK;
    for (int i = threadIdx.x; i < slots; i += kThreads) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) v += (&s_part[w][0][0])[i];
        bt.partials[(int64_t)blockIdx.x * slots + i] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(bt.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // rows of partials are dealt out to the warps (row = warp, warp + 8, ...), a lane owns slot lane + 32 j; eight loads
    // in flight per lane; double accumulation in a fixed order, then the eight warp sums in warp order: reproducible
    for (int base = 0; base < slots; base += 32) {
        const int i = base + lane;
        double acc = 0.0;
        if (i < slots) {
            const float* src = bt.partials + i;
            int c = warp;
            for (; c + 7 * kWarps < (int)gridDim.x; c += 8 * kWarps) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = ld_cg(src + (int64_t)(c + u * kWarps) * slots);
#pragma unroll
                for (int u = 0; u < 8; ++u) acc += (double)v[u];
            }
            for (; c < (int)gridDim.x; c += kWarps) acc += (double)ld_cg(src + (int64_t)c * slots);
            s_acc[warp][i] = acc;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < slots; i += kThreads) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) acc += s_acc[w][i];
        const int pi = i / kMaxK, k = i - pi * kMaxK;
        s_loss[pi][k] = (k < bt.p[pi].K) ? (float)acc / bt.p[pi].norm[k] : 0.f;
    }
    __syncthreads();
    if (threadIdx.x < bt.n) {
        const Problem& pr = bt.p[threadIdx.x];
        float tot = 0.f;
        for (int k = 0; k < pr.K; ++k) {
            pr.loss_out[1 + k] = s_loss[threadIdx.x][k];
            tot += s_loss[threadIdx.x][k] * pr.coef[k];
        }
        pr.loss_out[0] = tot;
    }
    if (threadIdx.x == 0) *bt.ticket = 0u;
}

struct ScaleBatch {
    int n;
    float* buf[kMaxProblems];
    int64_t len[kMaxProblems];
    const float* scale[kMaxProblems];
};

// buf_i *= *scale_i for every buffer of the list (the upstream gradients of the losses of one multi launch); a buffer
// whose scalar is exactly 1 is not touched
__global__ void __launch_bounds__(256) multi_scale(const __grid_constant__ ScaleBatch sb) {
    for (int j = 0; j < sb.n; ++j) {
        const float s = __ldg(sb.scale[j]);
        if (s == 1.0f) continue;
        float* buf = sb.buf[j];
        const int64_t n = sb.len[j];
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) buf[i] *= s;
    }
}

}  // namespace multi
}  // namespace pea

using namespace pea;

extern "C" size_t pea_affs_loss_multi_workspace(const pea_loss_problem* problems, int n) {
    if (!problems || n < 1 || n > PEA_MULTI_MAX_PROBLEMS) return 0;
    return (size_t)148 * 8 * (size_t)n * PEA_MULTI_MAX_OFFSETS * sizeof(float);
}

extern "C" int pea_affs_loss_multi(const pea_loss_problem* problems, int n, int C, void* workspace, size_t ws_bytes,
                                   uint32_t* ticket, void* stream) {
    if (!problems || !workspace || !ticket) return PEA_E_NULL;
    if (n < 1 || n > PEA_MULTI_MAX_PROBLEMS) return PEA_E_SHAPE;
    if (C != 4 && C != 8 && C != 16 && C != 32) return PEA_E_CHANNELS;
    if ((reinterpret_cast<uintptr_t>(workspace) % 16) || (reinterpret_cast<uintptr_t>(ticket) % 4)) return PEA_E_ALIGN;
    multi::Batch bt;
    bt.n = n;
    int64_t chunks = 0;
    for (int i = 0; i < n; ++i) {
        const pea_loss_problem& s = problems[i];
        multi::Problem& p = bt.p[i];
        if (!s.emb || !s.target || !s.weight || !s.loss_out || !s.offsets || !s.loss_norm || !s.loss_coef) return PEA_E_NULL;
        if (s.B <= 0 || s.D <= 0 || s.H <= 0 || s.W <= 0) return PEA_E_SHAPE;
        if (s.K < 1 || s.K > PEA_MULTI_MAX_OFFSETS) return PEA_E_OFFSETS;
        if (s.border != PEA_BORDER_WRAP && s.border != PEA_BORDER_ZERO) return PEA_E_MODE;
        if (s.mask_dtype != PEA_MASK_NONE && s.mask_dtype != PEA_MASK_U8 && s.mask_dtype != PEA_MASK_F32) return PEA_E_DTYPE;
        if (s.mask_dtype != PEA_MASK_NONE && !s.mask) return PEA_E_NULL;
        p.emb = s.emb;
        p.par = s.partner ? s.partner : s.emb;
        p.cross = s.partner ? 1 : 0;
        p.target = s.target; p.weight = s.weight; p.mask = s.mask; p.affs = s.affs; p.grad = s.grad_emb; p.loss_out = s.loss_out;
        p.B = s.B; p.D = s.D; p.H = s.H; p.W = s.W; p.K = s.K;
        p.wrap = s.border == PEA_BORDER_WRAP;
        p.mask_dtype = s.mask_dtype;
        p.upstream = 0;
        for (int k = 0; k < s.K; ++k) {
            const int32_t oz = s.offsets[3 * k], oy = s.offsets[3 * k + 1], ox = s.offsets[3 * k + 2];
            if (oz < -127 || oz > 127 || oy < -127 || oy > 127 || ox < -127 || ox > 127) return PEA_E_OFFSETS;
            p.dz[k] = (signed char)oz; p.dy[k] = (signed char)oy; p.dx[k] = (signed char)ox;
            p.norm[k] = s.loss_norm[k];
            p.coef[k] = s.loss_coef[k];
            p.gcoef[k] = (float)(2.0 * (double)s.loss_coef[k] / (double)s.loss_norm[k]);
        }
        const int64_t total = (int64_t)s.B * s.D * s.H * s.W;
        if (total >= ((int64_t)1 << 31) - 64) return PEA_E_SHAPE;  // this entry point is for SMALL problems (int32 voxel indices)
        p.chunk_begin = (int)chunks;
        chunks += (total + 31) / 32;
        if (chunks > ((int64_t)1 << 30)) return PEA_E_SHAPE;
    }
    bt.total_chunks = (int)chunks;
    int64_t g = (chunks + multi::kWarps - 1) / multi::kWarps;
    int dev = 0, nsm = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int cap = nsm > 0 && nsm <= 592 ? 2 * nsm : 296;  // two resident CTAs per SM walk the chunk list
    const int grid = (int)(g < 1 ? 1 : (g > cap ? cap : g));
    if ((size_t)grid * (size_t)n * PEA_MULTI_MAX_OFFSETS * sizeof(float) > ws_bytes) return PEA_E_WORKSPACE;
    // (2^31 voxels per problem bound above keeps every in-kernel voxel index inside int32)
    bt.partials = static_cast<float*>(workspace);
    bt.ticket = ticket;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (C) {
        case 4: multi::multi_fused<4><<<grid, multi::kThreads, 0, st>>>(bt); break;
        case 8: multi::multi_fused<8><<<grid, multi::kThreads, 0, st>>>(bt); break;
        case 16: multi::multi_fused<16><<<grid, multi::kThreads, 0, st>>>(bt); break;
        default: multi::multi_fused<32><<<grid, multi::kThreads, 0, st>>>(bt); break;
    }
    return (int)cudaGetLastError();
}

// K5 through the work-list kernel: grad_emb = d loss / d emb for the upstream gradient grad_affs of the affinities of ONE
// volume (pea_affs_bwd); gradient w.r.t. the anchor view only.  Returns PEA_E_MODE when the problem does not fit the
// kernel's limits (the caller then takes the table-driven kernel).
namespace pea {
int launch_multi_upstream(const float* emb, const float* partner, int B, int C, int D, int H, int W, const int32_t* offsets, int K,
                          int border, const float* grad_affs, float* grad_emb, cudaStream_t st) {
    if (K < 1 || K > PEA_MULTI_MAX_OFFSETS) return PEA_E_MODE;
    if (C != 4 && C != 8 && C != 16 && C != 32) return PEA_E_MODE;
    const int64_t total = (int64_t)B * D * H * W;
    if (total >= ((int64_t)1 << 31) - 64) return PEA_E_MODE;
    multi::Batch bt;
    bt.n = 1;
    multi::Problem& p = bt.p[0];
    p.emb = emb; p.par = partner ? partner : emb; p.cross = partner ? 1 : 0;
    p.target = grad_affs; p.weight = nullptr; p.mask = nullptr; p.affs = nullptr; p.grad = grad_emb; p.loss_out = nullptr;
    p.B = B; p.D = D; p.H = H; p.W = W; p.K = K; p.wrap = border == PEA_BORDER_WRAP; p.mask_dtype = PEA_MASK_NONE;
    p.upstream = 1;
    for (int k = 0; k < K; ++k) {
        const int32_t oz = offsets[3 * k], oy = offsets[3 * k + 1], ox = offsets[3 * k + 2];
        if (oz < -127 || oz > 127 || oy < -127 || oy > 127 || ox < -127 || ox > 127) return PEA_E_MODE;
        p.dz[k] = (signed char)oz; p.dy[k] = (signed char)oy; p.dx[k] = (signed char)ox;
        p.norm[k] = 1.f; p.coef[k] = 1.f; p.gcoef[k] = 1.f;
    }
    p.chunk_begin = 0;
    bt.total_chunks = (int)((total + 31) / 32);
    bt.partials = nullptr;
    bt.ticket = nullptr;
    int64_t g = (bt.total_chunks + multi::kWarps - 1) / multi::kWarps;
    int dev = 0, nsm = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const int cap = nsm > 0 && nsm <= 592 ? 2 * nsm : 296;
    const int grid = (int)(g < 1 ? 1 : (g > cap ? cap : g));
    switch (C) {
        case 4: multi::multi_fused<4><<<grid, multi::kThreads, 0, st>>>(bt); break;
        case 8: multi::multi_fused<8><<<grid, multi::kThreads, 0, st>>>(bt); break;
        case 16: multi::multi_fused<16><<<grid, multi::kThreads, 0, st>>>(bt); break;
        default: multi::multi_fused<32><<<grid, multi::kThreads, 0, st>>>(bt); break;
    }
    return (int)cudaGetLastError();
}
}  // namespace pea

extern "C" int pea_scale_multi(float* const* bufs, const int64_t* lens, const float* const* scales, int n, void* stream) {
    if (!bufs || !lens || !scales) return PEA_E_NULL;
    if (n < 1 || n > PEA_MULTI_MAX_PROBLEMS) return PEA_E_SHAPE;
    multi::ScaleBatch sb;
    sb.n = n;
    int64_t longest = 0;
    for (int i = 0; i < n; ++i) {
        if (!bufs[i] || !scales[i]) return PEA_E_NULL;
        if (lens[i] <= 0) return PEA_E_SHAPE;
        sb.buf[i] = bufs[i];
        sb.len[i] = lens[i];
        sb.scale[i] = scales[i];
        longest = lens[i] > longest ? lens[i] : longest;
    }
    int64_t g = (longest + 255) / 256;
    const int grid = (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
    multi::multi_scale<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(sb);
    return (int)cudaGetLastError();
}
