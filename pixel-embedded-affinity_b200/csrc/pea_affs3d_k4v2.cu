// K4 v2: fused 3-D kernel for the AC3/AC4 12-offset set (embedding_loss_norm5,
// scripts_ac3ac4/loss/loss_embedding_mse.py:143-194): embedding -> 12 affinities -> weighted MSE -> d loss / d embedding
// and the final loss scalars, ONE launch, ONE pass over HBM.
//
// Same march as fused3d_norm5_tm (pea_affs3d_tiled.cu: a CTA owns TY x 32 voxel columns and walks z, z rings in Tensor
// Memory, raw neighbour vectors staged per slice by TMA).  What changed against it, each from a measurement in
// profiles/ (round 1 digest: math-pipe throttle + not-selected + wait are the top stalls, 1640 instructions per column
// and slice, active 260 k of 291 k elapsed cycles):
//   * LANE-packed in-plane pairs.  The pair anchored at p (partner p - s) and the pair anchored at p + s (partner p,
//     gather-form backward) share one FFMA2 per channel: X_c = (x[p - s][c], x[p + s][c]).  No horizontal sums, the
//     scalar tail (affinity, difference, weight, gradient factor) runs packed as well.  Every shared-memory operand is
//     base register + immediate: the target / weight tiles of the "+ s" pairs sit in the SAME TMA box as the tile's own
//     (y channels: 2 TY rows, x channels: 64 columns), so there is no run-time choice of a buffer.
//   * BALANCED tiles.  12-row tiles give 280 CTAs on 148 SMs: 132 SMs run two CTAs, 16 run one, and everything waits
//     for 24 rows x D slices.  Here the number of CTAs is rounded up to a multiple of the SM count and the rows of
//     every strip are dealt out evenly (cfg3: 296 CTAs of 10-12 rows, taller tiles first so that the dynamic scheduler
//     pairs them with the short ones): 22-23 rows per SM.  Warps beyond the tile's height leave the march at once.
//   * The loss is FINALIZED IN THE KERNEL by the last CTA to arrive (threadfence + ticket, partials read back in a
//     fixed order, double accumulation: bit-reproducible), which removes the separate finalize launch.
#include "pea_k4_common.cuh"

namespace pea {
namespace t3 {

template <int TY>
struct alignas(128) StageV2 {
    float sY[kC][TY + 2 * HY][TX];  // rows y0-9 .. y0+TY+8
    float sAm[kC][TY][TX];          // rows y0-27 ..
    float sAp[kC][TY][TX];          // rows y0+27 ..
    float sX[kC][TY][XW];           // cols x0-28 .. x0+59
    float sTWy[2][4][2 * TY][TX];   // [target, weight][channel 1, 4, 7, 10][rows y0 .. y0+2TY-1]
    float sTWx[2][4][TY][2 * TX];   // [target, weight][channel 2, 5, 8, 11][cols x0 .. x0+63]
    float sFar[2][TY][TX];          // channel 10 (y shift 27), rows y0+27 ..: [target, weight]
};

struct ParamsV2 {
    int B, D, H, W;
    int strips_x;  // 32-column strips per row
    int mA, nA;    // the first nA strips (b-major, then x) hold mA tiles each, the others mA + 1
    const float* target;
    const float* weight;
    float* affs;      // may be null
    float* grad;      // may be null
    float* partials;  // [gridDim.x][12]
    float gcoef[12];
    // in-kernel finalize (ticket != null): loss_out[0] = sum_k coef_k * S_k / norm_k, loss_out[1 + k] = S_k / norm_k
    uint32_t* ticket;  // zero on entry, left zero
    float* loss_out;
    float norm[12], coef[12];
};

// Forward-only pair (no gradient requested): a = <e^_p, e^_q>, returns w (a - t)^2
__device__ __forceinline__ float pair_fwd(const float* __restrict__ nb, int chs, const f32x2 (&e)[kC2], float t, float w,
                                          float& a_out) {
    f32x2 xq[kC2];
#pragma unroll
    for (int i = 0; i < kC2; ++i) xq[i] = pack2(nb[(2 * i) * chs], nb[(2 * i + 1) * chs]);
    f32x2 ss0 = 0ull, ss1 = 0ull, d0 = 0ull, d1 = 0ull;
#pragma unroll
    for (int i = 0; i < kC2; i += 2) {
        ss0 = fma2(xq[i], xq[i], ss0);
        ss1 = fma2(xq[i + 1], xq[i + 1], ss1);
        d0 = fma2(e[i], xq[i], d0);
        d1 = fma2(e[i + 1], xq[i + 1], d1);
    }
    const float a = hsum2(d0, d1) * inv_norm(hsum2(ss0, ss1));
    a_out = a;
    const float diff = a - t;
    return w * diff * diff;
}

// Two in-plane neighbours of the own voxel p at once, lane-packed: A = p - s (pair anchored at p: affinity and loss term
// returned), B = p + s (pair anchored at B whose partner is p: gradient only).  GL[c] accumulates
// (g_A x_A[c], g_B x_B[c]); the two halves are added once per slice.
__device__ __forceinline__ float pair2(const float* __restrict__ nbA, const float* __restrict__ nbB, int chs,
                                       const f32x2 (&e)[kC2], f32x2 (&GL)[kC], float tA, float wA, float tB, float wB,
                                       float gco, float& a_out) {
    f32x2 X[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) X[c] = pack2(nbA[c * chs], nbB[c * chs]);
    f32x2 ss0 = 0ull, ss1 = 0ull, d0 = 0ull, d1 = 0ull;
#pragma unroll
    for (int i = 0; i < kC2; ++i) {
        float elo, ehi;
        unpack2(e[i], elo, ehi);
        ss0 = fma2(X[2 * i], X[2 * i], ss0);
        ss1 = fma2(X[2 * i + 1], X[2 * i + 1], ss1);
        d0 = fma2(pack2(elo, elo), X[2 * i], d0);
        d1 = fma2(pack2(ehi, ehi), X[2 * i + 1], d1);
    }
    float ssA, ssB;
    unpack2(add2(ss0, ss1), ssA, ssB);
    const f32x2 rq2 = pack2(inv_norm(ssA), inv_norm(ssB));
    const f32x2 a2 = mul2(add2(d0, d1), rq2);
    const f32x2 diff2 = add2(a2, pack2(-tA, -tB));
    const f32x2 wd2 = mul2(pack2(wA, wB), diff2);
    const f32x2 gr2 = mul2(mul2(pack2(gco, gco), wd2), rq2);
#pragma unroll
    for (int c = 0; c < kC; ++c) GL[c] = fma2(gr2, X[c], GL[c]);
    float aA, aB, wdA, wdB, dA, dB;
    unpack2(a2, aA, aB);
    unpack2(wd2, wdA, wdB);
    unpack2(diff2, dA, dB);
    a_out = aA;
    return wdA * dA;
}

// Hybrid: norms and dot products lane-packed as in pair2 (no horizontal sums; the dot takes the cheap scalar-broadcast
// FFMA2 form), but the gradient update in SCALAR FFMA on the halves of X_c -- a packed FMA with three distinct register
// pairs costs 3.5 issue clocks of the math port against 2 x 1.37 for two scalar FMAs (profiles/r1_microbench_pair_mix.txt).
__device__ __forceinline__ float pair2h(const float* __restrict__ nbA, const float* __restrict__ nbB, int chs,
                                        const f32x2 (&e)[kC2], float (&G)[kC], float tA, float wA, float tB, float wB,
                                        float gco, float& a_out) {
    f32x2 X[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) X[c] = pack2(nbA[c * chs], nbB[c * chs]);
    f32x2 ss0 = 0ull, ss1 = 0ull, d0 = 0ull, d1 = 0ull;
#pragma unroll
    for (int i = 0; i < kC2; ++i) {
        float elo, ehi;
        unpack2(e[i], elo, ehi);
        ss0 = fma2(X[2 * i], X[2 * i], ss0);
        ss1 = fma2(X[2 * i + 1], X[2 * i + 1], ss1);
        d0 = fma2(pack2(elo, elo), X[2 * i], d0);
        d1 = fma2(pack2(ehi, ehi), X[2 * i + 1], d1);
    }
    float ssA, ssB;
    unpack2(add2(ss0, ss1), ssA, ssB);
    const f32x2 rq2 = pack2(inv_norm(ssA), inv_norm(ssB));
    const f32x2 a2 = mul2(add2(d0, d1), rq2);
    const f32x2 diff2 = add2(a2, pack2(-tA, -tB));
    const f32x2 wd2 = mul2(pack2(wA, wB), diff2);
    const f32x2 gr2 = mul2(mul2(pack2(gco, gco), wd2), rq2);
    float gA, gB;
    unpack2(gr2, gA, gB);
#pragma unroll
    for (int c = 0; c < kC; ++c) {
        float xa, xb;
        unpack2(X[c], xa, xb);
        G[c] = fmaf(gA, xa, G[c]);
    }
#pragma unroll
    for (int c = 0; c < kC; ++c) {
        float xa, xb;
        unpack2(X[c], xa, xb);
        G[c] = fmaf(gB, xb, G[c]);
    }
    float aA, aB, wdA, wdB, dA, dB;
    unpack2(a2, aA, aB);
    unpack2(wd2, wdA, wdB);
    unpack2(diff2, dA, dB);
    a_out = aA;
    return wdA * dA;
}

// Channel-packed pair (the round-1 form): a = <e^_p, e^_q>, g = gcoef * w * (a - t) and G += g * e^_q.
__device__ __forceinline__ float pair1(const float* __restrict__ nb, int chs, const f32x2 (&e)[kC2], f32x2 (&G)[kC2], float t,
                                       float w, float gco, float& a_out) {
    f32x2 xq[kC2];
#pragma unroll
    for (int i = 0; i < kC2; ++i) xq[i] = pack2(nb[(2 * i) * chs], nb[(2 * i + 1) * chs]);
    f32x2 ss0 = 0ull, ss1 = 0ull, d0 = 0ull, d1 = 0ull;
#pragma unroll
    for (int i = 0; i < kC2; i += 2) {
        ss0 = fma2(xq[i], xq[i], ss0);
        ss1 = fma2(xq[i + 1], xq[i + 1], ss1);
        d0 = fma2(e[i], xq[i], d0);
        d1 = fma2(e[i + 1], xq[i + 1], d1);
    }
    const float rq = inv_norm(hsum2(ss0, ss1));
    const float a = hsum2(d0, d1) * rq;
    a_out = a;
    const float diff = a - t;
    const float gr = gco * w * diff * rq;
    const f32x2 gr2 = pack2(gr, gr);
#pragma unroll
    for (int i = 0; i < kC2; ++i) G[i] = fma2(gr2, xq[i], G[i]);
    return w * diff * diff;
}

#ifndef PEA_K4V2_PAIR
#define PEA_K4V2_PAIR 0  // in-plane pair evaluation: 0 channel-packed (round 1), 1 lane-packed, 2 lane-packed + scalar update
#endif

__device__ __forceinline__ float ld_cg(const float* p) {
    float v;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

template <int TY, bool GRAD>
__global__ void __launch_bounds__(TY * 32, 1)
fused3d_norm5_v2(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapA,
                 const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapTy,
                 const __grid_constant__ CUtensorMap mapWy, const __grid_constant__ CUtensorMap mapTx,
                 const __grid_constant__ CUtensorMap mapWx, const __grid_constant__ CUtensorMap mapT1,
                 const __grid_constant__ CUtensorMap mapW1, const ParamsV2 P) {
    static_assert(TY == 12, "12 warps: four lane quarters x three 128-column groups of Tensor Memory; 12 channel sums");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    StageV2<TY>& S = *reinterpret_cast<StageV2<TY>*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    __shared__ uint64_t barY, barX;
    __shared__ uint32_t cntY, cntX, s_tmem, s_last;
    __shared__ float s_red[12][TY];
    __shared__ float s_loss[12];

    constexpr uint32_t kBytesY = sizeof(S.sY) + sizeof(S.sAm) + sizeof(S.sAp) + sizeof(S.sTWy) + sizeof(S.sFar);
    constexpr uint32_t kBytesX = sizeof(S.sX) + sizeof(S.sTWx);
    constexpr int CHS_Y = (TY + 2 * HY) * TX, CHS_A = TY * TX, CHS_X = TY * XW;
    constexpr int CHS_TY = 2 * TY * TX, CHS_TX = TY * 2 * TX;  // channel planes of the target / weight boxes
    constexpr int kWy = 4 * CHS_TY, kWx = 4 * CHS_TX;         // weight = target + kW

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int D = P.D, H = P.H, W = P.W;
    // ---- balanced tiling: strip (b, x strip) and the j-th of its m tiles
    int strip, j, m;
    {
        const int bid = blockIdx.x, nfirst = P.nA * P.mA;
        if (bid < nfirst) {
            strip = bid / P.mA;
            j = bid - strip * P.mA;
            m = P.mA;
        } else {
            const int r = bid - nfirst;
            strip = P.nA + r / (P.mA + 1);
            j = r % (P.mA + 1);
            m = P.mA + 1;
        }
    }
    const int b = strip / P.strips_x, x0 = (strip - b * P.strips_x) * TX;
    const int y0 = (int)(((int64_t)j * H) / m);
    const int rows = (int)(((int64_t)(j + 1) * H) / m) - y0;  // 1 .. TY live rows (= participating warps)
    const int x = x0 + tx, y = y0 + ty;
    const bool rowlive = ty < rows;
    const bool live = rowlive && (x < W);
    const int64_t HW = (int64_t)H * W, N = HW * D;
    const int64_t pix = (int64_t)(live ? y : 0) * W + (live ? x : 0);

    auto issueY = [&](int z) {
        mbar_arrive_expect_tx(&barY, kBytesY);
        tma_load_5d(&S.sY[0][0][0], &mapY, &barY, x0, y0 - HY, z, 0, b);
        tma_load_5d(&S.sTWy[0][0][0][0], &mapTy, &barY, x0, y0, z, 1, b);
        tma_load_5d(&S.sTWy[1][0][0][0], &mapWy, &barY, x0, y0, z, 1, b);
        tma_load_5d(&S.sFar[0][0][0], &mapT1, &barY, x0, y0 + FAR, z, 10, b);
        tma_load_5d(&S.sFar[1][0][0], &mapW1, &barY, x0, y0 + FAR, z, 10, b);
        tma_load_5d(&S.sAm[0][0][0], &mapA, &barY, x0, y0 - FAR, z, 0, b);
        tma_load_5d(&S.sAp[0][0][0], &mapA, &barY, x0, y0 + FAR, z, 0, b);
    };
    auto issueX = [&](int z) {
        mbar_arrive_expect_tx(&barX, kBytesX);
        tma_load_5d(&S.sX[0][0][0], &mapX, &barX, x0 - XPAD, y0, z, 0, b);
        tma_load_5d(&S.sTWx[0][0][0][0], &mapTx, &barX, x0, y0, z, 2, b);
        tma_load_5d(&S.sTWx[1][0][0][0], &mapWx, &barX, x0, y0, z, 2, b);
    };

    if (threadIdx.x == 0) {
        mbar_init(&barY, 1);
        mbar_init(&barX, 1);
        cntY = cntX = 0;
        mbar_fence_init();
        issueY(0);
        issueX(0);
    }
    if (ty == 0) {  // one warp allocates the whole Tensor Memory of this SM (one CTA per SM anyway: shared memory)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    // this thread's private strip: lane = 32 * (warp % 4) + lane id, 128 columns starting at 128 * (warp / 4);
    // slot s (slice z with z & 3 == s): columns [32 s, 32 s + 16) unit vector, [32 s + 16, 32 s + 32) gradient accumulator
    const uint32_t tq = tmem_base + (((uint32_t)(ty & 3) * 32u) << 16) + (uint32_t)(ty >> 2) * 128u;

    float lossacc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) lossacc[k] = 0.f;

    if (rowlive) {
        const uint32_t nwarps = (uint32_t)rows;
        float rr1 = 0.f, rr2 = 0.f, rr3 = 0.f, rr4 = 0.f;  // 1/|x| of slices z-1 .. z-4 (for the normalisation backward)
        const float* tgt_b = P.target + (int64_t)b * 12 * N + pix;
        const float* wgt_b = P.weight + (int64_t)b * 12 * N + pix;
        float* affs_b = P.affs ? P.affs + (int64_t)b * 12 * N + pix : nullptr;
        float* grad_b = P.grad ? P.grad + (int64_t)b * kC * N + pix : nullptr;

        // back through the normalisation:  (G - e^ <e^, G>) / |x|   (eps branch of F.normalize: G / eps)
        auto retire = [&](const f32x2 (&e)[kC2], const f32x2 (&G)[kC2], float r, int zf) {
            if (!GRAD || !live) return;
            f32x2 p0 = 0ull, p1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                p0 = fma2(e[i], G[i], p0);
                p1 = fma2(e[i + 1], G[i + 1], p1);
            }
            float eg = hsum2(p0, p1);
            if (!(r < 1e12f)) eg = 0.f;
            const f32x2 neg = pack2(-eg, -eg), r2 = pack2(r, r);
            float* dst = grad_b + (int64_t)zf * HW;
#pragma unroll
            for (int i = 0; i < kC2; ++i) {
                float lo, hi;
                unpack2(mul2(fma2(e[i], neg, G[i]), r2), lo, hi);
                dst[(int64_t)(2 * i) * N] = lo;
                dst[(int64_t)(2 * i + 1) * N] = hi;
            }
        };

        const float* oY = &S.sY[0][ty + HY][tx];   // own voxel in the y-column box
        const float* oAm = &S.sAm[0][ty][tx];
        const float* oAp = &S.sAp[0][ty][tx];
        const float* oX = &S.sX[0][ty][tx + XPAD];
        const float* tY = &S.sTWy[0][0][ty][tx];
        const float* tX = &S.sTWx[0][0][ty][tx];
        const float* tF = &S.sFar[0][ty][tx];

        for (int z = 0; z < D; ++z) {
            const uint32_t par = (uint32_t)(z & 1);
            const int64_t zoff = (int64_t)z * HW;
            // own target / weight of the four z channels (0, 3, 6, 9): needed at the end of the slice, requested now
            float tz[4], wz[4];
#pragma unroll
            for (int jz = 0; jz < 4; ++jz) {
                const int64_t o = (int64_t)(3 * jz) * N + zoff;
                tz[jz] = ldg_early(tgt_b + o, live);
                wz[jz] = ldg_early(wgt_b + o, live);
            }

            f32x2 e[kC2], G[kC2];
#if PEA_K4V2_PAIR == 1
            f32x2 GL[kC];
#elif PEA_K4V2_PAIR == 2
            float GS[kC];
#endif
            float rp;
            // ---------------- y part
            mbar_wait(&barY, par);
            {
                f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
                for (int i = 0; i < kC2; ++i) e[i] = pack2(oY[(2 * i) * CHS_Y], oY[(2 * i + 1) * CHS_Y]);
#pragma unroll
                for (int i = 0; i < kC2; i += 2) {
                    s0 = fma2(e[i], e[i], s0);
                    s1 = fma2(e[i + 1], e[i + 1], s1);
                }
                rp = inv_norm(hsum2(s0, s1));
                const f32x2 rp2 = pack2(rp, rp);
#pragma unroll
                for (int i = 0; i < kC2; ++i) e[i] = mul2(e[i], rp2);
                if (GRAD) {
#if PEA_K4V2_PAIR == 1
#pragma unroll
                    for (int c = 0; c < kC; ++c) GL[c] = 0ull;
#elif PEA_K4V2_PAIR == 2
#pragma unroll
                    for (int c = 0; c < kC; ++c) GS[c] = 0.f;
#else
#pragma unroll
                    for (int i = 0; i < kC2; ++i) G[i] = 0ull;
#endif
                }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int s = shift_yx(g);
                const int k = 3 * g + 1;
                const float* nb_m = (g < 3) ? oY - s * TX : oAm;
                const float* nb_p = (g < 3) ? oY + s * TX : oAp;
                const int chs = (g < 3) ? CHS_Y : CHS_A;
                const bool vf = y >= s;  // partner above the volume: pair dropped (its raw vector is TMA zero fill: a == 0)
                const float t_f = tY[g * CHS_TY], w_f = vf ? tY[kWy + g * CHS_TY] : 0.f;
                float a;
                if (GRAD) {
                    // pair anchored at y + s: rows below the volume are zero fill in the weight box == pair dropped
                    const float t_r = (g < 3) ? tY[g * CHS_TY + s * TX] : tF[0];
                    const float w_r = (g < 3) ? tY[kWy + g * CHS_TY + s * TX] : tF[CHS_A];
#if PEA_K4V2_PAIR == 1
                    lossacc[k] += pair2(nb_m, nb_p, chs, e, GL, t_f, w_f, t_r, w_r, P.gcoef[k], a);
#elif PEA_K4V2_PAIR == 2
                    lossacc[k] += pair2h(nb_m, nb_p, chs, e, GS, t_f, w_f, t_r, w_r, P.gcoef[k], a);
#else
                    float a_r;
                    lossacc[k] += pair1(nb_m, chs, e, G, t_f, w_f, P.gcoef[k], a);
                    pair1(nb_p, chs, e, G, t_r, w_r, P.gcoef[k], a_r);
#endif
                } else {
                    lossacc[k] += pair_fwd(nb_m, chs, e, t_f, w_f, a);
                }
                if (affs_b && live) affs_b[(int64_t)k * N + zoff] = a;
            }
            __syncwarp();
            if (tx == 0 && arrive_count(&cntY) == nwarps * (uint32_t)(z + 1) - 1 && z + 1 < D) issueY(z + 1);

            // ---------------- x part
            mbar_wait(&barX, par);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int s = shift_yx(g);
                const int k = 3 * g + 2;
                const bool vf = x >= s;
                const float t_f = tX[g * CHS_TX], w_f = vf ? tX[kWx + g * CHS_TX] : 0.f;
                float a;
                if (GRAD) {
                    const float t_r = tX[g * CHS_TX + s], w_r = tX[kWx + g * CHS_TX + s];
#if PEA_K4V2_PAIR == 1
                    lossacc[k] += pair2(oX - s, oX + s, CHS_X, e, GL, t_f, w_f, t_r, w_r, P.gcoef[k], a);
#elif PEA_K4V2_PAIR == 2
                    lossacc[k] += pair2h(oX - s, oX + s, CHS_X, e, GS, t_f, w_f, t_r, w_r, P.gcoef[k], a);
#else
                    float a_r;
                    lossacc[k] += pair1(oX - s, CHS_X, e, G, t_f, w_f, P.gcoef[k], a);
                    pair1(oX + s, CHS_X, e, G, t_r, w_r, P.gcoef[k], a_r);
#endif
                } else {
                    lossacc[k] += pair_fwd(oX - s, CHS_X, e, t_f, w_f, a);
                }
                if (affs_b && live) affs_b[(int64_t)k * N + zoff] = a;
            }
            __syncwarp();
            if (tx == 0 && arrive_count(&cntX) == nwarps * (uint32_t)(z + 1) - 1 && z + 1 < D) issueX(z + 1);
#if PEA_K4V2_PAIR == 1
            if (GRAD) {  // the two halves of every lane-packed accumulator -> channel-packed G
#pragma unroll
                for (int i = 0; i < kC2; ++i) {
                    float a0, a1, b0, b1;
                    unpack2(GL[2 * i], a0, a1);
                    unpack2(GL[2 * i + 1], b0, b1);
                    G[i] = pack2(a0 + a1, b0 + b1);
                }
            }
#elif PEA_K4V2_PAIR == 2
            if (GRAD) {
#pragma unroll
                for (int i = 0; i < kC2; ++i) G[i] = pack2(GS[2 * i], GS[2 * i + 1]);
            }
#endif

            // ---------------- z part: pairs (z, z - j); slice z - j comes out of its Tensor-Memory slot, its updated
            // gradient accumulator goes back (j < 4) or -- all pairs of slice z - 4 seen -- straight through the
            // normalisation backward to global memory
#pragma unroll
            for (int jz = 1; jz <= 4; ++jz) {
                const int k = 3 * (jz - 1);
                if (z >= jz) {  // CTA-uniform
                    const uint32_t slot = tq + ((uint32_t)((z - jz) & 3) << 5);
                    f32x2 ej[kC2], Gj[kC2];
                    if (GRAD) tmem_ld16x2(slot, ej, Gj);
                    else tmem_ld16(slot, ej);
                    f32x2 d0 = 0ull, d1 = 0ull;
#pragma unroll
                    for (int i = 0; i < kC2; i += 2) {
                        d0 = fma2(e[i], ej[i], d0);
                        d1 = fma2(e[i + 1], ej[i + 1], d1);
                    }
                    const float a = hsum2(d0, d1);
                    const float diff = a - tz[jz - 1];
                    const float wv = wz[jz - 1];  // 0 on dead lanes
                    lossacc[k] += wv * diff * diff;
                    if (affs_b && live) affs_b[(int64_t)k * N + zoff] = a;
                    if (GRAD) {
                        const float gg = P.gcoef[k] * wv * diff;
                        const f32x2 gg2 = pack2(gg, gg);
#pragma unroll
                        for (int i = 0; i < kC2; ++i) {
                            G[i] = fma2(gg2, ej[i], G[i]);
                            Gj[i] = fma2(gg2, e[i], Gj[i]);
                        }
                        if (jz < 4) tmem_st16(slot + 16, Gj);
                        else retire(ej, Gj, rr4, z - 4);
                    }
                } else if (affs_b && live) {
                    affs_b[(int64_t)k * N + zoff] = 0.f;
                }
            }
            // this slice takes the slot of slice z - 4 (retired above)
            {
                const uint32_t slot = tq + ((uint32_t)(z & 3) << 5);
                tmem_st16(slot, e);
                if (GRAD) tmem_st16(slot + 16, G);
                tmem_wait_st();
            }
            rr4 = rr3;
            rr3 = rr2;
            rr2 = rr1;
            rr1 = rp;
        }
        // drain: slices D-4 .. D-1 (those >= 0) still hold their gradient in Tensor Memory; rr_j belongs to slice D - j
        if (GRAD) {
#pragma unroll
            for (int jz = 4; jz >= 1; --jz) {
                if (D - jz >= 0) {
                    f32x2 ej[kC2], Gj[kC2];
                    tmem_ld16x2(tq + ((uint32_t)((D - jz) & 3) << 5), ej, Gj);
                    retire(ej, Gj, jz == 4 ? rr4 : (jz == 3 ? rr3 : (jz == 2 ? rr2 : rr1)), D - jz);
                }
            }
        }
    }

    // ---------------- loss partials: thread -> warp -> CTA row (warps beyond the tile's height contribute zeros)
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const float v = warp_sum(lossacc[k]);
        if (tx == 0) s_red[k][ty] = v;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (ty == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
    if (threadIdx.x < 12) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < TY; ++i) v += s_red[threadIdx.x][i];
        P.partials[(int64_t)blockIdx.x * 12 + threadIdx.x] = v;
    }
    if (P.ticket == nullptr) return;  // the caller runs the separate finalize kernel

    // ---------------- in-kernel finalize by the last CTA to arrive
    if (ty == 0) {  // the partials were written by lanes 0..11 of this warp: fence, then one ticket per CTA
        __threadfence();
        __syncwarp();
        if (tx == 0) s_last = (atomicAdd(P.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        // warp k owns channel k: every lane keeps four independent loads in flight, rows lane, lane + 32, ...; double
        // accumulation in a fixed order, then a fixed shuffle tree => run-to-run reproducible
        const int k = ty, n = (int)gridDim.x;
        const float* src = P.partials + k;
        double acc = 0.0;
        int i = tx;
        for (; i + 96 < n; i += 128) {
            const float v0 = ld_cg(src + (int64_t)i * 12), v1 = ld_cg(src + (int64_t)(i + 32) * 12);
            const float v2 = ld_cg(src + (int64_t)(i + 64) * 12), v3 = ld_cg(src + (int64_t)(i + 96) * 12);
            acc += (double)v0;
            acc += (double)v1;
            acc += (double)v2;
            acc += (double)v3;
        }
        for (; i < n; i += 32) acc += (double)ld_cg(src + (int64_t)i * 12);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (tx == 0) s_loss[k] = (float)acc / P.norm[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.f;
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            P.loss_out[1 + k] = s_loss[k];
            total += s_loss[k] * P.coef[k];
        }
        P.loss_out[0] = total;
        *P.ticket = 0u;  // ready for the next launch on this stream
    }
}

}  // namespace t3

// ---- host side -------------------------------------------------------------------------------------------------
#ifndef PEA_K4V2_BALANCE
#define PEA_K4V2_BALANCE 0  // 0: ceil(H / 12) tiles of 12 rows per strip (A/B)
#endif
namespace {
constexpr int kTY = 12;

struct TilePlan {
    int strips_x, mA, nA, grid;
};

// Deal the rows of every (sample, 32-column strip) out to tiles of at most kTY rows such that the number of CTAs is a
// multiple of the SM count whenever that keeps the tiles reasonably tall (>= 9 rows on average).
TilePlan plan_tiles(int B, int H, int W) {
    TilePlan p;
    p.strips_x = (W + t3::TX - 1) / t3::TX;
    const int64_t S = (int64_t)B * p.strips_x;
    const int m_min = (H + kTY - 1) / kTY;
    int64_t n = S * m_min;
    int dev = 0, nsm = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    if (nsm > 0 && PEA_K4V2_BALANCE) {
        const int64_t n_up = (n + nsm - 1) / nsm * nsm;
        const int64_t m_hi = (n_up + S - 1) / S;  // tiles of the most finely cut strips
        if (m_hi > 0 && (double)H / (double)m_hi >= 9.0) n = n_up;
    }
    p.mA = (int)(n / S);
    const int nB = (int)(n % S);
    p.nA = (int)S - nB;
    p.grid = (int)n;
    return p;
}
}  // namespace

int fused3d_norm5_v2_blocks(int B, int H, int W) { return plan_tiles(B, H, W).grid; }

int launch_fused3d_norm5_v2(const float* emb, const Dims& d, const OffsetTable& tab, const float* target,
                            const float* weight, float* affs, float* partials, float* grad, uint32_t* ticket,
                            float* loss_out, const float* norm, const float* coef, cudaStream_t st) {
    constexpr int TY = kTY;
    CUtensorMap mY, mA, mX, mTy, mWy, mTx, mWx, mT1, mW1;
    int rc;
    if ((rc = encode_volume_map(&mY, emb, d.B, 16, d.D, d.H, d.W, t3::TX, TY + 2 * t3::HY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mA, emb, d.B, 16, d.D, d.H, d.W, t3::TX, TY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mX, emb, d.B, 16, d.D, d.H, d.W, t3::XW, TY, 1, 16))) return rc;
    // every third channel from the start coordinate: 4 planes (c0, c0 + 3, c0 + 6, c0 + 9) per copy
    if ((rc = encode_volume_map(&mTy, target, d.B, 12, d.D, d.H, d.W, t3::TX, 2 * TY, 1, 10, 0, 3))) return rc;
    if ((rc = encode_volume_map(&mWy, weight, d.B, 12, d.D, d.H, d.W, t3::TX, 2 * TY, 1, 10, 0, 3))) return rc;
    if ((rc = encode_volume_map(&mTx, target, d.B, 12, d.D, d.H, d.W, 2 * t3::TX, TY, 1, 10, 0, 3))) return rc;
    if ((rc = encode_volume_map(&mWx, weight, d.B, 12, d.D, d.H, d.W, 2 * t3::TX, TY, 1, 10, 0, 3))) return rc;
    if ((rc = encode_volume_map(&mT1, target, d.B, 12, d.D, d.H, d.W, t3::TX, TY, 1, 1))) return rc;
    if ((rc = encode_volume_map(&mW1, weight, d.B, 12, d.D, d.H, d.W, t3::TX, TY, 1, 1))) return rc;
    const TilePlan tp = plan_tiles(d.B, d.H, d.W);
    t3::ParamsV2 P;
    P.B = d.B; P.D = d.D; P.H = d.H; P.W = d.W;
    P.strips_x = tp.strips_x; P.mA = tp.mA; P.nA = tp.nA;
    P.target = target; P.weight = weight; P.affs = affs; P.grad = grad; P.partials = partials;
    P.ticket = ticket; P.loss_out = loss_out;
    for (int k = 0; k < 12; ++k) {
        P.gcoef[k] = tab.gcoef[k];
        P.norm[k] = norm ? norm[k] : 1.f;
        P.coef[k] = coef ? coef[k] : 1.f;
    }
    const size_t smem = sizeof(t3::StageV2<TY>) + 128;
    cudaError_t e;
    if (grad) {
        auto kern = t3::fused3d_norm5_v2<TY, true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<tp.grid, TY * 32, smem, st>>>(mY, mA, mX, mTy, mWy, mTx, mWx, mT1, mW1, P);
    } else {
        auto kern = t3::fused3d_norm5_v2<TY, false>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<tp.grid, TY * 32, smem, st>>>(mY, mA, mX, mTy, mWy, mTx, mWx, mT1, mW1, P);
    }
    return (int)cudaGetLastError();
}

}  // namespace pea
