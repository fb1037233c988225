// Tiled 3-D kernels for the AC3/AC4 12-offset set (scripts_ac3ac4/loss/loss_embedding_mse.py): a CTA owns TY x 32 voxel
// columns of one sample and MARCHES along z; raw neighbour vectors are staged per slice by TMA (cp.async.bulk.tensor,
// 5-D maps, zero fill outside the volume == the reference's "valid region only" rule), the z pairs never touch memory.
//
//   fused3d_norm5_tm   (default, embedding_loss_norm5 :169-194 and its label-driven flavour)  loss + backward, ONE pass:
//       z rings (unit vectors + gradient accumulators of the last four slices) in TENSOR MEMORY (tcgen05.st / ld as a
//       second register file, slot = z & 3), 12 warps per SM; y-column box + two y+-27 arm boxes + x-row box of raw
//       vectors, neighbour 1/|x| recomputed in registers; target / weight by channel-strided TMA groups refilled by the
//       last warp that has finished reading them (release / acquire counters, no __syncthreads in the slice loop);
//       gather-form backward (20 pair evaluations per voxel); loss partials -> fixed-order finalize kernel.
//   cross3d_norm5_tm   (ema_embedding_loss_norm5 :263-289)  two views, second one constant: no reverse pairs, partner
//       ring in Tensor Memory.
//   fwd3d_norm5        (inf_embedding_loss_norm5 :212-234, z-sharded volumes)  forward only, 4-slice register ring,
//       optional external halo (peer GPU memory) for the ring warm-up, relu epilogue.
//   fused3d_norm5 / cross3d_norm5  the round-1 register-ring variants (PEA_K4_TMEM = 0 / PEA_K4X_TMEM = 0), kept as the
//       A/B reference point of DESIGN.md's history: 250 registers, 8 warps per SM.
// What round 2 measured and rejected for the one-view kernel (lane-packed pairs, scalar gradient update, balanced tile
// heights, in-kernel loss finalize; commit 160d351 holds that kernel) is written up in DESIGN.md section 4.
#include "pea_k4_common.cuh"

#ifndef PEA_TILE_Y3D
#define PEA_TILE_Y3D 8
#endif

namespace pea {

namespace t3 {


template <int TY>
struct alignas(1024) Stage {
    float sY[kC][TY + 2 * HY][TX];  // rows y0-9 .. y0+TY+8
    float sAm[kC][TY][TX];          // rows y0-27 ..
    float sAp[kC][TY][TX];          // rows y0+27 ..
    float sX[kC][TY][XW];           // cols x0-28 .. x0+59
    float sT[2][12][TY][TX];        // own target tile, double-buffered over z
    float sW[2][12][TY][TX];        // own weight tile
};

struct Params {
    int B, D, H, W;
    const float* target;
    const float* weight;
    float* affs;      // may be null
    float* grad;      // may be null
    float* partials;  // [gridDim.x][12]
    float gcoef[12];
    const float* wtab;  // label-driven flavour: [B][12][2] weight where the target is 1 / 0 (pea_gt_pair_weights)
    int accumulate;     // two-view kernel: grad += instead of grad = (PEA_EPI_GRAD_ACCUMULATE: the one-view loss of the same
                        // training step already wrote its gradient there)
};


// One in-plane neighbour q of the own voxel p.  `nb` points at channel 0 of q in shared memory (raw values),
// channels are `chs` floats apart.  Computes a = <e^_p, e^_q>, g = gcoef * w * (a - t) and G += g * e^_q.
template <bool GRAD>
__device__ __forceinline__ float inplane_pair(const float* __restrict__ nb, int chs, const f32x2 (&e)[kC2],
                                              f32x2 (&G)[kC2], float t, float w, float gco, float& a_out) {
    f32x2 xq[kC2];
#pragma unroll
    for (int i = 0; i < kC2; ++i) xq[i] = pack2(nb[(2 * i) * chs], nb[(2 * i + 1) * chs]);
#ifdef PEA_DOT_SCALAR
    // the dot product reads three distinct register pairs per FFMA2 (3.5 clk on the register-bandwidth-bound packed
    // pipe, tools/microbench_mix.cu): two scalar FFMAs take 2.8 clk
    f32x2 ss0 = 0ull, ss1 = 0ull;
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int i = 0; i < kC2; i += 2) {
        ss0 = fma2(xq[i], xq[i], ss0);
        ss1 = fma2(xq[i + 1], xq[i + 1], ss1);
    }
#pragma unroll
    for (int i = 0; i < kC2; ++i) {
        float e0, e1, x0, x1;
        unpack2(e[i], e0, e1);
        unpack2(xq[i], x0, x1);
        d0 = fmaf(e0, x0, d0);
        d1 = fmaf(e1, x1, d1);
    }
    const float rq = inv_norm(hsum2(ss0, ss1));
    const float a = (d0 + d1) * rq;
#else
    f32x2 ss0 = 0ull, ss1 = 0ull, d0 = 0ull, d1 = 0ull;  // bit pattern of (+0.f, +0.f)
#pragma unroll
    for (int i = 0; i < kC2; i += 2) {
        ss0 = fma2(xq[i], xq[i], ss0);
        ss1 = fma2(xq[i + 1], xq[i + 1], ss1);
        d0 = fma2(e[i], xq[i], d0);
        d1 = fma2(e[i + 1], xq[i + 1], d1);
    }
    const float rq = inv_norm(hsum2(ss0, ss1));
    const float a = hsum2(d0, d1) * rq;
#endif
    a_out = a;
    const float diff = a - t;
    if (GRAD) {
        const float gr = gco * w * diff * rq;
        const f32x2 gr2 = pack2(gr, gr);
#pragma unroll
        for (int i = 0; i < kC2; ++i) G[i] = fma2(gr2, xq[i], G[i]);
    }
    return w * diff * diff;
}


// Two in-plane neighbours of the own voxel p at once, LANE-packed: X_c = (x_A[c], x_B[c]) shares one FFMA2 per channel
// for the two norms, the two dot products (own unit vector as the scalar-broadcast operand) and the two gradient
// updates, so there are no horizontal sums and the scalar tail runs packed.  A is the forward pair (anchored at p:
// affinity and loss term returned), B the pair anchored at the other neighbour (gradient only).  GL[c] accumulates
// (g_A x_A[c], g_B x_B[c]); the halves are added once per slice.
__device__ __forceinline__ float inplane_pair2(const float* __restrict__ nbA, const float* __restrict__ nbB, int chs,
                                               const f32x2 (&e)[kC2], f32x2 (&GL)[kC], float tA, float wA, float tB,
                                               float wB, float gco, float& a_out) {
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

template <int TY, bool GRAD>
__global__ void __launch_bounds__(TY * 32, (TY <= 4) ? 2 : 1)
fused3d_norm5(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapA,
              const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapT,
              const __grid_constant__ CUtensorMap mapW, const Params P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // align inside the dynamic window with pointer arithmetic (keeps the shared address space visible to the compiler)
    Stage<TY>& S = *reinterpret_cast<Stage<TY>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    __shared__ uint64_t barY, barX, barT[2];
    __shared__ uint32_t cntY, cntX, cntZ;
    __shared__ float s_red[12][TY];

    constexpr uint32_t kBytesY = sizeof(S.sY) + sizeof(S.sAm) + sizeof(S.sAp);
    constexpr uint32_t kBytesX = sizeof(S.sX);
    constexpr uint32_t kBytesT = sizeof(S.sT[0]) + sizeof(S.sW[0]);
    constexpr int CHS_Y = (TY + 2 * HY) * TX, CHS_A = TY * TX, CHS_X = TY * XW;
    constexpr uint32_t kWarps = TY;

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int D = P.D, H = P.H, W = P.W;
    const int tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
    int bid = blockIdx.x;
    const int tix = bid % tiles_x;
    bid /= tiles_x;
    const int tiy = bid % tiles_y;
    const int b = bid / tiles_y;
    const int x0 = tix * TX, y0 = tiy * TY;
    const int x = x0 + tx, y = y0 + ty;
    const bool live = (x < W) && (y < H);
    const int64_t HW = (int64_t)H * W, N = HW * D;
    const int64_t pix = (int64_t)y * W + x;

    auto issueY = [&](int z) {
        mbar_arrive_expect_tx(&barY, kBytesY);
        tma_load_5d(&S.sY[0][0][0], &mapY, &barY, x0, y0 - HY, z, 0, b);
        tma_load_5d(&S.sAm[0][0][0], &mapA, &barY, x0, y0 - FAR, z, 0, b);
        tma_load_5d(&S.sAp[0][0][0], &mapA, &barY, x0, y0 + FAR, z, 0, b);
    };
    auto issueX = [&](int z) {
        mbar_arrive_expect_tx(&barX, kBytesX);
        tma_load_5d(&S.sX[0][0][0], &mapX, &barX, x0 - XPAD, y0, z, 0, b);
    };
    auto issueT = [&](int z) {
        mbar_arrive_expect_tx(&barT[z & 1], kBytesT);
        tma_load_5d(&S.sT[z & 1][0][0][0], &mapT, &barT[z & 1], x0, y0, z, 0, b);
        tma_load_5d(&S.sW[z & 1][0][0][0], &mapW, &barT[z & 1], x0, y0, z, 0, b);
    };

    if (threadIdx.x == 0) {
        mbar_init(&barY, 1);
        mbar_init(&barX, 1);
        mbar_init(&barT[0], 1);
        mbar_init(&barT[1], 1);
        cntY = cntX = cntZ = 0;
        mbar_fence_init();
        issueY(0);
        issueX(0);
        issueT(0);
        if (D > 1) issueT(1);
    }
    __syncthreads();

    // register rings over z: slot j holds slice z - j (slot 0 = the slice being processed); shifted once per slice
    f32x2 re[5][kC2], rG[5][kC2];
    float rr[5];
    float lossacc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) lossacc[k] = 0.f;
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        rr[s] = 0.f;
#pragma unroll
        for (int i = 0; i < kC2; ++i) re[s][i] = rG[s][i] = 0ull;
    }

    const float* tgt_b = P.target + (int64_t)b * 12 * N + pix;
    const float* wgt_b = P.weight + (int64_t)b * 12 * N + pix;
    float* affs_b = P.affs ? P.affs + (int64_t)b * 12 * N + pix : nullptr;
    float* grad_b = P.grad ? P.grad + (int64_t)b * kC * N + pix : nullptr;

    // back through the normalisation:  (G - e^ <e^, G>) / |x|   (eps branch of F.normalize: G / eps)
    auto finalize = [&](const f32x2 (&e)[kC2], const f32x2 (&G)[kC2], float r, int zf) {
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

    for (int z = 0; z < D; ++z) {
        const uint32_t par = (uint32_t)(z & 1);
        const int64_t zoff = (int64_t)z * HW;
        const int tb = z & 1;
        // target / weight of the pairs anchored at the +s neighbours (gather-form backward): issued now, used later
        float t2[8], w2[8];
        if (GRAD) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int s = shift_yx(g);
                const bool vy = live && (y + s < H), vx = live && (x + s < W);
                const int64_t oy = (int64_t)(3 * g + 1) * N + zoff + (int64_t)s * W;
                const int64_t ox = (int64_t)(3 * g + 2) * N + zoff + s;
                t2[g] = vy ? __ldg(tgt_b + oy) : 0.f;
                w2[g] = vy ? __ldg(wgt_b + oy) : 0.f;
                t2[4 + g] = vx ? __ldg(tgt_b + ox) : 0.f;
                w2[4 + g] = vx ? __ldg(wgt_b + ox) : 0.f;
            }
        }

        f32x2 (&e)[kC2] = re[0];
        f32x2 (&G)[kC2] = rG[0];
        // ---------------- y part (needs the y-column box, the arms and the own target/weight tile)
        mbar_wait(&barY, par);
        {
            const float* own = &S.sY[0][ty + HY][tx];
            f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; ++i) e[i] = pack2(own[(2 * i) * CHS_Y], own[(2 * i + 1) * CHS_Y]);
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                s0 = fma2(e[i], e[i], s0);
                s1 = fma2(e[i + 1], e[i + 1], s1);
            }
            const float rp = inv_norm(hsum2(s0, s1));
            rr[0] = rp;
            const f32x2 rp2 = pack2(rp, rp);
#pragma unroll
            for (int i = 0; i < kC2; ++i) {
                e[i] = mul2(e[i], rp2);
                G[i] = 0ull;
            }
        }
        mbar_wait(&barT[tb], (uint32_t)((z >> 1) & 1));
        const float* sTo = &S.sT[tb][0][ty][tx];
        const float* sWo = &S.sW[tb][0][ty][tx];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int s = shift_yx(g);
            const int k = 3 * g + 1;
            const float* nb_m = (g < 3) ? &S.sY[0][ty + HY - s][tx] : &S.sAm[0][ty][tx];
            const float* nb_p = (g < 3) ? &S.sY[0][ty + HY + s][tx] : &S.sAp[0][ty][tx];
            const int chs = (g < 3) ? CHS_Y : CHS_A;
            // pair anchored here: partner y - s
            const bool vf = live && (y >= s);
            float a;
            const float l = inplane_pair<GRAD>(nb_m, chs, e, G, sTo[k * CHS_A], vf ? sWo[k * CHS_A] : 0.f, P.gcoef[k], a);
            lossacc[k] += vf ? l : 0.f;
            if (affs_b && live) affs_b[(int64_t)k * N + zoff] = vf ? a : 0.f;
            // pair anchored at y + s whose partner is this voxel
            if (GRAD) {
                float a2;
                inplane_pair<true>(nb_p, chs, e, G, t2[g], w2[g], P.gcoef[k], a2);
            }
        }
        // every lane of this warp is done reading sY / arms; the last warp to get here refills them for z + 1
        __syncwarp();
        if (tx == 0 && arrive_count(&cntY) == kWarps * (uint32_t)(z + 1) - 1 && z + 1 < D) issueY(z + 1);

        // ---------------- x part
        mbar_wait(&barX, par);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int s = shift_yx(g);
            const int k = 3 * g + 2;
            const float* nb_m = &S.sX[0][ty][tx + XPAD - s];
            const float* nb_p = &S.sX[0][ty][tx + XPAD + s];
            const bool vf = live && (x >= s);
            float a;
            const float l = inplane_pair<GRAD>(nb_m, CHS_X, e, G, sTo[k * CHS_A], vf ? sWo[k * CHS_A] : 0.f, P.gcoef[k], a);
            lossacc[k] += vf ? l : 0.f;
            if (affs_b && live) affs_b[(int64_t)k * N + zoff] = vf ? a : 0.f;
            if (GRAD) {
                float a2;
                inplane_pair<true>(nb_p, CHS_X, e, G, t2[4 + g], w2[4 + g], P.gcoef[k], a2);
            }
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntX) == kWarps * (uint32_t)(z + 1) - 1 && z + 1 < D) issueX(z + 1);

        // ---------------- z part: pairs (z, z - j), j = 1..4, both ends in registers (each pair evaluated once)
#pragma unroll
        for (int j = 1; j <= 4; ++j) {
            const int k = 3 * (j - 1);
            const bool v = live && (z >= j);
            f32x2 d0 = 0ull, d1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                d0 = fma2(re[0][i], re[j][i], d0);
                d1 = fma2(re[0][i + 1], re[j][i + 1], d1);
            }
            const float a = hsum2(d0, d1);
            const float diff = a - sTo[k * CHS_A];
            const float wv = v ? sWo[k * CHS_A] : 0.f;
            lossacc[k] += wv * diff * diff;
            if (affs_b && live) affs_b[(int64_t)k * N + zoff] = v ? a : 0.f;
            if (GRAD) {
                const float gg = P.gcoef[k] * wv * diff;
                const f32x2 gg2 = pack2(gg, gg);
#pragma unroll
                for (int i = 0; i < kC2; ++i) {
                    rG[0][i] = fma2(gg2, re[j][i], rG[0][i]);
                    rG[j][i] = fma2(gg2, re[0][i], rG[j][i]);
                }
            }
        }
        // the own target/weight tile of this slice is consumed: the last warp refills that buffer for z + 2
        __syncwarp();
        if (tx == 0 && arrive_count(&cntZ) == kWarps * (uint32_t)(z + 1) - 1 && z + 2 < D) issueT(z + 2);

        // slice z - 4 has now seen all of its pairs: retire it, then shift the rings
        if (z >= 4) finalize(re[4], rG[4], rr[4], z - 4);
#pragma unroll
        for (int s = 4; s >= 1; --s) {
            rr[s] = rr[s - 1];
#pragma unroll
            for (int i = 0; i < kC2; ++i) {
                re[s][i] = re[s - 1][i];
                rG[s][i] = rG[s - 1][i];
            }
        }
    }
    // drain: after the last shift slot j holds slice D - j; slices D-4 .. D-1 (those >= 0) are still pending
#pragma unroll
    for (int j = 4; j >= 1; --j)
        if (D - j >= 0) finalize(re[j], rG[j], rr[j], D - j);

    // ---------------- loss partials: thread -> warp -> CTA row
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const float v = warp_sum(lossacc[k]);
        if (tx == 0) s_red[k][ty] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < TY; ++i) v += s_red[threadIdx.x][i];
        P.partials[(int64_t)blockIdx.x * 12 + threadIdx.x] = v;
    }
}


// ================================================================================================================
// Fused kernel, TMEM variant (the default for the loss + backward path).
//
// fused3d_norm5 above keeps the z rings (5 slices x (unit vector + gradient accumulator) = 165 registers per thread) in
// the register file: 250 registers, 8 warps per SM -- and measurements show the kernel's time is inversely
// proportional to the number of resident warps (latency-bound: TY = 6 -> 240 us, TY = 8 -> 187 us at cfg3).
// Blackwell's Tensor Memory (256 KB per SM, 512 columns x 128 lanes x 32 bit) is a second register file when no MMA
// uses it: tcgen05.st / tcgen05.ld with the 32x32b shape move 16 or 32 consecutive columns of ONE lane to / from the
// registers of ONE thread.  Here the four older ring slots live in TMEM (4 x 32 columns per thread, slot of slice z =
// z & 3, addressed at run time: no ring shifting at all), which brings the kernel to <= 168 registers and lets a
// CTA of 12 warps (12 x 32 voxel columns) run per SM.  Shared memory (230 KB) holds the embedding boxes (178 KB) and the
// single-buffered target / weight groups; only the z channels' target / weight are per-thread loads.

#ifdef PEA_K4_ONE_LOSS  // register-pressure experiment: one loss accumulator (per-channel losses wrong)
#define PEA_K4_LK(k) 0
#else
#define PEA_K4_LK(k) (k)
#endif
#ifndef PEA_K4_LANEPACK
#define PEA_K4_LANEPACK 0  // forward + reverse in-plane pair of a shift evaluated together, lane-packed (inplane_pair2)
#endif
#ifndef PEA_K4_DUP
#define PEA_K4_DUP 0  // timing experiments only: extra warps that repeat the work of rows 0.. (no stores)
#endif
#ifdef PEA_K4_NOREFILL  // timing experiments only: boxes loaded once, never refilled (wrong results)
#define PEA_K4_REFILL 0
#else
#define PEA_K4_REFILL 1
#endif
#ifndef PEA_K4_NG
#define PEA_K4_NG 4  // timing experiments only: < 4 drops the far in-plane shifts (wrong results)
#endif

template <int TY>
struct alignas(128) StageM {
    float sY[kC][TY + 2 * HY][TX];  // rows y0-9 .. y0+TY+8
    float sAm[kC][TY][TX];          // rows y0-27 ..
    float sAp[kC][TY][TX];          // rows y0+27 ..
    float sX[kC][TY][XW];           // cols x0-28 .. x0+59
    // target / weight tiles [group][target, weight][4 channels], each group refilled as soon as the part of the slice that
    // consumes it is over (like the embedding boxes).  Group 0: y channels (1, 4, 7, 10) of the tile; 1: x channels
    // (2, 5, 8, 11) of the tile; 2: x channels of the NEXT 32 columns and 3: y channels of the TY rows BELOW the tile
    // (pairs anchored at p + s beyond the tile, gather-form backward).  The z channels come straight from global memory.
    float sTW[4][2][4][TY][TX];
    float sFar[2][TY][TX];  // channel 10 (y shift 27), rows y0 + 27 ..: [target, weight]
};

// LAB: target and weight are derived in the kernel from the int32 segmentation (seg_to_aff semantics: same positive id)
// and the per-channel weight pairs of pea_gt_pair_weights -- mapT / mapW then describe the label volume with the
// footprints of a y-column box (TY + 54 rows) and of the x-row box, staged in the space of the target / weight groups;
// the z channels compare against a four-slice label ring in registers.
template <int TY, bool GRAD, bool LAB = false>
#ifdef PEA_K4_MAXNREG
__global__ void __maxnreg__(PEA_K4_MAXNREG)
#else
__global__ void __launch_bounds__(TY * 32, 1)
#endif
fused3d_norm5_tm(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapA,
                 const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapT,
                 const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapT1,
                 const __grid_constant__ CUtensorMap mapW1, const Params P) {
    static_assert(TY <= 16, "four lane quarters x (512 / 128) column groups");
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    StageM<TY>& S = *reinterpret_cast<StageM<TY>*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    __shared__ uint64_t barY, barX;
    __shared__ uint32_t cntY, cntX, s_tmem;
    __shared__ float s_red[12][TY];
    __shared__ float s_w[12][2];  // LAB: weight where the channel's target is 1 / 0

    constexpr uint32_t kBytesTW = sizeof(S.sTW[0]);
    constexpr int LYR = TY + 2 * FAR;  // rows of the label y-column box, first row y0 - 27
    constexpr uint32_t kBytesLY = LYR * TX * 4, kBytesLX = TY * XW * 4;
    static_assert(kBytesLY <= 2 * kBytesTW && kBytesLX <= kBytesTW, "label boxes live in the target / weight groups' space");
    constexpr uint32_t kBytesY = sizeof(S.sY) + sizeof(S.sAm) + sizeof(S.sAp) +
                                 (LAB ? kBytesLY : kBytesTW + (GRAD ? kBytesTW + sizeof(S.sFar) : 0));
    constexpr uint32_t kBytesX = sizeof(S.sX) + (LAB ? kBytesLX : kBytesTW + (GRAD ? kBytesTW : 0));
    const int32_t* lY = reinterpret_cast<const int32_t*>(&S.sTW[0][0][0][0][0]);  // [LYR][TX]
    const int32_t* lX = reinterpret_cast<const int32_t*>(&S.sTW[2][0][0][0][0]);  // [TY][XW], cols x0 - 28 ..
    constexpr int CHS_Y = (TY + 2 * HY) * TX, CHS_A = TY * TX, CHS_X = TY * XW;
    constexpr uint32_t kWarps = TY + PEA_K4_DUP;

    const int tx = threadIdx.x & 31, wid = threadIdx.x >> 5, ty = wid % TY;
    const int D = P.D, H = P.H, W = P.W;
    const int tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
    int bid = blockIdx.x;
    const int tix = bid % tiles_x;
    bid /= tiles_x;
    const int tiy = bid % tiles_y;
    const int b = bid / tiles_y;
    const int x0 = tix * TX, y0 = tiy * TY;
    const int x = x0 + tx, y = y0 + ty;
    const bool live = (x < W) && (y < H) && (wid < TY);
    const int64_t HW = (int64_t)H * W, N = HW * D;
    const int64_t pix = (int64_t)(live ? y : 0) * W + (live ? x : 0);

    // channels 3 g + c0 (g = 0..3) of target and weight in one strided copy each
    auto issueTW = [&](int part, int c0, int z, uint64_t* bar) {
        tma_load_5d(&S.sTW[part][0][0][0][0], &mapT, bar, x0, y0, z, c0, b);
        tma_load_5d(&S.sTW[part][1][0][0][0], &mapW, bar, x0, y0, z, c0, b);
    };
    auto issueY = [&](int z) {
        mbar_arrive_expect_tx(&barY, kBytesY);
        tma_load_5d(&S.sY[0][0][0], &mapY, &barY, x0, y0 - HY, z, 0, b);
        if (LAB) tma_load_5d(&S.sTW[0][0][0][0][0], &mapT, &barY, x0, y0 - FAR, z, 0, b);
        else issueTW(0, 1, z, &barY);
        if (GRAD && !LAB) {  // rows below the tile for the pairs anchored at y + s (zero fill below the volume == pair dropped)
            tma_load_5d(&S.sTW[3][0][0][0][0], &mapT, &barY, x0, y0 + TY, z, 1, b);
            tma_load_5d(&S.sTW[3][1][0][0][0], &mapW, &barY, x0, y0 + TY, z, 1, b);
            tma_load_5d(&S.sFar[0][0][0], &mapT1, &barY, x0, y0 + FAR, z, 10, b);
            tma_load_5d(&S.sFar[1][0][0], &mapW1, &barY, x0, y0 + FAR, z, 10, b);
        }
        tma_load_5d(&S.sAm[0][0][0], &mapA, &barY, x0, y0 - FAR, z, 0, b);
        tma_load_5d(&S.sAp[0][0][0], &mapA, &barY, x0, y0 + FAR, z, 0, b);
    };
    auto issueX = [&](int z) {
        mbar_arrive_expect_tx(&barX, kBytesX);
        tma_load_5d(&S.sX[0][0][0], &mapX, &barX, x0 - XPAD, y0, z, 0, b);
        if (LAB) tma_load_5d(&S.sTW[2][0][0][0][0], &mapW, &barX, x0 - XPAD, y0, z, 0, b);
        else issueTW(1, 2, z, &barX);
        if (GRAD && !LAB) {
            // the same four x channels of the NEXT 32 columns: target / weight of the pairs anchored at x + s beyond the
            // tile (TMA wants a 16-byte aligned start coordinate, so a box shifted by s itself is not an option)
            tma_load_5d(&S.sTW[2][0][0][0][0], &mapT, &barX, x0 + TX, y0, z, 2, b);
            tma_load_5d(&S.sTW[2][1][0][0][0], &mapW, &barX, x0 + TX, y0, z, 2, b);
        }
    };

    if (threadIdx.x == 0) {
        mbar_init(&barY, 1);
        mbar_init(&barX, 1);
        cntY = cntX = 0;
        mbar_fence_init();
        issueY(0);
        issueX(0);
    }
    if (LAB && threadIdx.x < 24) (&s_w[0][0])[threadIdx.x] = P.wtab[(int64_t)b * 24 + threadIdx.x];
    if (wid == 0) {  // one warp allocates the whole Tensor Memory of this SM (one CTA per SM anyway: shared memory)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
#ifdef PEA_PDL_TRIGGER
    // the loss-finalize kernel behind this one is a programmatic dependent (griddepcontrol.wait at its top): letting
    // it become resident now takes its launch latency off the end of the step
    asm volatile("griddepcontrol.launch_dependents;");
#endif
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    // this thread's private strip: lane = 32 * (warp % 4) + lane id, 128 columns starting at 128 * (warp / 4);
    // slot s (slice z with z & 3 == s): columns [32 s, 32 s + 16) unit vector, [32 s + 16, 32 s + 32) gradient accumulator
    const uint32_t tq = tmem_base + (((uint32_t)(wid & 3) * 32u) << 16) + (uint32_t)(wid >> 2) * 128u;

    float rr1 = 0.f, rr2 = 0.f, rr3 = 0.f, rr4 = 0.f;  // 1/|x| of slices z-1 .. z-4 (for the normalisation backward)
    float lossacc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) lossacc[k] = 0.f;

    const float* tgt_b = LAB ? nullptr : P.target + (int64_t)b * 12 * N + pix;
    const float* wgt_b = LAB ? nullptr : P.weight + (int64_t)b * 12 * N + pix;
    int32_t lz1 = 0, lz2 = 0, lz3 = 0, lz4 = 0, lp = 0;  // LAB: labels of this column at z-1 .. z-4, and at z
    // seg_to_aff: both ends carry the same positive id; the weight map takes one of two values per channel
    auto lab_t = [&](int32_t a, int32_t c) { return a > 0 && c > 0 && a == c; };
#ifdef PEA_K4_NOSTORE
    float* affs_b = nullptr;
#else
    float* affs_b = P.affs ? P.affs + (int64_t)b * 12 * N + pix : nullptr;
#endif
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
#ifdef PEA_K4_NOSTORE
        if (eg != 12345.678f) return;
#endif
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

    // own target / weight of the four z channels (0, 3, 6, 9): needed only at the end of the slice, requested at its start
    float tz[4], wz[4];
    auto load_tz = [&](int z) {
        if constexpr (!LAB) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t o = (int64_t)(3 * j) * N + (int64_t)z * HW;
            tz[j] = ldg_early(tgt_b + o, live);
            wz[j] = ldg_early(wgt_b + o, live);
        }
        }
    };

    const float* sTy = &S.sTW[0][0][0][ty][tx];
    const float* sTx = &S.sTW[1][0][0][ty][tx];
    const float* sTn = &S.sTW[2][0][0][ty][tx];  // next 32 columns of the x channels
    const float* sTb = &S.sTW[3][0][0][ty][tx];  // next TY rows of the y channels
    constexpr int kW = 4 * CHS_A;  // weight plane of a group = target plane + kW

    for (int z = 0; z < D; ++z) {
        const uint32_t par = (uint32_t)(z & 1);
        const int64_t zoff = (int64_t)z * HW;
        load_tz(z);

        f32x2 e[kC2], G[kC2];
        f32x2 GL[kC];  // lane-packed accumulators of the in-plane parts
        float rp;
        // ---------------- y part
#ifdef PEA_K4_NOREFILL
        if (z == 0)
#endif
        mbar_wait(&barY, par);
        {
            const float* own = &S.sY[0][ty + HY][tx];
            f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; ++i) e[i] = pack2(own[(2 * i) * CHS_Y], own[(2 * i + 1) * CHS_Y]);
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                s0 = fma2(e[i], e[i], s0);
                s1 = fma2(e[i + 1], e[i + 1], s1);
            }
            rp = inv_norm(hsum2(s0, s1));
            const f32x2 rp2 = pack2(rp, rp);
#pragma unroll
            for (int i = 0; i < kC2; ++i) {
                e[i] = mul2(e[i], rp2);
                G[i] = 0ull;
            }
            if (PEA_K4_LANEPACK && GRAD) {
#pragma unroll
                for (int c = 0; c < kC; ++c) GL[c] = 0ull;
            }
        }
#pragma unroll
        for (int g = 0; g < PEA_K4_NG; ++g) {
            const int s = shift_yx(g);
            const int k = 3 * g + 1;
            const float* nb_m = (g < 3) ? &S.sY[0][ty + HY - s][tx] : &S.sAm[0][ty][tx];
            const float* nb_p = (g < 3) ? &S.sY[0][ty + HY + s][tx] : &S.sAp[0][ty][tx];
            const int chs = (g < 3) ? CHS_Y : CHS_A;
            const bool vf = y >= s;  // out-of-volume voxels have zero weight (TMA zero fill)
            float a;
            if (LAB) {
                const int32_t* lq = lY + (ty + FAR) * TX + tx;
                if (g == 0) lp = lq[0];
                const bool t_f = lab_t(lp, lq[-s * TX]), t_r = lab_t(lq[s * TX], lp);
                const float w_f = (live && vf) ? s_w[k][t_f ? 0 : 1] : 0.f;
                lossacc[PEA_K4_LK(k)] += inplane_pair<GRAD>(nb_m, chs, e, G, t_f ? 1.f : 0.f, w_f, P.gcoef[k], a);
                if (affs_b && live) affs_b[(int64_t)k * N + zoff] = a;
                if (GRAD) {
                    const float w_r = (y + s < H) ? s_w[k][t_r ? 0 : 1] : 0.f;  // anchor below the volume: pair dropped
                    float a2;
                    inplane_pair<true>(nb_p, chs, e, G, t_r ? 1.f : 0.f, w_r, P.gcoef[k], a2);
                }
            } else if (PEA_K4_LANEPACK && GRAD) {
                const float* pr = (g == 3) ? &S.sFar[0][ty][tx]
                                           : (ty + s < TY ? sTy + (g * CHS_A + s * TX) : sTb + (g * CHS_A + (s - TY) * TX));
                lossacc[PEA_K4_LK(k)] += inplane_pair2(nb_m, nb_p, chs, e, GL, sTy[g * CHS_A], vf ? sTy[kW + g * CHS_A] : 0.f, pr[0],
                                            pr[(g == 3) ? CHS_A : kW], P.gcoef[k], a);
                if (affs_b && live) affs_b[(int64_t)k * N + zoff] = a;
            } else {
            lossacc[PEA_K4_LK(k)] += inplane_pair<GRAD>(nb_m, chs, e, G, sTy[g * CHS_A], vf ? sTy[kW + g * CHS_A] : 0.f, P.gcoef[k], a);
            if (affs_b && live) affs_b[(int64_t)k * N + zoff] = a;  // partner outside the volume: zero fill, a == 0
            if (GRAD) {
                // pair anchored at y + s: its target / weight sit in the own tile group when that row belongs to this CTA,
                // else in the group of the TY rows below; shift 27 always lands in the far box (rows beyond the image
                // are zero fill everywhere == pair dropped).  Warp-uniform choice.
                const float* pr = (g == 3) ? &S.sFar[0][ty][tx]
                                           : (ty + s < TY ? sTy + (g * CHS_A + s * TX) : sTb + (g * CHS_A + (s - TY) * TX));
                const float t_r = pr[0];
                const float w_r = pr[(g == 3) ? CHS_A : kW];
                float a2;
                inplane_pair<true>(nb_p, chs, e, G, t_r, w_r, P.gcoef[k], a2);
            }
            }
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntY) == kWarps * (uint32_t)(z + 1) - 1 && z + 1 < D && PEA_K4_REFILL) issueY(z + 1);

        // ---------------- x part
#ifdef PEA_K4_NOREFILL
        if (z == 0)
#endif
        mbar_wait(&barX, par);
#pragma unroll
        for (int g = 0; g < PEA_K4_NG; ++g) {
            const int s = shift_yx(g);
            const int k = 3 * g + 2;
            const bool vf = x >= s;
            float a;
            if (LAB) {
                const int32_t* lq = lX + ty * XW + tx + XPAD;
                const bool t_f = lab_t(lp, lq[-s]), t_r = lab_t(lq[s], lp);
                const float w_f = (live && vf) ? s_w[k][t_f ? 0 : 1] : 0.f;
                lossacc[PEA_K4_LK(k)] += inplane_pair<GRAD>(&S.sX[0][ty][tx + XPAD - s], CHS_X, e, G, t_f ? 1.f : 0.f, w_f,
                                                            P.gcoef[k], a);
                if (affs_b && live) affs_b[(int64_t)k * N + zoff] = a;
                if (GRAD) {
                    const float w_r = (x + s < W) ? s_w[k][t_r ? 0 : 1] : 0.f;
                    float a2;
                    inplane_pair<true>(&S.sX[0][ty][tx + XPAD + s], CHS_X, e, G, t_r ? 1.f : 0.f, w_r, P.gcoef[k], a2);
                }
            } else if (PEA_K4_LANEPACK && GRAD) {
                const float* pr = (tx + s < TX) ? sTx + (g * CHS_A + s) : sTn + (g * CHS_A + s - TX);
                lossacc[PEA_K4_LK(k)] += inplane_pair2(&S.sX[0][ty][tx + XPAD - s], &S.sX[0][ty][tx + XPAD + s], CHS_X, e, GL,
                                            sTx[g * CHS_A], vf ? sTx[kW + g * CHS_A] : 0.f, pr[0], pr[kW], P.gcoef[k], a);
                if (affs_b && live) affs_b[(int64_t)k * N + zoff] = a;
            } else {
            lossacc[PEA_K4_LK(k)] += inplane_pair<GRAD>(&S.sX[0][ty][tx + XPAD - s], CHS_X, e, G, sTx[g * CHS_A],
                                             vf ? sTx[kW + g * CHS_A] : 0.f, P.gcoef[k], a);
            if (affs_b && live) affs_b[(int64_t)k * N + zoff] = a;
            if (GRAD) {
                // pair anchored at x + s: its target / weight sit in the own tile group, or in the next 32 columns
                const float* pr = (tx + s < TX) ? sTx + (g * CHS_A + s) : sTn + (g * CHS_A + s - TX);
                float a2;
                inplane_pair<true>(&S.sX[0][ty][tx + XPAD + s], CHS_X, e, G, pr[0], pr[kW], P.gcoef[k], a2);
            }
            }
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntX) == kWarps * (uint32_t)(z + 1) - 1 && z + 1 < D && PEA_K4_REFILL) issueX(z + 1);
        if (PEA_K4_LANEPACK && GRAD) {  // the two halves of every lane-packed accumulator -> channel-packed G
#pragma unroll
            for (int i = 0; i < kC2; ++i) {
                float a0, a1, b0, b1;
                unpack2(GL[2 * i], a0, a1);
                unpack2(GL[2 * i + 1], b0, b1);
                G[i] = pack2(a0 + a1, b0 + b1);
            }
        }

        // ---------------- z part: pairs (z, z - j); slice z - j comes out of its Tensor-Memory slot, its updated
        // gradient accumulator goes back (j < 4) or -- all pairs of slice z - 4 seen -- straight through the
        // normalisation backward to global memory
#pragma unroll
        for (int j = 1; j <= 4; ++j) {
            const int k = 3 * (j - 1);
            if (z >= j) {  // CTA-uniform
                const uint32_t slot = tq + ((uint32_t)((z - j) & 3) << 5);
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
                float tzj, wv;
                if (LAB) {
                    const bool t = lab_t(lp, j == 1 ? lz1 : (j == 2 ? lz2 : (j == 3 ? lz3 : lz4)));
                    tzj = t ? 1.f : 0.f;
                    wv = live ? s_w[k][t ? 0 : 1] : 0.f;
                } else {
                    tzj = tz[j - 1];
                    wv = wz[j - 1];  // 0 on dead lanes
                }
                const float diff = a - tzj;
                lossacc[PEA_K4_LK(k)] += wv * diff * diff;
                if (affs_b && live) affs_b[(int64_t)k * N + zoff] = a;
                if (GRAD) {
                    const float gg = P.gcoef[k] * wv * diff;
                    const f32x2 gg2 = pack2(gg, gg);
#pragma unroll
                    for (int i = 0; i < kC2; ++i) {
                        G[i] = fma2(gg2, ej[i], G[i]);
                        Gj[i] = fma2(gg2, e[i], Gj[i]);
                    }
                    if (j < 4) tmem_st16(slot + 16, Gj);
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
        if (LAB) {
            lz4 = lz3;
            lz3 = lz2;
            lz2 = lz1;
            lz1 = lp;
        }
    }
    // drain: slices D-4 .. D-1 (those >= 0) still hold their gradient in Tensor Memory; rr_j belongs to slice D - j
    if (GRAD) {
#pragma unroll
        for (int j = 4; j >= 1; --j) {
            if (D - j >= 0) {
                f32x2 ej[kC2], Gj[kC2];
                tmem_ld16x2(tq + ((uint32_t)((D - j) & 3) << 5), ej, Gj);
                retire(ej, Gj, j == 4 ? rr4 : (j == 3 ? rr3 : (j == 2 ? rr2 : rr1)), D - j);
            }
        }
    }

    // ---------------- loss partials: thread -> warp -> CTA row
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const float v = warp_sum(lossacc[k]);
        if (tx == 0 && wid < TY) s_red[k][ty] = v;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (wid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
    if (threadIdx.x < 12) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < TY; ++i) v += s_red[threadIdx.x][i];
        P.partials[(int64_t)blockIdx.x * 12 + threadIdx.x] = v;
    }
}


// ================================================================================================================
// Forward only: inf_embedding_loss_norm5 (scripts_ac3ac4/loss/loss_embedding_mse.py:212-234) and the z-sharded
// full-volume inference.  Same z-march, but only the "-shift" side of every offset is needed: a 4-slice register
// ring of unit vectors, a y-column box with 9 rows of upper halo, the y-27 arm and an x-row box with 28 columns of
// left halo.  `z_skip` leading slices are halo from the previous z-shard: they only warm the ring up (read straight
// from global, no tiles), nothing is written for them.
template <int TY>
struct alignas(1024) StageF {
    float sY[kC][TY + HY][TX];    // rows y0-9 .. y0+TY-1
    float sAm[kC][TY][TX];        // rows y0-27 ..
    float sX[kC][TY][TX + XPAD];  // cols x0-28 .. x0+31
};

struct ParamsF {
    int B, D, H, W, z_skip;
    int64_t emb_plane, affs_plane;  // elements between channel planes of emb / affs
    const float* emb;
    float* affs;
    // optional external halo: the previous z-shard's last `halo_n` (<= 4) slices, typically PEER memory mapped over
    // NVLink (symmetric memory): read in place by the ring warm-up, so the halo "exchange" is fused into this kernel
    const float* halo;
    int halo_n;
    int64_t halo_plane;
    int relu;  // store max(a, 0): the caller-side F.relu(pred) folded into the kernel
    // optional STAGED halo (pea_zshard_stage_halo): a local copy of the same four slices that a copy engine fills band
    // by band (bands of `stage_band_rows` image rows, starting at row `stage_row0`) while this kernel runs; flag j
    // reaches `stage_epoch` when band j has landed.  Tiles above stage_row0 read `halo` directly (their CTAs start
    // before the first band can have arrived), the others wait for their band's flag and read the local copy.
    int rot_ctas;  // CTAs with blockIdx.x below this (odd ones) visit the halo-dependent slices last
    const float* stage;
    int64_t stage_plane;
    const uint32_t* stage_flags;
    uint32_t stage_epoch;
    int stage_row0, stage_band_rows;
};

#ifndef PEA_K3_ORDER
// tile -> CTA order of the forward kernel: 0 row-major, 1 column-major, 2 blocks of PEA_K3_BLK tile columns (row-major
// inside a block).  A wave of 296 resident CTAs then covers a 512 x 148 pixel region instead of 1024 x 74: less of its
// boundary lies above it, where the y halo (9 rows) and the y-27 arm reach into rows an earlier wave read long ago.
// cfg4 on one GPU: order 0 3.65 ms, order 2 with blocks of 4 / 8 / 16 columns 3.56 / 3.54 / 3.52 ms, order 1 3.74 ms.
#define PEA_K3_ORDER 2
#endif
#ifndef PEA_K3_BLK
#define PEA_K3_BLK 16
#endif
#ifndef PEA_K3_ROTSPREAD
#define PEA_K3_ROTSPREAD 0  // 1: D - 3 starting slices instead of two phases (measured at 2 x 13 slices: 0.571 vs 0.563 ms, no gain)
#endif
#ifndef PEA_K3_STCS
#define PEA_K3_STCS 0   // 1: affinities leave with st.global.cs (evict-first): write-once data does not displace the halos in L2
#endif
__device__ __forceinline__ void st_out(float* p, float v) {
#if PEA_K3_STCS
    __stcs(p, v);
#else
    *p = v;
#endif
}

constexpr uint64_t kStageWaitNs = 40000;  // a staged band later than this is read from the peer in place

template <int TY>
#ifdef PEA_K3_MAXNREG   // timing experiments (tools/build_variant.sh): an explicit register cap instead of the launch bounds
__global__ void __maxnreg__(PEA_K3_MAXNREG)
#else
__global__ void __launch_bounds__(TY * 32, (TY <= 8) ? 2 : 1)
#endif
fwd3d_norm5(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapA,
            const __grid_constant__ CUtensorMap mapX, const ParamsF P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    StageF<TY>& S = *reinterpret_cast<StageF<TY>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    __shared__ uint64_t barY, barX;
    __shared__ uint32_t cntY, cntX;
    constexpr uint32_t kBytesY = sizeof(S.sY) + sizeof(S.sAm);
    constexpr uint32_t kBytesX = sizeof(S.sX);
    constexpr int CHS_Y = (TY + HY) * TX, CHS_A = TY * TX, XWF = TX + XPAD, CHS_X = TY * XWF;
    constexpr uint32_t kWarps = TY;

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int D = P.D, H = P.H, W = P.W, zs = P.z_skip;
    const int tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
    int bid = blockIdx.x;
#if PEA_K3_ORDER == 1      // column-major: vertically adjacent tiles (the y halo and the y-27 arm) are launched back to back
    const int tiy = bid % tiles_y;
    bid /= tiles_y;
    const int tix = bid % tiles_x;
    const int b = bid / tiles_x;
#elif PEA_K3_ORDER == 2    // blocks of PEA_K3_BLK tile columns, column-major inside a block
    const int per_b = tiles_x * tiles_y;
    const int b = bid / per_b;
    bid -= b * per_b;
    const int blk = bid / (PEA_K3_BLK * tiles_y), r = bid - blk * (PEA_K3_BLK * tiles_y);
    const int wblk = min(PEA_K3_BLK, tiles_x - blk * PEA_K3_BLK);  // the last block may be narrower
    const int tiy = r / wblk, tix = blk * PEA_K3_BLK + r % wblk;
#else
    const int tix = bid % tiles_x;
    bid /= tiles_x;
    const int tiy = bid % tiles_y;
    const int b = bid / tiles_y;
#endif
    const int x0 = tix * TX, y0 = tiy * TY;
    const int x = x0 + tx, y = y0 + ty;
    const bool live = (x < W) && (y < H);
    const int64_t HW = (int64_t)H * W;
    const int64_t pix = (int64_t)(live ? y : 0) * W + (live ? x : 0);

    auto issueY = [&](int z) {
        mbar_arrive_expect_tx(&barY, kBytesY);
        tma_load_5d(&S.sY[0][0][0], &mapY, &barY, x0, y0 - HY, z, 0, b);
        tma_load_5d(&S.sAm[0][0][0], &mapA, &barY, x0, y0 - FAR, z, 0, b);
    };
    auto issueX = [&](int z) {
        mbar_arrive_expect_tx(&barX, kBytesX);
        tma_load_5d(&S.sX[0][0][0], &mapX, &barX, x0 - XPAD, y0, z, 0, b);
    };
    // Slice order.  With an external (peer) halo every CTA has to pull 4 slices x 16 channels over NVLink before it can
    // produce slice 0; if all resident CTAs do that at the same moment the link is the bottleneck while the SMs idle.
    // Odd CTAs therefore start with the part of the slab that needs no halo (slices 4 .. D-1, ring warmed up on the
    // slab's own first slices) and visit slices 0 .. 3 last: NVLink demand is spread over the whole kernel.
    const bool ext = P.halo != nullptr;
    // staged halo: the copy engine runs ahead of the tile rows, the CTA reads local memory and needs no rotation
    // (decided per CTA below: a band that has not landed within kStageWaitNs is read from the peer instead -- always
    // correct, the peer's tail is complete before the launch -- so the kernel cannot hang on a copy that never runs)
    bool staged = ext && P.stage_flags != nullptr && y0 >= P.stage_row0;
    __shared__ int stage_ok;
#if PEA_K3_ROTSPREAD
    // experiment (round 2): D - 3 different starting slices (0, i.e. halo first, or 4 .. D-1) dealt out by block index, so
    // that the resident CTAs ask for their halo at D - 3 different points of their lifetime instead of two.  No gain:
    // the link carries 268 MB in the time the slab takes either way
    int rot = 0;
    if (ext && zs == 0 && D > 4) {
        const int j = (int)(blockIdx.x % (unsigned)(D - 3));
        rot = j == 0 ? 0 : 3 + j;
    }
#else
    const int rot = (ext && !staged && zs == 0 && D > 4 && (blockIdx.x & 1) && (int)blockIdx.x < P.rot_ctas) ? 4 : 0;
#endif
    const int nz = D - zs;
    auto slice_of = [&](int i) { const int r = i + rot; return zs + (r >= nz ? r - nz : r); };
    if (threadIdx.x == 0) {
        mbar_init(&barY, 1);
        mbar_init(&barX, 1);
        cntY = cntX = 0;
        mbar_fence_init();
        if (zs < D) {
            issueY(slice_of(0));
            issueX(slice_of(0));
        }
        if (staged) {   // the band of this tile's rows has landed when its flag carries this step's epoch
            const uint32_t* f = P.stage_flags + (y0 - P.stage_row0) / P.stage_band_rows;
            const uint64_t t0 = globaltimer_ns();
            int ok = 1;
            while ((int32_t)(ld_acquire_sys_u32(f) - P.stage_epoch) < 0) {
                if (globaltimer_ns() - t0 > kStageWaitNs) { ok = 0; break; }
                __nanosleep(256);
            }
            stage_ok = ok;
        }
    }
    __syncthreads();   // also carries thread 0's acquire to the threads that read the staged rows
    if (staged) staged = stage_ok != 0;   // CTA-uniform

    f32x2 re[5][kC2];  // slot j: unit vector of slice z - j
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
        for (int i = 0; i < kC2; ++i) re[s][i] = 0ull;

    // ---- ring warm-up: the (up to four) slices that precede the next output slice, own column straight from global --
    //      the leading z_skip slices of emb, the slab's own slices (rotated order), or the external (peer) halo buffer.
    //      All of them are requested at once, straight into their ring slots (slot j = slice -j): over NVLink one round
    //      trip costs microseconds, four dependent ones dominated the CTA.
    // issue: volatile loads (the compiler may not sink them), raw values land in the ring slots; finish: normalise them.
    // Between the two the y and x parts of the next slice run (they need only the own vector and shared memory), so
    // the NVLink round trip of the halo is hidden behind ~80 % of a slice.
    auto warm_issue = [&](const float* src, int64_t wplane, int nwarm, bool fresh = false) {
        const float* warm_b = src + (int64_t)b * kC * wplane + pix;
#pragma unroll
        for (int j = 1; j <= 4; ++j) {
            const bool on = live && j <= nwarm;
            const float* p = warm_b + (int64_t)(on ? nwarm - j : 0) * HW;
#pragma unroll
            for (int i = 0; i < kC2; ++i) {
                // `fresh`: the rows were written by a copy engine while this kernel was running -> L2-coherent loads
                const float lo = fresh ? ldg_early_cg(p, on) : ldg_early(p, on);
                const float hi = fresh ? ldg_early_cg(p + wplane, on) : ldg_early(p + wplane, on);
                p += 2 * wplane;
                re[j][i] = pack2(lo, hi);
            }
        }
    };
    auto warm_finish = [&]() {
#pragma unroll
        for (int j = 1; j <= 4; ++j) {
            // same arithmetic order as the tile path below, so slabs reproduce the whole-volume result bit for bit
            f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                s0 = fma2(re[j][i], re[j][i], s0);
                s1 = fma2(re[j][i + 1], re[j][i + 1], s1);
            }
            const float r = inv_norm(hsum2(s0, s1));
            const f32x2 r2 = pack2(r, r);
#pragma unroll
            for (int i = 0; i < kC2; ++i) re[j][i] = mul2(re[j][i], r2);
        }
    };
    bool pending = true;  // raw halo values sit in the ring slots, to be normalised before the next z part
    if (rot) warm_issue(P.emb + (int64_t)(rot - 4) * HW, P.emb_plane, 4);   // slices rot-4 .. rot-1 of the slab precede slice rot
    else if (staged) warm_issue(P.stage, P.stage_plane, P.halo_n, true);
    else if (ext) warm_issue(P.halo, P.halo_plane, P.halo_n);
    else warm_issue(P.emb + (int64_t)(zs - (zs > 4 ? 4 : zs)) * HW, P.emb_plane, zs > 4 ? 4 : zs);

    const int zprev = ext ? P.halo_n : 0;  // slices that precede emb's slice 0 (zero border applies when none do)
    float* affs_b = P.affs + (int64_t)b * 12 * P.affs_plane + pix;
    for (int it = 0; it < nz; ++it) {
        const int z = slice_of(it);
        const bool more = it + 1 < nz;
        const int znext = more ? slice_of(it + 1) : 0;
        if (rot && z == zs) {   // rotated order: back at slice 0, now the peer halo
            warm_issue(P.halo, P.halo_plane, P.halo_n);
            pending = true;
        }
        const uint32_t par = (uint32_t)(it & 1);
        float* out = affs_b + (int64_t)(z - zs) * HW;
        f32x2 (&e)[kC2] = re[0];
        // ---------------- y part
        mbar_wait(&barY, par);
        {
            const float* own = &S.sY[0][ty + HY][tx];
            f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; ++i) e[i] = pack2(own[(2 * i) * CHS_Y], own[(2 * i + 1) * CHS_Y]);
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                s0 = fma2(e[i], e[i], s0);
                s1 = fma2(e[i + 1], e[i + 1], s1);
            }
            const float rp = inv_norm(hsum2(s0, s1));
            const f32x2 rp2 = pack2(rp, rp);
#pragma unroll
            for (int i = 0; i < kC2; ++i) e[i] = mul2(e[i], rp2);
        }
        f32x2 dummyG[kC2];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int s = shift_yx(g);
            const float* nb = (g < 3) ? &S.sY[0][ty + HY - s][tx] : &S.sAm[0][ty][tx];
            float a;
            inplane_pair<false>(nb, (g < 3) ? CHS_Y : CHS_A, e, dummyG, 0.f, 0.f, 0.f, a);
            if (live) st_out(out + (int64_t)(3 * g + 1) * P.affs_plane, (y >= s) ? (P.relu ? fmaxf(a, 0.f) : a) : 0.f);
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntY) == kWarps * (uint32_t)(it + 1) - 1 && more) issueY(znext);
        // ---------------- x part
        mbar_wait(&barX, par);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int s = shift_yx(g);
            float a;
            inplane_pair<false>(&S.sX[0][ty][tx + XPAD - s], CHS_X, e, dummyG, 0.f, 0.f, 0.f, a);
            if (live) st_out(out + (int64_t)(3 * g + 2) * P.affs_plane, (x >= s) ? (P.relu ? fmaxf(a, 0.f) : a) : 0.f);
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntX) == kWarps * (uint32_t)(it + 1) - 1 && more) issueX(znext);
        // ---------------- z part: ring
        if (pending) {  // CTA-uniform
            warm_finish();
            pending = false;
        }
#pragma unroll
        for (int j = 1; j <= 4; ++j) {
            f32x2 d0 = 0ull, d1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                d0 = fma2(re[0][i], re[j][i], d0);
                d1 = fma2(re[0][i + 1], re[j][i + 1], d1);
            }
            const float az = hsum2(d0, d1);
            if (live) st_out(out + (int64_t)(3 * (j - 1)) * P.affs_plane, (z + zprev >= j) ? (P.relu ? fmaxf(az, 0.f) : az) : 0.f);
        }
#ifndef PEA_K3_NOSHIFT  // timing experiment only (wrong results): what the ring shift costs
#pragma unroll
        for (int s = 4; s >= 1; --s)
#pragma unroll
            for (int i = 0; i < kC2; ++i) re[s][i] = re[s - 1][i];
#endif
    }
}


// ================================================================================================================
// Cross (two-view) fused loss: ema_embedding_loss_norm5 (scripts_ac3ac4/loss/loss_embedding_mse.py:263-289) with the
// second view treated as a constant, as both shipped training loops do (convert_consistency_flip detaches it,
// scripts_ac3ac4/main.py:214 / data_consistency.py:36).  a_i[p] = <e^[p], u^[p - s_i]>: only the "-shift" side exists,
// dL/de^[p] = sum_i g_i[p] u^[p - s_i] is complete as soon as slice z has been visited -- no gradient ring, no lag.
// Tiles (y-column with upper halo, y-27 arm, x-row with left halo) come from the PARTNER; the own vector from a
// centre box of the anchor view.
template <int TY>
struct alignas(1024) StageX {
    float sY[kC][TY + HY][TX];    // partner rows y0-9 .. y0+TY-1
    float sAm[kC][TY][TX];        // partner rows y0-27 ..
    float sX[kC][TY][TX + XPAD];  // partner cols x0-28 .. x0+31
    float sO[kC][TY][TX];         // anchor view, tile centre
    float sT[2][12][TY][TX];
    float sW[2][12][TY][TX];
};

template <int TY, bool GRAD>
__global__ void __launch_bounds__(TY * 32, (TY <= 4) ? 2 : 1)
cross3d_norm5(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapA,
              const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapO,
              const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapW, const Params P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    StageX<TY>& S = *reinterpret_cast<StageX<TY>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    __shared__ uint64_t barY, barX, barO, barT[2];
    __shared__ uint32_t cntY, cntX, cntO, cntZ;
    __shared__ float s_red[12][TY];
    constexpr uint32_t kBytesY = sizeof(S.sY) + sizeof(S.sAm);
    constexpr uint32_t kBytesX = sizeof(S.sX);
    constexpr uint32_t kBytesO = sizeof(S.sO);
    constexpr uint32_t kBytesT = sizeof(S.sT[0]) + sizeof(S.sW[0]);
    constexpr int CHS_Y = (TY + HY) * TX, CHS_A = TY * TX, XWF = TX + XPAD, CHS_X = TY * XWF;
    constexpr uint32_t kWarps = TY;

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int D = P.D, H = P.H, W = P.W;
    const int tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
    int bid = blockIdx.x;
    const int tix = bid % tiles_x;
    bid /= tiles_x;
    const int tiy = bid % tiles_y;
    const int b = bid / tiles_y;
    const int x0 = tix * TX, y0 = tiy * TY;
    const int x = x0 + tx, y = y0 + ty;
    const bool live = (x < W) && (y < H);
    const int64_t HW = (int64_t)H * W, N = HW * D;
    const int64_t pix = (int64_t)(live ? y : 0) * W + (live ? x : 0);

    auto issueY = [&](int z) {
        mbar_arrive_expect_tx(&barY, kBytesY);
        tma_load_5d(&S.sY[0][0][0], &mapY, &barY, x0, y0 - HY, z, 0, b);
        tma_load_5d(&S.sAm[0][0][0], &mapA, &barY, x0, y0 - FAR, z, 0, b);
    };
    auto issueX = [&](int z) {
        mbar_arrive_expect_tx(&barX, kBytesX);
        tma_load_5d(&S.sX[0][0][0], &mapX, &barX, x0 - XPAD, y0, z, 0, b);
    };
    auto issueO = [&](int z) {
        mbar_arrive_expect_tx(&barO, kBytesO);
        tma_load_5d(&S.sO[0][0][0], &mapO, &barO, x0, y0, z, 0, b);
    };
    auto issueT = [&](int z) {
        mbar_arrive_expect_tx(&barT[z & 1], kBytesT);
        tma_load_5d(&S.sT[z & 1][0][0][0], &mapT, &barT[z & 1], x0, y0, z, 0, b);
        tma_load_5d(&S.sW[z & 1][0][0][0], &mapW, &barT[z & 1], x0, y0, z, 0, b);
    };
    if (threadIdx.x == 0) {
        mbar_init(&barY, 1);
        mbar_init(&barX, 1);
        mbar_init(&barO, 1);
        mbar_init(&barT[0], 1);
        mbar_init(&barT[1], 1);
        cntY = cntX = cntO = cntZ = 0;
        mbar_fence_init();
        issueO(0);
        issueY(0);
        issueX(0);
        issueT(0);
        if (D > 1) issueT(1);
    }
    __syncthreads();

    f32x2 ru[5][kC2];  // slot j >= 1: partner unit vector of slice z - j
    float lossacc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) lossacc[k] = 0.f;
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
        for (int i = 0; i < kC2; ++i) ru[s][i] = 0ull;

    float* affs_b = P.affs ? P.affs + (int64_t)b * 12 * N + pix : nullptr;
    float* grad_b = P.grad ? P.grad + (int64_t)b * kC * N + pix : nullptr;

    for (int z = 0; z < D; ++z) {
        const uint32_t par = (uint32_t)(z & 1);
        const int64_t zoff = (int64_t)z * HW;
        const int tb = z & 1;
        f32x2 e[kC2], G[kC2];
        float rp;
        // ---------------- own unit vector (anchor view)
        mbar_wait(&barO, par);
        {
            const float* own = &S.sO[0][ty][tx];
            f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; ++i) e[i] = pack2(own[(2 * i) * CHS_A], own[(2 * i + 1) * CHS_A]);
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                s0 = fma2(e[i], e[i], s0);
                s1 = fma2(e[i + 1], e[i + 1], s1);
            }
            rp = inv_norm(hsum2(s0, s1));
            const f32x2 rp2 = pack2(rp, rp);
#pragma unroll
            for (int i = 0; i < kC2; ++i) {
                e[i] = mul2(e[i], rp2);
                G[i] = 0ull;
            }
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntO) == kWarps * (uint32_t)(z + 1) - 1 && z + 1 < D) issueO(z + 1);

        // ---------------- y part; the partner's own column (ring entry of this slice) comes from the same box
        mbar_wait(&barY, par);
        {
            const float* pc = &S.sY[0][ty + HY][tx];
            f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; ++i) ru[0][i] = pack2(pc[(2 * i) * CHS_Y], pc[(2 * i + 1) * CHS_Y]);
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                s0 = fma2(ru[0][i], ru[0][i], s0);
                s1 = fma2(ru[0][i + 1], ru[0][i + 1], s1);
            }
            const float ru_r = inv_norm(hsum2(s0, s1));
            const f32x2 r2 = pack2(ru_r, ru_r);
#pragma unroll
            for (int i = 0; i < kC2; ++i) ru[0][i] = mul2(ru[0][i], r2);
        }
        mbar_wait(&barT[tb], (uint32_t)((z >> 1) & 1));
        const float* sTo = &S.sT[tb][0][ty][tx];
        const float* sWo = &S.sW[tb][0][ty][tx];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int s = shift_yx(g);
            const int k = 3 * g + 1;
            const float* nb = (g < 3) ? &S.sY[0][ty + HY - s][tx] : &S.sAm[0][ty][tx];
            const bool vf = live && (y >= s);
            float a;
            const float l = inplane_pair<GRAD>(nb, (g < 3) ? CHS_Y : CHS_A, e, G, sTo[k * CHS_A], vf ? sWo[k * CHS_A] : 0.f,
                                               P.gcoef[k], a);
            lossacc[k] += vf ? l : 0.f;
            if (affs_b && live) affs_b[(int64_t)k * N + zoff] = vf ? a : 0.f;
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntY) == kWarps * (uint32_t)(z + 1) - 1 && z + 1 < D) issueY(z + 1);

        // ---------------- x part
        mbar_wait(&barX, par);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int s = shift_yx(g);
            const int k = 3 * g + 2;
            const bool vf = live && (x >= s);
            float a;
            const float l = inplane_pair<GRAD>(&S.sX[0][ty][tx + XPAD - s], CHS_X, e, G, sTo[k * CHS_A],
                                               vf ? sWo[k * CHS_A] : 0.f, P.gcoef[k], a);
            lossacc[k] += vf ? l : 0.f;
            if (affs_b && live) affs_b[(int64_t)k * N + zoff] = vf ? a : 0.f;
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntX) == kWarps * (uint32_t)(z + 1) - 1 && z + 1 < D) issueX(z + 1);

        // ---------------- z part: partner ring
#pragma unroll
        for (int j = 1; j <= 4; ++j) {
            const int k = 3 * (j - 1);
            const bool v = live && (z >= j);
            f32x2 d0 = 0ull, d1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                d0 = fma2(e[i], ru[j][i], d0);
                d1 = fma2(e[i + 1], ru[j][i + 1], d1);
            }
            const float a = hsum2(d0, d1);
            const float diff = a - sTo[k * CHS_A];
            const float wv = v ? sWo[k * CHS_A] : 0.f;
            lossacc[k] += wv * diff * diff;
            if (affs_b && live) affs_b[(int64_t)k * N + zoff] = v ? a : 0.f;
            if (GRAD) {
                const float gg = P.gcoef[k] * wv * diff;
                const f32x2 gg2 = pack2(gg, gg);
#pragma unroll
                for (int i = 0; i < kC2; ++i) G[i] = fma2(gg2, ru[j][i], G[i]);
            }
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntZ) == kWarps * (uint32_t)(z + 1) - 1 && z + 2 < D) issueT(z + 2);

        // ---------------- gradient of this slice is complete: back through the normalisation
        if (GRAD && live) {
            f32x2 p0 = 0ull, p1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                p0 = fma2(e[i], G[i], p0);
                p1 = fma2(e[i + 1], G[i + 1], p1);
            }
            float eg = hsum2(p0, p1);
            if (!(rp < 1e12f)) eg = 0.f;
            const f32x2 neg = pack2(-eg, -eg), r2 = pack2(rp, rp);
            float* dst = grad_b + zoff;
#pragma unroll
            for (int i = 0; i < kC2; ++i) {
                float lo, hi;
                unpack2(mul2(fma2(e[i], neg, G[i]), r2), lo, hi);
                dst[(int64_t)(2 * i) * N] = lo;
                dst[(int64_t)(2 * i + 1) * N] = hi;
            }
        }
#pragma unroll
        for (int s = 4; s >= 1; --s)
#pragma unroll
            for (int i = 0; i < kC2; ++i) ru[s][i] = ru[s - 1][i];
    }

#pragma unroll
    for (int k = 0; k < 12; ++k) {
        const float v = warp_sum(lossacc[k]);
        if (tx == 0) s_red[k][ty] = v;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < TY; ++i) v += s_red[threadIdx.x][i];
        P.partials[(int64_t)blockIdx.x * 12 + threadIdx.x] = v;
    }
}

// Two-view kernel, Tensor-Memory variant: the partner ring (4 older slices x 16 floats) lives in TMEM, which brings the
// kernel from 202 to <= 168 registers and lets a CTA of 12 warps (12 x 32 voxel columns) run per SM.
template <int TY, bool GRAD>
__global__ void __launch_bounds__(TY * 32, 1)
cross3d_norm5_tm(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapA,
              const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapO,
              const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapW, const Params P) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    StageX<TY>& S = *reinterpret_cast<StageX<TY>*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    __shared__ uint64_t barY, barX, barO, barT[2];
    __shared__ uint32_t cntY, cntX, cntO, cntZ, s_tmem;
    __shared__ float s_red[12][TY];
    constexpr uint32_t kBytesY = sizeof(S.sY) + sizeof(S.sAm);
    constexpr uint32_t kBytesX = sizeof(S.sX);
    constexpr uint32_t kBytesO = sizeof(S.sO);
    constexpr uint32_t kBytesT = sizeof(S.sT[0]) + sizeof(S.sW[0]);
    constexpr int CHS_Y = (TY + HY) * TX, CHS_A = TY * TX, XWF = TX + XPAD, CHS_X = TY * XWF;
    constexpr uint32_t kWarps = TY;

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int D = P.D, H = P.H, W = P.W;
    const int tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
    int bid = blockIdx.x;
    const int tix = bid % tiles_x;
    bid /= tiles_x;
    const int tiy = bid % tiles_y;
    const int b = bid / tiles_y;
    const int x0 = tix * TX, y0 = tiy * TY;
    const int x = x0 + tx, y = y0 + ty;
    const bool live = (x < W) && (y < H);
    const int64_t HW = (int64_t)H * W, N = HW * D;
    const int64_t pix = (int64_t)(live ? y : 0) * W + (live ? x : 0);

    auto issueY = [&](int z) {
        mbar_arrive_expect_tx(&barY, kBytesY);
        tma_load_5d(&S.sY[0][0][0], &mapY, &barY, x0, y0 - HY, z, 0, b);
        tma_load_5d(&S.sAm[0][0][0], &mapA, &barY, x0, y0 - FAR, z, 0, b);
    };
    auto issueX = [&](int z) {
        mbar_arrive_expect_tx(&barX, kBytesX);
        tma_load_5d(&S.sX[0][0][0], &mapX, &barX, x0 - XPAD, y0, z, 0, b);
    };
    auto issueO = [&](int z) {
        mbar_arrive_expect_tx(&barO, kBytesO);
        tma_load_5d(&S.sO[0][0][0], &mapO, &barO, x0, y0, z, 0, b);
    };
    auto issueT = [&](int z) {
        mbar_arrive_expect_tx(&barT[z & 1], kBytesT);
        tma_load_5d(&S.sT[z & 1][0][0][0], &mapT, &barT[z & 1], x0, y0, z, 0, b);
        tma_load_5d(&S.sW[z & 1][0][0][0], &mapW, &barT[z & 1], x0, y0, z, 0, b);
    };
    if (threadIdx.x == 0) {
        mbar_init(&barY, 1);
        mbar_init(&barX, 1);
        mbar_init(&barO, 1);
        mbar_init(&barT[0], 1);
        mbar_init(&barT[1], 1);
        cntY = cntX = cntO = cntZ = 0;
        mbar_fence_init();
        issueO(0);
        issueY(0);
        issueX(0);
        issueT(0);
        if (D > 1) issueT(1);
    }
    if (ty == 0) {  // Tensor Memory holds the partner ring (see fused3d_norm5_tm): 4 slots x 16 columns per thread
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = s_tmem;
    const uint32_t tq = tmem_base + (((uint32_t)(ty & 3) * 32u) << 16) + (uint32_t)(ty >> 2) * 64u;

    float lossacc[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) lossacc[k] = 0.f;

    float* affs_b = P.affs ? P.affs + (int64_t)b * 12 * N + pix : nullptr;
    float* grad_b = P.grad ? P.grad + (int64_t)b * kC * N + pix : nullptr;

    for (int z = 0; z < D; ++z) {
        const uint32_t par = (uint32_t)(z & 1);
        const int64_t zoff = (int64_t)z * HW;
        const int tb = z & 1;
        f32x2 e[kC2], G[kC2], u0[kC2];  // u0: partner unit vector of this slice (goes into its Tensor-Memory slot)
        float rp;
        // ---------------- own unit vector (anchor view)
        mbar_wait(&barO, par);
        {
            const float* own = &S.sO[0][ty][tx];
            f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; ++i) e[i] = pack2(own[(2 * i) * CHS_A], own[(2 * i + 1) * CHS_A]);
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                s0 = fma2(e[i], e[i], s0);
                s1 = fma2(e[i + 1], e[i + 1], s1);
            }
            rp = inv_norm(hsum2(s0, s1));
            const f32x2 rp2 = pack2(rp, rp);
#pragma unroll
            for (int i = 0; i < kC2; ++i) {
                e[i] = mul2(e[i], rp2);
                G[i] = 0ull;
            }
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntO) == kWarps * (uint32_t)(z + 1) - 1 && z + 1 < D) issueO(z + 1);

        // ---------------- y part; the partner's own column (ring entry of this slice) comes from the same box
        mbar_wait(&barY, par);
        {
            const float* pc = &S.sY[0][ty + HY][tx];
            f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; ++i) u0[i] = pack2(pc[(2 * i) * CHS_Y], pc[(2 * i + 1) * CHS_Y]);
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                s0 = fma2(u0[i], u0[i], s0);
                s1 = fma2(u0[i + 1], u0[i + 1], s1);
            }
            const float ru_r = inv_norm(hsum2(s0, s1));
            const f32x2 r2 = pack2(ru_r, ru_r);
#pragma unroll
            for (int i = 0; i < kC2; ++i) u0[i] = mul2(u0[i], r2);
        }
        mbar_wait(&barT[tb], (uint32_t)((z >> 1) & 1));
        const float* sTo = &S.sT[tb][0][ty][tx];
        const float* sWo = &S.sW[tb][0][ty][tx];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int s = shift_yx(g);
            const int k = 3 * g + 1;
            const float* nb = (g < 3) ? &S.sY[0][ty + HY - s][tx] : &S.sAm[0][ty][tx];
            const bool vf = live && (y >= s);
            float a;
            const float l = inplane_pair<GRAD>(nb, (g < 3) ? CHS_Y : CHS_A, e, G, sTo[k * CHS_A], vf ? sWo[k * CHS_A] : 0.f,
                                               P.gcoef[k], a);
            lossacc[k] += vf ? l : 0.f;
            if (affs_b && live) affs_b[(int64_t)k * N + zoff] = vf ? a : 0.f;
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntY) == kWarps * (uint32_t)(z + 1) - 1 && z + 1 < D) issueY(z + 1);

        // ---------------- x part
        mbar_wait(&barX, par);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int s = shift_yx(g);
            const int k = 3 * g + 2;
            const bool vf = live && (x >= s);
            float a;
            const float l = inplane_pair<GRAD>(&S.sX[0][ty][tx + XPAD - s], CHS_X, e, G, sTo[k * CHS_A],
                                               vf ? sWo[k * CHS_A] : 0.f, P.gcoef[k], a);
            lossacc[k] += vf ? l : 0.f;
            if (affs_b && live) affs_b[(int64_t)k * N + zoff] = vf ? a : 0.f;
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntX) == kWarps * (uint32_t)(z + 1) - 1 && z + 1 < D) issueX(z + 1);

        // ---------------- z part: partner vectors of slices z - 1 .. z - 4 come out of Tensor Memory
#pragma unroll
        for (int j = 1; j <= 4; ++j) {
            const int k = 3 * (j - 1);
            if (z >= j) {  // CTA-uniform
                f32x2 uj[kC2];
                tmem_ld16(tq + ((uint32_t)((z - j) & 3) << 4), uj);
                f32x2 d0 = 0ull, d1 = 0ull;
#pragma unroll
                for (int i = 0; i < kC2; i += 2) {
                    d0 = fma2(e[i], uj[i], d0);
                    d1 = fma2(e[i + 1], uj[i + 1], d1);
                }
                const float a = hsum2(d0, d1);
                const float diff = a - sTo[k * CHS_A];
                const float wv = live ? sWo[k * CHS_A] : 0.f;
                lossacc[k] += wv * diff * diff;
                if (affs_b && live) affs_b[(int64_t)k * N + zoff] = a;
                if (GRAD) {
                    const float gg = P.gcoef[k] * wv * diff;
                    const f32x2 gg2 = pack2(gg, gg);
#pragma unroll
                    for (int i = 0; i < kC2; ++i) G[i] = fma2(gg2, uj[i], G[i]);
                }
            } else if (affs_b && live) {
                affs_b[(int64_t)k * N + zoff] = 0.f;
            }
        }
        tmem_st16(tq + ((uint32_t)(z & 3) << 4), u0);  // slot of slice z - 4, no longer needed
        tmem_wait_st();
        __syncwarp();
        if (tx == 0 && arrive_count(&cntZ) == kWarps * (uint32_t)(z + 1) - 1 && z + 2 < D) issueT(z + 2);

        // ---------------- gradient of this slice is complete: back through the normalisation
        if (GRAD && live) {
            f32x2 p0 = 0ull, p1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                p0 = fma2(e[i], G[i], p0);
                p1 = fma2(e[i + 1], G[i + 1], p1);
            }
            float eg = hsum2(p0, p1);
            if (!(rp < 1e12f)) eg = 0.f;
            const f32x2 neg = pack2(-eg, -eg), r2 = pack2(rp, rp);
            float* dst = grad_b + zoff;
            if (P.accumulate) {  // CTA-uniform: add to what the one-view kernel left there (same thread, same address)
                float prev[kC];
#pragma unroll
                for (int c = 0; c < kC; ++c) prev[c] = dst[(int64_t)c * N];
#pragma unroll
                for (int i = 0; i < kC2; ++i) {
                    float lo, hi;
                    unpack2(mul2(fma2(e[i], neg, G[i]), r2), lo, hi);
                    dst[(int64_t)(2 * i) * N] = prev[2 * i] + lo;
                    dst[(int64_t)(2 * i + 1) * N] = prev[2 * i + 1] + hi;
                }
            } else {
#pragma unroll
            for (int i = 0; i < kC2; ++i) {
                float lo, hi;
                unpack2(mul2(fma2(e[i], neg, G[i]), r2), lo, hi);
                dst[(int64_t)(2 * i) * N] = lo;
                dst[(int64_t)(2 * i + 1) * N] = hi;
            }
            }
        }
    }

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
}


}  // namespace t3

// ---------------------------------------------------------------------------------------------------------------
bool tiled3d_norm5_eligible(int C, int D, int H, int W, const OffsetTable& tab, int border, int mask_dtype, bool cross,
                            const float* emb, const float* target, const float* weight) {
    static const int ref[12][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {-2, 0, 0}, {0, -3, 0}, {0, 0, -3},
                                   {-3, 0, 0}, {0, -9, 0}, {0, 0, -9}, {-4, 0, 0}, {0, -27, 0}, {0, 0, -27}};
    (void)cross;  // both the self and the two-view variant have a tiled kernel (gradient w.r.t. the anchor view only)
    if (C != 16 || tab.K != 12 || border != PEA_BORDER_ZERO || mask_dtype != PEA_MASK_NONE) return false;
    // TMA: 16-byte row pitch and base alignment for all three tensors
    if (W % 4 != 0 || ((reinterpret_cast<uintptr_t>(emb) | reinterpret_cast<uintptr_t>(target) |
                        reinterpret_cast<uintptr_t>(weight)) & 15)) return false;
    for (int k = 0; k < 12; ++k)
        if (tab.dz[k] != ref[k][0] || tab.dy[k] != ref[k][1] || tab.dx[k] != ref[k][2]) return false;
    (void)D; (void)H;
    return true;
}

constexpr int kTileY = PEA_TILE_Y3D;  // rows per CTA of the fused 3-D kernels (tuned on B200)

int tiled3d_norm5_blocks(int B, int H, int W) {
    constexpr int TY = kTileY;
    return B * ((H + TY - 1) / TY) * ((W + t3::TX - 1) / t3::TX);
}

int launch_fused3d_norm5_rf(const float* emb, const Dims& d, const OffsetTable& tab, const float* target,
                         const float* weight, float* affs, float* partials, float* grad, cudaStream_t st) {
    constexpr int TY = kTileY;
    CUtensorMap mY, mA, mX, mT, mW;
    int rc;
    if ((rc = encode_volume_map(&mY, emb, d.B, 16, d.D, d.H, d.W, t3::TX, TY + 2 * t3::HY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mA, emb, d.B, 16, d.D, d.H, d.W, t3::TX, TY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mX, emb, d.B, 16, d.D, d.H, d.W, t3::XW, TY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mT, target, d.B, 12, d.D, d.H, d.W, t3::TX, TY, 1, 12))) return rc;
    if ((rc = encode_volume_map(&mW, weight, d.B, 12, d.D, d.H, d.W, t3::TX, TY, 1, 12))) return rc;
    t3::Params P;
    P.B = d.B; P.D = d.D; P.H = d.H; P.W = d.W;
    P.target = target; P.weight = weight; P.affs = affs; P.grad = grad; P.partials = partials; P.wtab = nullptr; P.accumulate = 0;
    for (int k = 0; k < 12; ++k) P.gcoef[k] = tab.gcoef[k];
    const int grid = tiled3d_norm5_blocks(d.B, d.H, d.W);
    const size_t smem = sizeof(t3::Stage<TY>) + 1024;
    cudaError_t e;
    if (grad) {
        auto kern = t3::fused3d_norm5<TY, true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, TY * 32, smem, st>>>(mY, mA, mX, mT, mW, P);
    } else {
        auto kern = t3::fused3d_norm5<TY, false>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, TY * 32, smem, st>>>(mY, mA, mX, mT, mW, P);
    }
    return (int)cudaGetLastError();
}


#ifndef PEA_K4_TMEM
#define PEA_K4_TMEM 1
#endif
#ifndef PEA_TILE_Y3D_TM
#define PEA_TILE_Y3D_TM 12
#endif
constexpr int kTileYTm = PEA_TILE_Y3D_TM;  // rows (= warps) per CTA of the Tensor-Memory variant

// number of CTAs (= rows of loss partials) of the self (one-view) fused kernel
int fused3d_norm5_blocks(int B, int H, int W) {
#if PEA_K4_TMEM
    return B * ((H + kTileYTm - 1) / kTileYTm) * ((W + t3::TX - 1) / t3::TX);
#else
    return tiled3d_norm5_blocks(B, H, W);
#endif
}

int launch_fused3d_norm5(const float* emb, const Dims& d, const OffsetTable& tab, const float* target,
                         const float* weight, float* affs, float* partials, float* grad, cudaStream_t st) {
#if !PEA_K4_TMEM
    return launch_fused3d_norm5_rf(emb, d, tab, target, weight, affs, partials, grad, st);
#else
    constexpr int TY = kTileYTm;
    CUtensorMap mY, mA, mX, mT, mW;
    int rc;
    if ((rc = encode_volume_map(&mY, emb, d.B, 16, d.D, d.H, d.W, t3::TX, TY + 2 * t3::HY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mA, emb, d.B, 16, d.D, d.H, d.W, t3::TX, TY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mX, emb, d.B, 16, d.D, d.H, d.W, t3::XW, TY, 1, 16))) return rc;
    // every third channel from the start coordinate: 4 planes (c0, c0 + 3, c0 + 6, c0 + 9) per copy
    if ((rc = encode_volume_map(&mT, target, d.B, 12, d.D, d.H, d.W, t3::TX, TY, 1, 10, 0, 3))) return rc;
    if ((rc = encode_volume_map(&mW, weight, d.B, 12, d.D, d.H, d.W, t3::TX, TY, 1, 10, 0, 3))) return rc;
    CUtensorMap mT1, mW1;
    if ((rc = encode_volume_map(&mT1, target, d.B, 12, d.D, d.H, d.W, t3::TX, TY, 1, 1))) return rc;
    if ((rc = encode_volume_map(&mW1, weight, d.B, 12, d.D, d.H, d.W, t3::TX, TY, 1, 1))) return rc;
    t3::Params P;
    P.B = d.B; P.D = d.D; P.H = d.H; P.W = d.W;
    P.target = target; P.weight = weight; P.affs = affs; P.grad = grad; P.partials = partials; P.wtab = nullptr; P.accumulate = 0;
    for (int k = 0; k < 12; ++k) P.gcoef[k] = tab.gcoef[k];
    const int grid = fused3d_norm5_blocks(d.B, d.H, d.W);
    const size_t smem = sizeof(t3::StageM<TY>) + 128;
    cudaError_t e;
    if (grad) {
        auto kern = t3::fused3d_norm5_tm<TY, true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, (TY + PEA_K4_DUP) * 32, smem, st>>>(mY, mA, mX, mT, mW, mT1, mW1, P);
    } else {
        auto kern = t3::fused3d_norm5_tm<TY, false>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, (TY + PEA_K4_DUP) * 32, smem, st>>>(mY, mA, mX, mT, mW, mT1, mW1, P);
    }
    return (int)cudaGetLastError();
#endif
}

// Label-driven flavour (SURVEY.md 8(f) item 1 for the 3-D path): target = seg_to_aff(seg, NHOOD_NORM5, pad='') and its
// per-channel weight_binary_ratio are derived inside the loss kernel from the int32 segmentation.
int launch_fused3d_norm5_labels(const float* emb, const int32_t* seg, const float* wtab, const Dims& d, const OffsetTable& tab,
                                float* affs, float* partials, float* grad, cudaStream_t st) {
    constexpr int TY = kTileYTm;
    const float* lab = reinterpret_cast<const float*>(seg);  // 4-byte elements, copied bit for bit
    CUtensorMap mY, mA, mX, mLy, mLx;
    int rc;
    if ((rc = encode_volume_map(&mY, emb, d.B, 16, d.D, d.H, d.W, t3::TX, TY + 2 * t3::HY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mA, emb, d.B, 16, d.D, d.H, d.W, t3::TX, TY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mX, emb, d.B, 16, d.D, d.H, d.W, t3::XW, TY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mLy, lab, d.B, 1, d.D, d.H, d.W, t3::TX, TY + 2 * t3::FAR, 1, 1))) return rc;
    if ((rc = encode_volume_map(&mLx, lab, d.B, 1, d.D, d.H, d.W, t3::XW, TY, 1, 1))) return rc;
    t3::Params P;
    P.B = d.B; P.D = d.D; P.H = d.H; P.W = d.W;
    P.target = nullptr; P.weight = nullptr; P.affs = affs; P.grad = grad; P.partials = partials; P.wtab = wtab; P.accumulate = 0;
    for (int k = 0; k < 12; ++k) P.gcoef[k] = tab.gcoef[k];
    const int grid = fused3d_norm5_blocks(d.B, d.H, d.W);
    const size_t smem = sizeof(t3::StageM<TY>) + 128;
    cudaError_t e;
    if (grad) {
        auto kern = t3::fused3d_norm5_tm<TY, true, true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, TY * 32, smem, st>>>(mY, mA, mX, mLy, mLx, mLy, mLx, P);
    } else {
        auto kern = t3::fused3d_norm5_tm<TY, false, true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, TY * 32, smem, st>>>(mY, mA, mX, mLy, mLx, mLy, mLx, P);
    }
    return (int)cudaGetLastError();
}

bool tiled3d_fwd_eligible(int C, int W, const OffsetTable& tab, int border, bool cross, const float* emb,
                          int64_t emb_plane) {
    static const int ref[12][3] = {{-1, 0, 0}, {0, -1, 0}, {0, 0, -1}, {-2, 0, 0}, {0, -3, 0}, {0, 0, -3},
                                   {-3, 0, 0}, {0, -9, 0}, {0, 0, -9}, {-4, 0, 0}, {0, -27, 0}, {0, 0, -27}};
    if (C != 16 || tab.K != 12 || border != PEA_BORDER_ZERO || cross) return false;
    if (W % 4 != 0 || (reinterpret_cast<uintptr_t>(emb) & 15) || (emb_plane % 4) != 0) return false;
    for (int k = 0; k < 12; ++k)
        if (tab.dz[k] != ref[k][0] || tab.dy[k] != ref[k][1] || tab.dx[k] != ref[k][2]) return false;
    return true;
}

// Waves of resident CTAs (2 per SM) whose odd CTAs use the rotated slice order with an external halo.  Only the FIRST
// wave starts in lock step (296 CTAs asking for 19 MB of halo at t = 0); later CTAs start whenever an earlier one ends,
// so their halo requests are spread already and the rotation only costs them a second ring warm-up (134 MB of re-reads
// per 13-slice slab at 1024 x 1024).  2 x B200, 13-slice slabs, full step (tools/zshard_staged.py 26): no rotation
// 0.580 ms, every CTA (the rule up to ABI 6) 0.561 - 0.566, first wave only 0.547, two / three / five waves 0.556 /
// 0.555 / 0.560; start slices spread over D - 3 values need a run-time rotation (128 registers): slower throughout.
constexpr int kRotWavesDefault = 1;

int launch_fwd3d_norm5(const float* emb, const Dims& d, int z_skip, int64_t emb_plane, int64_t affs_plane, float* affs,
                       const float* halo, int halo_n, int64_t halo_plane, int relu, cudaStream_t st,
                       const StagedHalo* stage) {
#ifndef PEA_TILE_Y3D_FWD
#define PEA_TILE_Y3D_FWD 8
#endif
    constexpr int TY = PEA_TILE_Y3D_FWD;
    CUtensorMap mY, mA, mX;
    int rc;
    if ((rc = encode_volume_map(&mY, emb, d.B, 16, d.D, d.H, d.W, t3::TX, TY + t3::HY, 1, 16, emb_plane))) return rc;
    if ((rc = encode_volume_map(&mA, emb, d.B, 16, d.D, d.H, d.W, t3::TX, TY, 1, 16, emb_plane))) return rc;
    if ((rc = encode_volume_map(&mX, emb, d.B, 16, d.D, d.H, d.W, t3::TX + t3::XPAD, TY, 1, 16, emb_plane))) return rc;
    t3::ParamsF P;
    P.B = d.B; P.D = d.D; P.H = d.H; P.W = d.W; P.z_skip = z_skip;
    P.emb_plane = emb_plane; P.affs_plane = affs_plane; P.emb = emb; P.affs = affs;
    P.halo = halo; P.halo_n = halo_n; P.halo_plane = halo_plane; P.relu = relu;
    {   // which CTAs take the rotated slice order (kernel comment "Slice order"): see kRotWavesDefault
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int knob = debug_knob(PEA_KNOB_K3_ROT_WAVES);
        const int waves = knob == 0 ? kRotWavesDefault : knob;
        P.rot_ctas = waves < 0 ? 0 : (waves >= (1 << 20) / (2 * sms) ? 1 << 30 : waves * 2 * sms);
    }
    P.stage = nullptr; P.stage_plane = 0; P.stage_flags = nullptr; P.stage_epoch = 0; P.stage_row0 = 0; P.stage_band_rows = 1;
    if (stage && halo) {   // bands must start on tile rows: a tile waits for exactly one flag
        if (stage->row0 % TY || stage->band_rows % TY || stage->band_rows <= 0) return PEA_E_SHAPE;
        P.stage = stage->data; P.stage_plane = stage->plane; P.stage_flags = stage->flags; P.stage_epoch = stage->epoch;
        P.stage_row0 = stage->row0; P.stage_band_rows = stage->band_rows;
    }
    const int grid = d.B * ((d.H + TY - 1) / TY) * ((d.W + t3::TX - 1) / t3::TX);
    const size_t smem = sizeof(t3::StageF<TY>) + 1024;
    // (round 2 measured a Tensor-Memory ring + halo prefetch for this kernel as well: bit-identical, but 4.11 ms against
    // 3.70 ms -- four dependent tcgen05.ld per slice cost more than the register ring's moves)
    auto kern = t3::fwd3d_norm5<TY>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, TY * 32, smem, st>>>(mY, mA, mX, P);
    return (int)cudaGetLastError();
}

#ifndef PEA_K4X_TMEM
#define PEA_K4X_TMEM 1
#endif
constexpr int kTileYX = PEA_K4X_TMEM ? 12 : kTileY;  // rows per CTA of the two-view kernel

int cross3d_norm5_blocks(int B, int H, int W) { return B * ((H + kTileYX - 1) / kTileYX) * ((W + t3::TX - 1) / t3::TX); }

int launch_cross3d_norm5(const float* emb, const float* partner, const Dims& d, const OffsetTable& tab,
                         const float* target, const float* weight, float* affs, float* partials, float* grad,
                         int accumulate, cudaStream_t st) {
    constexpr int TY = kTileYX;
    if (reinterpret_cast<uintptr_t>(partner) & 15) return PEA_E_ALIGN;
    CUtensorMap mY, mA, mX, mO, mT, mW;
    int rc;
    if ((rc = encode_volume_map(&mY, partner, d.B, 16, d.D, d.H, d.W, t3::TX, TY + t3::HY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mA, partner, d.B, 16, d.D, d.H, d.W, t3::TX, TY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mX, partner, d.B, 16, d.D, d.H, d.W, t3::TX + t3::XPAD, TY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mO, emb, d.B, 16, d.D, d.H, d.W, t3::TX, TY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mT, target, d.B, 12, d.D, d.H, d.W, t3::TX, TY, 1, 12))) return rc;
    if ((rc = encode_volume_map(&mW, weight, d.B, 12, d.D, d.H, d.W, t3::TX, TY, 1, 12))) return rc;
    t3::Params P;
    P.B = d.B; P.D = d.D; P.H = d.H; P.W = d.W;
    P.target = target; P.weight = weight; P.affs = affs; P.grad = grad; P.partials = partials; P.wtab = nullptr; P.accumulate = (accumulate && grad) ? 1 : 0;
    for (int k = 0; k < 12; ++k) P.gcoef[k] = tab.gcoef[k];
    const int grid = cross3d_norm5_blocks(d.B, d.H, d.W);
    const size_t smem = sizeof(t3::StageX<TY>) + 1024;
    cudaError_t e;
#if PEA_K4X_TMEM
    if (grad) {
        auto kern = t3::cross3d_norm5_tm<TY, true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, TY * 32, smem, st>>>(mY, mA, mX, mO, mT, mW, P);
    } else {
        auto kern = t3::cross3d_norm5_tm<TY, false>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, TY * 32, smem, st>>>(mY, mA, mX, mO, mT, mW, P);
    }
    return (int)cudaGetLastError();
#endif
    if (grad) {
        auto kern = t3::cross3d_norm5<TY, true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, TY * 32, smem, st>>>(mY, mA, mX, mO, mT, mW, P);
    } else {
        auto kern = t3::cross3d_norm5<TY, false>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, TY * 32, smem, st>>>(mY, mA, mX, mO, mT, mW, P);
    }
    return (int)cudaGetLastError();
}

}  // namespace pea
