// Fused 2-D kernel for AXIS-ALIGNED offset tables -- multi_offset(shifts, neighbor=4), every shipped 2-D config
// (scripts_cvppp/utils/affinity_ours.py:4-15, cvppp.yaml:69-75, bbbc039v1.yaml:65-71): embedding -> K affinities ->
// masked weighted MSE -> d loss / d embedding in one pass (embedding_loss / ema_embedding_loss,
// scripts_cvppp/loss/loss_embedding_mse.py:18-47, 79-95).
//
// Design (DESIGN.md "K2a"): the 3-D kernel's recipe without the z march --
//   * a CTA owns TY x 32 pixels of one sample; the RAW channel-planar embedding of a y-column box (TY + 2 PMAX rows)
//     and an x-row box (32 + 2 PX columns) arrives by TMA (zero fill outside the image == zero border); nothing is
//     normalised in shared memory: a neighbour's 1/|x| is recomputed in registers (8 FFMA2 + MUFU), which is cheaper
//     than a fill phase with a barrier and keeps the whole staging asynchronous;
//   * circular border (torch.roll): only the CTAs whose boxes cross the image edge (~15 % at 520 x 696) fill their boxes
//     by hand through the wrap rule; interior CTAs take the TMA path;
//   * runtime offset table, two compact loops (y offsets, x offsets): per offset one forward pair and -- self variant
//     with gradient -- one reverse pair (gather-form backward), target / weight / mask fetched two offsets ahead;
//   * per-offset loss terms -> shared scratch -> fixed-order CTA sum -> one row of partials per CTA.
#include <cstdlib>
#include "pea_common.cuh"
#include "pea_tma.cuh"

#ifndef PEA_AXIS_DEPTH
#define PEA_AXIS_DEPTH 3  // rotating target / weight / mask fetch buffers of the persistent kernel (3 or 4)
#endif

namespace pea {
namespace a2 {

// one affinity, predicated without a branch (a predicated-off store touches no memory, so a null base is fine): the `if`
// form ends a basic block after every pair evaluation
#ifndef PEA_AXIS_BRANCHY_STORE
#define PEA_AXIS_BRANCHY_STORE 0
#endif
__device__ __forceinline__ void st_affs(float* p, float v, bool ok) {
#if PEA_AXIS_BRANCHY_STORE
    if (ok) *p = v;
#else
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"((int)ok));
#endif
}

constexpr int kC = 16;
constexpr int kC2 = 8;
constexpr int TX = 32;
constexpr int kMaxK = 12;  // offsets this kernel accepts (longer tables: fused2d_tiled)

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float hsum2(f32x2 a, f32x2 b) {
    float a0, a1, b0, b1;
    unpack2(a, a0, a1);
    unpack2(b, b0, b1);
    return (a0 + a1) + (b0 + b1);
}
__device__ __forceinline__ float inv_norm(float ss) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(ss));
    return fminf(r, 1e12f);  // == 1 / max(sqrt(ss), 1e-12)
}

// predicated read-only loads the compiler may not sink towards their use (volatile): issued two offsets ahead
__device__ __forceinline__ float ldg_f32_early(const float* p, bool pred) {
    float v;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}"
        : "=f"(v)
        : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ int32_t ldg_i32_early(const int32_t* p, bool pred) {
    int32_t v;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.s32 %0, 0;\n\t@q ld.global.nc.s32 %0, [%1];\n\t}"
        : "=r"(v)
        : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ uint32_t ldg_u8_early(const uint8_t* p, bool pred) {
    uint32_t v;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.u32 %0, 0;\n\t@q ld.global.nc.u8 %0, [%1];\n\t}"
        : "=r"(v)
        : "l"(p), "r"((int)pred));
    return v;
}

// v mod n for -n <= v < 2n
__device__ __forceinline__ int wrap1(int v, int n) {
    v += (v < 0) ? n : 0;
    v -= (v >= n) ? n : 0;
    return v;
}
// -n <= v < 4n
__device__ __forceinline__ int wrap3(int v, int n) {
    v = wrap1(v, n);
    v -= (v >= n) ? n : 0;
    v -= (v >= n) ? n : 0;
    return v;
}

struct Params {
    int B, H, W, K;
    int wrap;
    int ny, nx;               // number of y offsets / x offsets
    const float* emb;         // anchor view
    const float* par;         // staged view (== emb for the self variant)
    const float* target;
    const float* weight;
    const void* mask;
    float* affs;              // may be null
    float* grad;              // may be null
    float* partials;          // [gridDim.x][K]
    // y offsets first, then x offsets: channel id, signed shift, 2 * coef / norm of that channel
    int chan[kMaxK], shift[kMaxK];
    float gcoef[kMaxK];
    // label-driven flavour (LAB): target / mask / weight derived in the kernel from the int32 label plane (gen_affs_ours
    // semantics) and the per-(sample, channel) weight pairs of pea_gt_pair_weights; target / weight / mask are null
    const int32_t* labels;
    const float* wtab;  // [B][K][2]: weight where the target is 1 / 0
    int ignore;
};

template <int TY, int PMAX>
struct Geo {
    static constexpr int PX = (PMAX + 3) / 4 * 4;  // x halo rounded up to 16 bytes (TMA start coordinate)
    static constexpr int YR = TY + 2 * PMAX;
    static constexpr int XW = TX + 2 * PX;
    static constexpr int CHS_Y = YR * TX;
    static constexpr int CHS_X = TY * XW;
    static constexpr int NT = TY * 32;
};

template <int TY, int PMAX>
struct alignas(128) Stage {
    float sY[kC][Geo<TY, PMAX>::YR][TX];   // rows y0 - PMAX .. y0 + TY + PMAX - 1
    float sX[kC][TY][Geo<TY, PMAX>::XW];   // cols x0 - PX .. x0 + 31 + PX
    float sL[kMaxK][TY * 32];              // per-offset loss terms of every pixel
};

// One neighbour q of the own pixel p: a = <e^_p, e^_q>, G += gcoef * w * (a - t) / |x_q| * x_q; returns w * (a - t)^2.
template <int CHS, bool GRAD>
__device__ __forceinline__ float pair(const float* __restrict__ nb, const f32x2 (&e)[kC2], f32x2 (&G)[kC2], float t, float w,
                                      float gco, float& a_out) {
    f32x2 xq[kC2];
#pragma unroll
    for (int i = 0; i < kC2; ++i) xq[i] = pack2(nb[(2 * i) * CHS], nb[(2 * i + 1) * CHS]);
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
    if (GRAD) {
        const float gr = gco * w * diff * rq;
        const f32x2 gr2 = pack2(gr, gr);
#pragma unroll
        for (int i = 0; i < kC2; ++i) G[i] = fma2(gr2, xq[i], G[i]);
    }
    return w * diff * diff;
}

struct TrueT { static constexpr bool value = true; };
struct FalseT { static constexpr bool value = false; };

template <int MASK>
__device__ __forceinline__ float mask_sq(uint32_t raw) {
    if (MASK == PEA_MASK_NONE) return 1.f;
    const float m = MASK == PEA_MASK_U8 ? (float)raw : __uint_as_float(raw);
    return m * m;
}

// unpredicated flavours for CTAs whose pixels and neighbours all lie inside the image
__device__ __forceinline__ float ldg_f32_now(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ldg_u8_now(const uint8_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// target / weight / mask of the pair anchored at the own pixel (channel of the offset) and of the pair anchored at
// p - o whose partner is the own pixel, for one offset
struct KV {
    float t, w, t2, w2;
    uint32_t m, m2;
};

template <int TY, int PMAX, bool CROSS, int MASK, bool GRAD>
__global__ void __launch_bounds__(TY * 32, (TY <= 8) ? 2 : 1)
fused2d_axis(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapX, const Params P) {
    using G_ = Geo<TY, PMAX>;
    constexpr int PX = G_::PX, YR = G_::YR, XW = G_::XW, CHS_Y = G_::CHS_Y, CHS_X = G_::CHS_X, NT = G_::NT;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Stage<TY, PMAX>& S = *reinterpret_cast<Stage<TY, PMAX>*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    __shared__ uint64_t barY, barX;

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int H = P.H, W = P.W, K = P.K;
    const int tix = blockIdx.x, tiy = blockIdx.y, b = blockIdx.z;  // 3-D grid: no div / mod chain
    const int lin_block = (b * (int)gridDim.y + tiy) * (int)gridDim.x + tix;
    const int x0 = tix * TX, y0 = tiy * TY;
    const int x = x0 + tx, y = y0 + ty;
    const bool live = (x < W) && (y < H);
    const bool wrap = P.wrap != 0;
    const uint32_t Nu = (uint32_t)(H * W);
    const uint32_t pix = (uint32_t)((live ? y : 0) * W + (live ? x : 0));
    const int ny = P.ny, ntot = P.ny + P.nx;

    // ---- staging.  `inner`: the tile and both boxes lie inside the image -- no border rule, no dead lanes.  TMA is used
    //      whenever zero fill is the right answer outside the image (always under the zero border); under the circular
    //      border the CTAs on the image edge fill their boxes by hand through the wrap rule.
    const bool inner = y0 >= PMAX && y0 + TY + PMAX <= H && x0 >= PX && x0 + TX + PX <= W;
    const bool manual = wrap && !inner;
    if (threadIdx.x == 0) {
        mbar_init(&barY, 1);
        mbar_init(&barX, 1);
        mbar_fence_init();
        if (!manual) {
            mbar_arrive_expect_tx(&barY, (uint32_t)sizeof(S.sY));
            tma_load_5d(&S.sY[0][0][0], &mapY, &barY, x0, y0 - PMAX, 0, 0, b);
            mbar_arrive_expect_tx(&barX, (uint32_t)sizeof(S.sX));
            tma_load_5d(&S.sX[0][0][0], &mapX, &barX, x0 - PX, y0, 0, 0, b);
        }
    }
    const float* tgt_b = P.target + (int64_t)b * K * Nu + pix;
    const float* wgt_b = P.weight + (int64_t)b * K * Nu + pix;
    const uint8_t* m8_b = MASK == PEA_MASK_U8 ? reinterpret_cast<const uint8_t*>(P.mask) + (int64_t)b * K * Nu + pix : nullptr;
    const float* mf_b = MASK == PEA_MASK_F32 ? reinterpret_cast<const float*>(P.mask) + (int64_t)b * K * Nu + pix : nullptr;

    // fast flavour: everything valid, the reverse anchor is a fixed step away
    auto fetch_inner = [&](int i) -> KV {
        KV v = {0.f, 0.f, 0.f, 0.f, 0u, 0u};
        if (i >= ntot) return v;
        const int s = P.shift[i];
        const uint32_t koff = (uint32_t)P.chan[i] * Nu;
        const int64_t back = i >= ny ? (int64_t)s : (int64_t)s * W;  // p - o in elements
        v.t = ldg_f32_now(tgt_b + koff);
        v.w = ldg_f32_now(wgt_b + koff);
        if (MASK == PEA_MASK_U8) v.m = ldg_u8_now(m8_b + koff);
        if (MASK == PEA_MASK_F32) v.m = __float_as_uint(ldg_f32_now(mf_b + koff));
        if (GRAD && !CROSS) {
            v.t2 = ldg_f32_now(tgt_b + koff - back);
            v.w2 = ldg_f32_now(wgt_b + koff - back);
            if (MASK == PEA_MASK_U8) v.m2 = ldg_u8_now(m8_b + koff - back);
            if (MASK == PEA_MASK_F32) v.m2 = __float_as_uint(ldg_f32_now(mf_b + koff - back));
        }
        return v;
    };
    // general flavour: border rule and dead lanes
    auto fetch_edge = [&](int i) -> KV {
        KV v = {0.f, 0.f, 0.f, 0.f, 0u, 0u};
        if (i >= ntot) return v;
        const bool isx = i >= ny;
        const int s = P.shift[i];
        const uint32_t koff = (uint32_t)P.chan[i] * Nu;
        bool vf = live;
        if (!wrap) vf = vf && (isx ? (unsigned)(x + s) < (unsigned)W : (unsigned)(y + s) < (unsigned)H);
        v.t = ldg_f32_early(tgt_b + koff, vf);
        v.w = ldg_f32_early(wgt_b + koff, vf);
        if (MASK == PEA_MASK_U8) v.m = ldg_u8_early(m8_b + koff, vf);
        if (MASK == PEA_MASK_F32) v.m = __float_as_uint(ldg_f32_early(mf_b + koff, vf));
        if (GRAD && !CROSS) {
            int qy = y, qx = x;
            bool vr = live;
            if (isx) {
                qx = x - s;
                if (wrap) qx = wrap1(qx, W);
                else vr = vr && (unsigned)qx < (unsigned)W;
            } else {
                qy = y - s;
                if (wrap) qy = wrap1(qy, H);
                else vr = vr && (unsigned)qy < (unsigned)H;
            }
            const int64_t i2 = (int64_t)koff + (vr ? (int64_t)(qy * W + qx) - (int64_t)pix : 0);
            v.t2 = ldg_f32_early(tgt_b + i2, vr);
            v.w2 = ldg_f32_early(wgt_b + i2, vr);
            if (MASK == PEA_MASK_U8) v.m2 = ldg_u8_early(m8_b + i2, vr);
            if (MASK == PEA_MASK_F32) v.m2 = __float_as_uint(ldg_f32_early(mf_b + i2, vr));
        }
        return v;
    };

    if (manual) {
        // two staged pixels per iteration: 32 independent loads in flight (the chain of L2 round trips is what an edge
        // CTA waits for)
        const float* src = (CROSS ? P.par : P.emb) + (int64_t)b * kC * Nu;
        float* sYf = &S.sY[0][0][0];
        float* sXf = &S.sX[0][0][0];
        for (int idx = threadIdx.x; idx < YR * TX; idx += 2 * NT) {
            const int idx2 = idx + NT;
            const bool two = idx2 < YR * TX;
            const int row = idx >> 5, col = idx & 31, row2 = (two ? idx2 : idx) >> 5;
            const uint32_t gx = (uint32_t)wrap3(x0 + col, W);
            const uint32_t o = (uint32_t)(wrap3(y0 - PMAX + row, H) * W) + gx;
            const uint32_t o2 = (uint32_t)(wrap3(y0 - PMAX + row2, H) * W) + gx;
            float v[kC], v2[kC];
#pragma unroll
            for (int c = 0; c < kC; ++c) {
                v[c] = __ldg(src + (o + (uint32_t)c * Nu));
                v2[c] = __ldg(src + (o2 + (uint32_t)c * Nu));
            }
#pragma unroll
            for (int c = 0; c < kC; ++c) {
                sYf[c * CHS_Y + idx] = v[c];
                if (two) sYf[c * CHS_Y + idx2] = v2[c];
            }
        }
        for (int idx = threadIdx.x; idx < TY * XW; idx += 2 * NT) {
            const int idx2 = idx + NT;
            const bool two = idx2 < TY * XW;
            const int row = idx / XW, col = idx - row * XW;
            const int j2 = two ? idx2 : idx, row2 = j2 / XW, col2 = j2 - row2 * XW;
            const uint32_t o = (uint32_t)(wrap3(y0 + row, H) * W + wrap3(x0 - PX + col, W));
            const uint32_t o2 = (uint32_t)(wrap3(y0 + row2, H) * W + wrap3(x0 - PX + col2, W));
            float v[kC], v2[kC];
#pragma unroll
            for (int c = 0; c < kC; ++c) {
                v[c] = __ldg(src + (o + (uint32_t)c * Nu));
                v2[c] = __ldg(src + (o2 + (uint32_t)c * Nu));
            }
#pragma unroll
            for (int c = 0; c < kC; ++c) {
                sXf[c * CHS_X + idx] = v[c];
                if (two) sXf[c * CHS_X + idx2] = v2[c];
            }
        }
    }
    const float* ownY = &S.sY[0][ty + PMAX][tx];
    const float* ownX = &S.sX[0][ty][tx + PX];
    float* affs_b = P.affs ? P.affs + (int64_t)b * K * Nu + pix : nullptr;
    f32x2 e[kC2], G[kC2];
    float rp = 0.f;

    // ---- everything after the staging, in the two flavours of the target / weight / mask fetch
    auto run = [&](auto inner_c) {
        constexpr bool INNER = decltype(inner_c)::value;
        auto fetch = [&](int i) -> KV { return INNER ? fetch_inner(i) : fetch_edge(i); };
        KV kv0 = fetch(0), kv1 = fetch(1), kv2;
        __syncthreads();  // barrier init visible / hand-filled boxes complete
        if (!manual) mbar_wait(&barY, 0);
        // own unit vector
        if (CROSS) {
            const float* eb = P.emb + (int64_t)b * kC * Nu;
#pragma unroll
            for (int i = 0; i < kC2; ++i)
                e[i] = pack2(live ? __ldg(eb + (pix + (uint32_t)(2 * i) * Nu)) : 0.f,
                             live ? __ldg(eb + (pix + (uint32_t)(2 * i + 1) * Nu)) : 0.f);
        } else {
#pragma unroll
            for (int i = 0; i < kC2; ++i) e[i] = pack2(ownY[(2 * i) * CHS_Y], ownY[(2 * i + 1) * CHS_Y]);
        }
        {
            f32x2 s0 = 0ull, s1 = 0ull;
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
        // one offset: forward pair (anchored here, partner p + o; zero border: a partner outside the image is zero fill,
        // a == 0, and its weight was fetched as 0) and -- self variant with gradient -- the pair anchored at p - o
        auto body = [&](int i, const KV& cur) {
            const int c = P.chan[i], s = P.shift[i];
            const float gco = P.gcoef[i];
            float a, lt;
            if (i >= ny) {  // CTA-uniform
                lt = pair<CHS_X, GRAD>(ownX + s, e, G, cur.t, cur.w * mask_sq<MASK>(cur.m), gco, a);
                if (GRAD && !CROSS) {
                    float a2;
                    pair<CHS_X, true>(ownX - s, e, G, cur.t2, cur.w2 * mask_sq<MASK>(cur.m2), gco, a2);
                }
            } else {
                lt = pair<CHS_Y, GRAD>(ownY + s * TX, e, G, cur.t, cur.w * mask_sq<MASK>(cur.m), gco, a);
                if (GRAD && !CROSS) {
                    float a2;
                    pair<CHS_Y, true>(ownY - s * TX, e, G, cur.t2, cur.w2 * mask_sq<MASK>(cur.m2), gco, a2);
                }
            }
            st_affs(affs_b + (uint32_t)c * Nu, a, affs_b != nullptr && (INNER || live));
            S.sL[c][threadIdx.x] = lt;
        };
        // three rotating fetch buffers, statically named (a register-to-register rotation would wait on the loads)
        bool xready = manual;
#pragma unroll 1
        for (int i = 0; i < ntot; i += 3) {
            if (!xready && i + 2 >= ny) {
                mbar_wait(&barX, 0);
                xready = true;
            }
            kv2 = fetch(i + 2);
            body(i, kv0);
            if (i + 1 < ntot) {
                kv0 = fetch(i + 3);
                body(i + 1, kv1);
            }
            if (i + 2 < ntot) {
                kv1 = fetch(i + 4);
                body(i + 2, kv2);
            }
        }
    };
    if (inner) run(TrueT{});
    else run(FalseT{});

    // ---- gradient back through the normalisation
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
        float* gb = P.grad + (int64_t)b * kC * Nu + pix;
#pragma unroll
        for (int i = 0; i < kC2; ++i) {
            float lo, hi;
            unpack2(mul2(fma2(e[i], neg, G[i]), r2), lo, hi);
            gb[(uint32_t)(2 * i) * Nu] = lo;
            gb[(uint32_t)(2 * i + 1) * Nu] = hi;
        }
    }

    // ---- loss partials: fixed-order CTA sum per channel
    __syncthreads();
    for (int k = ty; k < K; k += TY) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < TY; ++i) v += S.sL[k][i * 32 + tx];
        v = warp_sum(v);
        if (tx == 0) P.partials[(int64_t)lin_block * K + k] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent flavour (the default): the same tile body, but a CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x,
// ... and the TMA boxes of its NEXT tile are requested as soon as the last warp has finished reading the current ones
// (y-column box after the last y offset, x-row box after the last x offset: release / acquire counters as in the 3-D
// kernels, no CTA barrier in the loop).  The one-tile kernel above waits ~2 us for its boxes at the start of every
// 8 us CTA; here that latency hides behind the other half of the previous tile, the gradient stores and the next
// tile's first target / weight / mask fetches.  Circular border: TMA zero-fills what lies outside the image; on the
// tiles whose boxes cross the image edge exactly those elements are patched by hand afterwards (one CTA barrier per box).  Loss terms accumulate per thread and offset in shared memory over all the
// CTA's tiles and are summed in a fixed order at the end: gridDim.x rows of partials.
__device__ __forceinline__ uint32_t arrive_count(uint32_t* cnt) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(cnt)) : "memory");
    return old;
}

template <int TY, int PMAX, bool CROSS, int MASK, bool GRAD, bool LAB = false>
__global__ void __launch_bounds__(TY * 32, (TY <= 8) ? 2 : 1)
fused2d_axis_p(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapX, const Params P,
               const int tiles_x, const int tiles_y, const int n_tiles) {
    using G_ = Geo<TY, PMAX>;
    constexpr int PX = G_::PX, YR = G_::YR, XW = G_::XW, CHS_Y = G_::CHS_Y, CHS_X = G_::CHS_X, NT = G_::NT;
    constexpr uint32_t kWarps = TY;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Stage<TY, PMAX>& S = *reinterpret_cast<Stage<TY, PMAX>*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    __shared__ uint64_t barY, barX;
    __shared__ uint32_t cntY, cntX;

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int H = P.H, W = P.W, K = P.K;
    const bool wrap = P.wrap != 0;
    const uint32_t Nu = (uint32_t)(H * W);
    const int ny = P.ny, ntot = P.ny + P.nx;

    auto tile_origin = [&](int t, int& b, int& y0, int& x0) {
        const int r = t / tiles_x;
        x0 = (t - r * tiles_x) * TX;
        b = r / tiles_y;
        y0 = (r - b * tiles_y) * TY;
    };
    auto tile_inner = [&](int y0, int x0) { return y0 >= PMAX && y0 + TY + PMAX <= H && x0 >= PX && x0 + TX + PX <= W; };
    auto issueY = [&](int b, int y0, int x0) {
        mbar_arrive_expect_tx(&barY, (uint32_t)sizeof(S.sY));
        tma_load_5d(&S.sY[0][0][0], &mapY, &barY, x0, y0 - PMAX, 0, 0, b);
    };
    auto issueX = [&](int b, int y0, int x0) {
        mbar_arrive_expect_tx(&barX, (uint32_t)sizeof(S.sX));
        tma_load_5d(&S.sX[0][0][0], &mapX, &barX, x0 - PX, y0, 0, 0, b);
    };

#pragma unroll
    for (int k = 0; k < kMaxK; ++k) S.sL[k][threadIdx.x] = 0.f;  // own column only: no barrier needed before the += below
    int t = blockIdx.x;
    if (threadIdx.x == 0) {
        mbar_init(&barY, 1);
        mbar_init(&barX, 1);
        cntY = cntX = 0;
        mbar_fence_init();
        int b, y0, x0;
        tile_origin(t, b, y0, x0);
        issueY(b, y0, x0);
        issueX(b, y0, x0);
    }
    __syncthreads();
    uint32_t phY = 0, phX = 0, it = 0;

    for (; t < n_tiles; t += (int)gridDim.x, ++it) {
        int b, y0, x0;
        tile_origin(t, b, y0, x0);
        int bn = 0, y0n = 0, x0n = 0;
        const bool next_tma = t + (int)gridDim.x < n_tiles;
        if (next_tma) tile_origin(t + (int)gridDim.x, bn, y0n, x0n);
        const int x = x0 + tx, y = y0 + ty;
        const bool live = (x < W) && (y < H);
        const uint32_t pix = (uint32_t)((live ? y : 0) * W + (live ? x : 0));
        const bool inner = tile_inner(y0, x0);
        // circular border: TMA zero-fills what lies outside the image; those elements are patched by hand afterwards
        const bool patch = wrap && !inner;
        const uint32_t goal = kWarps * (it + 1) - 1;

        const float* tgt_b = LAB ? nullptr : P.target + (int64_t)b * K * Nu + pix;
        const float* wgt_b = LAB ? nullptr : P.weight + (int64_t)b * K * Nu + pix;
        // LAB: the pixel's own label; partners' labels are fetched per offset like target / weight / mask otherwise
        const int32_t* lab_b = LAB ? P.labels + (int64_t)b * Nu + pix : nullptr;
        const int32_t lp = LAB ? ldg_i32_early(lab_b, live) : 0;
        // gen_affs_ours (scripts_cvppp/utils/affinity_ours.py:17-39): target = same label (ignore: and both non-zero); a
        // pair whose other end lies outside the image is masked out (mask 0; the target's padding value is immaterial) --
        // the ground truth never wraps, whatever the border rule of the affinities is
        auto fetch_lab = [&](int i) -> KV {
            KV v = {0.f, 0.f, 0.f, 0.f, 1u, 1u};
            if (i >= ntot) return v;
            const bool isx = i >= ny;
            const int s = P.shift[i];
            const int qx = isx ? x + s : x, qy = isx ? y : y + s;  // partner of the pair anchored at this pixel
            const int rx = isx ? x - s : x, ry = isx ? y : y - s;  // anchor of the pair whose partner is this pixel
            const bool vf = live && (unsigned)qx < (unsigned)W && (unsigned)qy < (unsigned)H;
            const bool vr = GRAD && live && (unsigned)rx < (unsigned)W && (unsigned)ry < (unsigned)H;
            const int32_t lq = ldg_i32_early(lab_b + (vf ? (int64_t)(qy * W + qx) - (int64_t)pix : 0), vf);
            const int32_t lr = ldg_i32_early(lab_b + (vr ? (int64_t)(ry * W + rx) - (int64_t)pix : 0), vr);
            const bool sf = lp == lq && (!P.ignore || (lp != 0 && lq != 0));
            const bool sr = lr == lp && (!P.ignore || (lp != 0 && lr != 0));
            const float* wt = P.wtab + ((int64_t)b * K + P.chan[i]) * 2;
            v.t = sf ? 1.f : 0.f;
            v.w = ldg_f32_early(wt + (sf ? 0 : 1), vf);
            v.t2 = sr ? 1.f : 0.f;
            v.w2 = ldg_f32_early(wt + (sr ? 0 : 1), vr);
            return v;
        };
        const uint8_t* m8_b = MASK == PEA_MASK_U8 ? reinterpret_cast<const uint8_t*>(P.mask) + (int64_t)b * K * Nu + pix : nullptr;
        const float* mf_b = MASK == PEA_MASK_F32 ? reinterpret_cast<const float*>(P.mask) + (int64_t)b * K * Nu + pix : nullptr;

        auto fetch_inner = [&](int i) -> KV {
            KV v = {0.f, 0.f, 0.f, 0.f, 0u, 0u};
            if (i >= ntot) return v;
            const int s = P.shift[i];
            const uint32_t koff = (uint32_t)P.chan[i] * Nu;
            const int64_t back = i >= ny ? (int64_t)s : (int64_t)s * W;  // p - o in elements
            v.t = ldg_f32_now(tgt_b + koff);
            v.w = ldg_f32_now(wgt_b + koff);
            if (MASK == PEA_MASK_U8) v.m = ldg_u8_now(m8_b + koff);
            if (MASK == PEA_MASK_F32) v.m = __float_as_uint(ldg_f32_now(mf_b + koff));
            if (GRAD && !CROSS) {
                v.t2 = ldg_f32_now(tgt_b + koff - back);
                v.w2 = ldg_f32_now(wgt_b + koff - back);
                if (MASK == PEA_MASK_U8) v.m2 = ldg_u8_now(m8_b + koff - back);
                if (MASK == PEA_MASK_F32) v.m2 = __float_as_uint(ldg_f32_now(mf_b + koff - back));
            }
            return v;
        };
        auto fetch_edge = [&](int i) -> KV {
            KV v = {0.f, 0.f, 0.f, 0.f, 0u, 0u};
            if (i >= ntot) return v;
            const bool isx = i >= ny;
            const int s = P.shift[i];
            const uint32_t koff = (uint32_t)P.chan[i] * Nu;
            bool vf = live;
            if (!wrap) vf = vf && (isx ? (unsigned)(x + s) < (unsigned)W : (unsigned)(y + s) < (unsigned)H);
            v.t = ldg_f32_early(tgt_b + koff, vf);
            v.w = ldg_f32_early(wgt_b + koff, vf);
            if (MASK == PEA_MASK_U8) v.m = ldg_u8_early(m8_b + koff, vf);
            if (MASK == PEA_MASK_F32) v.m = __float_as_uint(ldg_f32_early(mf_b + koff, vf));
            if (GRAD && !CROSS) {
                int qy = y, qx = x;
                bool vr = live;
                if (isx) {
                    qx = x - s;
                    if (wrap) qx = wrap1(qx, W);
                    else vr = vr && (unsigned)qx < (unsigned)W;
                } else {
                    qy = y - s;
                    if (wrap) qy = wrap1(qy, H);
                    else vr = vr && (unsigned)qy < (unsigned)H;
                }
                const int64_t i2 = (int64_t)koff + (vr ? (int64_t)(qy * W + qx) - (int64_t)pix : 0);
                v.t2 = ldg_f32_early(tgt_b + i2, vr);
                v.w2 = ldg_f32_early(wgt_b + i2, vr);
                if (MASK == PEA_MASK_U8) v.m2 = ldg_u8_early(m8_b + i2, vr);
                if (MASK == PEA_MASK_F32) v.m2 = __float_as_uint(ldg_f32_early(mf_b + i2, vr));
            }
            return v;
        };

        auto patchY = [&]() {
            const float* src = (CROSS ? P.par : P.emb) + (int64_t)b * kC * Nu;
            float* sYf = &S.sY[0][0][0];
            for (int idx = threadIdx.x; idx < YR * TX; idx += NT) {
                const int gy = y0 - PMAX + (idx >> 5), gx = x0 + (idx & 31);
                if ((unsigned)gy < (unsigned)H && gx < W) continue;
                const uint32_t o = (uint32_t)(wrap3(gy, H) * W + wrap3(gx, W));
                float v[kC];
#pragma unroll
                for (int c = 0; c < kC; ++c) v[c] = __ldg(src + (o + (uint32_t)c * Nu));
#pragma unroll
                for (int c = 0; c < kC; ++c) sYf[c * CHS_Y + idx] = v[c];
            }
            __syncthreads();  // patched box complete
        };
        auto patchX = [&]() {
            const float* src = (CROSS ? P.par : P.emb) + (int64_t)b * kC * Nu;
            float* sXf = &S.sX[0][0][0];
            for (int idx = threadIdx.x; idx < TY * XW; idx += NT) {
                const int row = idx / XW, col = idx - row * XW;
                const int gy = y0 + row, gx = x0 - PX + col;
                if (gy < H && (unsigned)gx < (unsigned)W) continue;
                const uint32_t o = (uint32_t)(wrap3(gy, H) * W + wrap3(gx, W));
                float v[kC];
#pragma unroll
                for (int c = 0; c < kC; ++c) v[c] = __ldg(src + (o + (uint32_t)c * Nu));
#pragma unroll
                for (int c = 0; c < kC; ++c) sXf[c * CHS_X + idx] = v[c];
            }
            __syncthreads();
        };
        const float* ownY = &S.sY[0][ty + PMAX][tx];
        const float* ownX = &S.sX[0][ty][tx + PX];
#ifdef PEA_AXIS_NOSTORE  // timing experiments only
        float* affs_b = nullptr;
#else
        float* affs_b = P.affs ? P.affs + (int64_t)b * K * Nu + pix : nullptr;
#endif
        f32x2 e[kC2], G[kC2];
        float rp = 0.f;

        // the last warp to leave a box requests the next tile's copy of it
        auto releaseY = [&]() {
            __syncwarp();
            if (tx == 0 && arrive_count(&cntY) == goal && next_tma) issueY(bn, y0n, x0n);
        };
        auto releaseX = [&]() {
            __syncwarp();
            if (tx == 0 && arrive_count(&cntX) == goal && next_tma) issueX(bn, y0n, x0n);
        };

        auto run = [&](auto inner_c) {
            constexpr bool INNER = decltype(inner_c)::value;
#ifdef PEA_AXIS_NOFETCH  // timing experiments only
            auto fetch = [&](int i) -> KV { KV v = {0.5f, 1.25f, 0.25f, 0.75f, 1u, 1u}; return v; };
#else
            auto fetch = [&](int i) -> KV { return LAB ? fetch_lab(i) : (INNER ? fetch_inner(i) : fetch_edge(i)); };
#endif
#if PEA_AXIS_DEPTH == 4
            KV kv0 = fetch(0), kv1 = fetch(1), kv2 = fetch(2), kv3;
#else
            KV kv0 = fetch(0), kv1 = fetch(1), kv2;
#endif
            mbar_wait(&barY, phY & 1u);
            if (patch) patchY();
            if (CROSS) {
                const float* eb = P.emb + (int64_t)b * kC * Nu;
#pragma unroll
                for (int i = 0; i < kC2; ++i)
                    e[i] = pack2(live ? __ldg(eb + (pix + (uint32_t)(2 * i) * Nu)) : 0.f,
                                 live ? __ldg(eb + (pix + (uint32_t)(2 * i + 1) * Nu)) : 0.f);
            } else {
#pragma unroll
                for (int i = 0; i < kC2; ++i) e[i] = pack2(ownY[(2 * i) * CHS_Y], ownY[(2 * i + 1) * CHS_Y]);
            }
            {
                f32x2 s0 = 0ull, s1 = 0ull;
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
            auto body = [&](int i, const KV& cur) {
                const int c = P.chan[i], s = P.shift[i];
                const float gco = P.gcoef[i];
                float a, lt;
                if (i >= ny) {  // CTA-uniform
                    lt = pair<CHS_X, GRAD>(ownX + s, e, G, cur.t, cur.w * mask_sq<MASK>(cur.m), gco, a);
                    if (GRAD && !CROSS) {
                        float a2;
                        pair<CHS_X, true>(ownX - s, e, G, cur.t2, cur.w2 * mask_sq<MASK>(cur.m2), gco, a2);
                    }
                } else {
                    lt = pair<CHS_Y, GRAD>(ownY + s * TX, e, G, cur.t, cur.w * mask_sq<MASK>(cur.m), gco, a);
                    if (GRAD && !CROSS) {
                        float a2;
                        pair<CHS_Y, true>(ownY - s * TX, e, G, cur.t2, cur.w2 * mask_sq<MASK>(cur.m2), gco, a2);
                    }
                }
                st_affs(affs_b + (uint32_t)c * Nu, a, affs_b != nullptr && (INNER || live));
                S.sL[c][threadIdx.x] += lt;
            };
            bool xready = false, yfree = false;
#if PEA_AXIS_DEPTH == 4
            // four rotating fetch buffers: target / weight / mask are requested three offsets (~1 us) before their use
#pragma unroll 1
            for (int i = 0; i < ntot; i += 4) {
                if (!yfree && i >= ny) {  // every y offset done: the y-column box is free
                    releaseY();
                    yfree = true;
                }
                if (!xready && i + 3 >= ny) {
                    mbar_wait(&barX, phX & 1u);
                    if (patch) patchX();
                    xready = true;
                }
                kv3 = fetch(i + 3);
                body(i, kv0);
                if (i + 1 < ntot) {
                    kv0 = fetch(i + 4);
                    body(i + 1, kv1);
                }
                if (i + 2 < ntot) {
                    kv1 = fetch(i + 5);
                    body(i + 2, kv2);
                }
                if (i + 3 < ntot) {
                    kv2 = fetch(i + 6);
                    body(i + 3, kv3);
                }
            }
#else
#pragma unroll 1
            for (int i = 0; i < ntot; i += 3) {
                if (!yfree && i >= ny) {  // every y offset done: the y-column box is free
                    releaseY();
                    yfree = true;
                }
                if (!xready && i + 2 >= ny) {
                    mbar_wait(&barX, phX & 1u);
                    if (patch) patchX();
                    xready = true;
                }
                kv2 = fetch(i + 2);
                body(i, kv0);
                if (i + 1 < ntot) {
                    kv0 = fetch(i + 3);
                    body(i + 1, kv1);
                }
                if (i + 2 < ntot) {
                    kv1 = fetch(i + 4);
                    body(i + 2, kv2);
                }
            }
#endif
            if (!yfree) releaseY();
            if (!xready) mbar_wait(&barX, phX & 1u);  // (no x offsets) consume the phase all the same
            releaseX();
        };
        if (inner) run(TrueT{});
        else run(FalseT{});
        ++phY;
        ++phX;

        // ---- gradient back through the normalisation
        if (GRAD && live) {
            f32x2 p0 = 0ull, p1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                p0 = fma2(e[i], G[i], p0);
                p1 = fma2(e[i + 1], G[i + 1], p1);
            }
            float eg = hsum2(p0, p1);
            if (!(rp < 1e12f)) eg = 0.f;
#ifdef PEA_AXIS_NOSTORE
            if (eg != 12345.678f) continue;
#endif
            const f32x2 neg = pack2(-eg, -eg), r2 = pack2(rp, rp);
            float* gb = P.grad + (int64_t)b * kC * Nu + pix;
#pragma unroll
            for (int i = 0; i < kC2; ++i) {
                float lo, hi;
                unpack2(mul2(fma2(e[i], neg, G[i]), r2), lo, hi);
                gb[(uint32_t)(2 * i) * Nu] = lo;
                gb[(uint32_t)(2 * i + 1) * Nu] = hi;
            }
        }
    }

    // ---- loss partials: fixed-order CTA sum per channel
    __syncthreads();
    for (int k = ty; k < K; k += TY) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < TY; ++i) v += S.sL[k][i * 32 + tx];
        v = warp_sum(v);
        if (tx == 0) P.partials[(int64_t)blockIdx.x * K + k] = v;
    }
}

}  // namespace a2

// ---------------------------------------------------------------------------------------------------------------
struct PlanAxis {
    int pmax;  // 12 or 28
    int ty;
    int ny, nx;
    int chan[a2::kMaxK], shift[a2::kMaxK];
};

static bool plan_axis(const OffsetTable& tab, int H, int W, int wrap, PlanAxis& p) {
    if (tab.K > a2::kMaxK) return false;
    int P = 0;
    p.ny = p.nx = 0;
    int ych[a2::kMaxK], ysh[a2::kMaxK], xch[a2::kMaxK], xsh[a2::kMaxK];
    for (int k = 0; k < tab.K; ++k) {
        if (tab.dz[k] != 0) return false;
        // under the circular border an offset acts modulo the image extent (C remainder keeps the sign)
        const int oy = wrap ? tab.dy[k] % H : tab.dy[k], ox = wrap ? tab.dx[k] % W : tab.dx[k];
        if ((oy != 0) == (ox != 0)) return false;  // axis-aligned, non-zero
        const int a = oy ? (oy < 0 ? -oy : oy) : (ox < 0 ? -ox : ox);
        if (a > P) P = a;
        if (oy) { ych[p.ny] = k; ysh[p.ny++] = oy; }
        else    { xch[p.nx] = k; xsh[p.nx++] = ox; }
    }
    if (P > 28) return false;
    p.pmax = P <= 12 ? 12 : 28;
#ifndef PEA_AXIS_TY
#define PEA_AXIS_TY 8
#endif
    p.ty = P <= 12 ? PEA_AXIS_TY : 12;
    // the hand fill of edge CTAs maps coordinates through wrap3 (-n <= v < 4n); TMA needs 16-byte row pitch
    if (W % 4 != 0 || H < 32 || W < 64) return false;
    if ((int64_t)(tab.K > 16 ? tab.K : 16) * H * W >= ((int64_t)1 << 31)) return false;
    for (int i = 0; i < p.ny; ++i) { p.chan[i] = ych[i]; p.shift[i] = ysh[i]; }
    for (int i = 0; i < p.nx; ++i) { p.chan[p.ny + i] = xch[i]; p.shift[p.ny + i] = xsh[i]; }
    return true;
}

bool axis2d_eligible(int C, int D, int H, int W, const OffsetTable& tab, int wrap, bool need_grad_partner, const float* emb,
                     const float* partner) {
    if (C != 16 || D != 1 || need_grad_partner) return false;
    if ((reinterpret_cast<uintptr_t>(emb) | reinterpret_cast<uintptr_t>(partner)) & 15) return false;
    if (H > 65535 * 8) return false;  // grid.y
    PlanAxis p;
    return plan_axis(tab, H, W, wrap, p);
}

#ifndef PEA_AXIS_PERSIST
#define PEA_AXIS_PERSIST 1  // 0: one CTA per tile (fused2d_axis), for A/B runs
#endif

// CTAs of the persistent kernel: every SM's resident set (2 CTAs of 8 warps, or 1 of 12), never more than tiles
static int persistent_grid(int n_tiles, int ty) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        sms = 148;
    int g = sms * (ty <= 8 ? 2 : 1);
    // test knob (pea_debug_configure): a small grid makes every CTA walk many tiles even on small images
    if (const int c = debug_knob(PEA_KNOB_MAX_CTAS)) {
        if (c > 0 && c < g) g = c;
    }
    return n_tiles < g ? n_tiles : g;
}

// rows of loss partials the kernel writes
int axis2d_blocks(int B, int H, int W, const OffsetTable& tab, int wrap) {
    PlanAxis p;
    if (!plan_axis(tab, H, W, wrap, p)) return 0;
    const int n_tiles = B * ((H + p.ty - 1) / p.ty) * ((W + a2::TX - 1) / a2::TX);
    return PEA_AXIS_PERSIST ? persistent_grid(n_tiles, p.ty) : n_tiles;
}

template <int TY, int PMAX, bool CROSS, int MASK>
static int launch_axis_g(const CUtensorMap& mY, const CUtensorMap& mX, const a2::Params& P, int, cudaStream_t st) {
    const dim3 grid((unsigned)((P.W + a2::TX - 1) / a2::TX), (unsigned)((P.H + TY - 1) / TY), (unsigned)P.B);
    const size_t smem = sizeof(a2::Stage<TY, PMAX>) + 128;
    cudaError_t e;
#if PEA_AXIS_PERSIST
    {
        const int tiles_x = (int)grid.x, tiles_y = (int)grid.y, n_tiles = tiles_x * tiles_y * P.B;
        const int g = persistent_grid(n_tiles, TY);
        if (P.grad) {
            auto kern = a2::fused2d_axis_p<TY, PMAX, CROSS, MASK, true>;
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
            kern<<<g, TY * 32, smem, st>>>(mY, mX, P, tiles_x, tiles_y, n_tiles);
        } else {
            auto kern = a2::fused2d_axis_p<TY, PMAX, CROSS, MASK, false>;
            if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
            kern<<<g, TY * 32, smem, st>>>(mY, mX, P, tiles_x, tiles_y, n_tiles);
        }
        return (int)cudaGetLastError();
    }
#endif
    if (P.grad) {
        auto kern = a2::fused2d_axis<TY, PMAX, CROSS, MASK, true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, TY * 32, smem, st>>>(mY, mX, P);
    } else {
        auto kern = a2::fused2d_axis<TY, PMAX, CROSS, MASK, false>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, TY * 32, smem, st>>>(mY, mX, P);
    }
    return (int)cudaGetLastError();
}

template <int TY, int PMAX>
static int launch_axis_p(const float* src, const Dims& d, const a2::Params& P, bool cross, int mask_dtype, int grid,
                         cudaStream_t st) {
    using G_ = a2::Geo<TY, PMAX>;
    CUtensorMap mY, mX;
    int rc;
    if ((rc = encode_volume_map(&mY, src, d.B, 16, 1, d.H, d.W, a2::TX, G_::YR, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mX, src, d.B, 16, 1, d.H, d.W, G_::XW, TY, 1, 16))) return rc;
    if (cross) {
        if (mask_dtype == PEA_MASK_U8) return launch_axis_g<TY, PMAX, true, PEA_MASK_U8>(mY, mX, P, grid, st);
        if (mask_dtype == PEA_MASK_F32) return launch_axis_g<TY, PMAX, true, PEA_MASK_F32>(mY, mX, P, grid, st);
        return launch_axis_g<TY, PMAX, true, PEA_MASK_NONE>(mY, mX, P, grid, st);
    }
    if (mask_dtype == PEA_MASK_U8) return launch_axis_g<TY, PMAX, false, PEA_MASK_U8>(mY, mX, P, grid, st);
    if (mask_dtype == PEA_MASK_F32) return launch_axis_g<TY, PMAX, false, PEA_MASK_F32>(mY, mX, P, grid, st);
    return launch_axis_g<TY, PMAX, false, PEA_MASK_NONE>(mY, mX, P, grid, st);
}

template <int TY, int PMAX>
static int launch_axis_lab(const float* src, const Dims& d, const a2::Params& P, cudaStream_t st) {
    using G_ = a2::Geo<TY, PMAX>;
    CUtensorMap mY, mX;
    int rc;
    if ((rc = encode_volume_map(&mY, src, d.B, 16, 1, d.H, d.W, a2::TX, G_::YR, 1, 16))) return rc;
    if ((rc = encode_volume_map(&mX, src, d.B, 16, 1, d.H, d.W, G_::XW, TY, 1, 16))) return rc;
    const int tiles_x = (P.W + a2::TX - 1) / a2::TX, tiles_y = (P.H + TY - 1) / TY, n_tiles = tiles_x * tiles_y * P.B;
    const int g = persistent_grid(n_tiles, TY);
    const size_t smem = sizeof(a2::Stage<TY, PMAX>) + 128;
    cudaError_t e;
    if (P.grad) {
        auto kern = a2::fused2d_axis_p<TY, PMAX, false, PEA_MASK_NONE, true, true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<g, TY * 32, smem, st>>>(mY, mX, P, tiles_x, tiles_y, n_tiles);
    } else {
        auto kern = a2::fused2d_axis_p<TY, PMAX, false, PEA_MASK_NONE, false, true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<g, TY * 32, smem, st>>>(mY, mX, P, tiles_x, tiles_y, n_tiles);
    }
    return (int)cudaGetLastError();
}

// fused loss (+ gradient w.r.t. the anchor view) for an axis-aligned table; target / weight must be given
int launch_axis2d(const float* emb, const float* partner, const Dims& d, const OffsetTable& tab, int wrap,
                  const float* target, const float* weight, const void* mask, int mask_dtype, float* affs, float* partials,
                  float* grad, cudaStream_t st) {
    PlanAxis pl;
    if (!plan_axis(tab, d.H, d.W, wrap, pl)) return PEA_E_MODE;
    a2::Params P;
    P.B = d.B; P.H = d.H; P.W = d.W; P.K = tab.K; P.wrap = wrap;
    P.ny = pl.ny; P.nx = pl.nx;
    P.emb = emb; P.par = partner ? partner : emb;
    P.target = target; P.weight = weight; P.mask = mask;
    P.affs = affs; P.grad = grad; P.partials = partials;
    P.labels = nullptr; P.wtab = nullptr; P.ignore = 0;
    for (int i = 0; i < pl.ny + pl.nx; ++i) {
        P.chan[i] = pl.chan[i];
        P.shift[i] = pl.shift[i];
        P.gcoef[i] = tab.gcoef[pl.chan[i]];
    }
    for (int i = pl.ny + pl.nx; i < a2::kMaxK; ++i) { P.chan[i] = 0; P.shift[i] = 0; P.gcoef[i] = 0.f; }
    const int grid = d.B * ((d.H + pl.ty - 1) / pl.ty) * ((d.W + a2::TX - 1) / a2::TX);
    const float* src = partner ? partner : emb;
    if (pl.pmax == 12) return launch_axis_p<PEA_AXIS_TY, 12>(src, d, P, partner != nullptr, mask_dtype, grid, st);
    return launch_axis_p<12, 28>(src, d, P, partner != nullptr, mask_dtype, grid, st);
}

// Label-driven flavour for every axis-aligned table (e.g. CVPPP's multi_offset([1, 3, 5, 9, 27], 4), which the canonical
// kernel does not take): target = gen_affs_ours(labels, offsets, ignore, padding)[0], mask = [1], weight from the pair
// table of pea_gt_pair_weights -- none of the three tensors exists.
int launch_axis2d_labels(const float* emb, const int32_t* labels, const float* wtab, int ignore, const Dims& d,
                         const OffsetTable& tab, int wrap, float* affs, float* partials, float* grad, cudaStream_t st) {
    PlanAxis pl;
    if (!plan_axis(tab, d.H, d.W, wrap, pl)) return PEA_E_MODE;
    a2::Params P;
    P.B = d.B; P.H = d.H; P.W = d.W; P.K = tab.K; P.wrap = wrap;
    P.ny = pl.ny; P.nx = pl.nx;
    P.emb = emb; P.par = emb;
    P.target = nullptr; P.weight = nullptr; P.mask = nullptr;
    P.affs = affs; P.grad = grad; P.partials = partials;
    P.labels = labels; P.wtab = wtab; P.ignore = ignore;
    for (int i = 0; i < pl.ny + pl.nx; ++i) {
        P.chan[i] = pl.chan[i];
        P.shift[i] = pl.shift[i];
        P.gcoef[i] = tab.gcoef[pl.chan[i]];
    }
    for (int i = pl.ny + pl.nx; i < a2::kMaxK; ++i) { P.chan[i] = 0; P.shift[i] = 0; P.gcoef[i] = 0.f; }
    return pl.pmax == 12 ? launch_axis_lab<PEA_AXIS_TY, 12>(emb, d, P, st) : launch_axis_lab<12, 28>(emb, d, P, st);
}

}  // namespace pea
