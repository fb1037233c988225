// Tiled kernel for the 2-D family (embedding_loss / ema_embedding_loss / embedding2affs,
// scripts_cvppp/loss/loss_embedding_mse.py:18-95): runtime offset table, circular wrap (torch.roll) or zero border,
// uint8 / float32 / no mask, self or cross (two-view) affinities, forward-only or fused loss + backward.
//
// Design (DESIGN.md "K2"):
//   * a CTA owns a TY x 32 pixel tile of one sample; one thread per pixel;
//   * the CTA first fills shared memory with the NORMALISED embedding of the tile plus its halo -- a y-column
//     (TY+2P rows x 32) and an x-row (TY rows x 32+2P) for axis-aligned offset sets, or the full (TY+2P) x (32+2P)
//     rectangle when the table holds diagonal offsets.  The fill maps halo coordinates through the border rule
//     (modulo for wrap, zeros outside for the zero border), so wrap-around needs no special tiles, and every staged
//     pixel is normalised exactly once;
//   * the offset loop is a compact runtime loop: per offset one forward neighbour (pair anchored here) and, for the
//     self variant with gradient, one reverse neighbour (pair anchored at p - o), 16 LDS + 16 packed FMAs each;
//   * per-offset loss terms go to shared memory and are reduced in a fixed order -> one row of partials per CTA.
#include "pea_common.cuh"

namespace pea {
namespace t2 {

constexpr int kC = 16;
constexpr int kC2 = 8;
constexpr int TX = 32;
constexpr int kMaxK = 24;  // offsets the tiled path accepts (loss scratch = K * threads floats, aliased on the tile)

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

struct Params {
    int B, H, W, K, P;        // P = halo = max |offset component|
    int wrap, rect, mask_dtype;
    const float* emb;         // anchor view
    const float* par;         // partner view (== emb for the self variant)
    const float* target;      // null => forward only
    const float* weight;
    const void* mask;
    float* affs;              // may be null
    float* grad;              // may be null
    float* partials;          // [gridDim.x][K], may be null when target is null
    int dy[kMaxK], dx[kMaxK];
    float gcoef[kMaxK];
};

// Normalised embedding of one staged pixel -> channel-planar shared memory (stride `chs` floats between channels).
__device__ __forceinline__ void stage_pixel(const float* __restrict__ src, int64_t N, bool inside, float* dst, int chs,
                                            float* rinv_out) {
    float v[kC];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kC; ++c) {
        v[c] = inside ? __ldg(src + (int64_t)c * N) : 0.f;
        ss = fmaf(v[c], v[c], ss);
    }
    const float r = inv_norm(ss);
#pragma unroll
    for (int c = 0; c < kC; ++c) dst[c * chs] = v[c] * r;
    if (rinv_out) *rinv_out = r;
}

template <int TY, bool CROSS>
__global__ void __launch_bounds__(TY * 32)
fused2d_tiled(const Params P) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_rinv[TY][TX];
    constexpr int NT = TY * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int H = P.H, W = P.W, K = P.K, HP = P.P;
    const int tiles_x = (W + TX - 1) / TX, tiles_y = (H + TY - 1) / TY;
    int bid = blockIdx.x;
    const int tix = bid % tiles_x;
    bid /= tiles_x;
    const int tiy = bid % tiles_y;
    const int b = bid / tiles_y;
    const int x0 = tix * TX, y0 = tiy * TY;
    const int x = x0 + tx, y = y0 + ty;
    const bool live = (x < W) && (y < H);
    const int64_t N = (int64_t)H * W;
    const bool wrap = P.wrap != 0;

    // ---- shared-memory geometry
    const int YR = TY + 2 * HP;          // rows of the y-column / rectangle
    const int XW = TX + 2 * HP;          // cols of the x-row / rectangle
    float* sY;                           // [c][YR][TX]      (plus mode)   or [c][YR][XW] (rect mode)
    float* sX;                           // [c][TY][XW]      (plus mode only)
    int chsY, pitchY, chsX;
    if (P.rect) {
        pitchY = XW; chsY = YR * XW; sY = smem; sX = smem; chsX = chsY;
    } else {
        pitchY = TX; chsY = YR * TX; sY = smem; chsX = TY * XW; sX = smem + kC * chsY;
    }
    const float* stage_src = (CROSS ? P.par : P.emb) + (int64_t)b * kC * N;

    // ---- fill: every staged pixel is fetched through the border rule and normalised once
    {
        const int colsY = P.rect ? XW : TX;
        const int chunksY = (colsY + 31) / 32;
        for (int r = ty; r < YR * chunksY; r += TY) {
            const int row = r / chunksY, col = (r - row * chunksY) * 32 + tx;
            if (col < colsY) {
                int gy = y0 - HP + row, gx = (P.rect ? x0 - HP : x0) + col;
                bool inside = true;
                if (wrap) { gy = pmod(gy, H); gx = pmod(gx, W); }
                else inside = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
                // the own 1/|x| of the tile centre is kept for the final projection (self variant)
                const int cy = row - HP, cx = P.rect ? col - HP : col;
                float* rv = (!CROSS && (unsigned)cy < (unsigned)TY && (unsigned)cx < (unsigned)TX) ? &s_rinv[cy][cx] : nullptr;
                stage_pixel(stage_src + (int64_t)gy * W + gx, N, inside, sY + row * pitchY + col, chsY, rv);
            }
        }
        if (!P.rect) {
            const int chunksX = (XW + 31) / 32;
            for (int r = ty; r < TY * chunksX; r += TY) {
                const int row = r / chunksX, col = (r - row * chunksX) * 32 + tx;
                if (col < XW) {
                    int gy = y0 + row, gx = x0 - HP + col;
                    bool inside = true;
                    if (wrap) { gy = pmod(gy, H); gx = pmod(gx, W); }
                    else inside = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
                    stage_pixel(stage_src + (int64_t)gy * W + gx, N, inside, sX + row * XW + col, chsX, nullptr);
                }
            }
        }
    }
    __syncthreads();

    // ---- own unit vector
    f32x2 e[kC2], G[kC2];
    float rp;
    if (CROSS) {
        const float* src = P.emb + (int64_t)b * kC * N + (int64_t)(live ? y : 0) * W + (live ? x : 0);
        float v[kC], ss = 0.f;
#pragma unroll
        for (int c = 0; c < kC; ++c) {
            v[c] = live ? __ldg(src + (int64_t)c * N) : 0.f;
            ss = fmaf(v[c], v[c], ss);
        }
        rp = inv_norm(ss);
#pragma unroll
        for (int i = 0; i < kC2; ++i) e[i] = pack2(v[2 * i] * rp, v[2 * i + 1] * rp);
    } else {
        const float* own = P.rect ? sY + (ty + HP) * pitchY + (tx + HP) : sY + (ty + HP) * pitchY + tx;
#pragma unroll
        for (int i = 0; i < kC2; ++i) e[i] = pack2(own[(2 * i) * chsY], own[(2 * i + 1) * chsY]);
        rp = s_rinv[ty][tx];
    }
#pragma unroll
    for (int i = 0; i < kC2; ++i) G[i] = 0ull;

    const bool want_loss = P.target != nullptr;
    const bool want_grad = P.grad != nullptr;
    const int64_t pix = (int64_t)(live ? y : 0) * W + (live ? x : 0);
    const int64_t kbase = (int64_t)b * K * N;
    // per-offset loss terms of this thread; the scratch lives behind the tile when it fits, else it aliases the
    // tile after the offset loop (see below) -- here we keep them in a small local array indexed dynamically
    float lterm[kMaxK];

    for (int k = 0; k < K; ++k) {
        const int dy = P.dy[k], dx = P.dx[k];
        // ---------------- pair anchored here: partner p + o_k
        const float* nb;
        int chs;
        if (P.rect) { nb = sY + (ty + HP + dy) * pitchY + (tx + HP + dx); chs = chsY; }
        else if (dx == 0) { nb = sY + (ty + HP + dy) * pitchY + tx; chs = chsY; }
        else { nb = sX + ty * XW + (tx + HP + dx); chs = chsX; }
        f32x2 xq[kC2];
#pragma unroll
        for (int i = 0; i < kC2; ++i) xq[i] = pack2(nb[(2 * i) * chs], nb[(2 * i + 1) * chs]);
        f32x2 d0 = 0ull, d1 = 0ull;
#pragma unroll
        for (int i = 0; i < kC2; i += 2) {
            d0 = fma2(e[i], xq[i], d0);
            d1 = fma2(e[i + 1], xq[i + 1], d1);
        }
        float a = hsum2(d0, d1);
        bool valid = live;
        if (!wrap) valid = valid && (unsigned)(y + dy) < (unsigned)H && (unsigned)(x + dx) < (unsigned)W;
        if (!valid) a = 0.f;
        const int64_t idx = kbase + (int64_t)k * N + pix;
        if (P.affs && live) P.affs[idx] = a;
        float lt = 0.f;
        if (want_loss) {
            float g = 0.f;
            if (valid) {
                const float t = __ldg(P.target + idx), w = __ldg(P.weight + idx), m = load_mask(P.mask, P.mask_dtype, idx);
                const float dd = a * m - t * m;
                lt = w * dd * dd;
                g = P.gcoef[k] * w * m * m * (a - t);
            }
            if (want_grad) {
#pragma unroll
                for (int i = 0; i < kC2; ++i) G[i] = fma2(pack2(g, g), xq[i], G[i]);
            }
            // ---------------- pair anchored at p - o_k, whose partner is this pixel (self variant only)
            if (!CROSS && want_grad) {
                const float* nb2;
                if (P.rect) nb2 = sY + (ty + HP - dy) * pitchY + (tx + HP - dx);
                else if (dx == 0) nb2 = sY + (ty + HP - dy) * pitchY + tx;
                else nb2 = sX + ty * XW + (tx + HP - dx);
                f32x2 xr[kC2];
#pragma unroll
                for (int i = 0; i < kC2; ++i) xr[i] = pack2(nb2[(2 * i) * chs], nb2[(2 * i + 1) * chs]);
                f32x2 r0 = 0ull, r1 = 0ull;
#pragma unroll
                for (int i = 0; i < kC2; i += 2) {
                    r0 = fma2(e[i], xr[i], r0);
                    r1 = fma2(e[i + 1], xr[i + 1], r1);
                }
                const float a2 = hsum2(r0, r1);
                int qy = y - dy, qx = x - dx;
                bool v2 = live;
                if (wrap) { qy = pmod(qy, H); qx = pmod(qx, W); }
                else v2 = v2 && (unsigned)qy < (unsigned)H && (unsigned)qx < (unsigned)W;
                float g2 = 0.f;
                if (v2) {
                    const int64_t idx2 = kbase + (int64_t)k * N + (int64_t)qy * W + qx;
                    const float t2 = __ldg(P.target + idx2), w2 = __ldg(P.weight + idx2),
                                m2 = load_mask(P.mask, P.mask_dtype, idx2);
                    g2 = P.gcoef[k] * w2 * m2 * m2 * (a2 - t2);
                }
#pragma unroll
                for (int i = 0; i < kC2; ++i) G[i] = fma2(pack2(g2, g2), xr[i], G[i]);
            }
        }
        lterm[k] = lt;
    }

    // ---- gradient back through the normalisation
    if (want_grad && live) {
        f32x2 p0 = 0ull, p1 = 0ull;
#pragma unroll
        for (int i = 0; i < kC2; i += 2) {
            p0 = fma2(e[i], G[i], p0);
            p1 = fma2(e[i + 1], G[i + 1], p1);
        }
        float eg = hsum2(p0, p1);
        if (!(rp < 1e12f)) eg = 0.f;
        const f32x2 neg = pack2(-eg, -eg), r2 = pack2(rp, rp);
        float* dst = P.grad + (int64_t)b * kC * N + pix;
#pragma unroll
        for (int i = 0; i < kC2; ++i) {
            float lo, hi;
            unpack2(mul2(fma2(e[i], neg, G[i]), r2), lo, hi);
            dst[(int64_t)(2 * i) * N] = lo;
            dst[(int64_t)(2 * i + 1) * N] = hi;
        }
    }

    // ---- loss partials: per-offset terms -> shared scratch (aliases the tile) -> fixed-order CTA sum
    if (want_loss && P.partials) {
        __syncthreads();  // everyone is done with the tile
        float* scratch = smem;  // [K][NT]
        for (int k = 0; k < K; ++k) scratch[k * NT + threadIdx.x] = lterm[k];
        __syncthreads();
        for (int k = ty; k < K; k += TY) {
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < TY; ++i) v += scratch[k * NT + i * 32 + tx];
            v = warp_sum(v);
            if (tx == 0) P.partials[(int64_t)blockIdx.x * K + k] = v;
        }
    }
}

}  // namespace t2

// ---------------------------------------------------------------------------------------------------------------
struct Plan2D {
    int TY;
    int P;
    int rect;
    size_t smem;
};

static bool plan2d(const OffsetTable& tab, int H, int W, Plan2D& plan) {
    if (tab.K > t2::kMaxK) return false;
    int P = 0, rect = 0;
    for (int k = 0; k < tab.K; ++k) {
        if (tab.dz[k] != 0) return false;
        const int ay = tab.dy[k] < 0 ? -tab.dy[k] : tab.dy[k], ax = tab.dx[k] < 0 ? -tab.dx[k] : tab.dx[k];
        if (ay > P) P = ay;
        if (ax > P) P = ax;
        if (ay != 0 && ax != 0) rect = 1;
    }
    if (P > 64) return false;
    (void)H; (void)W;
    const int cands[3] = {16, 12, 8};
    for (int i = 0; i < 3; ++i) {
        const int TY = cands[i];
        const size_t tile = rect ? (size_t)16 * (TY + 2 * P) * (t2::TX + 2 * P) * 4
                                 : (size_t)16 * ((size_t)(TY + 2 * P) * t2::TX + (size_t)TY * (t2::TX + 2 * P)) * 4;
        const size_t scratch = (size_t)tab.K * TY * 32 * 4;
        const size_t smem = tile > scratch ? tile : scratch;
        // prefer two resident CTAs per SM (113 KB each); accept one CTA (up to ~220 KB) for the largest halos
        if (smem <= 110 * 1024 || (i == 2 && smem <= 220 * 1024) || (smem <= 220 * 1024 && TY == 12 && P > 16)) {
            plan.TY = TY; plan.P = P; plan.rect = rect; plan.smem = smem;
            return true;
        }
    }
    return false;
}

bool tiled2d_eligible(int C, int D, int H, int W, const OffsetTable& tab, bool need_grad_partner) {
    if (C != 16 || D != 1 || need_grad_partner) return false;
    Plan2D plan;
    return plan2d(tab, H, W, plan);
}

int tiled2d_blocks(int B, int H, int W, const OffsetTable& tab) {
    Plan2D plan;
    if (!plan2d(tab, H, W, plan)) return 0;
    return B * ((H + plan.TY - 1) / plan.TY) * ((W + t2::TX - 1) / t2::TX);
}

template <int TY>
static int launch2d_ty(const t2::Params& P, bool cross, int grid, size_t smem, cudaStream_t st) {
    cudaError_t e;
    if (cross) {
        auto kern = t2::fused2d_tiled<TY, true>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, TY * 32, smem, st>>>(P);
    } else {
        auto kern = t2::fused2d_tiled<TY, false>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, TY * 32, smem, st>>>(P);
    }
    return (int)cudaGetLastError();
}

// target == nullptr: forward only (affs must be given).  partner == nullptr: self variant.
int launch_tiled2d(const float* emb, const float* partner, const Dims& d, const OffsetTable& tab, int wrap,
                   const float* target, const float* weight, const void* mask, int mask_dtype, float* affs,
                   float* partials, float* grad, cudaStream_t st) {
    Plan2D plan;
    if (!plan2d(tab, d.H, d.W, plan)) return PEA_E_MODE;
    t2::Params P;
    P.B = d.B; P.H = d.H; P.W = d.W; P.K = tab.K; P.P = plan.P;
    P.wrap = wrap; P.rect = plan.rect; P.mask_dtype = mask_dtype;
    P.emb = emb; P.par = partner ? partner : emb;
    P.target = target; P.weight = weight; P.mask = mask;
    P.affs = affs; P.grad = grad; P.partials = partials;
    for (int k = 0; k < tab.K; ++k) { P.dy[k] = tab.dy[k]; P.dx[k] = tab.dx[k]; P.gcoef[k] = tab.gcoef[k]; }
    const int grid = d.B * ((d.H + plan.TY - 1) / plan.TY) * ((d.W + t2::TX - 1) / t2::TX);
    switch (plan.TY) {
        case 16: return launch2d_ty<16>(P, partner != nullptr, grid, plan.smem, st);
        case 12: return launch2d_ty<12>(P, partner != nullptr, grid, plan.smem, st);
        default: return launch2d_ty<8>(P, partner != nullptr, grid, plan.smem, st);
    }
}

}  // namespace pea
