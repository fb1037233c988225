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
//   * shared memory is PIXEL-MAJOR (16 channels = 64 B per staged pixel, 16-byte chunks XOR-swizzled by the pixel
//     index): a neighbour vector is four conflict-free LDS.128 with immediate offsets whatever the halo size, and
//     lands in registers as aligned pairs ready for packed FFMA2;
//   * the offset loop is a compact runtime loop: per offset one forward neighbour (pair anchored here) and, for the
//     self variant with gradient, one reverse neighbour (pair anchored at p - o), 4 LDS.128 + 16 packed FMAs each;
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

// v mod n for -n <= v < 4n (offsets are pre-reduced on the host and the plan requires H >= TY/2, W >= 16).
// Written as straight-line selects: a loop here is recognised as a remainder and compiled to an integer division.
__device__ __forceinline__ int wrapn(int v, int n) {
    v += (v < 0) ? n : 0;
    v -= (v >= n) ? n : 0;
    v -= (v >= n) ? n : 0;
    v -= (v >= n) ? n : 0;
    return v;
}

// mask element as raw bits (no conversion at the load, so the load can stay in flight) and its later decoding
__device__ __forceinline__ uint32_t load_mask_raw(const void* mask, int mask_dtype, int64_t i) {
    if (mask_dtype == PEA_MASK_U8) return (uint32_t)__ldg(reinterpret_cast<const uint8_t*>(mask) + i);
    if (mask_dtype == PEA_MASK_F32) return __float_as_uint(__ldg(reinterpret_cast<const float*>(mask) + i));
    return mask_dtype == PEA_MASK_NONE ? 0x3f800000u : 0u;
}
__device__ __forceinline__ float decode_mask(uint32_t raw, int mask_dtype) {
    return mask_dtype == PEA_MASK_U8 ? (float)raw : __uint_as_float(raw);
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

// Pixel-major staging: pixel q occupies floats [16 q, 16 q + 16); its 16-byte chunk j sits at slot j ^ ((q >> 1) & 3),
// which makes a warp's LDS.128 / STS.128 over consecutive pixels bank-conflict free.
__device__ __forceinline__ int swz(int q) { return (q >> 1) & 3; }

__device__ __forceinline__ void load_px(const float* __restrict__ tile, int q, f32x2 (&v)[kC2]) {
    const float4* base = reinterpret_cast<const float4*>(tile) + (size_t)q * 4;
    const int sw = swz(q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 c = base[j ^ sw];
        v[2 * j] = pack2(c.x, c.y);
        v[2 * j + 1] = pack2(c.z, c.w);
    }
}

// Normalised embedding of one staged pixel (fetched channel-planar from global) -> shared memory.
__device__ __forceinline__ void stage_pixel(const float* __restrict__ src, int64_t N, bool inside, float* tile, int q,
                                            float* rinv_out) {
    float v[kC];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kC; ++c) {
        v[c] = inside ? __ldg(src + (int64_t)c * N) : 0.f;
        ss = fmaf(v[c], v[c], ss);
    }
    const float r = inv_norm(ss);
    float4* base = reinterpret_cast<float4*>(tile) + (size_t)q * 4;
    const int sw = swz(q);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        base[j ^ sw] = make_float4(v[4 * j] * r, v[4 * j + 1] * r, v[4 * j + 2] * r, v[4 * j + 3] * r);
    if (rinv_out) *rinv_out = r;
}

template <int TY, bool CROSS>
__global__ void __launch_bounds__(TY * 32, (TY <= 12) ? 2 : 1)
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
    // staged pixel index: y-column (or rectangle) first -- q = row * pitchY + col -- then the x-row at qX0 + row * XW + col
    const int pitchY = P.rect ? XW : TX;
    const int qX0 = P.rect ? 0 : YR * TX;
    const float* stage_src = (CROSS ? P.par : P.emb) + (int64_t)b * kC * N;

    // ---- fill: every staged pixel is fetched through the border rule and normalised once
    {
        const int colsY = P.rect ? XW : TX;
        for (int row = ty; row < YR; row += TY) {
            for (int col = tx; col < colsY; col += 32) {
                int gy = y0 - HP + row, gx = (P.rect ? x0 - HP : x0) + col;
                bool inside = true;
                if (wrap) { gy = wrapn(gy, H); gx = wrapn(gx, W); }
                else inside = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
                // the own 1/|x| of the tile centre is kept for the final projection (self variant)
                const int cy = row - HP, cx = P.rect ? col - HP : col;
                float* rv = (!CROSS && (unsigned)cy < (unsigned)TY && (unsigned)cx < (unsigned)TX) ? &s_rinv[cy][cx] : nullptr;
                stage_pixel(stage_src + (int64_t)gy * W + gx, N, inside, smem, row * pitchY + col, rv);
            }
        }
        if (!P.rect) {
            {
                const int row = ty;
                for (int col = tx; col < XW; col += 32) {
                    int gy = y0 + row, gx = x0 - HP + col;
                    bool inside = true;
                    if (wrap) { gy = wrapn(gy, H); gx = wrapn(gx, W); }
                    else inside = (unsigned)gy < (unsigned)H && (unsigned)gx < (unsigned)W;
                    stage_pixel(stage_src + (int64_t)gy * W + gx, N, inside, smem, qX0 + row * XW + col, nullptr);
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
        load_px(smem, (ty + HP) * pitchY + (P.rect ? tx + HP : tx), e);
        rp = s_rinv[ty][tx];
    }
#pragma unroll
    for (int i = 0; i < kC2; ++i) G[i] = 0ull;

    const bool want_loss = P.target != nullptr;
    const bool want_grad = P.grad != nullptr;
    const bool want_rev = !CROSS && want_grad;
    const int64_t pix = (int64_t)(live ? y : 0) * W + (live ? x : 0);
    const int64_t kbase = (int64_t)b * K * N;
    float lterm[kMaxK];  // per-offset loss terms of this pixel (reduced over the CTA after the loop)

    // target / weight / mask of the pair anchored here and of the pair anchored at p - o_k, fetched ONE OFFSET AHEAD
    struct KV { float t, w, t2, w2; uint32_t m, m2; };
    auto fetch = [&](int k) -> KV {
        KV v = {0.f, 0.f, 0.f, 0.f, 0u, 0u};
        if (!want_loss || k >= K) return v;
        const int dy = P.dy[k], dx = P.dx[k];
        bool valid = live;
        if (!wrap) valid = valid && (unsigned)(y + dy) < (unsigned)H && (unsigned)(x + dx) < (unsigned)W;
        const int64_t idx = kbase + (int64_t)k * N + pix;
        if (valid) { v.t = __ldg(P.target + idx); v.w = __ldg(P.weight + idx); v.m = load_mask_raw(P.mask, P.mask_dtype, idx); }
        if (want_rev) {
            int qy = y - dy, qx = x - dx;
            bool v2 = live;
            if (wrap) { qy = wrapn(qy, H); qx = wrapn(qx, W); }
            else v2 = v2 && (unsigned)qy < (unsigned)H && (unsigned)qx < (unsigned)W;
            if (v2) {
                const int64_t idx2 = kbase + (int64_t)k * N + (int64_t)qy * W + qx;
                v.t2 = __ldg(P.target + idx2); v.w2 = __ldg(P.weight + idx2); v.m2 = load_mask_raw(P.mask, P.mask_dtype, idx2);
            }
        }
        return v;
    };

    KV cur = fetch(0);
    for (int k = 0; k < K; ++k) {
        const KV nxt = fetch(k + 1);
        const int dy = P.dy[k], dx = P.dx[k];
        // ---------------- pair anchored here: partner p + o_k
        int q;
        if (P.rect) q = (ty + HP + dy) * pitchY + (tx + HP + dx);
        else if (dx == 0) q = (ty + HP + dy) * pitchY + tx;
        else q = qX0 + ty * XW + (tx + HP + dx);
        f32x2 xq[kC2];
        load_px(smem, q, xq);
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
        if (P.affs && live) P.affs[kbase + (int64_t)k * N + pix] = a;
        float lt = 0.f;
        if (want_loss) {
            // invalid pairs were fetched as t = w = m = 0: they contribute nothing
            const float m = decode_mask(cur.m, P.mask_dtype);
            const float dd = a * m - cur.t * m;
            lt = cur.w * dd * dd;
            if (want_grad) {
                const float g = P.gcoef[k] * cur.w * m * m * (a - cur.t);
#pragma unroll
                for (int i = 0; i < kC2; ++i) G[i] = fma2(pack2(g, g), xq[i], G[i]);
            }
            // ---------------- pair anchored at p - o_k, whose partner is this pixel (self variant only)
            if (want_rev) {
                int q2;
                if (P.rect) q2 = (ty + HP - dy) * pitchY + (tx + HP - dx);
                else if (dx == 0) q2 = (ty + HP - dy) * pitchY + tx;
                else q2 = qX0 + ty * XW + (tx + HP - dx);
                f32x2 xr[kC2];
                load_px(smem, q2, xr);
                f32x2 r0 = 0ull, r1 = 0ull;
#pragma unroll
                for (int i = 0; i < kC2; i += 2) {
                    r0 = fma2(e[i], xr[i], r0);
                    r1 = fma2(e[i + 1], xr[i + 1], r1);
                }
                const float a2 = hsum2(r0, r1);
                const float m2 = decode_mask(cur.m2, P.mask_dtype);
                const float g2 = P.gcoef[k] * cur.w2 * m2 * m2 * (a2 - cur.t2);
#pragma unroll
                for (int i = 0; i < kC2; ++i) G[i] = fma2(pack2(g2, g2), xr[i], G[i]);
            }
        }
        lterm[k] = lt;
        cur = nxt;
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

// Under the circular border an offset acts modulo the image extent: reduce it once on the host so the kernel's halo
// and its index arithmetic stay within one period (C remainder: |r| < n, sign of the dividend).
static inline int eff_off(int o, int n, int wrap) { return wrap ? o % n : o; }

static bool plan2d(const OffsetTable& tab, int H, int W, int wrap, Plan2D& plan) {
    if (tab.K > t2::kMaxK) return false;
    int P = 0, rect = 0;
    for (int k = 0; k < tab.K; ++k) {
        if (tab.dz[k] != 0) return false;
        const int oy = eff_off(tab.dy[k], H, wrap), ox = eff_off(tab.dx[k], W, wrap);
        const int ay = oy < 0 ? -oy : oy, ax = ox < 0 ? -ox : ox;
        if (ay > P) P = ay;
        if (ax > P) P = ax;
        if (ay != 0 && ax != 0) rect = 1;
    }
    if (P > 64) return false;
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

bool tiled2d_eligible(int C, int D, int H, int W, const OffsetTable& tab, int wrap, bool need_grad_partner) {
    if (C != 16 || D != 1 || need_grad_partner) return false;
    if (H < 8 || W < 16) return false;  // wrapn() covers at most three periods beyond the image
    Plan2D plan;
    return plan2d(tab, H, W, wrap, plan);
}

int tiled2d_blocks(int B, int H, int W, const OffsetTable& tab, int wrap) {
    Plan2D plan;
    if (!plan2d(tab, H, W, wrap, plan)) return 0;
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
    if (!plan2d(tab, d.H, d.W, wrap, plan)) return PEA_E_MODE;
    t2::Params P;
    P.B = d.B; P.H = d.H; P.W = d.W; P.K = tab.K; P.P = plan.P;
    P.wrap = wrap; P.rect = plan.rect; P.mask_dtype = mask_dtype;
    P.emb = emb; P.par = partner ? partner : emb;
    P.target = target; P.weight = weight; P.mask = mask;
    P.affs = affs; P.grad = grad; P.partials = partials;
    for (int k = 0; k < tab.K; ++k) {
        P.dy[k] = eff_off(tab.dy[k], d.H, wrap);
        P.dx[k] = eff_off(tab.dx[k], d.W, wrap);
        P.gcoef[k] = tab.gcoef[k];
    }
    const int grid = d.B * ((d.H + plan.TY - 1) / plan.TY) * ((d.W + t2::TX - 1) / t2::TX);
    switch (plan.TY) {
        case 16: return launch2d_ty<16>(P, partner != nullptr, grid, plan.smem, st);
        case 12: return launch2d_ty<12>(P, partner != nullptr, grid, plan.smem, st);
        default: return launch2d_ty<8>(P, partner != nullptr, grid, plan.smem, st);
    }
}

}  // namespace pea
