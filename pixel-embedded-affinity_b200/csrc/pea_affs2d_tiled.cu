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
#include <cstdlib>

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
    int B, H, W, K, P;        // P = halo = max |offset component| (after reduction modulo the image under wrap)
    int wrap, rect;
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

// v mod n for -n <= v < 2n: straight-line selects (a loop here is recognised as a remainder -> integer division)
__device__ __forceinline__ int wrap1(int v, int n) {
    v += (v < 0) ? n : 0;
    v -= (v >= n) ? n : 0;
    return v;
}
// -n <= v < 4n: fill coordinates of tiles hanging over the image edge
__device__ __forceinline__ int wrap3(int v, int n) {
    v = wrap1(v, n);
    v -= (v >= n) ? n : 0;
    v -= (v >= n) ? n : 0;
    return v;
}

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

// Normalised embedding of one staged pixel (fetched channel-planar from global) -> shared memory; returns 1/|x|.
// `base` is the sample's channel-0 plane; `off` the 32-bit element offset of the pixel (16 H W < 2^31 by the plan).
__device__ __forceinline__ float stage_pixel(const float* __restrict__ base, uint32_t off, uint32_t Nu, bool inside,
                                             float* tile, int q) {
    float v[kC];
    if (inside) {
#pragma unroll
        for (int c = 0; c < kC; ++c) v[c] = __ldg(base + (off + (uint32_t)c * Nu));
    } else {
#pragma unroll
        for (int c = 0; c < kC; ++c) v[c] = 0.f;
    }
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int c = 0; c < kC; c += 2) {
        s0 = fmaf(v[c], v[c], s0);
        s1 = fmaf(v[c + 1], v[c + 1], s1);
    }
    const float r = inv_norm(s0 + s1);
    float4* dst = reinterpret_cast<float4*>(tile) + (size_t)q * 4;
    const int sw = swz(q);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        dst[j ^ sw] = make_float4(v[4 * j] * r, v[4 * j + 1] * r, v[4 * j + 2] * r, v[4 * j + 3] * r);
    return r;
}

template <int MASK>
__device__ __forceinline__ uint32_t mask_raw(const void* mask, uint32_t i) {
    if (MASK == PEA_MASK_U8) return (uint32_t)__ldg(reinterpret_cast<const uint8_t*>(mask) + i);
    if (MASK == PEA_MASK_F32) return __float_as_uint(__ldg(reinterpret_cast<const float*>(mask) + i));
    return 0x3f800000u;
}
template <int MASK>
__device__ __forceinline__ float mask_val(uint32_t raw) {
    return MASK == PEA_MASK_U8 ? (float)raw : __uint_as_float(raw);
}

// per-offset table, decoded once per CTA: which staged region holds the neighbour and its pixel-index delta
struct KTab {
    int sel[kMaxK];    // 0: y-column / rectangle, 1: x-row
    int delta[kMaxK];  // staged-pixel index of (p + o) minus that of p within the region
};

// KU > 0: the offset loop is fully unrolled over KU slots (K <= KU), every per-offset quantity is statically indexed
// (registers, constant-bank immediates); KU == 0: runtime loop for longer tables.
template <int TY, bool CROSS, int MASK, int KU>
__global__ void __launch_bounds__(TY * 32, (TY <= 12) ? 2 : 1)
fused2d_tiled(const Params P) {
    extern __shared__ __align__(16) float smem[];
    __shared__ float s_rinv[TY][TX];
    __shared__ KTab s_k;
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

    // ---- shared-memory geometry (staged pixel index q): y-column (or rectangle) first, q = row * pitchY + col,
    //      then the x-row at qX0 + row * XW + col
    const int YR = TY + 2 * HP;
    const int XW = TX + 2 * HP;
    const int pitchY = P.rect ? XW : TX;
    const int qX0 = P.rect ? 0 : YR * TX;
    if (threadIdx.x < K) {
        const int k = threadIdx.x, dy = P.dy[k], dx = P.dx[k];
        const bool inX = !P.rect && dx != 0;
        s_k.sel[k] = inX ? 1 : 0;
        s_k.delta[k] = inX ? dx : dy * pitchY + dx;
    }
    const float* stage_src = (CROSS ? P.par : P.emb) + (int64_t)b * kC * N;

    // ---- fill: every staged pixel is fetched through the border rule and normalised once
    {
        const int colsY = P.rect ? XW : TX;
        const int gx_base = P.rect ? x0 - HP : x0;
        for (int row = ty; row < YR; row += TY) {
            int gy = y0 - HP + row;
            const bool rin = wrap || (unsigned)gy < (unsigned)H;
            if (wrap) gy = wrap3(gy, H);
            const int cy = row - HP;
            for (int col = tx; col < colsY; col += 32) {
                int gx = gx_base + col;
                const bool inside = rin && (wrap || (unsigned)gx < (unsigned)W);
                if (wrap) gx = wrap3(gx, W);
                const float r = stage_pixel(stage_src, (uint32_t)(gy * W + gx), (uint32_t)N, inside, smem, row * pitchY + col);
                // the own 1/|x| of the tile centre is kept for the final projection (self variant)
                const int cx = P.rect ? col - HP : col;
                if (!CROSS && (unsigned)cy < (unsigned)TY && (unsigned)cx < (unsigned)TX) s_rinv[cy][cx] = r;
            }
        }
        if (!P.rect) {
            int gy = y0 + ty;
            const bool rin = wrap || gy < H;
            if (wrap) gy = wrap3(gy, H);
            for (int col = tx; col < XW; col += 32) {
                int gx = x0 - HP + col;
                const bool inside = rin && (wrap || (unsigned)gx < (unsigned)W);
                if (wrap) gx = wrap3(gx, W);
                stage_pixel(stage_src, (uint32_t)(gy * W + gx), (uint32_t)N, inside, smem, qX0 + ty * XW + col);
            }
        }
    }
    __syncthreads();

    // ---- own unit vector
    const int qY = (ty + HP) * pitchY + (P.rect ? tx + HP : tx);  // own pixel in the y-column / rectangle
    const int qXo = qX0 + ty * XW + tx + HP;                        // own pixel in the x-row
    f32x2 e[kC2], G[kC2];
    float rp;
    if (CROSS) {
        const float* eb = P.emb + (int64_t)b * kC * N;
        const uint32_t po = (uint32_t)((live ? y : 0) * W + (live ? x : 0));
        float v[kC], ss = 0.f;
#pragma unroll
        for (int c = 0; c < kC; ++c) {
            v[c] = live ? __ldg(eb + (po + (uint32_t)c * (uint32_t)N)) : 0.f;
            ss = fmaf(v[c], v[c], ss);
        }
        rp = inv_norm(ss);
#pragma unroll
        for (int i = 0; i < kC2; ++i) e[i] = pack2(v[2 * i] * rp, v[2 * i + 1] * rp);
    } else {
        load_px(smem, qY, e);
        rp = s_rinv[ty][tx];
    }
#pragma unroll
    for (int i = 0; i < kC2; ++i) G[i] = 0ull;

    const bool want_loss = P.target != nullptr;
    const bool want_grad = P.grad != nullptr;
    const bool want_rev = !CROSS && want_grad;
    // per-sample bases (64-bit) + 32-bit element offsets inside the sample (K * N < 2^31 is a plan requirement)
    const uint32_t Nu = (uint32_t)N;
    const uint32_t pix = (uint32_t)((live ? y : 0) * W + (live ? x : 0));
    const float* tgt_b = want_loss ? P.target + (int64_t)b * K * N : nullptr;
    const float* wgt_b = want_loss ? P.weight + (int64_t)b * K * N : nullptr;
    const void* msk_b = nullptr;
    if (MASK == PEA_MASK_U8) msk_b = reinterpret_cast<const uint8_t*>(P.mask) + (int64_t)b * K * N;
    if (MASK == PEA_MASK_F32) msk_b = reinterpret_cast<const float*>(P.mask) + (int64_t)b * K * N;
    float* affs_b = P.affs ? P.affs + (int64_t)b * K * N : nullptr;
    float lterm[KU > 0 ? KU : kMaxK];  // per-offset loss terms of this pixel (reduced over the CTA after the loop)

    // target / weight / mask of the pair anchored here and of the pair anchored at p - o_k, fetched ONE OFFSET AHEAD
    // (raw mask bits: converting at the load would stall the warp on it)
    struct KV { float t, w, t2, w2; uint32_t m, m2; };
    auto fetch = [&](int k) -> KV {
        KV v = {0.f, 0.f, 0.f, 0.f, 0u, 0u};
        if (!want_loss || k >= K) return v;
        const int dy = P.dy[k], dx = P.dx[k];
        const uint32_t koff = (uint32_t)k * Nu;
        bool valid = live;
        if (!wrap) valid = valid && (unsigned)(y + dy) < (unsigned)H && (unsigned)(x + dx) < (unsigned)W;
        if (valid) {
            v.t = __ldg(tgt_b + koff + pix);
            v.w = __ldg(wgt_b + koff + pix);
            v.m = mask_raw<MASK>(msk_b, koff + pix);
        }
        if (want_rev) {
            int qy = y - dy, qx = x - dx;
            bool v2 = live;
            if (wrap) { qy = wrap1(qy, H); qx = wrap1(qx, W); }
            else v2 = v2 && (unsigned)qy < (unsigned)H && (unsigned)qx < (unsigned)W;
            if (v2) {
                const uint32_t i2 = koff + (uint32_t)(qy * W + qx);
                v.t2 = __ldg(tgt_b + i2);
                v.w2 = __ldg(wgt_b + i2);
                v.m2 = mask_raw<MASK>(msk_b, i2);
            }
        }
        return v;
    };

    KV cur = fetch(0);
#pragma unroll(KU > 0 ? KU : 1)
    for (int k = 0; k < (KU > 0 ? KU : K); ++k) {
        if (KU > 0 && k >= K) break;
        const KV nxt = fetch(k + 1);
        const int sel = s_k.sel[k], delta = s_k.delta[k];
        const int qo = sel ? qXo : qY;
        // ---------------- pair anchored here: partner p + o_k
        f32x2 xq[kC2];
        load_px(smem, qo + delta, xq);
        f32x2 d0 = 0ull, d1 = 0ull;
#pragma unroll
        for (int i = 0; i < kC2; i += 2) {
            d0 = fma2(e[i], xq[i], d0);
            d1 = fma2(e[i + 1], xq[i + 1], d1);
        }
        float a = hsum2(d0, d1);
        if (!wrap) {
            const bool valid = (unsigned)(y + P.dy[k]) < (unsigned)H && (unsigned)(x + P.dx[k]) < (unsigned)W;
            if (!valid) a = 0.f;
        }
        if (affs_b && live) affs_b[(uint32_t)k * Nu + pix] = a;
        float lt = 0.f;
        if (want_loss) {
            // invalid pairs were fetched as t = w = m = 0: they contribute nothing
            const float m = mask_val<MASK>(cur.m);
            const float dd = a * m - cur.t * m;
            lt = cur.w * dd * dd;
            if (want_grad) {
                const float g = P.gcoef[k] * cur.w * m * m * (a - cur.t);
#pragma unroll
                for (int i = 0; i < kC2; ++i) G[i] = fma2(pack2(g, g), xq[i], G[i]);
            }
            // ---------------- pair anchored at p - o_k, whose partner is this pixel (self variant only)
            if (want_rev) {
                f32x2 xr[kC2];
                load_px(smem, qo - delta, xr);
                f32x2 r0 = 0ull, r1 = 0ull;
#pragma unroll
                for (int i = 0; i < kC2; i += 2) {
                    r0 = fma2(e[i], xr[i], r0);
                    r1 = fma2(e[i + 1], xr[i + 1], r1);
                }
                const float a2 = hsum2(r0, r1);
                const float m2 = mask_val<MASK>(cur.m2);
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
        float* gb = P.grad + (int64_t)b * kC * N;
#pragma unroll
        for (int i = 0; i < kC2; ++i) {
            float lo, hi;
            unpack2(mul2(fma2(e[i], neg, G[i]), r2), lo, hi);
            gb[pix + (uint32_t)(2 * i) * Nu] = lo;
            gb[pix + (uint32_t)(2 * i + 1) * Nu] = hi;
        }
    }

    // ---- loss partials: per-offset terms -> shared scratch (aliases the tile) -> fixed-order CTA sum
    if (want_loss && P.partials) {
        __syncthreads();  // everyone is done with the tile
        float* scratch = smem;  // [K][NT]
#pragma unroll(KU > 0 ? KU : 1)
        for (int k = 0; k < (KU > 0 ? KU : K); ++k)
            if (k < K) scratch[k * NT + threadIdx.x] = lterm[k];
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
    if (P > 64 || (int64_t)(tab.K > 16 ? tab.K : 16) * H * W >= ((int64_t)1 << 31)) return false;
    const int cands[3] = {16, 12, 8};
    const int force = debug_knob(PEA_KNOB_TUNE_TY2D);  // tuning aid only (pea_debug_configure): pin the tile height
    for (int i = 0; i < 3; ++i) {
        const int TY = cands[i];
        if (force && force != TY) continue;
        const size_t tile = rect ? (size_t)16 * (TY + 2 * P) * (t2::TX + 2 * P) * 4
                                 : (size_t)16 * ((size_t)(TY + 2 * P) * t2::TX + (size_t)TY * (t2::TX + 2 * P)) * 4;
        const size_t scratch = (size_t)tab.K * TY * 32 * 4;
        const size_t smem = tile > scratch ? tile : scratch;
        // prefer two resident CTAs per SM (113 KB each); accept one CTA (up to ~220 KB) for the largest halos
        if (smem <= 110 * 1024 || (i == 2 && smem <= 220 * 1024) || (smem <= 220 * 1024 && TY == 12 && P > 16) ||
            (force && smem <= 220 * 1024)) {
            plan.TY = TY; plan.P = P; plan.rect = rect; plan.smem = smem;
            return true;
        }
    }
    return false;
}

bool tiled2d_eligible(int C, int D, int H, int W, const OffsetTable& tab, int wrap, bool need_grad_partner) {
    if (C != 16 || D != 1 || need_grad_partner) return false;
    if (H < 8 || W < 16) return false;  // wrap3() covers at most three periods beyond the image
    Plan2D plan;
    return plan2d(tab, H, W, wrap, plan);
}

int tiled2d_blocks(int B, int H, int W, const OffsetTable& tab, int wrap) {
    Plan2D plan;
    if (!plan2d(tab, H, W, wrap, plan)) return 0;
    return B * ((H + plan.TY - 1) / plan.TY) * ((W + t2::TX - 1) / t2::TX);
}

constexpr int kStaticK = 10;  // multi_offset(5 shifts, neighbor 4): every shipped 2-D config and its deep-supervision prefixes

template <int TY, bool CROSS, int MASK, int KU>
static int launch2d_k(const t2::Params& P, int grid, size_t smem, cudaStream_t st) {
    auto kern = t2::fused2d_tiled<TY, CROSS, MASK, KU>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, TY * 32, smem, st>>>(P);
    return (int)cudaGetLastError();
}

template <int TY, bool CROSS, int MASK>
static int launch2d_one(const t2::Params& P, int grid, size_t smem, cudaStream_t st) {
    if (P.K <= kStaticK) return launch2d_k<TY, CROSS, MASK, kStaticK>(P, grid, smem, st);
    return launch2d_k<TY, CROSS, MASK, 0>(P, grid, smem, st);
}

template <int TY>
static int launch2d_ty(const t2::Params& P, bool cross, int mask_dtype, int grid, size_t smem, cudaStream_t st) {
    if (cross) {
        if (mask_dtype == PEA_MASK_U8) return launch2d_one<TY, true, PEA_MASK_U8>(P, grid, smem, st);
        if (mask_dtype == PEA_MASK_F32) return launch2d_one<TY, true, PEA_MASK_F32>(P, grid, smem, st);
        return launch2d_one<TY, true, PEA_MASK_NONE>(P, grid, smem, st);
    }
    if (mask_dtype == PEA_MASK_U8) return launch2d_one<TY, false, PEA_MASK_U8>(P, grid, smem, st);
    if (mask_dtype == PEA_MASK_F32) return launch2d_one<TY, false, PEA_MASK_F32>(P, grid, smem, st);
    return launch2d_one<TY, false, PEA_MASK_NONE>(P, grid, smem, st);
}

// target == nullptr: forward only (affs must be given).  partner == nullptr: self variant.
int launch_tiled2d(const float* emb, const float* partner, const Dims& d, const OffsetTable& tab, int wrap,
                   const float* target, const float* weight, const void* mask, int mask_dtype, float* affs,
                   float* partials, float* grad, cudaStream_t st) {
    Plan2D plan;
    if (!plan2d(tab, d.H, d.W, wrap, plan)) return PEA_E_MODE;
    t2::Params P;
    P.B = d.B; P.H = d.H; P.W = d.W; P.K = tab.K; P.P = plan.P;
    P.wrap = wrap; P.rect = plan.rect;
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
        case 16: return launch2d_ty<16>(P, partner != nullptr, mask_dtype, grid, plan.smem, st);
        case 12: return launch2d_ty<12>(P, partner != nullptr, mask_dtype, grid, plan.smem, st);
        default: return launch2d_ty<8>(P, partner != nullptr, mask_dtype, grid, plan.smem, st);
    }
}

}  // namespace pea
