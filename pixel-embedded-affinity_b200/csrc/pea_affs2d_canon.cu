// Fused 2-D kernel for the CANONICAL offset tables multi_offset([s0..s4], neighbor=4) with shifts <= 12 -- channel 2j
// is the y shift [-s_j, 0], channel 2j + 1 the x shift [0, -s_j] (scripts_cvppp/utils/affinity_ours.py:4-15; the
// BBBC039V1 table [1, 3, 5, 9, 11] of bbbc039v1.yaml:65-71): embedding -> 10 affinities -> masked weighted MSE ->
// gradient, in three flavours of one template:
//   SELF   embedding_loss      (scripts_cvppp/loss/loss_embedding_mse.py:18-47), described below;
//   CROSS  ema_embedding_loss  (:79-95) with the second view constant: only the pair anchored at p exists -- upper / left
//          halo only, tile-sized target / weight groups, the anchor view's own tile in a box of its own;
//   LAB    embedding_loss from instance labels (the data provider's gen_affs_ours + weight_binary_ratio folded in,
//          SURVEY.md 8(f) item 1): the int32 label plane is staged instead of the target / weight groups.
//
// Design (DESIGN.md "K2c"): the 3-D kernel's in-plane recipe on a persistent 2-D tile walk.
//   * a CTA of 16 warps owns 16 x 32 pixels; it walks tiles t = blockIdx.x, + gridDim.x, ...; per tile the RAW
//     channel-planar embedding of a y-column box (16 + 2 * 12 rows) and an x-row box (32 + 2 * 12 columns) AND the
//     target / weight planes arrive by TMA: the five y channels of the tile plus the 12 rows below it in one
//     channel-strided copy per tensor (the rows below hold the pairs anchored at p + s whose partner is p: gather-form
//     backward), the five x channels of the tile plus the next 16 columns in another;
//   * the copies of the NEXT tile are requested by the last warp that has finished reading a box (release / acquire
//     counters, no CTA barrier in the loop), so they land while the other half of the tile is computed;
//   * the offset loops are fully unrolled (static channels, shifts as uniform run-time values): every shared-memory
//     operand is base + immediate, loss terms live in 10 registers; only the mask -- whose uint8 row pitch is not a
//     multiple of 16 bytes at W = 696, so TMA cannot carry it -- is fetched per thread, one part of the tile ahead;
//   * circular border (torch.roll): TMA zero-fills what lies outside the image; on the tiles whose boxes cross the image
//     edge (17 % at 520 x 696) exactly those elements are then patched by hand through the wrap rule.
// The table-driven kernel (pea_affs2d_axis.cu) executes ~2400 instructions per pixel, this one ~2000 (ncu: 0.373 ms one
// tile per CTA, 0.356 ms persistent, 0.280 ms here at B = 8, 520 x 696).
#include <cstdlib>
#include "pea_common.cuh"
#include "pea_tma.cuh"

#ifndef PEA_CANON_MASK_STAGE
#define PEA_CANON_MASK_STAGE 0  // 1: uint8 masks staged in shared memory one tile ahead instead of per-thread byte loads
#endif

namespace pea {
namespace c2 {

constexpr int kC = 16, kC2 = 8;
constexpr int TX = 32, TY = 16, PM = 12, NS = 5, K = 2 * NS;
constexpr int NT = TY * 32;
// One view (SELF): every offset is evaluated in both directions (gather-form backward), so the boxes carry a halo on
// both sides and the target / weight groups the anchors s below / to the right of the tile.  Two views (CROSS, second
// view constant): only the pair anchored at p exists -- upper / left halo only, tile-sized target / weight groups, and
// the own vector of the anchor view in a box of its own.
template <bool CROSS>
struct Geo {
    static constexpr int YR = CROSS ? TY + PM : TY + 2 * PM;      // rows of the y-column box, first row y0 - PM
    static constexpr int XW = CROSS ? TX + PM : TX + 2 * PM;      // columns of the x-row box, first column x0 - PM
    static constexpr int CHS_Y = YR * TX, CHS_X = TY * XW;
    static constexpr int GY = CROSS ? TY : TY + PM;               // rows of the y-channel target / weight group
    static constexpr int GX = CROSS ? TX : TX + 16;               // columns of the x-channel group
};

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
__device__ __forceinline__ uint32_t arrive_count(uint32_t* cnt) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(cnt)) : "memory");
    return old;
}
// predicated read-only loads the compiler may not sink towards their use (volatile)
__device__ __forceinline__ uint32_t ldg_u8_early(const uint8_t* p, bool pred) {
    uint32_t v;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.u32 %0, 0;\n\t@q ld.global.nc.u8 %0, [%1];\n\t}"
        : "=r"(v)
        : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ uint32_t ldg_b32_early(const float* p, bool pred) {
    uint32_t v;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.u32 %0, 0;\n\t@q ld.global.nc.b32 %0, [%1];\n\t}"
        : "=r"(v)
        : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ int wrap1(int v, int n) {  // v mod n for -n <= v < 2n
    v += (v < 0) ? n : 0;
    v -= (v >= n) ? n : 0;
    return v;
}
__device__ __forceinline__ int wrap3(int v, int n) {  // -n <= v < 4n
    v = wrap1(v, n);
    v -= (v >= n) ? n : 0;
    v -= (v >= n) ? n : 0;
    return v;
}

struct Params {
    int B, H, W;
    int wrap;
    int tiles_x, tiles_y, n_tiles;
    int shift[NS];   // s_j > 0: channel 2j pairs p with the pixel s_j rows above, channel 2j + 1 with s_j columns to the left
    float gcoef[K];  // 2 * coef / norm of the channel
    const float* emb;      // anchor view
    const float* par;      // staged view (== emb for one view)
    const float* target;
    const float* weight;
    const void* mask;
    const int32_t* labels;  // LAB: (B, H, W) instance labels
    const float* wtab;      // LAB: [B][K][2] weight where the target is 1 / where it is 0 (pea_gt_pair_weights)
    int ignore;             // LAB: gen_affs_ours(ignore=...)
    float* affs;      // may be null
    float* grad;      // may be null
    float* partials;  // [gridDim.x][K]
};

// LAB: target / mask / weight are derived in the kernel from the instance labels (staged like the embedding boxes, one
// plane) and a per-(sample, channel) table of the two class-balance weight values: no target / weight groups at all.
template <bool CROSS, bool LAB>
struct alignas(128) Stage {
    float sY[kC][Geo<CROSS>::YR][TX];   // staged view, rows y0 - 12 .. (SELF: y0 + 27)
    float sX[kC][TY][Geo<CROSS>::XW];   // staged view, cols x0 - 12 .. (SELF: x0 + 43)
    float tY[LAB ? 1 : NS][LAB ? 1 : Geo<CROSS>::GY][TX];   // target, y channels (0, 2, .., 8), rows y0 .. (SELF: y0 + 27)
    float wY[LAB ? 1 : NS][LAB ? 1 : Geo<CROSS>::GY][TX];   // weight
    float tX[LAB ? 1 : NS][LAB ? 1 : TY][Geo<CROSS>::GX];   // target, x channels (1, 3, .., 9), cols x0 .. (SELF: x0 + 47)
    float wX[LAB ? 1 : NS][LAB ? 1 : TY][Geo<CROSS>::GX];
    float sO[CROSS ? kC : 1][CROSS ? TY : 1][TX];  // CROSS: the anchor view's own tile
    int32_t lY[LAB ? Geo<CROSS>::YR : 1][TX];      // LAB: labels of the y-column box
    int32_t lX[LAB ? TY : 1][Geo<CROSS>::XW];      // LAB: labels of the x-row box
    // uint8 masks of the target / weight groups' footprints, double-buffered over the tile walk (PEA_CANON_MASK_STAGE):
    // TMA cannot carry them (row pitch W bytes is not a multiple of 16 at W = 696), so every thread loads a few aligned
    // 32-bit words of the NEXT tile's masks and stores them before it releases the corresponding box
    uint32_t mY[PEA_CANON_MASK_STAGE && !LAB ? 2 : 1][PEA_CANON_MASK_STAGE && !LAB ? NS : 1]
              [PEA_CANON_MASK_STAGE && !LAB ? Geo<CROSS>::GY : 1][TX / 4];
    uint32_t mX[PEA_CANON_MASK_STAGE && !LAB ? 2 : 1][PEA_CANON_MASK_STAGE && !LAB ? NS : 1][PEA_CANON_MASK_STAGE && !LAB ? TY : 1]
              [Geo<CROSS>::GX / 4];
};

// One neighbour q of the own pixel p (raw values, channels CHS floats apart): a = <e^_p, e^_q>,
// G += gcoef * w * (a - t) / |x_q| * x_q; returns w * (a - t)^2.
template <int CHS, bool GRAD>
__device__ __forceinline__ float pair(const float* __restrict__ nb, const f32x2 (&e)[kC2], f32x2 (&G)[kC2], float t, float w,
                                      float gco, float& a_out) {
    f32x2 xq[kC2];
#pragma unroll
    for (int i = 0; i < kC2; ++i) xq[i] = pack2(nb[(2 * i) * CHS], nb[(2 * i + 1) * CHS]);
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

// Store of one affinity, predicated WITHOUT a branch: `if (affs_b && live) affs_b[o] = a;` ends a basic block after
// every pair evaluation, so the next pair's shared-memory loads cannot be scheduled above it (the forward-only 2-D kernel
// gained 7.5 % from the same change).  A predicated-off store touches no memory, so a null affs_b is fine.
#ifndef PEA_CANON_BRANCHY_STORE
#define PEA_CANON_BRANCHY_STORE 0   // 1: the `if` form (A/B)
#endif
__device__ __forceinline__ void st_affs(float* base, uint32_t off, float v, bool ok) {
#if PEA_CANON_BRANCHY_STORE
    if (ok) base[off] = v;
#else
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.f32 [%0], %1;\n\t}" ::"l"(base + off), "f"(v), "r"((int)ok));
#endif
}

template <int MASK>
__device__ __forceinline__ float mask_sq(uint32_t raw) {
    if (MASK == PEA_MASK_NONE) return 1.f;
    const float m = MASK == PEA_MASK_U8 ? (float)raw : __uint_as_float(raw);
    return m * m;
}

template <int MASK, bool GRAD, bool CROSS, bool LAB = false>
__global__ void __launch_bounds__(NT, 1)
fused2d_canon(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapX,
              const __grid_constant__ CUtensorMap mapTy, const __grid_constant__ CUtensorMap mapWy,
              const __grid_constant__ CUtensorMap mapTx, const __grid_constant__ CUtensorMap mapWx,
              const __grid_constant__ CUtensorMap mapO, const Params P) {
    using G_ = Geo<CROSS>;
    constexpr int YR = G_::YR, XW = G_::XW, CHS_Y = G_::CHS_Y, CHS_X = G_::CHS_X, GY = G_::GY, GX = G_::GX;
    constexpr bool REV = GRAD && !CROSS;  // the pair anchored at p + s, whose partner is p, contributes to grad(p)
    constexpr bool MST = PEA_CANON_MASK_STAGE && MASK == PEA_MASK_U8 && !LAB;  // masks staged in shared memory
    constexpr int NWY = NS * GY * (TX / 4), NWX = NS * TY * (GX / 4);          // 32-bit words of the two mask groups
    constexpr int RY = (NWY + NT - 1) / NT, RX = (NWX + NT - 1) / NT;          // words per thread
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    static_assert(!(LAB && (CROSS || MASK != PEA_MASK_NONE)), "label-driven flavour: one view, mask derived in the kernel");
    Stage<CROSS, LAB>& S = *reinterpret_cast<Stage<CROSS, LAB>*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    __shared__ float s_w[K][2];
    __shared__ uint64_t barY, barX;
    __shared__ uint32_t cntY, cntX;
    __shared__ float s_red[K][TY];
    constexpr uint32_t kWarps = TY;
    constexpr uint32_t kBytesY = sizeof(S.sY) + (LAB ? sizeof(S.lY) : sizeof(S.tY) + sizeof(S.wY)) + (CROSS ? sizeof(S.sO) : 0);
    constexpr uint32_t kBytesX = sizeof(S.sX) + (LAB ? sizeof(S.lX) : sizeof(S.tX) + sizeof(S.wX));

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int H = P.H, W = P.W;
    const bool wrap = P.wrap != 0;
    const uint32_t Nu = (uint32_t)(H * W);
    const int tiles_x = P.tiles_x, tiles_y = P.tiles_y, n_tiles = P.n_tiles;

    auto tile_origin = [&](int t, int& b, int& y0, int& x0) {
        const int r = t / tiles_x;
        x0 = (t - r * tiles_x) * TX;
        b = r / tiles_y;
        y0 = (r - b * tiles_y) * TY;
    };
    // every box of the tile lies inside the image (circular border: TMA's zero fill would be wrong outside)
    auto tile_inner = [&](int y0, int x0) {
        return y0 >= PM && y0 - PM + YR <= H && (LAB || y0 + GY <= H) && x0 >= PM && x0 - PM + XW <= W && (LAB || x0 + GX <= W);
    };
    auto issueY = [&](int b, int y0, int x0) {
        mbar_arrive_expect_tx(&barY, kBytesY);
        tma_load_5d(&S.sY[0][0][0], &mapY, &barY, x0, y0 - PM, 0, 0, b);
        if (LAB) {  // mapTy / mapTx describe the label plane with the embedding boxes' footprints
            tma_load_5d(&S.lY[0][0], &mapTy, &barY, x0, y0 - PM, 0, 0, b);
        } else {
            tma_load_5d(&S.tY[0][0][0], &mapTy, &barY, x0, y0, 0, 0, b);
            tma_load_5d(&S.wY[0][0][0], &mapWy, &barY, x0, y0, 0, 0, b);
        }
        if (CROSS) tma_load_5d(&S.sO[0][0][0], &mapO, &barY, x0, y0, 0, 0, b);
    };
    auto issueX = [&](int b, int y0, int x0) {
        mbar_arrive_expect_tx(&barX, kBytesX);
        tma_load_5d(&S.sX[0][0][0], &mapX, &barX, x0 - PM, y0, 0, 0, b);
        if (LAB) {
            tma_load_5d(&S.lX[0][0], &mapTx, &barX, x0 - PM, y0, 0, 0, b);
        } else {
            tma_load_5d(&S.tX[0][0][0], &mapTx, &barX, x0, y0, 0, 1, b);
            tma_load_5d(&S.wX[0][0][0], &mapWx, &barX, x0, y0, 0, 1, b);
        }
    };

    float lossacc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) lossacc[k] = 0.f;

    // word `idx` of a tile's y-channel (isx = false) or x-channel mask group: rows y0 .., columns x0 .. of channel
    // 2 j (+ 1); outside the image the circular border wraps (whole words: W % 4 == 0), the zero border reads 0
    auto mask_word = [&](bool isx, int idx, int b, int y0, int x0) -> uint32_t {
        const int wpr = isx ? GX / 4 : TX / 4, rows = isx ? TY : GY;
        const int j = idx / (rows * wpr), rem = idx - j * rows * wpr;
        const int row = rem / wpr, wc = rem - row * wpr;
        int gy = y0 + row, gx = x0 + 4 * wc;
        if (wrap) {
            gy = wrap3(gy, H);
            gx = wrap3(gx, W);
        } else if (gy >= H || gx >= W) {
            return 0u;
        }
        const uint8_t* m = reinterpret_cast<const uint8_t*>(P.mask) + (int64_t)b * K * Nu;
        return __ldg(reinterpret_cast<const uint32_t*>(m + ((uint32_t)(2 * j + (isx ? 1 : 0)) * Nu + (uint32_t)(gy * W + gx))));
    };

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
    if (MST && t < n_tiles) {  // the first tile's masks, synchronously
        int b, y0, x0;
        tile_origin(t, b, y0, x0);
        for (int idx = threadIdx.x; idx < NWY; idx += NT) (&S.mY[0][0][0][0])[idx] = mask_word(false, idx, b, y0, x0);
        for (int idx = threadIdx.x; idx < NWX; idx += NT) (&S.mX[0][0][0][0])[idx] = mask_word(true, idx, b, y0, x0);
    }
    __syncthreads();
    uint32_t phY = 0, phX = 0, it = 0;
    int b_tab = -1;

    for (; t < n_tiles; t += (int)gridDim.x, ++it) {
        int b, y0, x0;
        tile_origin(t, b, y0, x0);
        if (LAB && b != b_tab) {  // CTA-uniform: the sample's weight pairs (a CTA changes sample a handful of times)
            __syncthreads();
            if (threadIdx.x < 2 * K) (&s_w[0][0])[threadIdx.x] = P.wtab[(int64_t)b * 2 * K + threadIdx.x];
            __syncthreads();
            b_tab = b;
        }
        int bn = 0, y0n = 0, x0n = 0;
        const bool has_next = t + (int)gridDim.x < n_tiles;
        if (has_next) tile_origin(t + (int)gridDim.x, bn, y0n, x0n);
        const int x = x0 + tx, y = y0 + ty;
        const bool live = (x < W) && (y < H);
        const uint32_t pix = (uint32_t)((live ? y : 0) * W + (live ? x : 0));
        // circular border: TMA zero-fills what lies outside the image; those elements are patched by hand afterwards
        const bool patch = wrap && !tile_inner(y0, x0);
        const uint32_t goal = kWarps * (it + 1) - 1;
        const int64_t sample = (int64_t)b * K * Nu;

        // ---- mask of the pair anchored here and of the pair anchored s below / s to the right, for one axis
        uint32_t mky[NS][2] = {}, mkx[NS][2] = {};
        // `whole`: the tile and the anchors s below / to the right of it lie inside the image -- no border rule, no dead lanes
        const bool whole = y0 + TY + (CROSS ? 0 : PM) <= H && x0 + TX + (CROSS ? 0 : PM) <= W;
        auto fetch_masks = [&](bool isx, uint32_t (&mk)[NS][2]) {
            if constexpr (MASK != PEA_MASK_NONE) {
                const uint8_t* m8 = reinterpret_cast<const uint8_t*>(P.mask) + sample;
                const float* mf = reinterpret_cast<const float*>(P.mask) + sample;
                if (whole) {
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        const uint32_t o = (uint32_t)(2 * j + (isx ? 1 : 0)) * Nu + pix;
                        const uint32_t o2 = o + (uint32_t)(isx ? P.shift[j] : P.shift[j] * W);
                        if (MASK == PEA_MASK_U8) {
                            mk[j][0] = ldg_u8_early(m8 + o, true);
                            if (REV) mk[j][1] = ldg_u8_early(m8 + o2, true);
                        } else {
                            mk[j][0] = ldg_b32_early(mf + o, true);
                            if (REV) mk[j][1] = ldg_b32_early(mf + o2, true);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        const int s = P.shift[j];
                        const uint32_t koff = (uint32_t)(2 * j + (isx ? 1 : 0)) * Nu;
                        int qy = y + (isx ? 0 : s), qx = x + (isx ? s : 0);
                        bool vr = live;
                        if (wrap) {
                            qy = wrap1(qy, H);
                            qx = wrap1(qx, W);
                        } else {
                            vr = vr && qy < H && qx < W;
                        }
                        const uint32_t o2 = koff + (vr ? (uint32_t)(qy * W + qx) : 0u);
                        if (MASK == PEA_MASK_U8) {
                            mk[j][0] = ldg_u8_early(m8 + (koff + pix), live);
                            if (REV) mk[j][1] = ldg_u8_early(m8 + o2, vr);
                        } else {
                            mk[j][0] = ldg_b32_early(mf + (koff + pix), live);
                            if (REV) mk[j][1] = ldg_b32_early(mf + o2, vr);
                        }
                    }
                }
            }
        };
        uint32_t nextY[MST ? RY : 1], nextX[MST ? RX : 1];
        const int mb = (int)(it & 1u);  // this tile's mask buffer; the next tile's is mb ^ 1
        if (MST) {
            // the next tile's y group is requested now and stored before the y boxes are released (the release /
            // acquire chain of the box refill makes it visible); this tile's masks are read after the barrier wait
            if (has_next) {
#pragma unroll
                for (int r = 0; r < RY; ++r) {
                    const int idx = threadIdx.x + r * NT;
                    nextY[r] = idx < NWY ? mask_word(false, idx, bn, y0n, x0n) : 0u;
                }
            }
        } else {
            fetch_masks(false, mky);
        }

        mbar_wait(&barY, phY & 1u);
        if (patch) {
            const float* src = P.par + (int64_t)b * kC * Nu;
            float* sYf = &S.sY[0][0][0];
            for (int idx = threadIdx.x; idx < YR * TX; idx += NT) {
                const int gy = y0 - PM + (idx >> 5), gx = x0 + (idx & 31);
                if ((unsigned)gy < (unsigned)H && gx < W) continue;
                const uint32_t o = (uint32_t)(wrap3(gy, H) * W + wrap3(gx, W));
                float v[kC];
#pragma unroll
                for (int c = 0; c < kC; ++c) v[c] = __ldg(src + (o + (uint32_t)c * Nu));
#pragma unroll
                for (int c = 0; c < kC; ++c) sYf[c * CHS_Y + idx] = v[c];
            }
            const float* tg = P.target + sample;
            const float* wg = P.weight + sample;
            if (!LAB)  // labels are never patched: gen_affs_ours does not wrap, pairs across the image edge are masked out
            for (int idx = threadIdx.x; idx < GY * TX; idx += NT) {
                const int gy = y0 + (idx >> 5), gx = x0 + (idx & 31);
                if (gy < H && gx < W) continue;
                const uint32_t o = (uint32_t)(wrap3(gy, H) * W + wrap3(gx, W));
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    (&S.tY[j][0][0])[idx] = __ldg(tg + (o + (uint32_t)(2 * j) * Nu));
                    (&S.wY[j][0][0])[idx] = __ldg(wg + (o + (uint32_t)(2 * j) * Nu));
                }
            }
            __syncthreads();  // patched y boxes complete
        }

        if (MST) {  // after the wait on barY: the stores of the tile before are visible
            const uint8_t* my8 = reinterpret_cast<const uint8_t*>(&S.mY[MST ? mb : 0][0][0][0]) + ty * TX + tx;
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                mky[j][0] = my8[j * GY * TX];
                if (REV) mky[j][1] = my8[j * GY * TX + P.shift[j] * TX];
            }
        }
        const float* ownY = &S.sY[0][ty + PM][tx];
        const float* ownX = &S.sX[0][ty][tx + PM];
        float* affs_b = P.affs ? P.affs + sample + pix : nullptr;
        const bool st_ok = affs_b != nullptr && live;
        f32x2 e[kC2], G[kC2];
        float rp;
        {
            f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; ++i)
                e[i] = CROSS ? pack2(S.sO[CROSS ? 2 * i : 0][CROSS ? ty : 0][tx], S.sO[CROSS ? 2 * i + 1 : 0][CROSS ? ty : 0][tx])
                             : pack2(ownY[(2 * i) * CHS_Y], ownY[(2 * i + 1) * CHS_Y]);
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
        if (!MST) fetch_masks(true, mkx);  // needed after the y part: a third of a tile ahead

        // ---------------- y part: channel 2j pairs p with (y - s, x); the pair anchored at (y + s, x) has p as partner
        {
            const float* tq = &S.tY[0][ty][tx];
            const float* wq = &S.wY[0][ty][tx];
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const int s = P.shift[j];
                // zero border: a partner outside the image is zero fill (a == 0) and the pair is dropped
                const bool vf = live && (wrap || y >= s);
                float t_f, w_f, t_r = 0.f, w_r = 0.f;
                if (LAB) {
                    // gen_affs_ours: same label (ignore: and both non-zero); the mask drops pairs across the image edge
                    const int32_t* lq = &S.lY[LAB ? ty + PM : 0][tx];
                    const int32_t lp = lq[0], lm = lq[LAB ? -s * TX : 0], ll = lq[LAB ? s * TX : 0];
                    const bool same_f = lp == lm && (!P.ignore || (lp != 0 && lm != 0));
                    const bool same_r = ll == lp && (!P.ignore || (lp != 0 && ll != 0));
                    t_f = same_f ? 1.f : 0.f;
                    w_f = (live && y >= s) ? s_w[2 * j][same_f ? 0 : 1] : 0.f;
                    t_r = same_r ? 1.f : 0.f;
                    w_r = (y + s < H) ? s_w[2 * j][same_r ? 0 : 1] : 0.f;
                } else {
                    t_f = tq[j * GY * TX];
                    w_f = vf ? wq[j * GY * TX] * mask_sq<MASK>(mky[j][0]) : 0.f;
                    if (REV) {
                        t_r = tq[j * GY * TX + s * TX];
                        w_r = wq[j * GY * TX + s * TX] * mask_sq<MASK>(mky[j][1]);
                    }
                }
                float a;
                lossacc[2 * j] += pair<CHS_Y, GRAD>(ownY - s * TX, e, G, t_f, w_f, P.gcoef[2 * j], a);
                st_affs(affs_b, (uint32_t)(2 * j) * Nu, a, st_ok);
                if (REV) {
                    // rows below the image are zero fill in the weight box: pair dropped
                    float a2;
                    pair<CHS_Y, true>(ownY + s * TX, e, G, t_r, w_r, P.gcoef[2 * j], a2);
                }
            }
        }
        if (MST && has_next) {
#pragma unroll
            for (int r = 0; r < RY; ++r) {
                const int idx = threadIdx.x + r * NT;
                if (idx < NWY) (&S.mY[MST ? mb ^ 1 : 0][0][0][0])[idx] = nextY[r];
            }
#pragma unroll
            for (int r = 0; r < RX; ++r) {
                const int idx = threadIdx.x + r * NT;
                nextX[r] = idx < NWX ? mask_word(true, idx, bn, y0n, x0n) : 0u;
            }
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntY) == goal && has_next) issueY(bn, y0n, x0n);

        // ---------------- x part
        mbar_wait(&barX, phX & 1u);
        if (patch) {
            const float* src = P.par + (int64_t)b * kC * Nu;
            float* sXf = &S.sX[0][0][0];
            for (int idx = threadIdx.x; idx < TY * XW; idx += NT) {
                const int row = idx / XW, col = idx - row * XW;
                const int gy = y0 + row, gx = x0 - PM + col;
                if (gy < H && (unsigned)gx < (unsigned)W) continue;
                const uint32_t o = (uint32_t)(wrap3(gy, H) * W + wrap3(gx, W));
                float v[kC];
#pragma unroll
                for (int c = 0; c < kC; ++c) v[c] = __ldg(src + (o + (uint32_t)c * Nu));
#pragma unroll
                for (int c = 0; c < kC; ++c) sXf[c * CHS_X + idx] = v[c];
            }
            const float* tg = P.target + sample;
            const float* wg = P.weight + sample;
            if (!LAB)
            for (int idx = threadIdx.x; idx < TY * GX; idx += NT) {
                const int row = idx / GX, col = idx - row * GX;
                const int gy = y0 + row, gx = x0 + col;
                if (gy < H && gx < W) continue;
                const uint32_t o = (uint32_t)(wrap3(gy, H) * W + wrap3(gx, W));
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    (&S.tX[j][0][0])[idx] = __ldg(tg + (o + (uint32_t)(2 * j + 1) * Nu));
                    (&S.wX[j][0][0])[idx] = __ldg(wg + (o + (uint32_t)(2 * j + 1) * Nu));
                }
            }
            __syncthreads();  // patched x boxes complete
        }
        if (MST) {
            const uint8_t* mx8 = reinterpret_cast<const uint8_t*>(&S.mX[MST ? mb : 0][0][0][0]) + ty * GX + tx;
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                mkx[j][0] = mx8[j * TY * GX];
                if (REV) mkx[j][1] = mx8[j * TY * GX + P.shift[j]];
            }
        }
        {
            const float* tq = &S.tX[0][ty][tx];
            const float* wq = &S.wX[0][ty][tx];
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const int s = P.shift[j];
                const bool vf = live && (wrap || x >= s);
                float t_f, w_f, t_r = 0.f, w_r = 0.f;
                if (LAB) {
                    const int32_t* lq = &S.lX[LAB ? ty : 0][tx + PM];
                    const int32_t lp = lq[0], lm = lq[LAB ? -s : 0], ll = lq[LAB ? s : 0];
                    const bool same_f = lp == lm && (!P.ignore || (lp != 0 && lm != 0));
                    const bool same_r = ll == lp && (!P.ignore || (lp != 0 && ll != 0));
                    t_f = same_f ? 1.f : 0.f;
                    w_f = (live && x >= s) ? s_w[2 * j + 1][same_f ? 0 : 1] : 0.f;
                    t_r = same_r ? 1.f : 0.f;
                    w_r = (x + s < W) ? s_w[2 * j + 1][same_r ? 0 : 1] : 0.f;
                } else {
                    t_f = tq[j * TY * GX];
                    w_f = vf ? wq[j * TY * GX] * mask_sq<MASK>(mkx[j][0]) : 0.f;
                    if (REV) {
                        t_r = tq[j * TY * GX + s];
                        w_r = wq[j * TY * GX + s] * mask_sq<MASK>(mkx[j][1]);
                    }
                }
                float a;
                lossacc[2 * j + 1] += pair<CHS_X, GRAD>(ownX - s, e, G, t_f, w_f, P.gcoef[2 * j + 1], a);
                st_affs(affs_b, (uint32_t)(2 * j + 1) * Nu, a, st_ok);
                if (REV) {
                    float a2;
                    pair<CHS_X, true>(ownX + s, e, G, t_r, w_r, P.gcoef[2 * j + 1], a2);
                }
            }
        }
        if (MST && has_next) {
#pragma unroll
            for (int r = 0; r < RX; ++r) {
                const int idx = threadIdx.x + r * NT;
                if (idx < NWX) (&S.mX[MST ? mb ^ 1 : 0][0][0][0])[idx] = nextX[r];
            }
        }
        __syncwarp();
        if (tx == 0 && arrive_count(&cntX) == goal && has_next) issueX(bn, y0n, x0n);
        ++phY;
        ++phX;

        // ---------------- gradient back through the normalisation:  (G - e^ <e^, G>) / |x|
        if (GRAD && live) {
            f32x2 p0 = 0ull, p1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                p0 = fma2(e[i], G[i], p0);
                p1 = fma2(e[i + 1], G[i + 1], p1);
            }
            float eg = hsum2(p0, p1);
            if (!(rp < 1e12f)) eg = 0.f;  // eps branch of F.normalize: G / eps
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

    // ---------------- loss partials: thread -> warp -> CTA row (fixed order)
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float v = warp_sum(lossacc[k]);
        if (tx == 0) s_red[k][ty] = v;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < TY; ++i) v += s_red[threadIdx.x][i];
        P.partials[(int64_t)blockIdx.x * K + threadIdx.x] = v;
    }
}

}  // namespace c2

// ---------------------------------------------------------------------------------------------------------------
// channel 2j = [0, -s_j, 0], channel 2j + 1 = [0, 0, -s_j], 1 <= s_j <= 12: multi_offset(shifts, neighbor=4)
static bool canon_shifts(const OffsetTable& tab, int* shifts) {
    if (tab.K != c2::K) return false;
    for (int j = 0; j < c2::NS; ++j) {
        const int s = -tab.dy[2 * j];
        if (s < 1 || s > c2::PM) return false;
        if (tab.dz[2 * j] || tab.dx[2 * j] || tab.dz[2 * j + 1] || tab.dy[2 * j + 1] || tab.dx[2 * j + 1] != -s) return false;
        shifts[j] = s;
    }
    return true;
}

bool canon2d_eligible(int C, int D, int H, int W, const OffsetTable& tab, const float* emb, const float* partner,
                      const float* target, const float* weight) {
    int sh[c2::NS];
    if (C != 16 || D != 1 || !canon_shifts(tab, sh)) return false;
    if ((reinterpret_cast<uintptr_t>(emb) | reinterpret_cast<uintptr_t>(partner) | reinterpret_cast<uintptr_t>(target) |
         reinterpret_cast<uintptr_t>(weight)) & 15)
        return false;
    // the hand patch of edge tiles maps coordinates through wrap3 (-n <= v < 4n); TMA needs a 16-byte row pitch
    if (W % 4 != 0 || H < 32 || W < 64) return false;
    return (int64_t)16 * H * W < ((int64_t)1 << 31);
}

// staged uint8 masks are read as aligned 32-bit words
bool canon2d_mask_ok(const void* mask, int mask_dtype) {
    return !(PEA_CANON_MASK_STAGE && mask_dtype == PEA_MASK_U8 && (reinterpret_cast<uintptr_t>(mask) & 3));
}

static int canon_grid(int n_tiles) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        sms = 148;
    int g = sms;
    // test knob (pea_debug_configure): a small grid makes every CTA walk many tiles even on small images
    if (const int c = debug_knob(PEA_KNOB_MAX_CTAS)) {
        if (c > 0 && c < g) g = c;
    }
    return n_tiles < g ? n_tiles : g;
}

// rows of loss partials the kernel writes
int canon2d_blocks(int B, int H, int W) {
    return canon_grid(B * ((H + c2::TY - 1) / c2::TY) * ((W + c2::TX - 1) / c2::TX));
}

template <int MASK, bool CROSS, bool LAB = false>
static int launch_canon_m(const CUtensorMap* m, const c2::Params& P, int grid, cudaStream_t st) {
    const size_t smem = sizeof(c2::Stage<CROSS, LAB>) + 128;
    cudaError_t e;
    if (P.grad) {
        auto kern = c2::fused2d_canon<MASK, true, CROSS, LAB>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, c2::NT, smem, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], P);
    } else {
        auto kern = c2::fused2d_canon<MASK, false, CROSS, LAB>;
        if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
        kern<<<grid, c2::NT, smem, st>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], P);
    }
    return (int)cudaGetLastError();
}

template <bool CROSS>
static int launch_canon_c(const float* emb, const float* partner, const Dims& d, c2::Params& P, const float* target,
                          const float* weight, int mask_dtype, cudaStream_t st) {
    using G_ = c2::Geo<CROSS>;
    const float* staged = CROSS ? partner : emb;
    CUtensorMap m[7];
    int rc;
    if ((rc = encode_volume_map(&m[0], staged, d.B, 16, 1, d.H, d.W, c2::TX, G_::YR, 1, 16))) return rc;
    if ((rc = encode_volume_map(&m[1], staged, d.B, 16, 1, d.H, d.W, G_::XW, c2::TY, 1, 16))) return rc;
    // every second channel from the start coordinate: 5 planes (c0, c0 + 2, .., c0 + 8) per copy
    if ((rc = encode_volume_map(&m[2], target, d.B, c2::K, 1, d.H, d.W, c2::TX, G_::GY, 1, c2::K - 1, 0, 2))) return rc;
    if ((rc = encode_volume_map(&m[3], weight, d.B, c2::K, 1, d.H, d.W, c2::TX, G_::GY, 1, c2::K - 1, 0, 2))) return rc;
    if ((rc = encode_volume_map(&m[4], target, d.B, c2::K, 1, d.H, d.W, G_::GX, c2::TY, 1, c2::K - 1, 0, 2))) return rc;
    if ((rc = encode_volume_map(&m[5], weight, d.B, c2::K, 1, d.H, d.W, G_::GX, c2::TY, 1, c2::K - 1, 0, 2))) return rc;
    if ((rc = encode_volume_map(&m[6], emb, d.B, 16, 1, d.H, d.W, c2::TX, c2::TY, 1, 16))) return rc;  // CROSS: own tile
    P.par = staged;
    const int grid = canon_grid(P.n_tiles);
    if (mask_dtype == PEA_MASK_U8) return launch_canon_m<PEA_MASK_U8, CROSS>(m, P, grid, st);
    if (mask_dtype == PEA_MASK_F32) return launch_canon_m<PEA_MASK_F32, CROSS>(m, P, grid, st);
    return launch_canon_m<PEA_MASK_NONE, CROSS>(m, P, grid, st);
}

// `partner` != nullptr: two-view loss, gradient w.r.t. the anchor view only
int launch_canon2d(const float* emb, const float* partner, const Dims& d, const OffsetTable& tab, int wrap, const float* target,
                   const float* weight, const void* mask, int mask_dtype, float* affs, float* partials, float* grad,
                   cudaStream_t st) {
    c2::Params P;
    if (!canon_shifts(tab, P.shift)) return PEA_E_MODE;
    P.B = d.B; P.H = d.H; P.W = d.W; P.wrap = wrap;
    P.tiles_x = (d.W + c2::TX - 1) / c2::TX;
    P.tiles_y = (d.H + c2::TY - 1) / c2::TY;
    P.n_tiles = d.B * P.tiles_x * P.tiles_y;
    for (int k = 0; k < c2::K; ++k) P.gcoef[k] = tab.gcoef[k];
    P.emb = emb; P.target = target; P.weight = weight; P.mask = mask;
    P.labels = nullptr; P.wtab = nullptr; P.ignore = 0;
    P.affs = affs; P.grad = grad; P.partials = partials;
    if (partner) return launch_canon_c<true>(emb, partner, d, P, target, weight, mask_dtype, st);
    return launch_canon_c<false>(emb, partner, d, P, target, weight, mask_dtype, st);
}

// Label-driven flavour: target = gen_affs_ours(labels, offsets, ignore, padding)[0], mask = [1], weight =
// weight_binary_ratio per (sample, channel) -- all derived inside the loss kernel from the int32 label plane and the
// weight-pair table of pea_gt_pair_weights: 172 instead of 258 algorithmic bytes per pixel.
int launch_canon2d_labels(const float* emb, const int32_t* labels, const float* wtab, int ignore, const Dims& d,
                          const OffsetTable& tab, int wrap, float* affs, float* partials, float* grad, cudaStream_t st) {
    c2::Params P;
    if (!canon_shifts(tab, P.shift)) return PEA_E_MODE;
    using G_ = c2::Geo<false>;
    const float* lab = reinterpret_cast<const float*>(labels);  // 4-byte elements, copied bit for bit
    CUtensorMap m[7];
    int rc;
    if ((rc = encode_volume_map(&m[0], emb, d.B, 16, 1, d.H, d.W, c2::TX, G_::YR, 1, 16))) return rc;
    if ((rc = encode_volume_map(&m[1], emb, d.B, 16, 1, d.H, d.W, G_::XW, c2::TY, 1, 16))) return rc;
    if ((rc = encode_volume_map(&m[2], lab, d.B, 1, 1, d.H, d.W, c2::TX, G_::YR, 1, 1))) return rc;
    if ((rc = encode_volume_map(&m[4], lab, d.B, 1, 1, d.H, d.W, G_::XW, c2::TY, 1, 1))) return rc;
    m[3] = m[2]; m[5] = m[4]; m[6] = m[0];  // unused by this flavour
    P.B = d.B; P.H = d.H; P.W = d.W; P.wrap = wrap;
    P.tiles_x = (d.W + c2::TX - 1) / c2::TX;
    P.tiles_y = (d.H + c2::TY - 1) / c2::TY;
    P.n_tiles = d.B * P.tiles_x * P.tiles_y;
    for (int k = 0; k < c2::K; ++k) P.gcoef[k] = tab.gcoef[k];
    P.emb = emb; P.par = emb; P.target = nullptr; P.weight = nullptr; P.mask = nullptr;
    P.labels = labels; P.wtab = wtab; P.ignore = ignore;
    P.affs = affs; P.grad = grad; P.partials = partials;
    return launch_canon_m<PEA_MASK_NONE, false, true>(m, P, canon_grid(P.n_tiles), st);
}

}  // namespace pea
