// K1c: forward-only 2-D affinities for the axis-aligned tables multi_offset([s0..], neighbor=4) -- channel 2j is the y shift
// [-s_j, 0], channel 2j + 1 the x shift [0, -s_j] -- i.e. embedding2affs of scripts_cvppp/loss/loss_embedding_mse.py:58-66
// (helper embedding_single_offset :49-56) at every shipped 2-D table: BBBC039V1 [1, 3, 5, 9, 11] (bbbc039v1.yaml:65-71) and
// CVPPP [1, 3, 5, 9, 27] (cvppp.yaml:69-75).  This is the inference op of scripts_cvppp (validation in main.py, inference.py),
// optionally with the caller's F.relu (main.py:312) folded into the store.
//
// Until round 2 forward-only calls took the round-1 tiled kernel (a cooperative fill that normalises every staged pixel,
// run-time offset table): 0.260 ms at B = 8, 520 x 696 -- as long as the whole fused forward + backward.  Forward only,
// a pixel needs its own vector and K / 2 partners above it and K / 2 to its left: ~50 instructions per offset.
//   * persistent CTAs of 16 warps walk 16 x 32 tiles; per tile the RAW channel-planar y-column box (PM + 16 rows) and
//     x-row box (PM + 32 columns) arrive by TMA -- upper / left halo only;
//   * PM = 12 (shifts <= 12): two sets of boxes, the next tile's copies are in flight while the current tile is computed;
//     PM = 28 (CVPPP's shift 27): one set (151 KB), the copies of the next tile are issued when the tile is done;
//   * offset loops fully unrolled over a compile-time number of shifts, shifts themselves uniform run-time values: every
//     shared-memory operand is base + immediate; the partner's 1 / |x| is recomputed per pair (8 FFMA2 + MUFU), nothing
//     normalised is written anywhere;
//   * circular border (torch.roll): TMA zero-fills what lies above / left of the image; on those tiles the zero-filled
//     elements are patched by hand through the wrap rule (one CTA barrier).  Zero border: the zero fill IS the result.
#include "pea_common.cuh"
#include "pea_tma.cuh"

namespace pea {
#ifndef PEA_K1C_BRANCHY_STORE
#define PEA_K1C_BRANCHY_STORE 0
#endif

namespace f2 {

__device__ __forceinline__ void st_pred(float* p, float v, bool pred) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"((int)pred));
}

constexpr int kC = 16, kC2 = 8;
constexpr int TX = 32, TY = 16, NT = TY * 32, NSMAX = 6;

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
__device__ __forceinline__ int wrap_lo(int v, int n) { return v < 0 ? v + n : v; }  // -n <= v < n

struct Params {
    int B, H, W;
    int wrap, relu;
    int tiles_x, tiles_y, n_tiles;
    int ns;             // shifts in the table (K = 2 ns)
    int shift[NSMAX];   // s_j > 0
    const float* emb;
    float* affs;
};

template <int PM>
struct alignas(128) Boxes {
    float sY[kC][TY + PM][TX];  // rows y0 - PM .. y0 + 15
    float sX[kC][TY][TX + PM];  // columns x0 - PM .. x0 + 31
};

// <e^_p, e^_q> for the own UNIT vector e and the raw partner at nb (channels CHS floats apart)
template <int CHS>
__device__ __forceinline__ float affinity(const float* __restrict__ nb, const f32x2 (&e)[kC2]) {
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
    return hsum2(d0, d1) * inv_norm(hsum2(ss0, ss1));
}

template <int PM, int NBUF, int NS>
__global__ void __launch_bounds__(NT, 1)
fwd2d_canon(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapX, const Params P) {
    constexpr int YR = TY + PM, XW = TX + PM, CHS_Y = YR * TX, CHS_X = TY * XW;
    constexpr uint32_t kBytes = sizeof(Boxes<PM>);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Boxes<PM>* S = reinterpret_cast<Boxes<PM>*>(smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u));
    __shared__ uint64_t bar[NBUF];

    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int H = P.H, W = P.W;
    const bool wrap = P.wrap != 0;
    const uint32_t Nu = (uint32_t)(H * W);
    const int K = 2 * NS;

    auto tile_origin = [&](int t, int& b, int& y0, int& x0) {
        const int r = t / P.tiles_x;
        x0 = (t - r * P.tiles_x) * TX;
        b = r / P.tiles_y;
        y0 = (r - b * P.tiles_y) * TY;
    };
    auto issue = [&](int t, int buf) {
        int b, y0, x0;
        tile_origin(t, b, y0, x0);
        fence_proxy_async_smem();  // hand-patched elements of the set's last use are ordered before the copy
        mbar_arrive_expect_tx(&bar[buf], kBytes);
        tma_load_5d(&S[buf].sY[0][0][0], &mapY, &bar[buf], x0, y0 - PM, 0, 0, b);
        tma_load_5d(&S[buf].sX[0][0][0], &mapX, &bar[buf], x0 - PM, y0, 0, 0, b);
    };

    int t = blockIdx.x;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NBUF; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
        if (t < P.n_tiles) issue(t, 0);
    }
    __syncthreads();
    uint32_t it = 0;
    for (; t < P.n_tiles; t += (int)gridDim.x, ++it) {
        const int buf = NBUF == 2 ? (int)(it & 1u) : 0;
        const uint32_t phase = NBUF == 2 ? ((it >> 1) & 1u) : (it & 1u);
        const int tn = t + (int)gridDim.x;
        // two sets of boxes: the other set was last read by the tile before, which every warp has left (barrier below)
        if (NBUF == 2 && threadIdx.x == 0 && tn < P.n_tiles) issue(tn, buf ^ 1);
        int b, y0, x0;
        tile_origin(t, b, y0, x0);
        const int x = x0 + tx, y = y0 + ty;
        const bool live = x < W && y < H;
        Boxes<PM>& Sb = S[buf];
        mbar_wait(&bar[buf], phase);
        if (wrap && (y0 < PM || x0 < PM)) {  // CTA-uniform: rows above / columns left of the image were zero-filled
            const float* src = P.emb + (int64_t)b * kC * Nu;
            if (y0 < PM) {
                float* sYf = &Sb.sY[0][0][0];
                const int rows = PM - y0;  // box rows 0 .. rows - 1 lie above the image
                for (int idx = threadIdx.x; idx < rows * TX; idx += NT) {
                    const int gy = y0 - PM + (idx >> 5) + H, gx = x0 + (idx & 31);
                    if (gx < W) {
                        const uint32_t o = (uint32_t)(gy * W + gx);
#pragma unroll
                        for (int c = 0; c < kC; ++c) sYf[c * CHS_Y + idx] = __ldg(src + (o + (uint32_t)c * Nu));
                    }
                }
            }
            if (x0 < PM) {
                float* sXf = &Sb.sX[0][0][0];
                const int cols = PM - x0;  // box columns 0 .. cols - 1 lie left of the image
                for (int idx = threadIdx.x; idx < TY * cols; idx += NT) {
                    const int row = idx / cols, col = idx - row * cols;
                    const int gy = y0 + row, gx = x0 - PM + col + W;
                    if (gy < H) {
                        const uint32_t o = (uint32_t)(gy * W + gx);
#pragma unroll
                        for (int c = 0; c < kC; ++c) sXf[c * CHS_X + row * XW + col] = __ldg(src + (o + (uint32_t)c * Nu));
                    }
                }
            }
            __syncthreads();
        }
        const float* ownY = &Sb.sY[0][ty + PM][tx];
        const float* ownX = &Sb.sX[0][ty][tx + PM];
        f32x2 e[kC2];
        {
            f32x2 s0 = 0ull, s1 = 0ull;
#pragma unroll
            for (int i = 0; i < kC2; ++i) e[i] = pack2(ownY[(2 * i) * CHS_Y], ownY[(2 * i + 1) * CHS_Y]);
#pragma unroll
            for (int i = 0; i < kC2; i += 2) {
                s0 = fma2(e[i], e[i], s0);
                s1 = fma2(e[i + 1], e[i + 1], s1);
            }
            const float rp = inv_norm(hsum2(s0, s1));
            const f32x2 rp2 = pack2(rp, rp);
#pragma unroll
            for (int i = 0; i < kC2; ++i) asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(e[i]) : "l"(e[i]), "l"(rp2));
        }
        float* out = P.affs + (int64_t)b * K * Nu + (uint32_t)((live ? y : 0) * W + (live ? x : 0));
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const int s = P.shift[j];
            float ay = affinity<CHS_Y>(ownY - s * TX, e);
            float ax = affinity<CHS_X>(ownX - s, e);
            if (P.relu) {
                ay = fmaxf(ay, 0.f);
                ax = fmaxf(ax, 0.f);
            }
#if PEA_K1C_BRANCHY_STORE
            if (live) {
                out[(uint32_t)(2 * j) * Nu] = ay;
                out[(uint32_t)(2 * j + 1) * Nu] = ax;
            }
#else
            // predicated stores without a branch: `if (live)` makes ptxas sink both pair evaluations into the branch,
            // one reconvergence region per shift that nothing is scheduled across (`out` is valid for dead lanes)
            st_pred(out + (uint32_t)(2 * j) * Nu, ay, live);
            st_pred(out + (uint32_t)(2 * j + 1) * Nu, ax, live);
#endif
        }
        __syncthreads();  // every warp has left this set of boxes
        if (NBUF == 1 && threadIdx.x == 0 && tn < P.n_tiles) issue(tn, 0);
    }
}

}  // namespace f2

// channel 2j = [0, -s_j, 0], channel 2j + 1 = [0, 0, -s_j], 1 <= s_j <= 28, K = 2 .. 12
static int fwd_shifts(const OffsetTable& tab, int* shifts) {
    if (tab.K < 2 || tab.K > 2 * f2::NSMAX || (tab.K & 1)) return -1;
    int smax = 0;
    for (int j = 0; j < tab.K / 2; ++j) {
        const int s = -tab.dy[2 * j];
        if (s < 1 || s > 28) return -1;
        if (tab.dz[2 * j] || tab.dx[2 * j] || tab.dz[2 * j + 1] || tab.dy[2 * j + 1] || tab.dx[2 * j + 1] != -s) return -1;
        shifts[j] = s;
        smax = s > smax ? s : smax;
    }
    return smax;
}

bool fwd2d_canon_eligible(int C, int D, int H, int W, const OffsetTable& tab, const float* emb) {
    int sh[f2::NSMAX];
    if (C != 16 || D != 1 || fwd_shifts(tab, sh) < 0) return false;
    if (reinterpret_cast<uintptr_t>(emb) & 15) return false;
    if (W % 4 != 0 || H < 32 || W < 64) return false;  // 16-byte row pitch for TMA; the hand patch wraps once
    return (int64_t)16 * H * W < ((int64_t)1 << 31);
}

template <int PM, int NBUF, int NS>
static int launch_fwd_ns(const CUtensorMap* m, const f2::Params& P, int grid, cudaStream_t st) {
    const size_t smem = NBUF * sizeof(f2::Boxes<PM>) + 128;
    auto kern = f2::fwd2d_canon<PM, NBUF, NS>;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return (int)e;
    kern<<<grid, f2::NT, smem, st>>>(m[0], m[1], P);
    return (int)cudaGetLastError();
}

template <int PM, int NBUF>
static int launch_fwd_pm(const float* emb, const Dims& d, f2::Params& P, cudaStream_t st) {
    CUtensorMap m[2];
    int rc;
    if ((rc = encode_volume_map(&m[0], emb, d.B, 16, 1, d.H, d.W, f2::TX, f2::TY + PM, 1, 16))) return rc;
    if ((rc = encode_volume_map(&m[1], emb, d.B, 16, 1, d.H, d.W, f2::TX + PM, f2::TY, 1, 16))) return rc;
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        sms = 148;
    int grid = sms;
    if (const int c = debug_knob(PEA_KNOB_MAX_CTAS)) {
        if (c > 0 && c < grid) grid = c;
    }
    if (P.n_tiles < grid) grid = P.n_tiles;
    switch (P.ns) {
        case 1: return launch_fwd_ns<PM, NBUF, 1>(m, P, grid, st);
        case 2: return launch_fwd_ns<PM, NBUF, 2>(m, P, grid, st);
        case 3: return launch_fwd_ns<PM, NBUF, 3>(m, P, grid, st);
        case 4: return launch_fwd_ns<PM, NBUF, 4>(m, P, grid, st);
        case 5: return launch_fwd_ns<PM, NBUF, 5>(m, P, grid, st);
        default: return launch_fwd_ns<PM, NBUF, 6>(m, P, grid, st);
    }
}

int launch_fwd2d_canon(const float* emb, const Dims& d, const OffsetTable& tab, int wrap, int relu, float* affs, cudaStream_t st) {
    f2::Params P;
    const int smax = fwd_shifts(tab, P.shift);
    if (smax < 0) return PEA_E_MODE;
    P.B = d.B; P.H = d.H; P.W = d.W; P.wrap = wrap; P.relu = relu;
    P.tiles_x = (d.W + f2::TX - 1) / f2::TX;
    P.tiles_y = (d.H + f2::TY - 1) / f2::TY;
    P.n_tiles = d.B * P.tiles_x * P.tiles_y;
    P.ns = tab.K / 2;
    P.emb = emb; P.affs = affs;
    if (smax <= 12) return launch_fwd_pm<12, 2>(emb, d, P, st);
    return launch_fwd_pm<28, 1>(emb, d, P, st);
}

}  // namespace pea
