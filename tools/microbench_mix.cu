// What can one B200 SM sustain on the pair-evaluation instruction mix of the fused affinity kernels?
// One CTA of NW warps per SM (large dynamic shared memory forces 1 CTA / SM), every thread evaluates `iters` pairs
// against raw channel-planar neighbour vectors in shared memory, in several formulations:
//   0  full pair: 16 LDS.32 + ss (8 FFMA2) + dot (8 FFMA2) + rsqrt + scalar tail + G update (8 FFMA2)   [the kernels]
//   1  neighbour 1/|x| given (one more LDS instead of ss + rsqrt)
//   2  full pair with scalar FFMA (no packed math)
//   3  LDS only (16 per pair)              4  FFMA2 only (24 per pair)
//   5  full pair, two pairs interleaved by hand (ILP 2)
//   6  pair packing: two neighbours per FFMA2 (lane-wise), own vector as scalar broadcast, no horizontal sums
// Prints clocks per warp-pair per SM (lower is better) next to the pipe bounds.   Build: make -C tools microbench_mix
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float hsum2(f32x2 a, f32x2 b) { float a0, a1, b0, b1; unpack2(a, a0, a1); unpack2(b, b0, b1); return (a0 + a1) + (b0 + b1); }
__device__ __forceinline__ float inv_norm(float ss) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(ss)); return fminf(r, 1e12f); }

constexpr int ROWS = 64;            // rows of the staged plane
constexpr int CHS = ROWS * 32;      // channel stride in floats

template <int V>
__global__ void __launch_bounds__(512, 1) mix_kernel(float* out, int iters) {
    extern __shared__ float sm[];   // [16][ROWS][32] + one extra plane of 1/|x|
    for (int i = threadIdx.x; i < 17 * CHS; i += blockDim.x) sm[i] = 0.25f + (i % 97) * 0.01f;
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    f32x2 e[8], G[8], G2[8];
    float es[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { e[i] = pack2(0.1f * i + tx * 0.001f, 0.05f * i); G[i] = 0ull; G2[i] = 0ull; es[2 * i] = 0.1f * i + tx * 0.002f; es[2 * i + 1] = 0.05f * i + tx * 0.001f; }
    float Gs[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) Gs[i] = 0.f;
    float loss = 0.f;
    int row = ty;
    for (int it = 0; it < iters; ++it) {
        row = (row + 7) & (ROWS - 1);
        const float* nb = sm + row * 32 + tx;
        const float t = 0.5f, w = 1.25f, gco = 0.37f;
        if (V == 0 || V == 1) {
            f32x2 xq[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) xq[i] = pack2(nb[(2 * i) * CHS], nb[(2 * i + 1) * CHS]);
            f32x2 ss0 = 0ull, ss1 = 0ull, d0 = 0ull, d1 = 0ull;
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                if (V == 0) { ss0 = fma2(xq[i], xq[i], ss0); ss1 = fma2(xq[i + 1], xq[i + 1], ss1); }
                d0 = fma2(e[i], xq[i], d0); d1 = fma2(e[i + 1], xq[i + 1], d1);
            }
            const float rq = V == 0 ? inv_norm(hsum2(ss0, ss1)) : nb[16 * CHS];
            const float a = hsum2(d0, d1) * rq, diff = a - t, gr = gco * w * diff * rq;
            const f32x2 gr2 = pack2(gr, gr);
#pragma unroll
            for (int i = 0; i < 8; ++i) G[i] = fma2(gr2, xq[i], G[i]);
            loss += w * diff * diff;
        } else if (V == 7 || V == 8 || V == 9) {
            // hybrids: ss always packed (two distinct register pairs); 7: dot scalar, G packed; 8: dot + G scalar; 9: G scalar
            float xs[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) xs[i] = nb[i * CHS];
            f32x2 ss0 = 0ull, ss1 = 0ull;
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                ss0 = fma2(pack2(xs[2 * i], xs[2 * i + 1]), pack2(xs[2 * i], xs[2 * i + 1]), ss0);
                ss1 = fma2(pack2(xs[2 * i + 2], xs[2 * i + 3]), pack2(xs[2 * i + 2], xs[2 * i + 3]), ss1);
            }
            float dsum;
            if (V == 9) {
                f32x2 d0 = 0ull, d1 = 0ull;
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    d0 = fma2(e[i], pack2(xs[2 * i], xs[2 * i + 1]), d0);
                    d1 = fma2(e[i + 1], pack2(xs[2 * i + 2], xs[2 * i + 3]), d1);
                }
                dsum = hsum2(d0, d1);
            } else {
                float d0 = 0.f, d1 = 0.f;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    d0 = fmaf(es[i], xs[i], d0);
                    d1 = fmaf(es[i + 1], xs[i + 1], d1);
                }
                dsum = d0 + d1;
            }
            const float rq = inv_norm(hsum2(ss0, ss1));
            const float a = dsum * rq, diff = a - t, gr = gco * w * diff * rq;
            if (V == 7) {
                const f32x2 gr2 = pack2(gr, gr);
#pragma unroll
                for (int i = 0; i < 8; ++i) G[i] = fma2(gr2, pack2(xs[2 * i], xs[2 * i + 1]), G[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) Gs[i] = fmaf(gr, xs[i], Gs[i]);
            }
            loss += w * diff * diff;
        } else if (V == 10) {  // 24 FFMA2 of the form (x, x, acc): two distinct register pairs
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) G[i] = fma2(e[i], e[i], G[i]);
        } else if (V == 11) {  // 24 FFMA2 of the form (scalar broadcast, x, acc)
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) G[i] = fma2(pack2(es[i + r], es[i + r]), e[i], G[i]);
        } else if (V == 12) {  // 48 scalar FFMA with three distinct registers
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) Gs[i] = fmaf(es[i], es[(i + 1) & 15], Gs[i]);
        } else if (V == 2) {
            float xq[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) xq[i] = nb[i * CHS];
            float ss0 = 0.f, ss1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                ss0 = fmaf(xq[i], xq[i], ss0); ss1 = fmaf(xq[i + 1], xq[i + 1], ss1);
                d0 = fmaf(es[i], xq[i], d0); d1 = fmaf(es[i + 1], xq[i + 1], d1);
            }
            const float rq = inv_norm(ss0 + ss1);
            const float a = (d0 + d1) * rq, diff = a - t, gr = gco * w * diff * rq;
#pragma unroll
            for (int i = 0; i < 16; ++i) Gs[i] = fmaf(gr, xq[i], Gs[i]);
            loss += w * diff * diff;
        } else if (V == 3) {
#pragma unroll
            for (int i = 0; i < 16; ++i) loss += nb[i * CHS];
        } else if (V == 4) {
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) G[i] = fma2(e[i], e[(i + 1) & 7], G[i]);
        } else if (V == 5) {
            // two pairs per iteration (counted as two), interleaved
            const float* nb2 = nb + 13 * 32;
            f32x2 xq[8], xr[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { xq[i] = pack2(nb[(2 * i) * CHS], nb[(2 * i + 1) * CHS]); xr[i] = pack2(nb2[(2 * i) * CHS], nb2[(2 * i + 1) * CHS]); }
            f32x2 ss0 = 0ull, ss1 = 0ull, d0 = 0ull, d1 = 0ull, tt0 = 0ull, tt1 = 0ull, f0 = 0ull, f1 = 0ull;
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                ss0 = fma2(xq[i], xq[i], ss0); tt0 = fma2(xr[i], xr[i], tt0);
                ss1 = fma2(xq[i + 1], xq[i + 1], ss1); tt1 = fma2(xr[i + 1], xr[i + 1], tt1);
                d0 = fma2(e[i], xq[i], d0); f0 = fma2(e[i], xr[i], f0);
                d1 = fma2(e[i + 1], xq[i + 1], d1); f1 = fma2(e[i + 1], xr[i + 1], f1);
            }
            const float rq = inv_norm(hsum2(ss0, ss1)), rr = inv_norm(hsum2(tt0, tt1));
            const float a = hsum2(d0, d1) * rq, diff = a - t, gr = gco * w * diff * rq;
            const float a2 = hsum2(f0, f1) * rr, diff2 = a2 - t, gr_2 = gco * w * diff2 * rr;
            const f32x2 gq = pack2(gr, gr), gp = pack2(gr_2, gr_2);
#pragma unroll
            for (int i = 0; i < 8; ++i) { G[i] = fma2(gq, xq[i], G[i]); G2[i] = fma2(gp, xr[i], G2[i]); }
            loss += w * diff * diff + w * diff2 * diff2;
        } else if (V == 6) {
            // lane-wise packing of two neighbours (counted as two pairs): X_c = (xq_c, xr_c)
            const float* nb2 = nb + 13 * 32;
            f32x2 X[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) X[i] = pack2(nb[i * CHS], nb2[i * CHS]);
            f32x2 ss0 = 0ull, ss1 = 0ull, d0 = 0ull, d1 = 0ull;
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                ss0 = fma2(X[i], X[i], ss0); ss1 = fma2(X[i + 1], X[i + 1], ss1);
                d0 = fma2(pack2(es[i], es[i]), X[i], d0); d1 = fma2(pack2(es[i + 1], es[i + 1]), X[i + 1], d1);
            }
            float s_lo, s_hi;
            unpack2(add2(ss0, ss1), s_lo, s_hi);
            const f32x2 rq2 = pack2(inv_norm(s_lo), inv_norm(s_hi));
            const f32x2 a2 = mul2(add2(d0, d1), rq2);
            const f32x2 diff2 = add2(a2, pack2(-t, -t));
            const f32x2 gr2 = mul2(mul2(pack2(gco * w, gco * w), diff2), rq2);
            // G_c += gr_q * xq_c + gr_r * xr_c: lane-wise accumulators (16 x f32x2), halves summed once at the end
#pragma unroll
            for (int i = 0; i < 8; ++i) { G[i] = fma2(gr2, X[2 * i], G[i]); G2[i] = fma2(gr2, X[2 * i + 1], G2[i]); }
            float l_lo, l_hi;
            unpack2(mul2(diff2, diff2), l_lo, l_hi);
            loss += w * (l_lo + l_hi);
        }
    }
    float s = loss;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float a, b; unpack2(G[i], a, b); s += a + b; unpack2(G2[i], a, b); s += a + b; }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += Gs[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int V>
void run(const char* name, int sms, float* out, int pairs_per_iter) {
    const size_t smem = 17 * CHS * sizeof(float);
    CK(cudaFuncSetAttribute(mix_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int nw : {8, 12, 16}) {
        const int iters = 20000;
        mix_kernel<V><<<sms, nw * 32, smem>>>(out, 100);
        CK(cudaEventRecord(e0));
        mix_kernel<V><<<sms, nw * 32, smem>>>(out, iters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        int khz; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
        const double clk = ms * 1e-3 * khz * 1e3;
        printf("%-28s %2d warps: %7.2f clk per warp-pair per SM   (%.3f ms)\n", name, nw, clk / ((double)iters * pairs_per_iter * nw), ms);
    }
}

int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float* out; CK(cudaMalloc(&out, sms * 512 * sizeof(float)));
    printf("bounds per warp-pair: FMA pipe 24 FFMA2 x 2 clk / 4 = 12 clk (+ scalar tail), LDS 16 wavefronts = 16 clk, issue ~57 / 4 = 14 clk\n");
    run<0>("0 full pair (FFMA2)", sms, out, 1);
    run<1>("1 1/|x| given", sms, out, 1);
    run<2>("2 full pair (scalar FFMA)", sms, out, 1);
    run<3>("3 LDS.32 x16 only", sms, out, 1);
    run<4>("4 FFMA2 x24 only", sms, out, 1);
    run<5>("5 two pairs interleaved", sms, out, 2);
    run<6>("6 lane-packed two neighbours", sms, out, 2);
    run<7>("7 ss packed, dot scalar", sms, out, 1);
    run<8>("8 ss packed, dot+G scalar", sms, out, 1);
    run<9>("9 ss+dot packed, G scalar", sms, out, 1);
    run<10>("10 FFMA2 (x,x,acc) x24", sms, out, 1);
    run<11>("11 FFMA2 (s.F32,x,acc) x24", sms, out, 1);
    run<12>("12 FFMA x48 (3 distinct)", sms, out, 1);
    CK(cudaFree(out));
    return 0;
}
