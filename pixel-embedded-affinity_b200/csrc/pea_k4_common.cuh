// Device helpers shared by the fused 3-D kernels (pea_affs3d_tiled.cu, pea_affs3d_k4v2.cu): the 12-offset set's
// geometry constants, packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2), 1 / max(|x|, eps), the release / acquire
// arrival counter behind the barrier-free box refills, and Tensor Memory used as a second register file
// (tcgen05.st / tcgen05.ld, 32x32b shape: 16 or 32 columns of ONE lane <-> ONE thread).
#pragma once
#include "pea_common.cuh"
#include "pea_tma.cuh"

namespace pea {
namespace t3 {

constexpr int kC = 16;
constexpr int TX = 32;
constexpr int HY = 9;               // near y halo rows on each side of the tile in the y-column box
constexpr int FAR = 27;             // far shift served by the arm boxes / x-row box
constexpr int XPAD = 28;            // x halo columns on each side of the x-row box (27 rounded up to 16 B)
constexpr int XW = TX + 2 * XPAD;   // 88

// channel layout of the 12-offset set: k = 3*g + axis, shift of (group g, axis): z {1,2,3,4}, y/x {1,3,9,27}
__device__ __forceinline__ constexpr int shift_yx(int g) { return g == 0 ? 1 : (g == 1 ? 3 : (g == 2 ? 9 : 27)); }

// 1 / max(sqrt(ss), 1e-12) == min(rsqrt(ss), 1e12); one MUFU.RSQ (2 ulp), ss == 0 -> +inf -> 1e12
__device__ __forceinline__ float inv_norm(float ss) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(ss));
    return fminf(r, 1e12f);
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2): two channels per instruction, same rounding as fmaf
typedef unsigned long long f32x2;
constexpr int kC2 = kC / 2;
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
#ifdef PEA_HSUM_ADD2
    f32x2 t;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(a), "l"(b));
    float lo, hi;
    unpack2(t, lo, hi);
    return lo + hi;
#else
    float a0, a1, b0, b1;
    unpack2(a, a0, a1);
    unpack2(b, b0, b1);
    return (a0 + a1) + (b0 + b1);
#endif
}

// Arrive on a CTA-wide counter with release/acquire semantics; returns the value before the increment.
// The warp that observes the last arrival of a phase owns the refill of the buffer everyone has finished reading.
__device__ __forceinline__ uint32_t arrive_count(uint32_t* cnt) {
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(cnt)) : "memory");
    return old;
}

__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const f32x2 (&v)[kC2]) {
    float r[16];
#pragma unroll
    for (int i = 0; i < kC2; ++i) unpack2(v[i], r[2 * i], r[2 * i + 1]);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]), "f"(r[8]),
        "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15])
        : "memory");
}
// load + wait in one statement: the destination registers are valid when it returns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, f32x2 (&v)[kC2]) {
    float r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
          "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < kC2; ++i) v[i] = pack2(r[2 * i], r[2 * i + 1]);
}
// two 16-column loads (unit vector + gradient accumulator of one slot) behind a single wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t taddr, f32x2 (&a)[kC2], f32x2 (&b)[kC2]) {
    float r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
          "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15]), "=f"(r[16]),
          "=f"(r[17]), "=f"(r[18]), "=f"(r[19]), "=f"(r[20]), "=f"(r[21]), "=f"(r[22]), "=f"(r[23]), "=f"(r[24]),
          "=f"(r[25]), "=f"(r[26]), "=f"(r[27]), "=f"(r[28]), "=f"(r[29]), "=f"(r[30]), "=f"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < kC2; ++i) {
        a[i] = pack2(r[2 * i], r[2 * i + 1]);
        b[i] = pack2(r[16 + 2 * i], r[16 + 2 * i + 1]);
    }
}
// predicated read-only global load that the compiler may not sink towards its use (volatile): issued where it is
// written, i.e. half a slice before the value is needed
__device__ __forceinline__ float ldg_early(const float* p, bool pred) {
    float v;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@q ld.global.nc.f32 %0, [%1];\n\t}"
        : "=f"(v)
        : "l"(p), "r"((int)pred));
    return v;
}
// the same for memory another engine wrote while the kernel was running (staged z-shard halo): L2 only, no L1 / texture path
__device__ __forceinline__ float ldg_early_cg(const float* p, bool pred) {
    float v;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@q ld.global.cg.f32 %0, [%1];\n\t}"
        : "=f"(v)
        : "l"(p), "r"((int)pred));
    return v;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr uint32_t kTmemCols = 512;

}  // namespace t3
}  // namespace pea
