#include <atomic>
// extern "C" entry points for the affinity ops (declared in include/pea_b200.h): argument validation,
// workspace carving, kernel selection.  No allocation, no global state, everything on the caller's stream.
#include "pea_common.cuh"

namespace pea {

static std::atomic<int> g_knobs[PEA_KNOB_COUNT];
int debug_knob(int which) { return (which >= 0 && which < PEA_KNOB_COUNT) ? g_knobs[which].load(std::memory_order_relaxed) : 0; }

// pea_affs_generic.cu
int generic_partial_blocks(int64_t total);
int launch_rinv(const float* emb, int C, const Dims& d, float* rinv, cudaStream_t st);
int launch_fwd_generic(const float* emb, const float* par, const float* rinvE, const float* rinvU, int C, const Dims& d,
                       const OffsetTable& tab, int wrap, int z_skip, int64_t affs_plane, float* affs, int relu, cudaStream_t st);
int launch_fused_generic(const float* emb, const float* par, const float* rinvE, const float* rinvU, int C,
                         const Dims& d, const OffsetTable& tab, int wrap, int cross, const float* target,
                         const float* weight, const void* mask, int mask_dtype, const float* upstream, float* affs,
                         float* partials, float* grad_emb, float* grad_par, cudaStream_t st);

// pea_affs3d_tiled.cu
bool tiled3d_norm5_eligible(int C, int D, int H, int W, const OffsetTable& tab, int border, int mask_dtype, bool cross,
                            const float* emb, const float* target, const float* weight);
int tiled3d_norm5_blocks(int B, int H, int W);
int fused3d_norm5_blocks(int B, int H, int W);
int cross3d_norm5_blocks(int B, int H, int W);
bool tiled3d_fwd_eligible(int C, int W, const OffsetTable& tab, int border, bool cross, const float* emb,
                          int64_t emb_plane);
int launch_fwd3d_norm5(const float* emb, const Dims& d, int z_skip, int64_t emb_plane, int64_t affs_plane, float* affs,
                       const float* halo, int halo_n, int64_t halo_plane, int relu, cudaStream_t st,
                       const StagedHalo* stage = nullptr);
int launch_fused3d_norm5(const float* emb, const Dims& d, const OffsetTable& tab, const float* target,
                         const float* weight, float* affs, float* partials, float* grad, cudaStream_t st);
int launch_fused3d_norm5_labels(const float* emb, const int32_t* seg, const float* wtab, const Dims& d, const OffsetTable& tab,
                                float* affs, float* partials, float* grad, cudaStream_t st);

// pea_affs2d_tiled.cu
bool tiled2d_eligible(int C, int D, int H, int W, const OffsetTable& tab, int wrap, bool need_grad_partner);
int tiled2d_blocks(int B, int H, int W, const OffsetTable& tab, int wrap);
int launch_tiled2d(const float* emb, const float* partner, const Dims& d, const OffsetTable& tab, int wrap,
                   const float* target, const float* weight, const void* mask, int mask_dtype, float* affs,
                   float* partials, float* grad, cudaStream_t st);

// pea_affs2d_axis.cu
bool axis2d_eligible(int C, int D, int H, int W, const OffsetTable& tab, int wrap, bool need_grad_partner, const float* emb,
                     const float* partner);
int axis2d_blocks(int B, int H, int W, const OffsetTable& tab, int wrap);
bool canon2d_eligible(int C, int D, int H, int W, const OffsetTable& tab, const float* emb, const float* partner,
                      const float* target, const float* weight);
int canon2d_blocks(int B, int H, int W);
bool canon2d_mask_ok(const void* mask, int mask_dtype);
int launch_canon2d(const float* emb, const float* partner, const Dims& d, const OffsetTable& tab, int wrap, const float* target,
                   const float* weight, const void* mask, int mask_dtype, float* affs, float* partials, float* grad,
                   cudaStream_t st);
// pea_affs_multi.cu
int launch_multi_upstream(const float* emb, const float* partner, int B, int C, int D, int H, int W, const int32_t* offsets, int K,
                          int border, const float* grad_affs, float* grad_emb, cudaStream_t st);
// pea_affs2d_fwd.cu
bool fwd2d_canon_eligible(int C, int D, int H, int W, const OffsetTable& tab, const float* emb);
int launch_fwd2d_canon(const float* emb, const Dims& d, const OffsetTable& tab, int wrap, int relu, float* affs, cudaStream_t st);
int launch_canon2d_labels(const float* emb, const int32_t* labels, const float* wtab, int ignore, const Dims& d,
                          const OffsetTable& tab, int wrap, float* affs, float* partials, float* grad, cudaStream_t st);
int launch_axis2d_labels(const float* emb, const int32_t* labels, const float* wtab, int ignore, const Dims& d,
                         const OffsetTable& tab, int wrap, float* affs, float* partials, float* grad, cudaStream_t st);
int launch_axis2d(const float* emb, const float* partner, const Dims& d, const OffsetTable& tab, int wrap,
                  const float* target, const float* weight, const void* mask, int mask_dtype, float* affs, float* partials,
                  float* grad, cudaStream_t st);

int launch_cross3d_norm5(const float* emb, const float* partner, const Dims& d, const OffsetTable& tab,
                         const float* target, const float* weight, float* affs, float* partials, float* grad,
                         int accumulate, cudaStream_t st);

#ifndef PEA_PDL
#define PEA_PDL 1  // 0: plain stream-ordered launches of the two tail kernels (A/B)
#endif
// Programmatic dependent launch (sm_90+): the small tail kernels (loss finalize, gradient scale) are queued behind the
// preceding kernel with the "programmatic stream serialization" attribute, so their launch latency overlaps the tail of
// that kernel; they block in griddepcontrol.wait until it has completed and its writes are visible.
__device__ __forceinline__ void pdl_wait() {
#if PEA_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_tail(void (*kern)(KArgs...), int grid, int block, cudaStream_t st, Args... args) {
#if PEA_PDL
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
#else
    kern<<<grid, block, 0, st>>>(static_cast<KArgs>(args)...);
    return cudaGetLastError();
#endif
}

struct LossScale {
    float norm[PEA_MAX_OFFSETS];
    float coef[PEA_MAX_OFFSETS];
};

// partials[blk][k] -> loss_out[0] = sum_k coef_k * (S_k / norm_k), loss_out[1+k] = S_k / norm_k.
// One block: thread t owns channel t % K and every G-th row of partials (G = 1024 / K row groups; consecutive threads
// read consecutive addresses), accumulates in double, then one thread per channel adds the G group sums in a fixed
// order => run-to-run reproducible loss, and a few microseconds even for the ~10^4 CTAs of a 2-D batch.
__global__ void __launch_bounds__(1024) finalize_loss_kernel(const float* __restrict__ partials, int nblocks, int K,
                                                             LossScale sc, float* __restrict__ loss_out) {
    __shared__ double s_acc[1024];
    __shared__ float s_loss[PEA_MAX_OFFSETS];
    pdl_wait();  // launched as a programmatic dependent of the kernel that writes the partials
    const int G = 1024 / K;
    const int k = threadIdx.x % K, g = threadIdx.x / K;
    double acc = 0.0;
    if (g < G) {
        int b = g;
        for (; b + 3 * G < nblocks; b += 4 * G) {  // four independent loads in flight
            const float v0 = partials[(int64_t)b * K + k], v1 = partials[(int64_t)(b + G) * K + k];
            const float v2 = partials[(int64_t)(b + 2 * G) * K + k], v3 = partials[(int64_t)(b + 3 * G) * K + k];
            acc += (double)v0;
            acc += (double)v1;
            acc += (double)v2;
            acc += (double)v3;
        }
        for (; b < nblocks; b += G) acc += (double)partials[(int64_t)b * K + k];
        s_acc[g * K + k] = acc;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double tot = 0.0;
        for (int i = 0; i < G; ++i) tot += s_acc[i * K + threadIdx.x];
        s_loss[threadIdx.x] = (float)tot / sc.norm[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float total = 0.f;
        for (int kk = 0; kk < K; ++kk) {
            loss_out[1 + kk] = s_loss[kk];
            total += s_loss[kk] * sc.coef[kk];
        }
        loss_out[0] = total;
    }
}

__global__ void __launch_bounds__(256) scale_inplace_kernel(float* __restrict__ buf, int64_t n,
                                                            const float* __restrict__ scale) {
    pdl_wait();
    const float s = __ldg(scale);
    if (s == 1.0f) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(buf) & 15) == 0) {
        float4* b4 = reinterpret_cast<float4*>(buf);
        for (int64_t i = tid; i < n / 4; i += stride) {
            float4 v = b4[i];
            v.x *= s; v.y *= s; v.z *= s; v.w *= s;
            b4[i] = v;
        }
        for (int64_t i = (n / 4) * 4 + tid; i < n; i += stride) buf[i] *= s;
    } else {
        for (int64_t i = tid; i < n; i += stride) buf[i] *= s;
    }
}

struct Workspace {
    float* rinvE;
    float* rinvU;
    float* partials;
    size_t bytes;
};

static bool carve(void* ws, int B, int64_t N, int K, int cross, Workspace& out) {
    const size_t plane = align_up((size_t)B * (size_t)N * sizeof(float), 256);
    const size_t parts = align_up((size_t)generic_partial_blocks((int64_t)B * N) * (size_t)K * sizeof(float), 256);
    char* p = static_cast<char*>(ws);
    out.rinvE = reinterpret_cast<float*>(p);
    out.rinvU = cross ? reinterpret_cast<float*>(p + plane) : out.rinvE;
    out.partials = reinterpret_cast<float*>(p + plane * (cross ? 2 : 1));
    out.bytes = plane * (cross ? 2 : 1) + parts;
    return true;
}

// the tiled kernels write `blocks` rows of K loss partials at the head of the workspace
static bool partials_fit(int blocks, int K, const Workspace& ws) {
    return blocks > 0 && (size_t)blocks * (size_t)K * sizeof(float) <= ws.bytes;
}

static int check_volume(int B, int C, int D, int H, int W, int K, Dims& d) {
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0) return PEA_E_SHAPE;
    if (C != 4 && C != 8 && C != 16 && C != 32) return PEA_E_CHANNELS;
    if (K < 1 || K > PEA_MAX_OFFSETS) return PEA_E_OFFSETS;
    d.B = B; d.D = D; d.H = H; d.W = W;
    d.N = (int64_t)D * H * W;
    if (d.N > (int64_t)1 << 40) return PEA_E_SHAPE;
    return PEA_OK;
}

static void fill_offsets(const int32_t* offsets, int K, OffsetTable& tab) {
    tab.K = K;
    for (int k = 0; k < K; ++k) {
        tab.dz[k] = offsets[3 * k];
        tab.dy[k] = offsets[3 * k + 1];
        tab.dx[k] = offsets[3 * k + 2];
        tab.gcoef[k] = 0.f;
    }
}

}  // namespace pea

using namespace pea;

extern "C" int pea_abi_version(void) { return PEA_ABI_VERSION; }

extern "C" int pea_debug_configure(int knob, int value) {
    if (knob < 0 || knob >= PEA_KNOB_COUNT) return PEA_E_MODE;
    g_knobs[knob].store(value, std::memory_order_relaxed);
    return PEA_OK;
}

extern "C" const char* pea_strerror(int code) {
    if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
    switch (code) {
        case PEA_OK: return "ok";
        case PEA_E_NULL: return "a required pointer is NULL";
        case PEA_E_SHAPE: return "non-positive or overflowing extent";
        case PEA_E_CHANNELS: return "embedding channel count must be 4, 8, 16 or 32";
        case PEA_E_OFFSETS: return "number of offsets must be in [1, 64]";
        case PEA_E_WORKSPACE: return "workspace smaller than pea_workspace_bytes()";
        case PEA_E_DTYPE: return "unknown mask / label dtype code";
        case PEA_E_ALIGN: return "pointer not aligned for its element type";
        case PEA_E_MODE: return "unknown border / variant code";
        default: return "unknown pea error";
    }
}

extern "C" size_t pea_workspace_bytes(int B, int C, int D, int H, int W, int K, int cross) {
    Dims d;
    if (check_volume(B, C, D, H, W, K, d) != PEA_OK) return 0;
    Workspace w;
    carve(nullptr, B, d.N, K, cross, w);
    return w.bytes;
}

extern "C" int pea_affs_fwd(const float* emb, const float* partner, int B, int C, int D, int H, int W,
                            const int32_t* offsets, int K, int border, int impl, int z_skip, int64_t emb_plane,
                            int64_t affs_plane, const float* halo, int halo_slices, int64_t halo_plane, float* affs,
                            void* workspace, size_t ws_bytes, void* stream) {
    if (!emb || !offsets || !affs || !workspace) return PEA_E_NULL;
    const int relu = (impl & PEA_EPI_RELU) ? 1 : 0;
    impl &= ~PEA_EPI_RELU;
    if (impl != PEA_IMPL_AUTO && impl != PEA_IMPL_GENERIC && impl != PEA_IMPL_TILED) return PEA_E_MODE;
    Dims d;
    int rc = check_volume(B, C, D, H, W, K, d);
    if (rc) return rc;
    if (border != PEA_BORDER_WRAP && border != PEA_BORDER_ZERO) return PEA_E_MODE;
    if (z_skip < 0 || z_skip >= D) return PEA_E_SHAPE;
    if (halo && (z_skip != 0 || halo_slices < 1 || halo_slices > 4 || halo_plane < (int64_t)halo_slices * H * W))
        return PEA_E_SHAPE;
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return PEA_E_ALIGN;
    const int64_t HW = (int64_t)H * W;
    if (emb_plane == 0) emb_plane = d.N;
    if (affs_plane == 0) affs_plane = (int64_t)(D - z_skip) * HW;
    if (emb_plane < d.N || affs_plane < (int64_t)(D - z_skip) * HW) return PEA_E_SHAPE;
    const int cross = partner != nullptr;
    Workspace ws;
    carve(workspace, B, d.N, K, cross, ws);
    if (ws_bytes < ws.bytes) return PEA_E_WORKSPACE;
    OffsetTable tab;
    fill_offsets(offsets, K, tab);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool contiguous = emb_plane == d.N;
    const bool plain = contiguous && z_skip == 0 && affs_plane == d.N;
    const bool tiled2_ok = plain && !relu && tiled2d_eligible(C, D, H, W, tab, border == PEA_BORDER_WRAP, false);
    const bool tiled3_ok = tiled3d_fwd_eligible(C, W, tab, border, cross != 0, emb, emb_plane);
    if (impl == PEA_IMPL_TILED && !tiled2_ok && !tiled3_ok) return PEA_E_MODE;
    if (halo && (!tiled3_ok || impl == PEA_IMPL_GENERIC)) return PEA_E_MODE;  // only the tiled kernel reads an external halo
    // 2-D, one view, axis-aligned neighbor-4 tables (every shipped 2-D config): unrolled forward kernel, relu in the store.
    // PEA_KNOB_NO_CANON (tests) sends these calls to the older kernels below.
    if (impl != PEA_IMPL_GENERIC && plain && !cross && D == 1 && debug_knob(PEA_KNOB_NO_CANON) == 0 &&
        fwd2d_canon_eligible(C, D, H, W, tab, emb) && !(reinterpret_cast<uintptr_t>(affs) & 3))
        return launch_fwd2d_canon(emb, d, tab, border == PEA_BORDER_WRAP, relu, affs, st);
    if (impl != PEA_IMPL_GENERIC) {
        if (tiled3_ok)
            return launch_fwd3d_norm5(emb, d, z_skip, emb_plane, affs_plane, affs, halo, halo ? halo_slices : 0, halo_plane, relu,
                                      st);
        if (tiled2_ok)
            return launch_tiled2d(emb, partner, d, tab, border == PEA_BORDER_WRAP, nullptr, nullptr, nullptr,
                                  PEA_MASK_NONE, affs, nullptr, nullptr, st);
    }
    if (!contiguous) return PEA_E_MODE;  // the table-driven kernel reads contiguous volumes only
    if ((rc = launch_rinv(emb, C, d, ws.rinvE, st))) return rc;
    if (cross && (rc = launch_rinv(partner, C, d, ws.rinvU, st))) return rc;
    return launch_fwd_generic(emb, cross ? partner : emb, ws.rinvE, ws.rinvU, C, d, tab, border == PEA_BORDER_WRAP,
                              z_skip, affs_plane, affs, relu, st);
}

// affs[:, 0, :s] = affs[:, 0, s:2s] (z), affs[:, 1, :, :s] = affs[:, 1, :, s:2s] (y), affs[:, 2, :, :, :s] = ... (x)
__global__ void __launch_bounds__(256) replicate_first_kernel(float* __restrict__ affs, int B, int K, int D, int H, int W,
                                                              int64_t plane, int s) {
    const int64_t HW = (int64_t)H * W;
    const int64_t nz = D >= 2 * s ? (int64_t)s * HW : 0, ny = H >= 2 * s ? (int64_t)D * s * W : 0,
                  nx = W >= 2 * s ? (int64_t)D * H * s : 0;
    const int64_t per = nz + ny + nx, total = per * B;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / per;
        int64_t r = i - b * per;
        float* base = affs + b * K * plane;
        if (r < nz) {            // channel 0: slices [0, s) <- [s, 2s)
            base[r] = base[r + (int64_t)s * HW];
        } else if ((r -= nz) < ny) {  // channel 1: rows [0, s) <- [s, 2s)
            const int64_t z = r / ((int64_t)s * W), q = r - z * s * W;
            float* p = base + plane + z * HW + q;
            p[0] = p[(int64_t)s * W];
        } else {                 // channel 2: columns [0, s) <- [s, 2s)
            r -= ny;
            const int64_t zy = r / s, x = r - zy * s;
            float* p = base + 2 * plane + zy * W + x;
            p[0] = p[s];
        }
    }
}

extern "C" int pea_replicate_first(float* affs, int B, int K, int D, int H, int W, int64_t plane, int shift, void* stream) {
    if (!affs) return PEA_E_NULL;
    if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || shift <= 0) return PEA_E_SHAPE;
    if (K < 3 || K > PEA_MAX_OFFSETS) return PEA_E_OFFSETS;
    const int64_t N = (int64_t)D * H * W;
    if (plane == 0) plane = N;
    if (plane < N) return PEA_E_SHAPE;
    const int64_t HW = (int64_t)H * W;
    const int64_t per = (D >= 2 * shift ? shift * HW : 0) + (H >= 2 * shift ? (int64_t)D * shift * W : 0) +
                        (W >= 2 * shift ? (int64_t)D * H * shift : 0);
    if (per == 0) return PEA_OK;
    int64_t g = (per * B + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    replicate_first_kernel<<<(unsigned)g, 256, 0, static_cast<cudaStream_t>(stream)>>>(affs, B, K, D, H, W, plane, shift);
    return (int)cudaGetLastError();
}

extern "C" int pea_affs_loss_fwd_bwd(const float* emb, const float* partner, int B, int C, int D, int H, int W,
                                     const int32_t* offsets, int K, int border, int impl, const float* target,
                                     const float* weight, const void* mask, int mask_dtype, const float* loss_norm,
                                     const float* loss_coef, float* affs, float* loss_out, float* grad_emb,
                                     float* grad_partner, void* workspace, size_t ws_bytes, void* stream) {
    if (!emb || !offsets || !target || !weight || !loss_norm || !loss_coef || !loss_out || !workspace) return PEA_E_NULL;
    const int accumulate = (impl & PEA_EPI_GRAD_ACCUMULATE) ? 1 : 0;
    impl &= ~PEA_EPI_GRAD_ACCUMULATE;
    if (impl != PEA_IMPL_AUTO && impl != PEA_IMPL_GENERIC && impl != PEA_IMPL_TILED) return PEA_E_MODE;
    Dims d;
    int rc = check_volume(B, C, D, H, W, K, d);
    if (rc) return rc;
    if (border != PEA_BORDER_WRAP && border != PEA_BORDER_ZERO) return PEA_E_MODE;
    if (mask_dtype != PEA_MASK_NONE && mask_dtype != PEA_MASK_U8 && mask_dtype != PEA_MASK_F32) return PEA_E_DTYPE;
    if (mask_dtype != PEA_MASK_NONE && !mask) return PEA_E_NULL;
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return PEA_E_ALIGN;
    const int cross = partner != nullptr;
    if (!cross && grad_partner) return PEA_E_MODE;
    Workspace ws;
    carve(workspace, B, d.N, K, cross, ws);
    if (ws_bytes < ws.bytes) return PEA_E_WORKSPACE;
    OffsetTable tab;
    fill_offsets(offsets, K, tab);
    LossScale sc;
    for (int k = 0; k < K; ++k) {
        sc.norm[k] = loss_norm[k];
        sc.coef[k] = loss_coef[k];
        tab.gcoef[k] = (float)(2.0 * (double)loss_coef[k] / (double)loss_norm[k]);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // ---- tiled single-kernel path for the offset sets it specialises (same results, see tests)
    const bool tiled3_ok = tiled3d_norm5_eligible(C, D, H, W, tab, border, mask_dtype, cross != 0, emb, target, weight) &&
                           grad_partner == nullptr;
    const bool tiled2_ok = tiled2d_eligible(C, D, H, W, tab, border == PEA_BORDER_WRAP, grad_partner != nullptr);
    if (impl == PEA_IMPL_TILED && !tiled3_ok && !tiled2_ok) return PEA_E_MODE;
    // only the tiled two-view 3-D kernel adds into an existing gradient buffer
    if (accumulate && !(tiled3_ok && cross && impl != PEA_IMPL_GENERIC && grad_emb)) return PEA_E_MODE;
    // canonical multi_offset(shifts <= 12, neighbor=4) table (one view, or two views with the second one constant):
    // fully unrolled persistent kernel.
    // PEA_KNOB_NO_CANON: test knob (pea_debug_configure) that sends these calls to the table-driven axis kernel.
    if (tiled2_ok && !tiled3_ok && impl != PEA_IMPL_GENERIC && grad_partner == nullptr && debug_knob(PEA_KNOB_NO_CANON) == 0 &&
        canon2d_eligible(C, D, H, W, tab, emb, partner ? partner : emb, target, weight) && canon2d_mask_ok(mask, mask_dtype)) {
        float* parts = ws.rinvE;
        const int nblk = canon2d_blocks(B, H, W);  // one evaluation: the launcher and the finalize kernel agree by construction
        if (!partials_fit(nblk, K, ws)) return PEA_E_WORKSPACE;
        if ((rc = launch_canon2d(emb, partner, d, tab, border == PEA_BORDER_WRAP, target, weight, mask, mask_dtype, affs, parts,
                                 grad_emb, st)))
            return rc;
        return (int)launch_tail(finalize_loss_kernel, 1, 1024, st, parts, nblk, K, sc, loss_out);
    }
    if (tiled2_ok && !tiled3_ok && impl != PEA_IMPL_GENERIC &&
        B <= 65535 &&
        axis2d_eligible(C, D, H, W, tab, border == PEA_BORDER_WRAP, grad_partner != nullptr, emb, partner ? partner : emb) &&
        !((reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(weight)) & 3)) {
        // axis-aligned tables (every shipped 2-D config): TMA-staged kernel
        float* parts = ws.rinvE;
        const int nblk = axis2d_blocks(B, H, W, tab, border == PEA_BORDER_WRAP);
        if (!partials_fit(nblk, K, ws)) return PEA_E_WORKSPACE;
        if ((rc = launch_axis2d(emb, partner, d, tab, border == PEA_BORDER_WRAP, target, weight, mask, mask_dtype, affs, parts,
                                grad_emb, st)))
            return rc;
        return (int)launch_tail(finalize_loss_kernel, 1, 1024, st, parts, nblk, K, sc, loss_out);
    }
    if (tiled2_ok && !tiled3_ok && impl != PEA_IMPL_GENERIC) {
        float* parts = ws.rinvE;
        const int nblk = tiled2d_blocks(B, H, W, tab, border == PEA_BORDER_WRAP);
        if (!partials_fit(nblk, K, ws)) return PEA_E_WORKSPACE;
        if ((rc = launch_tiled2d(emb, partner, d, tab, border == PEA_BORDER_WRAP, target, weight, mask, mask_dtype, affs,
                                 parts, grad_emb, st)))
            return rc;
        return (int)launch_tail(finalize_loss_kernel, 1, 1024, st, parts, nblk, K, sc, loss_out);
    }
    const bool tiled_ok = tiled3_ok;
    if (tiled_ok && impl != PEA_IMPL_GENERIC) {
        // the per-voxel plane at the head of the workspace doubles as the (much smaller) partials array
        float* parts = ws.rinvE;
        const int nblk3 = cross ? cross3d_norm5_blocks(B, H, W) : fused3d_norm5_blocks(B, H, W);
        if (!partials_fit(nblk3, K, ws)) return PEA_E_WORKSPACE;
        rc = cross ? launch_cross3d_norm5(emb, partner, d, tab, target, weight, affs, parts, grad_emb, accumulate, st)
                   : launch_fused3d_norm5(emb, d, tab, target, weight, affs, parts, grad_emb, st);
        if (rc) return rc;
        return (int)launch_tail(finalize_loss_kernel, 1, 1024, st, parts, nblk3, K, sc, loss_out);
    }
    // Offset sets no specialised kernel covers (embedding_loss_norm1 / ema_embedding_loss_norm1 at full resolution:
    // TRAIN.embedding_mode == 1, scripts_ac3ac4/main.py:219-221; odd tables): the work-list kernel with ONE problem -- no
    // 1 / |x| plane pass, neighbours' norms inline, loss finalized in the kernel.  B = 4, 18 x 160 x 160, K = 3:
    // 415 -> 154 us.  The ticket word sits behind the partials in the caller's workspace and is zeroed in stream order.
    if (impl == PEA_IMPL_AUTO && K <= PEA_MULTI_MAX_OFFSETS && grad_partner == nullptr && !accumulate &&
        (int64_t)B * d.N < ((int64_t)1 << 31) - 64) {
        bool small = true;
        for (int k = 0; k < 3 * K; ++k) small = small && offsets[k] >= -127 && offsets[k] <= 127;
        pea_loss_problem pr;
        pr.emb = emb; pr.partner = partner; pr.target = target; pr.weight = weight; pr.mask = mask;
        pr.affs = affs; pr.grad_emb = grad_emb; pr.loss_out = loss_out;
        pr.offsets = offsets; pr.loss_norm = loss_norm; pr.loss_coef = loss_coef;
        pr.B = B; pr.D = D; pr.H = H; pr.W = W; pr.K = K; pr.border = border; pr.mask_dtype = mask_dtype;
        const size_t parts = align_up(pea_affs_loss_multi_workspace(&pr, 1), 256);
        if (small && parts + 16 <= ws.bytes) {
            uint32_t* ticket = reinterpret_cast<uint32_t*>(static_cast<char*>(workspace) + parts);
            cudaError_t e = cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st);
            if (e != cudaSuccess) return (int)e;
            return pea_affs_loss_multi(&pr, 1, C, workspace, parts, ticket, stream);
        }
    }
    if ((rc = launch_rinv(emb, C, d, ws.rinvE, st))) return rc;
    if (cross && (rc = launch_rinv(partner, C, d, ws.rinvU, st))) return rc;
    rc = launch_fused_generic(emb, cross ? partner : emb, ws.rinvE, ws.rinvU, C, d, tab, border == PEA_BORDER_WRAP, cross,
                              target, weight, mask, mask_dtype, nullptr, affs, ws.partials, grad_emb, grad_partner, st);
    if (rc) return rc;
    return (int)launch_tail(finalize_loss_kernel, 1, 1024, st, ws.partials, generic_partial_blocks((int64_t)B * d.N), K, sc, loss_out);
}

extern "C" int pea_affs_loss_labels_fwd_bwd(const float* emb, const int32_t* labels, int B, int C, int D, int H, int W,
                                            const int32_t* offsets, int K, int border, int semantics, int ignore,
                                            const float* wtab,
                                            const float* loss_norm, const float* loss_coef, float* affs, float* loss_out,
                                            float* grad_emb, void* workspace, size_t ws_bytes, void* stream) {
    if (!emb || !labels || !offsets || !wtab || !loss_norm || !loss_coef || !loss_out || !workspace) return PEA_E_NULL;
    Dims d;
    int rc = check_volume(B, C, D, H, W, K, d);
    if (rc) return rc;
    if (border != PEA_BORDER_WRAP && border != PEA_BORDER_ZERO) return PEA_E_MODE;
    if (semantics != 0 && semantics != 1) return PEA_E_MODE;
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return PEA_E_ALIGN;
    if ((reinterpret_cast<uintptr_t>(labels) | reinterpret_cast<uintptr_t>(wtab)) & 3) return PEA_E_ALIGN;
    Workspace ws;
    carve(workspace, B, d.N, K, 0, ws);
    if (ws_bytes < ws.bytes) return PEA_E_WORKSPACE;
    OffsetTable tab;
    fill_offsets(offsets, K, tab);
    LossScale sc;
    for (int k = 0; k < K; ++k) {
        sc.norm[k] = loss_norm[k];
        sc.coef[k] = loss_coef[k];
        tab.gcoef[k] = (float)(2.0 * (double)loss_coef[k] / (double)loss_norm[k]);
    }
    // the label plane is staged by TMA like the embedding: 16-byte aligned base
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* parts = ws.rinvE;
    if (semantics == 1) {  // 3-D: seg_to_aff targets of the 12-offset table, zero border
        const float* labf = reinterpret_cast<const float*>(labels);
        if (!tiled3d_norm5_eligible(C, D, H, W, tab, border, PEA_MASK_NONE, false, emb, labf, labf)) return PEA_E_MODE;
        const int nblk = fused3d_norm5_blocks(B, H, W);
        if (!partials_fit(nblk, K, ws)) return PEA_E_WORKSPACE;
        if ((rc = launch_fused3d_norm5_labels(emb, labels, wtab, d, tab, affs, parts, grad_emb, st))) return rc;
        return (int)launch_tail(finalize_loss_kernel, 1, 1024, st, parts, nblk, K, sc, loss_out);
    }
    if (D != 1) return PEA_E_MODE;
    if (!canon2d_eligible(C, 1, H, W, tab, emb, emb, emb, reinterpret_cast<const float*>(labels)) ||
        debug_knob(PEA_KNOB_NO_CANON) != 0) {
        // any other axis-aligned table (CVPPP's shift 27): the table-driven axis kernel's label-driven flavour
        if (B > 65535 || !axis2d_eligible(C, 1, H, W, tab, border == PEA_BORDER_WRAP, false, emb, emb)) return PEA_E_MODE;
        const int nb = axis2d_blocks(B, H, W, tab, border == PEA_BORDER_WRAP);
        if (!partials_fit(nb, K, ws)) return PEA_E_WORKSPACE;
        if ((rc = launch_axis2d_labels(emb, labels, wtab, ignore, d, tab, border == PEA_BORDER_WRAP, affs, parts, grad_emb, st)))
            return rc;
        return (int)launch_tail(finalize_loss_kernel, 1, 1024, st, parts, nb, K, sc, loss_out);
    }
    const int nblk = canon2d_blocks(B, H, W);
    if (!partials_fit(nblk, K, ws)) return PEA_E_WORKSPACE;
    if ((rc = launch_canon2d_labels(emb, labels, wtab, ignore, d, tab, border == PEA_BORDER_WRAP, affs, parts, grad_emb, st)))
        return rc;
    return (int)launch_tail(finalize_loss_kernel, 1, 1024, st, parts, nblk, K, sc, loss_out);
}

extern "C" int pea_affs_bwd(const float* emb, const float* partner, int B, int C, int D, int H, int W,
                            const int32_t* offsets, int K, int border, const float* grad_affs, float* grad_emb,
                            float* grad_partner, void* workspace, size_t ws_bytes, void* stream) {
    if (!emb || !offsets || !grad_affs || !workspace) return PEA_E_NULL;
    Dims d;
    int rc = check_volume(B, C, D, H, W, K, d);
    if (rc) return rc;
    if (border != PEA_BORDER_WRAP && border != PEA_BORDER_ZERO) return PEA_E_MODE;
    if (reinterpret_cast<uintptr_t>(workspace) % 16) return PEA_E_ALIGN;
    const int cross = partner != nullptr;
    if (!cross && grad_partner) return PEA_E_MODE;
    Workspace ws;
    carve(workspace, B, d.N, K, cross, ws);
    if (ws_bytes < ws.bytes) return PEA_E_WORKSPACE;
    OffsetTable tab;
    fill_offsets(offsets, K, tab);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // gradient for the anchor view only, at most 12 offsets: the work-list kernel in upstream mode (no 1 / |x| plane pass;
    // cfg2 1.31 -> see DESIGN.md).  PEA_KNOB_NO_CANON (tests) keeps the table-driven kernel.
    if (grad_emb && !grad_partner && debug_knob(PEA_KNOB_NO_CANON) == 0) {
        rc = launch_multi_upstream(emb, partner, B, C, D, H, W, offsets, K, border, grad_affs, grad_emb, st);
        if (rc != PEA_E_MODE) return rc;
    }
    if ((rc = launch_rinv(emb, C, d, ws.rinvE, st))) return rc;
    if (cross && (rc = launch_rinv(partner, C, d, ws.rinvU, st))) return rc;
    return launch_fused_generic(emb, cross ? partner : emb, ws.rinvE, ws.rinvU, C, d, tab, border == PEA_BORDER_WRAP,
                                cross, nullptr, nullptr, nullptr, PEA_MASK_NONE, grad_affs, nullptr, nullptr, grad_emb,
                                grad_partner, st);
}

extern "C" int pea_scale_inplace(float* buf, int64_t n, const float* scale, void* stream) {
    if (!buf || !scale) return PEA_E_NULL;
    if (n <= 0) return PEA_E_SHAPE;
    int64_t g = (n / 4 + 255) / 256;
    const int grid = (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
    return (int)launch_tail(scale_inplace_kernel, grid, 256, static_cast<cudaStream_t>(stream), buf, n, scale);
}
