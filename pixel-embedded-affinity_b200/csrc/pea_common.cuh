// Shared device/host helpers for the pea_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "pea_b200.h"

namespace pea {

constexpr float kNormEps = 1e-12f;  // F.normalize eps, scripts_cvppp/loss/loss_embedding_mse.py:20

// Offset table travels BY VALUE in the kernel parameter space: no __constant__ globals, so
// concurrent callers on different streams never share mutable state.
struct OffsetTable {
    int K;
    int dz[PEA_MAX_OFFSETS];
    int dy[PEA_MAX_OFFSETS];
    int dx[PEA_MAX_OFFSETS];
    float gcoef[PEA_MAX_OFFSETS];  // 2 * loss_coef[k] / loss_norm[k]
};

struct Dims {
    int B, D, H, W;
    int64_t N;  // D*H*W, voxels per (sample, channel) plane
};

// Local landing buffer of a z-shard's halo that a copy engine fills band by band while the forward kernel runs
// (pea_zshard_stage_halo / pea_zshard_affs_fwd, include/pea_b200.h).
struct StagedHalo {
    const float* data;       // (B, C, halo_slices, H, W), channel-plane stride `plane`
    int64_t plane;
    const uint32_t* flags;   // flag j >= epoch (mod 2^32): rows [row0 + j * band_rows, ...) have landed
    uint32_t epoch;
    int row0, band_rows;
};

__host__ __device__ __forceinline__ int pmod(int a, int n) {
    int r = a % n;
    return r < 0 ? r + n : r;
}

// neighbour coordinates under the border rule; returns false when the pair is dropped
__device__ __forceinline__ bool neighbour(int z, int y, int x, int dz, int dy, int dx, const Dims& d, bool wrap,
                                          int64_t& q) {
    int qz = z + dz, qy = y + dy, qx = x + dx;
    if (wrap) {
        qz = pmod(qz, d.D);
        qy = pmod(qy, d.H);
        qx = pmod(qx, d.W);
    } else if ((unsigned)qz >= (unsigned)d.D || (unsigned)qy >= (unsigned)d.H || (unsigned)qx >= (unsigned)d.W) {
        return false;
    }
    q = ((int64_t)qz * d.H + qy) * d.W + qx;
    return true;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float load_mask(const void* mask, int mask_dtype, int64_t i) {
    if (mask_dtype == PEA_MASK_U8) return (float)__ldg(reinterpret_cast<const uint8_t*>(mask) + i);
    if (mask_dtype == PEA_MASK_F32) return __ldg(reinterpret_cast<const float*>(mask) + i);
    return 1.0f;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Test-only dispatch knobs (pea_debug_configure, include/pea_b200.h): one relaxed atomic load per call, 0 in production.
int debug_knob(int which);

}  // namespace pea
