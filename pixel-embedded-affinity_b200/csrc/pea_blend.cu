// Full-volume inference epilogue on the device: the Gaussian-weighted accumulation of
// overlapping patch predictions that the reference does on the host after a device -> host copy of every patch
//   scripts_ac3ac4/data/provider_valid.py:306-348   add_vol:      out[:, pos + patch] += affs_vol * weight_vol
//                                                                 weight_map[pos + patch] += weight_vol
//                                                    get_results: out / weight_map
// and the `1 - affs` that scripts_ac3ac4/inference.py:196 hands to the Mutex Watershed.  Arithmetic is NumPy's:
// product rounded to float32, then added (no FMA contraction); the final division is IEEE round-to-nearest.
#include "pea_common.cuh"

namespace pea {

__global__ void __launch_bounds__(256) blend_add_kernel(const float* __restrict__ patch, const float* __restrict__ w, int K,
                                                        int d, int h, int wd, float* __restrict__ acc,
                                                        float* __restrict__ wmap, int D, int H, int W, int z0, int y0,
                                                        int x0) {
    const int64_t n = (int64_t)d * h * wd, N = (int64_t)D * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % wd), y = (int)((i / wd) % h), z = (int)(i / ((int64_t)wd * h));
        const int64_t o = ((int64_t)(z0 + z) * H + (y0 + y)) * W + (x0 + x);
        const float wv = __ldg(w + i);
        wmap[o] = __fadd_rn(wmap[o], wv);
        for (int k = 0; k < K; ++k) acc[(int64_t)k * N + o] = __fadd_rn(acc[(int64_t)k * N + o], __fmul_rn(__ldg(patch + (int64_t)k * n + i), wv));
    }
}

__global__ void __launch_bounds__(256) blend_finish_kernel(const float* __restrict__ acc, const float* __restrict__ wmap, int K,
                                                           int D, int H, int W, int pz, int py, int px, int one_minus,
                                                           float* __restrict__ out) {
    const int d = D - 2 * pz, h = H - 2 * py, w = W - 2 * px;
    const int64_t n = (int64_t)d * h * w, N = (int64_t)D * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % w), y = (int)((i / w) % h), z = (int)(i / ((int64_t)w * h));
        const int64_t o = ((int64_t)(z + pz) * H + (y + py)) * W + (x + px);
        const float wv = __ldg(wmap + o);
        for (int k = 0; k < K; ++k) {
            const float v = __fdiv_rn(__ldg(acc + (int64_t)k * N + o), wv);
            out[(int64_t)k * n + i] = one_minus ? __fsub_rn(1.0f, v) : v;
        }
    }
}

}  // namespace pea

using namespace pea;

extern "C" int pea_blend_add(const float* patch, const float* weight, int K, int d, int h, int w, float* acc, float* wmap,
                             int D, int H, int W, int z0, int y0, int x0, void* stream) {
    if (!patch || !weight || !acc || !wmap) return PEA_E_NULL;
    if (K < 1 || K > PEA_MAX_OFFSETS) return PEA_E_OFFSETS;
    if (d <= 0 || h <= 0 || w <= 0 || D <= 0 || H <= 0 || W <= 0) return PEA_E_SHAPE;
    if (z0 < 0 || y0 < 0 || x0 < 0 || z0 + d > D || y0 + h > H || x0 + w > W) return PEA_E_SHAPE;
    const int64_t n = (int64_t)d * h * w;
    int64_t g = (n + 255) / 256;
    const int grid = (int)(g > 148 * 8 ? 148 * 8 : g);
    blend_add_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(patch, weight, K, d, h, w, acc, wmap, D, H, W, z0, y0, x0);
    return (int)cudaGetLastError();
}

extern "C" int pea_blend_finish(const float* acc, const float* wmap, int K, int D, int H, int W, int pad_z, int pad_y,
                                int pad_x, int one_minus, float* out, void* stream) {
    if (!acc || !wmap || !out) return PEA_E_NULL;
    if (K < 1 || K > PEA_MAX_OFFSETS) return PEA_E_OFFSETS;
    if (D <= 0 || H <= 0 || W <= 0 || pad_z < 0 || pad_y < 0 || pad_x < 0) return PEA_E_SHAPE;
    if (D - 2 * pad_z <= 0 || H - 2 * pad_y <= 0 || W - 2 * pad_x <= 0) return PEA_E_SHAPE;
    const int64_t n = (int64_t)(D - 2 * pad_z) * (H - 2 * pad_y) * (W - 2 * pad_x);
    int64_t g = (n + 255) / 256;
    const int grid = (int)(g > 148 * 8 ? 148 * 8 : g);
    blend_finish_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(acc, wmap, K, D, H, W, pad_z, pad_y, pad_x, one_minus, out);
    return (int)cudaGetLastError();
}
