// convert_consistency_flip on the device in ONE pass: undo the per-sample flip / transpose
// augmentation of the second view's embedding before the two-view loss,
//   scripts_cvppp/data/data_consistency.py:19-45          2-D, rule = (x reflection, y reflection, xy transpose)
//   scripts_ac3ac4/utils/consistency_aug.py:58-77,217-228  3-D, rule = (z reflection, x reflection, y reflection, xy transpose)
// The reference does detach().clone(), a device -> host copy of `rules` (a synchronisation), up to four flip / permute
// passes per sample and a stack: about five passes over the embedding.  Here the rules stay on the device and every
// element moves once: out[z, y, x] = in[z', y', x'] (or in[z', x', y'] under the transpose) with the primed coordinates
// mirrored per rule; 32 x 32 tiles go through shared memory so that both the read and the write are coalesced even
// under the transpose.
#include "pea_common.cuh"

namespace pea {

// rules: int32 [B][4] as (rz, rx, ry, t); 2-D callers pass rz = 0, D = 1
__global__ void __launch_bounds__(256) flip_kernel(const float* __restrict__ in, const int32_t* __restrict__ rules, int C, int D,
                                                   int H, int W, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int tiles_x = (W + 31) / 32, tiles_y = (H + 31) / 32;
    int bid = blockIdx.x;
    const int tix = bid % tiles_x;
    bid /= tiles_x;
    const int tiy = bid % tiles_y;
    bid /= tiles_y;
    const int z = bid % D;
    bid /= D;
    const int c = bid % C, b = bid / C;
    const int rz = rules[b * 4 + 0] != 0, rx = rules[b * 4 + 1] != 0, ry = rules[b * 4 + 2] != 0, t = rules[b * 4 + 3] != 0;
    const int x0 = tix * 32, y0 = tiy * 32;
    const int zi = rz ? D - 1 - z : z;
    const float* src = in + (((int64_t)b * C + c) * D + zi) * (int64_t)H * W;
    float* dst = out + (((int64_t)b * C + c) * D + z) * (int64_t)H * W;
    // input tile: rows R0 .. R0+31, columns C0 .. C0+31 (the image of the output tile; H == W under the transpose)
    const int ylo = ry ? H - 1 - (y0 + 31) : y0, xlo = rx ? W - 1 - (x0 + 31) : x0;  // mirrored ranges of y', x'
    const int R0 = t ? xlo : ylo, C0 = t ? ylo : xlo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ri = R0 + ty + 8 * j, ci = C0 + tx;
        if (ri >= 0 && ri < H && ci >= 0 && ci < W) {
            const int yp = t ? ci : ri, xp = t ? ri : ci;  // (y', x') of this input element
            const int y = ry ? H - 1 - yp : yp, x = rx ? W - 1 - xp : xp;
            tile[y - y0][x - x0] = src[(int64_t)ri * W + ci];
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int y = y0 + ty + 8 * j, x = x0 + tx;
        if (y < H && x < W) dst[(int64_t)y * W + x] = tile[ty + 8 * j][tx];
    }
}

}  // namespace pea

using namespace pea;

extern "C" int pea_consistency_flip(const float* in, const int32_t* rules, int B, int C, int D, int H, int W, float* out,
                                    void* stream) {
    if (!in || !rules || !out) return PEA_E_NULL;
    if (B <= 0 || C <= 0 || D <= 0 || H <= 0 || W <= 0) return PEA_E_SHAPE;
    if (in == out) return PEA_E_MODE;  // not an in-place operation
    const int64_t blocks = (int64_t)B * C * D * ((H + 31) / 32) * ((W + 31) / 32);
    if (blocks > 0x7fffffff) return PEA_E_SHAPE;
    flip_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, rules, C, D, H, W, out);
    return (int)cudaGetLastError();
}
