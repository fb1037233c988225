// z-sharded volumes (BASELINE config 4): the staged flavour of the halo.
//
// pea_affs_fwd reads the previous shard's last four slices IN PLACE from peer memory: one launch, no exchange step --
// but every one of those loads crosses NVLink from inside the compute kernel (268 MB against a 0.47 ms slab at
// 100 x 1024 x 1024 over 8 GPUs), and the kernel pays for it: 0.56 ms per slab instead of 0.47.  Here a COPY ENGINE
// pulls the same slices into a local landing buffer while the kernel runs, in bands of image rows that arrive in the
// order the kernel's tile rows are scheduled; a 32-bit flag per band (stream-ordered cuStreamWriteValue32 behind the
// band's copy) tells the kernel's CTAs that their rows have landed.  The first tile rows -- whose CTAs start before any
// band can have arrived -- keep reading peer memory directly.  Still one compute launch; no SM ever waits on NVLink.
#include "pea_common.cuh"

#include <cuda.h>

namespace pea {
int launch_fwd3d_norm5(const float* emb, const Dims& d, int z_skip, int64_t emb_plane, int64_t affs_plane, float* affs,
                       const float* halo, int halo_n, int64_t halo_plane, int relu, cudaStream_t st,
                       const StagedHalo* stage);
bool tiled3d_fwd_eligible(int C, int W, const OffsetTable& tab, int border, bool cross, const float* emb,
                          int64_t emb_plane);

using WriteValue32Fn = CUresult (*)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static WriteValue32Fn write_value32_fn() {
    static WriteValue32Fn fn = []() -> WriteValue32Fn {   // resolved once per process, immutable afterwards
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<WriteValue32Fn>(p);
    }();
    return fn;
}
}  // namespace pea

using namespace pea;

extern "C" int pea_zshard_stage_halo(const float* tail, int64_t tail_plane, float* stage, int B, int C, int halo_slices,
                                     int H, int W, int row0, int band_rows, uint32_t* flags, uint32_t epoch,
                                     void* stream) {
    if (!tail || !stage || !flags) return PEA_E_NULL;
    if (B < 1 || C < 1 || halo_slices < 1 || halo_slices > 4 || H < 1 || W < 1) return PEA_E_SHAPE;
    const int64_t HW = (int64_t)H * W;
    if (row0 < 0 || row0 > H || band_rows < 1 || tail_plane < halo_slices * HW || tail_plane % HW) return PEA_E_SHAPE;
    WriteValue32Fn wv = write_value32_fn();
    if (!wv) return PEA_E_MODE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // One 3-D copy per band: "row" = the band's rows of one slice (contiguous: band_rows * W floats), "height" = the
    // halo slices of one channel plane, "depth" = the B * C channel planes.
    const size_t pitch = (size_t)HW * sizeof(float);
    for (int r = row0, j = 0; r < H; r += band_rows, ++j) {
        const int rows = (H - r < band_rows) ? H - r : band_rows;
        cudaMemcpy3DParms p = {};
        p.srcPtr = make_cudaPitchedPtr(const_cast<float*>(tail), pitch, pitch, (size_t)(tail_plane / HW));
        p.dstPtr = make_cudaPitchedPtr(stage, pitch, pitch, (size_t)halo_slices);
        p.srcPos = make_cudaPos((size_t)r * W * sizeof(float), 0, 0);
        p.dstPos = p.srcPos;
        p.extent = make_cudaExtent((size_t)rows * W * sizeof(float), (size_t)halo_slices, (size_t)B * C);
        p.kind = cudaMemcpyDeviceToDevice;
        cudaError_t e = cudaMemcpy3DAsync(&p, st);
        if (e != cudaSuccess) return (int)e;
        // default flags: the copy before it in the stream is visible to whoever observes the value
        CUresult cr = wv(reinterpret_cast<CUstream>(st), reinterpret_cast<CUdeviceptr>(flags + j), epoch, 0);
        if (cr != CUDA_SUCCESS) return (int)cr;
    }
    return PEA_OK;
}

extern "C" int pea_zshard_affs_fwd(const float* emb, int B, int C, int D, int H, int W, int64_t emb_plane,
                                   int64_t affs_plane, const float* halo, int halo_slices, int64_t halo_plane,
                                   const float* stage, const uint32_t* flags, uint32_t epoch, int row0, int band_rows,
                                   int relu, float* affs, void* stream) {
    if (!emb || !affs || !halo || !stage || !flags) return PEA_E_NULL;
    if (B < 1 || C != 16 || D < 1 || H < 1 || W < 1) return PEA_E_SHAPE;
    const int64_t HW = (int64_t)H * W, N = (int64_t)D * HW;
    if (halo_slices < 1 || halo_slices > 4 || halo_plane < (int64_t)halo_slices * HW) return PEA_E_SHAPE;
    if (emb_plane == 0) emb_plane = N;
    if (affs_plane == 0) affs_plane = N;
    if (emb_plane < N || affs_plane < N) return PEA_E_SHAPE;
    if (row0 < 0 || band_rows < 8 || row0 % 8 || band_rows % 8) return PEA_E_SHAPE;  // bands start on tile rows (8 rows)
    static const int32_t norm5[12 * 3] = {-1, 0, 0, 0, -1, 0, 0, 0, -1, -2, 0, 0, 0, -3, 0, 0, 0, -3,
                                          -3, 0, 0, 0, -9, 0, 0, 0, -9, -4, 0, 0, 0, -27, 0, 0, 0, -27};
    OffsetTable tab;
    tab.K = 12;
    for (int k = 0; k < 12; ++k) {
        tab.dz[k] = norm5[3 * k]; tab.dy[k] = norm5[3 * k + 1]; tab.dx[k] = norm5[3 * k + 2];
        tab.gcoef[k] = 0.f;
    }
    if (!tiled3d_fwd_eligible(C, W, tab, PEA_BORDER_ZERO, false, emb, emb_plane)) return PEA_E_MODE;
    Dims d;
    d.B = B; d.D = D; d.H = H; d.W = W; d.N = N;
    StagedHalo sh;
    sh.data = stage; sh.plane = (int64_t)halo_slices * HW; sh.flags = flags; sh.epoch = epoch; sh.row0 = row0; sh.band_rows = band_rows;
    return launch_fwd3d_norm5(emb, d, 0, emb_plane, affs_plane, affs, halo, halo_slices, halo_plane, relu ? 1 : 0,
                              static_cast<cudaStream_t>(stream), &sh);
}
