// Calibration microbenchmarks for the kernel design (DESIGN.md "machine numbers"): HBM stream, L2-hit read
// bandwidth, shared-memory read bandwidth, FFMA vs packed FFMA2 issue rate.  Build: make -C tools
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void read_kernel(const float4* __restrict__ src, size_t n4, int reps, float* sink) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride * 4) {
            float4 a = src[i];
            float4 b = (i + stride < n4) ? src[i + stride] : make_float4(0, 0, 0, 0);
            float4 c = (i + 2 * stride < n4) ? src[i + 2 * stride] : make_float4(0, 0, 0, 0);
            float4 d = (i + 3 * stride < n4) ? src[i + 3 * stride] : make_float4(0, 0, 0, 0);
            acc += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
        }
    }
    if (acc == 123.456f) *sink = acc;
}

__global__ void copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) dst[i] = src[i];
}

// each thread: ILP independent FFMA chains (3 distinct register operands)
template <int ILP>
__global__ void ffma_kernel(float* out, int iters, float a, float b) {
    float acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 0.001f + i;
    float x = a + threadIdx.x * 1e-6f, y = b;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], x, y);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void ffma2_kernel(float* out, int iters, float a, float b) {
    unsigned long long acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        float2 v = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
        acc[i] = *reinterpret_cast<unsigned long long*>(&v);
    }
    float2 xv = make_float2(a + threadIdx.x * 1e-6f, a), yv = make_float2(b, b * 0.5f);
    unsigned long long x = *reinterpret_cast<unsigned long long*>(&xv), y = *reinterpret_cast<unsigned long long*>(&yv);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(x), "l"(y));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float2 v = *reinterpret_cast<float2*>(&acc[i]); s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// shared-memory read bandwidth with LDS.128, conflict-free
__global__ void lds_kernel(float* out, int iters) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 v = sm[(idx + j * 256) & 4095];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        idx = (idx + 32) & 4095;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, L2 %d MB, clock attr %d MHz\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize >> 20, clk_khz / 1000);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int sms = prop.multiProcessorCount;
    float* sink; CK(cudaMalloc(&sink, 1 << 22));

    // ---- HBM: copy and read of 2 GiB
    {
        size_t bytes = (size_t)2 << 30;
        float4 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
        CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 0, bytes));
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            copy_kernel<<<sms * 8, 512>>>(a, b, bytes / 16);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        }
        printf("HBM copy 2GiB: %.1f GB/s (read+write)\n", 2.0 * bytes / time_ms(e0, e1) / 1e6);
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            read_kernel<<<sms * 8, 512>>>(a, bytes / 16, 1, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        }
        printf("HBM read 2GiB: %.1f GB/s\n", 1.0 * bytes / time_ms(e0, e1) / 1e6);
        // ---- L2-hit reads: working sets that fit
        for (size_t mb : {8, 16, 32, 48, 64, 96}) {
            size_t wb = mb << 20;
            read_kernel<<<sms * 8, 512>>>(a, wb / 16, 2, sink);  // warm
            int reps = 40;
            CK(cudaEventRecord(e0));
            read_kernel<<<sms * 8, 512>>>(a, wb / 16, reps, sink);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            printf("L2 read, %3zu MB working set: %.1f GB/s\n", mb, (double)wb * reps / time_ms(e0, e1) / 1e6);
        }
        CK(cudaFree(a)); CK(cudaFree(b));
    }
    // ---- FFMA issue rates
    {
        float* out; CK(cudaMalloc(&out, sms * 8 * 1024 * sizeof(float)));
        const int iters = 4096;
        for (int blocks_per_sm : {1, 2}) {
            for (int threads : {256, 512, 1024}) {
                ffma_kernel<8><<<sms * blocks_per_sm, threads>>>(out, 16, 1.0001f, 0.5f);
                CK(cudaEventRecord(e0));
                ffma_kernel<8><<<sms * blocks_per_sm, threads>>>(out, iters, 1.0001f, 0.5f);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                double fma = (double)sms * blocks_per_sm * threads * 8.0 * iters;
                double ms = time_ms(e0, e1);
                printf("FFMA  %d x %4d thr: %.2f TFMA/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", blocks_per_sm, threads, fma / ms / 1e9, fma / ms / 1e3 / sms / 1965e3);
                ffma2_kernel<8><<<sms * blocks_per_sm, threads>>>(out, 16, 1.0001f, 0.5f);
                CK(cudaEventRecord(e0));
                ffma2_kernel<8><<<sms * blocks_per_sm, threads>>>(out, iters, 1.0001f, 0.5f);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                ms = time_ms(e0, e1);
                printf("FFMA2 %d x %4d thr: %.2f TFMA/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", blocks_per_sm, threads, 2 * fma / ms / 1e9, 2 * fma / ms / 1e3 / sms / 1965e3);
            }
        }
        // ---- LDS.128 bandwidth
        for (int threads : {256, 512, 1024}) {
            const int it2 = 2048;
            CK(cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            lds_kernel<<<sms, threads, 65536>>>(out, 16);
            CK(cudaEventRecord(e0));
            lds_kernel<<<sms, threads, 65536>>>(out, it2);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            double bytes = (double)sms * threads * 8.0 * 16 * it2;
            double ms = time_ms(e0, e1);
            printf("LDS.128 %4d thr/SM: %.1f TB/s  (%.1f B/clk/SM at 1.965 GHz)\n", threads, bytes / ms / 1e9, bytes / ms / 1e3 / sms / 1965e3);
        }
        CK(cudaFree(out));
    }
    CK(cudaGetLastError());
    printf("done\n");
    return 0;
}
