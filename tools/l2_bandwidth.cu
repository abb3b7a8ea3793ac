// L2 -> SM read bandwidth of the device: every CTA streams an L2-resident buffer with LDG.128 (ld.global.nc), repeatedly, from a
// different starting block so that the CTAs do not walk in lock-step.  Used as the denominator of the L2 line of the roofline of
// the phase-aligned k_mc_accumulate (profiles/r02_acc_aligned.md).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2bw tools/l2_bandwidth.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void __launch_bounds__(256) k_read(const float4* __restrict__ buf, size_t n4, int reps, float* __restrict__ sink) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t start = ((size_t)blockIdx.x * 2654435761u) % n4;
    for (int r = 0; r < reps; r++) {
        size_t i = (start + (size_t)r * 977 * blockDim.x + threadIdx.x) % n4;
        for (size_t k = 0; k < n4 / stride; k += 4) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                v[u] = __ldg(buf + i);
                i += stride;
                if (i >= n4) i -= n4;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[0] = acc.x;
}

int main(int argc, char** argv) {
    int dev = 0, nsm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    float* sink; cudaMalloc(&sink, 4);
    printf("{\"sms\": %d, \"runs\": [", nsm);
    const int sizes_mb[] = {8, 16, 32, 48, 64, 96, 256, 1024};
    double best = 0; int best_mb = 0, first = 1;
    for (int si = 0; si < 8; si++) {
        const size_t bytes = (size_t)sizes_mb[si] << 20, n4 = bytes / 16;
        float4* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes);
        for (int cps = 4; cps <= 8; cps += 4) {
            const int grid = nsm * cps, reps = sizes_mb[si] <= 96 ? 40 : 4;
            k_read<<<grid, 256>>>(buf, n4, 2, sink);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            k_read<<<grid, 256>>>(buf, n4, reps, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const size_t per_rep = (n4 / ((size_t)grid * 256) + 3) / 4 * 4 * (size_t)grid * 256 * 16;
            const double gbs = (double)per_rep * reps / (ms * 1e-3) / 1e9;
            printf("%s{\"buffer_MB\": %d, \"ctas_per_sm\": %d, \"ms\": %.3f, \"GBs\": %.0f}", first ? "" : ", ", sizes_mb[si], cps, ms, gbs);
            first = 0;
            if (sizes_mb[si] <= 64 && gbs > best) { best = gbs; best_mb = sizes_mb[si]; }
        }
        cudaFree(buf);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("], \"l2_read_GBs\": %.0f, \"at_buffer_MB\": %d, \"cuda_error\": %d}\n", best, best_mb, (int)e);
    return e != cudaSuccess;
}
