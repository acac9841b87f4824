// L2 bandwidth microbenchmark: repeatedly read / RED / write an L2-resident buffer with 16-byte accesses.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2bw l2bw.cu && ./l2bw
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_read(const float4* __restrict__ p, size_t n4, int iters, float* out) {
    float acc = 0.f;
    for (int it = 0; it < iters; ++it)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            float4 v = __ldcg(p + ((i + (size_t)it * 977) % n4));
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) *out = acc;
}
__global__ void k_write(float4* __restrict__ p, size_t n4, int iters) {
    for (int it = 0; it < iters; ++it)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
            p[i] = make_float4(it, 1.f, 2.f, 3.f);
}
__global__ void k_red(float4* __restrict__ p, size_t n4, int iters) {
    for (int it = 0; it < iters; ++it)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
            atomicAdd(p + i, make_float4(1.f, 1.f, 1.f, 1.f));
}
__global__ void k_red_rand(float4* __restrict__ p, size_t n4, int iters) {
    for (int it = 0; it < iters; ++it)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            size_t j = (i * 2654435761ull + it * 40503ull) % n4;
            atomicAdd(p + j, make_float4(1.f, 1.f, 1.f, 1.f));
        }
}
__global__ void k_read_rand(const float4* __restrict__ p, size_t n4, int iters, float* out) {
    float acc = 0.f;
    for (int it = 0; it < iters; ++it)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            size_t j = (i * 2654435761ull + it * 40503ull) % n4;
            float4 v = __ldg(p + j);
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f) *out = acc;
}
int main() {
    for (size_t mb : {16, 32, 64, 256}) {
        size_t bytes = mb << 20, n4 = bytes / 16;
        float4* p; float* out; cudaMalloc(&p, bytes); cudaMalloc(&out, 4); cudaMemset(p, 0, bytes);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int iters = 20, grid = 148 * 8, block = 256;
        float ms;
        k_read<<<grid, block>>>(p, n4, 2, out);
        cudaEventRecord(e0); k_read<<<grid, block>>>(p, n4, iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("%4zu MB  read      %7.1f GB/s\n", mb, bytes * (double)iters / ms / 1e6);
        cudaEventRecord(e0); k_read_rand<<<grid, block>>>(p, n4, iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("%4zu MB  read rand %7.1f GB/s (16 B useful per access)\n", mb, bytes * (double)iters / ms / 1e6);
        cudaEventRecord(e0); k_write<<<grid, block>>>(p, n4, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("%4zu MB  write     %7.1f GB/s\n", mb, bytes * (double)iters / ms / 1e6);
        cudaEventRecord(e0); k_red<<<grid, block>>>(p, n4, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("%4zu MB  red.v4    %7.1f GB/s\n", mb, bytes * (double)iters / ms / 1e6);
        cudaEventRecord(e0); k_red_rand<<<grid, block>>>(p, n4, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("%4zu MB  red.v4 rand %5.1f GB/s\n", mb, bytes * (double)iters / ms / 1e6);
        cudaFree(p); cudaFree(out);
    }
    return 0;
}
