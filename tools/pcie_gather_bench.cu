// Measurement helper (not part of the library): throughput of small random reads of pinned host memory from a kernel,
// by request size (16..128 contiguous bytes per group of lanes), loads in flight per thread and grid size.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pcie_gather_bench tools/pcie_gather_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

template <int G, int U>   // G lanes x 16 B per request, U independent requests per thread and iteration
__global__ void gather(const float4* __restrict__ host, size_t n_chunks128, int iters, float* sink, uint32_t seed) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t group = tid / G, sub = tid % G;
    float acc = 0.f;
    uint32_t s = seed ^ (group * 2654435761u);
    for (int it = 0; it < iters; ++it) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = s * 1664525u + 1013904223u;
            const size_t chunk = (size_t)(s >> 4) % n_chunks128;             // a random 128-byte line
            const size_t off = chunk * 8 + ((s >> 1) % (8 / G)) * G + sub;     // G consecutive float4 inside it
            v[u] = __ldg(host + off);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 123.456f) sink[0] = acc;
}

template <int G, int U>
void run(const float4* dptr, size_t n_chunks, int ctas, int threads, float* sink) {
    const int iters = 64;
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    gather<G, U><<<ctas, threads>>>(dptr, n_chunks, 4, sink, 1u);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    gather<G, U><<<ctas, threads>>>(dptr, n_chunks, iters, sink, 7u);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double req = (double)ctas * threads / G * U * iters;
    printf("bytes/request %4d  in flight/thread %d  grid %3dx%3d : %7.1f M requests/s  %6.2f GB/s  (%.3f ms)\n", G * 16, U, ctas,
           threads, req / ms / 1e3, req * G * 16 / ms / 1e6, ms);
}

int main() {
    const size_t bytes = (size_t)2 << 30;
    float4* h = nullptr;
    if (cudaHostAlloc(&h, bytes, cudaHostAllocMapped) != cudaSuccess) return 1;
    for (size_t i = 0; i < bytes / 16; i += 64) h[i] = make_float4(1.f, 2.f, 3.f, 4.f);
    float4* d = nullptr;
    cudaHostGetDevicePointer(&d, h, 0);
    float* sink;
    cudaMalloc(&sink, 4);
    const size_t n_chunks = bytes / 128;
    for (int ctas : {4, 16, 148}) {
        run<1, 1>(d, n_chunks, ctas, 512, sink);
        run<1, 4>(d, n_chunks, ctas, 512, sink);
        run<1, 8>(d, n_chunks, ctas, 512, sink);
        run<2, 4>(d, n_chunks, ctas, 512, sink);
        run<2, 8>(d, n_chunks, ctas, 512, sink);
        run<4, 4>(d, n_chunks, ctas, 512, sink);
        run<4, 8>(d, n_chunks, ctas, 512, sink);
        run<8, 4>(d, n_chunks, ctas, 512, sink);
        run<8, 8>(d, n_chunks, ctas, 512, sink);
    }
    return 0;
}
