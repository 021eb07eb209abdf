// NVLink peer-memory microbenchmark (one process, devices 0 and 1): what can a kernel on GPU 0 move to /
// from GPU 1's HBM, by access width and pattern, against cudaMemcpyPeerAsync?  Decides between push
// (remote stores) and pull (remote loads) for the slab-FFT transposes.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a p2p.cu -o p2p && ./p2p
#include <cuda_runtime.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <typename V>
__global__ void copy_kernel(const V* __restrict__ src, V* __restrict__ dst, long long n, int seg, long long gap) {
    // seg = contiguous elements per segment, gap = distance between segments in dst (elements)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long s = i / seg, o = i % seg;
        dst[s * gap + o] = src[i];
    }
}
template <typename V, int U>
__global__ void copy_unrolled(const V* __restrict__ src, V* __restrict__ dst, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        V v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = src[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u) dst[i + u * stride] = v[u];
    }
    for (; i < n; i += stride) dst[i] = src[i];
}

int main() {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
    const size_t bytes = 256ull << 20;
    void *a0, *b0, *a1;
    CK(cudaSetDevice(1)); CK(cudaMalloc(&a1, 2 * bytes)); CK(cudaMemset(a1, 1, 2 * bytes));
    CK(cudaDeviceEnablePeerAccess(0, 0));
    CK(cudaSetDevice(0)); CK(cudaMalloc(&a0, 2 * bytes)); CK(cudaMalloc(&b0, 2 * bytes)); CK(cudaMemset(a0, 2, 2 * bytes));
    CK(cudaDeviceEnablePeerAccess(1, 0));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time = [&](const char* name, auto fn) {
        fn(); cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int r = 0; r < 5; ++r) fn();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-52s %8.1f GB/s\n", name, 5.0 * bytes / (ms * 1e-3) / 1e9);
        return 0;
    };
    const long long n16 = bytes / 16, n8 = bytes / 8;
    for (int blocks : {148 * 4, 148 * 16, 148 * 64}) {
        printf("-- grid %d x 256\n", blocks);
        time("local  copy 16B", [&] { copy_kernel<float4><<<blocks, 256>>>((float4*)a0, (float4*)b0, n16, 1 << 30, 0); });
        time("push   16B contiguous", [&] { copy_kernel<float4><<<blocks, 256>>>((float4*)a0, (float4*)a1, n16, 1 << 30, 0); });
        time("pull   16B contiguous", [&] { copy_kernel<float4><<<blocks, 256>>>((float4*)a1, (float4*)b0, n16, 1 << 30, 0); });
        time("push   16B unrolled x4", [&] { copy_unrolled<float4, 4><<<blocks, 256>>>((float4*)a0, (float4*)a1, n16); });
        time("pull   16B unrolled x4", [&] { copy_unrolled<float4, 4><<<blocks, 256>>>((float4*)a1, (float4*)b0, n16); });
        time("push   8B contiguous", [&] { copy_kernel<float2><<<blocks, 256>>>((float2*)a0, (float2*)a1, n8, 1 << 30, 0); });
        time("push   8B, 64-byte segments scattered (gap 2x)", [&] { copy_kernel<float2><<<blocks, 256>>>((float2*)a0, (float2*)a1, n8, 8, 16); });
        time("push   16B, 1040-byte rows (gap 2x)", [&] { copy_kernel<float4><<<blocks, 256>>>((float4*)a0, (float4*)a1, n16, 65, 130); });
        time("pull   8B, 64-byte segments (src contiguous)", [&] { copy_kernel<float2><<<blocks, 256>>>((float2*)a1, (float2*)b0, n8, 8, 16); });
    }
    time("cudaMemcpyPeerAsync 0 -> 1", [&] { cudaMemcpyPeerAsync(a1, 1, a0, 0, bytes); });
    time("cudaMemcpyPeerAsync 1 -> 0", [&] { cudaMemcpyPeerAsync(b0, 0, a1, 1, bytes); });
    // both directions at once (GPU 1 pushes to GPU 0 on its own stream)
    cudaStream_t s1; CK(cudaSetDevice(1)); CK(cudaStreamCreate(&s1)); CK(cudaSetDevice(0));
    time("push 16B 0->1 while 1->0 pushes too (per direction)", [&] {
        copy_kernel<float4><<<148 * 16, 256>>>((float4*)a0, (float4*)a1, n16, 1 << 30, 0);
        cudaSetDevice(1);
        copy_kernel<float4><<<148 * 16, 256, 0, s1>>>((float4*)((char*)a1 + bytes), (float4*)((char*)a0 + bytes), n16, 1 << 30, 0);
        cudaSetDevice(0);
    });
    cudaSetDevice(1); cudaDeviceSynchronize(); cudaSetDevice(0);
    return 0;
}
