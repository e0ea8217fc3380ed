// microbenchmark: dependent-load latency on B200 for the load flavours the step kernel uses
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>
template <int MODE> __device__ __forceinline__ uint32_t ld(const uint32_t* p) {
    uint32_t v;
    if (MODE == 0) v = *p;
    else if (MODE == 1) v = __ldcg(p);
    else if (MODE == 2) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else if (MODE == 3) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    else v = *(const volatile uint32_t*)p;
    return v;
}
template <int MODE> __global__ void chase(const uint32_t* a, int hops, long long* out, uint32_t* sink) {
    uint32_t i = threadIdx.x + blockIdx.x * 1024; // different chains per block
    long long t0 = clock64();
    for (int h = 0; h < hops; h++) i = ld<MODE>(a + i);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / hops;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = i;
}
__global__ void toucher(uint32_t* a, size_t n) { // rewrite every element from (other) SMs: same values
    for (size_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { uint32_t v = a[i]; __stcg(a + i, v); }
}
int main() {
    for (size_t mb : {1, 8, 64, 512}) {
        size_t n = mb * 1024 * 1024 / 4;
        std::vector<uint32_t> h(n); std::vector<uint32_t> perm(n);
        for (size_t i = 0; i < n; i++) perm[i] = (uint32_t)i;
        std::mt19937 g(1); std::shuffle(perm.begin(), perm.end(), g);
        for (size_t i = 0; i < n; i++) h[perm[i]] = perm[(i + 1) % n];
        uint32_t *d, *sink; long long* out;
        cudaMalloc(&d, n * 4); cudaMalloc(&sink, 4 * 4096); cudaMalloc(&out, 8 * 64);
        cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
        long long r[5];
        const char* names[5] = {"ld", "ld.cg", "ld.relaxed.gpu", "ld.acquire.gpu", "volatile"};
        for (int pass = 0; pass < 2; pass++) {
            for (int m = 0; m < 5; m++) {
                if (pass == 1) { toucher<<<148 * 4, 256>>>(d, n); cudaDeviceSynchronize(); }
                int hops = 2000;
                switch (m) {
                    case 0: chase<0><<<1, 1>>>(d, hops, out, sink); break;
                    case 1: chase<1><<<1, 1>>>(d, hops, out, sink); break;
                    case 2: chase<2><<<1, 1>>>(d, hops, out, sink); break;
                    case 3: chase<3><<<1, 1>>>(d, hops, out, sink); break;
                    case 4: chase<4><<<1, 1>>>(d, hops, out, sink); break;
                }
                cudaDeviceSynchronize();
                cudaMemcpy(&r[m], out, 8, cudaMemcpyDeviceToHost);
            }
            printf("%4zu MB %s:", mb, pass ? "after rewrite by all SMs" : "after H2D copy          ");
            for (int m = 0; m < 5; m++) printf("  %s=%lld", names[m], r[m]);
            printf("  cycles/hop\n");
        }
        // loaded: 148*8 warps chasing at once
        chase<1><<<148, 256>>>(d, 500, out, sink); cudaDeviceSynchronize();
        long long x; cudaMemcpy(&x, out, 8, cudaMemcpyDeviceToHost);
        printf("%4zu MB ld.cg with 148x8 warps (32 chains each) chasing: %lld cycles/hop\n", mb, x);
        cudaFree(d); cudaFree(sink); cudaFree(out);
    }
    return 0;
}
