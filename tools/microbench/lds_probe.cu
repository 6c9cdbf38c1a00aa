// lds_probe.cu -- what does a shared-memory load cost on B200, per warp instruction, when the lanes of a warp
// share addresses?  Decides whether "threshold-uniform warps + one broadcast LDS.128 for four chunks" (VERDICT r1,
// next-round item 4) can lift K1 off the shared-memory bound.  Not product code.
//
// Every variant runs 32 warps per SM (one 1024-thread CTA per SM, 148 CTAs) issuing ITER loads each from a 64 KB
// stage; the result is cycles per warp-level load instruction per SM (1.0 = one 128-byte wavefront per clock).
//   A  LDS.32   32 distinct consecutive words          (conflict-free, 128 B out)
//   B  LDS.32   all lanes one address                  (broadcast, 128 B out)
//   C  LDS.128  all lanes one 16-byte address          (broadcast, 512 B out)
//   D  LDS.128  2 distinct 16-byte addresses, bank groups apart
//   E  LDS.128  8 distinct 16-byte addresses, conflict-free bank groups
//   F  LDS.128  32 distinct consecutive 16-byte units  (conflict-free, 512 B unique)
//   G  LDS.64   all lanes one address
//   H  LDC      constant bank, warp-uniform index      (64 KB __constant__)
// Also: pinned H2D / D2H bandwidth of the box (one 64 MB copy each way), for the e2e floor.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__constant__ uint32_t c_tab[16384];

__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}

template <int V>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, long long* clk, int iters) {
    extern __shared__ __align__(128) uint32_t s[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) s[i] = i * 2654435761u;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(s);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t off;  // per-lane byte offset inside a 2 KB window
    switch (V) {
        case 0: off = lane * 4; break;
        case 1: off = 0; break;
        case 2: off = 0; break;
        case 3: off = (lane >> 4) * 16; break;            // 2 addresses, adjacent bank groups
        case 4: off = (lane >> 2) * 16; break;            // 8 addresses, 8 bank groups
        case 5: off = lane * 16; break;
        case 6: off = 0; break;
        default: off = 0; break;
    }
    uint32_t acc = 0;
    uint32_t a = base + off + warp * 64;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t addr = a + (uint32_t)((i + u) & 63) * 512u;  // walks the stage; address stays warp-shaped
            if (V == 0 || V == 1) acc ^= lds32(addr);
            else if (V == 6) { uint2 v = lds64(addr); acc ^= v.x ^ v.y; }
            else if (V == 7) acc ^= c_tab[((i + u) * 37 + warp) & 16383];
            else { uint4 v = lds128(addr); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int V>
int run(const char* name, uint32_t* out, long long* clk) {
    const int iters = 1 << 14;
    CK(cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    k<V><<<148, 1024, 65536>>>(out, clk, iters);
    k<V><<<148, 1024, 65536>>>(out, clk, iters);
    CK(cudaDeviceSynchronize());
    long long h[148];
    CK(cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost));
    double mean = 0;
    for (int i = 0; i < 148; ++i) mean += (double)h[i];
    mean /= 148;
    printf("{\"variant\": \"%s\", \"cycles_per_warp_load_per_sm\": %.3f}\n", name, mean / ((double)iters * 32));
    return 0;
}

int main() {
    uint32_t* out; long long* clk;
    CK(cudaMalloc(&out, 148 * 1024 * 4));
    CK(cudaMalloc(&clk, 148 * 8));
    uint32_t* hc = (uint32_t*)malloc(65536);
    for (int i = 0; i < 16384; ++i) hc[i] = i * 40503u;
    CK(cudaMemcpyToSymbol(c_tab, hc, 65536));
    if (run<0>("A LDS.32 distinct words", out, clk)) return 1;
    if (run<1>("B LDS.32 one address", out, clk)) return 1;
    if (run<2>("C LDS.128 one address", out, clk)) return 1;
    if (run<3>("D LDS.128 two addresses", out, clk)) return 1;
    if (run<4>("E LDS.128 eight addresses", out, clk)) return 1;
    if (run<5>("F LDS.128 32 distinct units", out, clk)) return 1;
    if (run<6>("G LDS.64 one address", out, clk)) return 1;
    if (run<7>("H LDC uniform index", out, clk)) return 1;
    // PCIe: pinned copies
    const size_t n = 64u << 20;
    void *h, *d;
    CK(cudaHostAlloc(&h, n, cudaHostAllocPortable));
    CK(cudaMalloc(&d, n));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int dir = 0; dir < 2; ++dir) {
        float best = 1e9f;
        for (int r = 0; r < 5; ++r) {
            CK(cudaEventRecord(e0));
            if (dir == 0) CK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice)); else CK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            best = ms < best ? ms : best;
        }
        printf("{\"pcie\": \"%s\", \"GBps\": %.1f}\n", dir == 0 ? "H2D pinned 64 MiB" : "D2H pinned 64 MiB", n / (best * 1e-3) / 1e9);
    }
    return 0;
}
