// read_probe.cu -- what does HBM deliver to a READ-MOSTLY kernel on this B200?  The roofline denominator in
// MEASURED_PEAKS.json is a copy (read + write counted); the status sweep reads 40 B per slot and writes ~0.1 B.
// Variants over 640 MiB: (a) plain grid-stride uint4 read, (b) the sweep's two streams (32 B record + 8 B hash per
// slot), (c) copy.  Not product code.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void __launch_bounds__(1024, 2) k_read(const uint4* __restrict__ p, size_t n, uint32_t* out) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const uint4 a = __ldg(p + i), b = __ldg(p + i + stride), c = __ldg(p + i + 2 * stride), d = __ldg(p + i + 3 * stride);
        acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
    }
    for (; i < n; i += stride) { const uint4 a = __ldg(p + i); acc ^= a.x ^ a.y ^ a.z ^ a.w; }
    if (acc == 0x12345678u) out[0] = acc;
}
__global__ void __launch_bounds__(1024, 2) k_read2(const uint4* __restrict__ rec, const unsigned long long* __restrict__ h, size_t nslots, uint32_t* out) {
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < nslots; s += stride) {
        const uint4 a = __ldg(rec + 2 * s), b = __ldg(rec + 2 * s + 1);
        const unsigned long long v = h[s];
        acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ (uint32_t)v ^ (uint32_t)(v >> 32);
    }
    if (acc == 0x12345678u) out[0] = acc;
}
__global__ void __launch_bounds__(1024, 2) k_copy(const uint4* __restrict__ p, uint4* __restrict__ q, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) q[i] = __ldg(p + i);
}

int main() {
    const size_t bytes = 640ull << 20, n = bytes / 16, nslots = bytes / 32;
    uint4 *a, *b; unsigned long long* h; uint32_t* out;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&h, nslots * 8)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes)); CK(cudaMemset(h, 3, nslots * 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int grid : {148 * 2, 148 * 8, 148 * 32}) {
        for (int v = 0; v < 3; ++v) {
            float best = 1e9f;
            for (int r = 0; r < 6; ++r) {
                CK(cudaEventRecord(e0));
                if (v == 0) k_read<<<grid, 1024>>>(a, n, out);
                else if (v == 1) k_read2<<<grid, 1024>>>(a, h, nslots, out);
                else k_copy<<<grid, 1024>>>(a, b, n);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (r > 0 && ms < best) best = ms;
            }
            const double moved = v == 0 ? (double)bytes : v == 1 ? (double)bytes + nslots * 8.0 : 2.0 * bytes;
            printf("{\"variant\": \"%s\", \"grid\": %d, \"us\": %.1f, \"GBps\": %.0f}\n", v == 0 ? "read uint4 grid-stride" : v == 1 ? "read 32B record + 8B hash per slot" : "copy (read+write)",
                   grid, best * 1e3, moved / (best * 1e-3) / 1e9);
        }
    }
    return 0;
}
