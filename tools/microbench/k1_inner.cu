// Micro-benchmark of candidate inner loops for the select grid kernel (one offer-score = one (row, offer)
// feasibility test + argmin update).  Not product code: used to pick the instruction mix.
//   V0: ((o - r) & G) == G -> SEL           (LOP3 + ISETP + SEL on ALU, IMAD.IADD on FMA)
//   V1: (~(o - r) & G) == 0 -> SEL          (LOP3 with predicate out + SEL on ALU)
//   V4: key = (~d & G) | (d & LOW); min3    (LOP3 + 1/2 VIMNMX3 on ALU)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int V, int R>
__global__ void __launch_bounds__(256, 3) k(const uint4* __restrict__ in, const uint32_t* __restrict__ rws, uint32_t* out,
                                            uint32_t guard, uint32_t low, int nchunks, int reps) {
    extern __shared__ uint4 tile[];
    for (int i = threadIdx.x; i < nchunks * 32; i += blockDim.x) tile[i] = in[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t rw[R], best[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { rw[r] = rws[(blockIdx.x * 256 + threadIdx.x) / 32 * R + r]; best[r] = 0xFFFFFFFFu; }
    for (int rep = 0; rep < reps; ++rep) {
        for (int ch = nchunks - 1; ch >= 0; --ch) {
            const uint4 o = tile[ch * 32 + lane];
            const uint32_t j0 = (uint32_t)(rep * nchunks + ch) * 128u + lane * 4u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (V == 0) {
                    if (((o.w - rw[r]) & guard) == guard) best[r] = j0 + 3;
                    if (((o.z - rw[r]) & guard) == guard) best[r] = j0 + 2;
                    if (((o.y - rw[r]) & guard) == guard) best[r] = j0 + 1;
                    if (((o.x - rw[r]) & guard) == guard) best[r] = j0;
                } else if (V == 1) {
                    if ((~(o.w - rw[r]) & guard) == 0) best[r] = j0 + 3;
                    if ((~(o.z - rw[r]) & guard) == 0) best[r] = j0 + 2;
                    if ((~(o.y - rw[r]) & guard) == 0) best[r] = j0 + 1;
                    if ((~(o.x - rw[r]) & guard) == 0) best[r] = j0;
                } else {
                    const uint32_t d0 = o.x - rw[r], d1 = o.y - rw[r], d2 = o.z - rw[r], d3 = o.w - rw[r];
                    const uint32_t k0 = (~d0 & guard) | (d0 & low), k1 = (~d1 & guard) | (d1 & low);
                    const uint32_t k2 = (~d2 & guard) | (d2 & low), k3 = (~d3 & guard) | (d3 & low);
                    best[r] = min(min(best[r], k0), k1);
                    best[r] = min(min(best[r], k2), k3);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        uint32_t m = __reduce_min_sync(0xFFFFFFFFu, best[r]);
        if (lane == 0) out[(blockIdx.x * 256 + threadIdx.x) / 32 * R + r] = m;
    }
}

template <int V, int R>
int run(const char* name, const uint4* in, const uint32_t* rws, uint32_t* out, uint32_t guard, uint32_t low) {
    const int nchunks = 128, reps = 64, grid = 148 * 3 * 8;
    const size_t smem = (size_t)nchunks * 32 * 16;
    CK(cudaFuncSetAttribute(k<V, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int w = 0; w < 2; ++w) k<V, R><<<grid, 256, smem>>>(in, rws, out, guard, low, nchunks, reps);
    CK(cudaEventRecord(a));
    const int iters = 5;
    for (int w = 0; w < iters; ++w) k<V, R><<<grid, 256, smem>>>(in, rws, out, guard, low, nchunks, reps);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= iters;
    const double scores = (double)grid * 8 * R * (double)nchunks * 128 * reps;
    printf("%-28s R=%2d  %.3f ms  %.3e offer-scores/s  (%.2f scores/clk/SM at 1965 MHz)\n", name, R, ms, scores / (ms * 1e-3),
           scores / (ms * 1e-3) / 148 / 1.965e9);
    return 0;
}

int main() {
    const int nchunks = 128;
    uint4* in; uint32_t *rws, *out;
    CK(cudaMalloc(&in, nchunks * 32 * 16)); CK(cudaMalloc(&rws, 148 * 3 * 8 * 8 * 16 * 4)); CK(cudaMalloc(&out, 148 * 3 * 8 * 8 * 16 * 4));
    CK(cudaMemset(in, 0x5A, nchunks * 32 * 16)); CK(cudaMemset(rws, 0x11, 148 * 3 * 8 * 8 * 16 * 4));
    const uint32_t guard = 0x80808000u, low = 0x3FFFu;
    if (run<0, 16>("V0 lop3+isetp+sel", in, rws, out, guard, low)) return 1;
    if (run<1, 16>("V1 lop3.p+sel", in, rws, out, guard, low)) return 1;
    if (run<4, 16>("V4 lop3+vimnmx3", in, rws, out, guard, low)) return 1;
    if (run<4, 8>("V4 lop3+vimnmx3", in, rws, out, guard, low)) return 1;
    if (run<1, 8>("V1 lop3.p+sel", in, rws, out, guard, low)) return 1;
    if (run<4, 24>("V4 lop3+vimnmx3", in, rws, out, guard, low)) return 1;
    return 0;
}
