// select.cu -- the P pods x G offers grid: Client.GetGPUTypes evaluated for every pod row at once
// (reference runpod_client.go:465-509, called per pod from :1281).
//
// Data flow per call (all on one stream):
//   k_pod_prep        pod columns -> per-row threshold word (rank compression), rows grouped by (cloud view,
//                     vcpu constrains, ram constrains), pos[] = none
//   k_select_bitmap   kind 4: grid = (row tile) x (offer segment).  The segment of the price-sorted, bit-sliced
//                     view (one 32-bit mask per column threshold per 32-offer chunk) is staged in shared
//                     memory with one bulk async copy (TMA, mbarrier completion); each lane owns RPL rows and
//                     ANDs its masks: one LOP3 = 32 (pod, offer) pairs.  Every pair is evaluated.
//   k_select_packed   kinds 3/2: one u32 of rank fields per offer, R rows per warp in registers, lanes stride
//                     over offers, row argmin by redux.sync.min over lanes.
//   k_select_wide     kind 1: int32 compares on (mem, vcpu, ram) for tables that do not rank-pack.
//   (all grid kernels) atomicMin merges segments; the last CTA of a row tile applies price < maxPrice to the
//                     winner (the bound is a prefix of the sorted order) and stores the offer index into
//                     every peer's assignment vector (the fused all-gather).
//   k_select_top5*    optional: the whole <=5 gpuTypeIds list per row (runpod_client.go:502-509).
//   k_peer_fence      one warp: signal / wait across GPUs after the fused gather.
#include <math_constants.h>

#include <cstdlib>
#include <cstring>

#include "rpk_internal.cuh"

namespace rpk {

__device__ __forceinline__ uint32_t lower_bound_i32(const int32_t* __restrict__ a, uint32_t n, int32_t x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid; }
    return lo;
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization attribute may start
// while its predecessor in the stream is still running; pdl_wait() blocks until the predecessor has completed and
// its writes are visible, pdl_trigger() in the predecessor lets the dependent start early.  Both are no-ops for
// launches without the attribute.  Used so that the grid kernel's launch latency hides behind k_pod_prep, and the
// fence kernel's behind the grid kernel's last wave.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// One row's result.  With several output vectors (the all-gather) only the local one is written here -- scattered
// 4-byte stores over NVLink cost ~45 us per million rows, measured -- and k_gather_push forwards the finished slice.
__device__ __forceinline__ void store_best(const SelectArgs& a, uint32_t row, int32_t b) {
    if (a.self_out >= 0) { a.best_out[a.self_out][a.row0 + row] = b; return; }
    for (int o = 0; o < a.n_out; ++o) a.best_out[o][a.row0 + row] = b;
}

// ---------------------------------------------------------------------------------------------------------
// K0: per-row preparation
// ---------------------------------------------------------------------------------------------------------
// Rows are grouped so that a tile is homogeneous in (cloud view, "vcpu request constrains", "ram request
// constrains"): group g = cloud*4 + needs_vcpu*2 + needs_ram.  A request whose rank threshold is 0 is met by
// every offer, so tiles of such rows never read that column's masks (for pods without vcpu/ram requests this
// is literally the reference's predicate, which has no such columns: runpod_client.go:478).
__global__ void __launch_bounds__(1024) k_pod_prep(SelectArgs a) {
    pdl_trigger();  // the grid kernel may be scheduled now; it waits (pdl_wait) before it reads anything written here
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    const bool valid = p < a.P;
    uint32_t grp = kGroups;  // none
    if (valid) {
        uint8_t c = a.cloud ? a.cloud[p] : (uint8_t)RPK_CLOUD_SECURE;
        a.pos[p] = kNone;
        uint32_t need = 3;  // kernels without rank thresholds always test both extension columns
        if (a.pk.bits || a.pk.bm_words) {
            uint32_t tm = lower_bound_i32(a.distinct[0], a.D[0], a.req_mem[p]) + 1;
            uint32_t tv = lower_bound_i32(a.distinct[1], a.D[1], a.req_vcpu ? a.req_vcpu[p] : 0);
            uint32_t tr = lower_bound_i32(a.distinct[2], a.D[2], a.req_ram ? a.req_ram[p] : 0);
            need = (tv != 0 ? 2u : 0u) | (tr != 0 ? 1u : 0u);
            if (a.pk.bm_words)  // bit-sliced view: the three mask-word indices of this row inside a chunk
                a.rw[p] = (tm - 1) | ((a.pk.bm_off_vcpu + tv) << 8) | ((a.pk.bm_off_ram + tr) << 16);
            else
                a.rw[p] = (tm << a.pk.sh_mem) | (tv << a.pk.sh_vcpu) | (tr << a.pk.sh_ram);
        }
        if (c <= 1) {
            grp = c * 4u + need;
        } else {  // neither SECURE nor COMMUNITY -> nothing feasible (runpod_client.go:469-475)
            store_best(a, p, -1);
            if (a.top5) for (int k = 0; k < RPK_TOPK; ++k) a.top5[(size_t)p * RPK_TOPK + k] = -1;
        }
    }
    // one atomic per (group, 1024-row block): the counters are shared by the whole grid
    __shared__ uint32_t s_cnt[kGroups][32], s_base[kGroups];
    const uint32_t warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t my_mask = 0;
#pragma unroll
    for (uint32_t g = 0; g < kGroups; ++g) {
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, grp == g);
        if (grp == g) my_mask = m;
        if (lane == 0) s_cnt[g][warp] = __popc(m);
    }
    __syncthreads();
    if (warp < kGroups) {  // warp g scans group g's per-warp counts
        const uint32_t c = lane < nwarps ? s_cnt[warp][lane] : 0u;
        uint32_t inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, d); if ((int)lane >= d) inc += n; }
        if (lane < nwarps) s_cnt[warp][lane] = inc - c;
        if (lane == 31) s_base[warp] = inc ? atomicAdd(&a.counts[warp], inc) : 0u;
    }
    __syncthreads();
    if (grp < kGroups)
        a.order[(size_t)grp * a.P + s_base[grp] + s_cnt[grp][warp] + __popc(my_mask & ((1u << lane) - 1))] = p;
}

// ---------------------------------------------------------------------------------------------------------
// bulk async copy global -> shared with mbarrier completion (cp.async.bulk, SASS UBLKCP)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// row-tile decoding shared by both grid kernels
// ---------------------------------------------------------------------------------------------------------
struct TileInfo { int cloud; uint32_t group; uint32_t row_base; uint32_t nrows; uint32_t total_tiles; bool valid; };

__device__ __forceinline__ TileInfo decode_tile(const SelectArgs& a, uint32_t tile, uint32_t rpc) {
    TileInfo t;
    t.valid = false; t.cloud = 0; t.group = 0; t.row_base = 0; t.nrows = 0;
    uint32_t first = 0;
    // tiles are handed out heaviest group first (3 mask words per (row, chunk), then 2, then 1): the CTAs that
    // finish the grid are the light ones, so the tail of the last wave is short
    constexpr uint32_t kHeavyFirst[kGroups] = {3, 7, 1, 2, 5, 6, 0, 4};
#pragma unroll
    for (uint32_t i = 0; i < kGroups; ++i) {
        const uint32_t g = a.tune_natural_order ? i : kHeavyFirst[i];
        const uint32_t n = a.counts[g];
        const uint32_t tg = (n + rpc - 1) / rpc;
        if (!t.valid && tile < first + tg) {
            t.valid = true; t.group = g; t.cloud = (int)(g >> 2);
            t.row_base = (tile - first) * rpc;
            t.nrows = min(rpc, n - t.row_base);
        }
        first += tg;
    }
    t.total_tiles = first;
    return t;
}
__device__ __forceinline__ uint32_t tile_row(const SelectArgs& a, const TileInfo& t, uint32_t slot) {
    return a.order[(size_t)t.group * a.P + t.row_base + slot];
}

// Merge the segment result, and let the last CTA of the tile finish the rows: price < maxPrice on the
// winner (strict, runpod_client.go:478) and sorted position -> offer index.
__device__ __forceinline__ void finish_tile(const SelectArgs& a, const TileInfo& t, uint32_t tile, uint32_t S, int* s_last) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        uint32_t ticket = atomicAdd(&a.tile_ctr[tile], 1u);
        *s_last = (ticket == S - 1);
    }
    __syncthreads();
    if (!*s_last) return;
    __threadfence();
    for (uint32_t slot = threadIdx.x; slot < t.nrows; slot += blockDim.x) {
        const uint32_t row = tile_row(a, t, slot);
        const uint32_t p = __ldcg(a.pos + row);
        int32_t b = -1;
        if (p != kNone) {
            const double pr = a.view[t.cloud].price[p];
            const double mx = a.max_price ? a.max_price[row] : RPK_DEFAULT_MAX_PRICE;
            if (pr < mx) b = a.view[t.cloud].perm[p];
        }
        store_best(a, row, b);
    }
    if (a.n_out > 1 && a.self_out < 0) __threadfence_system();  // direct peer stores are performed system-wide before this grid completes
    // The counters clean themselves (no memset launch per call): this CTA is the last one to touch the tile's
    // ticket, and the CTA that finishes the last tile of the grid clears the group counts.  Every CTA of a valid
    // tile read the counts before it took its ticket; CTAs past the last tile that read them late see fewer
    // tiles, never more, and still return.
    if (threadIdx.x == 0) {
        a.tile_ctr[tile] = 0u;
        if (atomicAdd(a.done, 1u) + 1u == t.total_tiles) {
#pragma unroll
            for (uint32_t g = 0; g < kGroups; ++g) a.counts[g] = 0u;
            *a.done = 0u;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K1 (packed): one u32 per offer.
//   kPos = true  (layout bits <= 18): 2.5 instructions per offer-score -- IMAD.IADD d = o - r (FMA pipe),
//                LOP3 key = (~d & guard) | (d & pos_mask) (ALU), half a VIMNMX3 (ALU).  The surviving key is
//                the lowest feasible position of the segment, or >= 2^kPosBits when nothing is feasible.
//   kPos = false (wider layouts): sub + LOP3-with-predicate + select; positions walked in descending order
//                so that a lower one overwrites.
// ---------------------------------------------------------------------------------------------------------
template <int R, bool kPos>
__global__ void __launch_bounds__(kCtaThreads, 3) k_select_packed(SelectArgs a, uint32_t S, uint32_t seg_len) {
    extern __shared__ __align__(128) uint32_t s_off[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_last;
    constexpr uint32_t RPC = kWarpsPerCta * R;
    const uint32_t tile = blockIdx.x / S, seg = blockIdx.x - tile * S;
    pdl_wait();     // k_pod_prep (counts, order, rw, pos) has completed
    pdl_trigger();  // a dependent (the peer fence) may be scheduled; it waits for this grid to complete
    const TileInfo t = decode_tile(a, tile, RPC);
    if (!t.valid) return;
    const uint32_t Gc = ((a.G + kChunk - 1) / kChunk) * kChunk;
    const uint32_t g0 = seg * seg_len;
    const uint32_t len = g0 < Gc ? min(seg_len, Gc - g0) : 0;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&s_bar, len * 4u);
        if (len) bulk_g2s(s_off, a.view[t.cloud].packed + g0, len * 4u, &s_bar);
    }
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rw[R], best[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t slot = warp * R + r;
        rw[r] = slot < t.nrows ? a.rw[tile_row(a, t, slot)] : 0u;
        best[r] = kNone;
    }
    const uint32_t guard = a.pk.guard;
    mbar_wait(&s_bar, 0);
    const uint4* s4 = reinterpret_cast<const uint4*>(s_off);
    if (kPos) {
        constexpr uint32_t kLow = (1u << kPosBits) - 1;
        for (int ch = (int)(len / kChunk) - 1; ch >= 0; --ch) {
            const uint4 o = s4[ch * 32 + lane];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t d0 = o.x - rw[r], d1 = o.y - rw[r], d2 = o.z - rw[r], d3 = o.w - rw[r];
                const uint32_t k0 = (~d0 & guard) | (d0 & kLow), k1 = (~d1 & guard) | (d1 & kLow);
                const uint32_t k2 = (~d2 & guard) | (d2 & kLow), k3 = (~d3 & guard) | (d3 & kLow);
                best[r] = min(min(best[r], k0), k1);
                best[r] = min(min(best[r], k2), k3);
            }
        }
    } else {
        for (int ch = (int)(len / kChunk) - 1; ch >= 0; --ch) {
            const uint4 o = s4[ch * 32 + lane];
            const uint32_t j0 = g0 + (uint32_t)ch * kChunk + lane * 4u;
            // descending positions: a lower position overwrites, so the lowest feasible one survives
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if ((~(o.w - rw[r]) & guard) == 0) best[r] = j0 + 3;
                if ((~(o.z - rw[r]) & guard) == 0) best[r] = j0 + 2;
                if ((~(o.y - rw[r]) & guard) == 0) best[r] = j0 + 1;
                if ((~(o.x - rw[r]) & guard) == 0) best[r] = j0;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        uint32_t m = __reduce_min_sync(0xFFFFFFFFu, best[r]);
        if (kPos) m = m < (1u << kPosBits) ? g0 + m : kNone;  // segments are 2^kPosBits-aligned
        const uint32_t slot = warp * R + r;
        if (lane == 0 && slot < t.nrows && m != kNone) atomicMin(&a.pos[tile_row(a, t, slot)], m);
    }
    finish_tile(a, t, tile, S, &s_last);
}

// ---------------------------------------------------------------------------------------------------------
// K1 (bit-sliced): the offer side of a 32-offer chunk is three precomputed 32-bit masks (one per column,
// picked by the row's rank thresholds), so ONE LOP3 evaluates the feasibility of 32 (pod, offer) pairs:
//     m = mask_mem[chunk][t_mem] & mask_vcpu[chunk][t_vcpu] & mask_ram[chunk][t_ram]
// Every pair of the grid is still evaluated (bit j of m IS the predicate of runpod_client.go:478 minus the
// price bound for offer j) -- there is no early exit and no deduplication of equal rows.  Each lane owns RPL
// pod rows; all lanes walk the same chunk, reading their own three words of its 128-byte row (distinct words
// = distinct banks, equal words = broadcast).  Chunks are walked in descending price order so the last hit
// is the cheapest; its mask is re-read once at the end for the bit position.
// ---------------------------------------------------------------------------------------------------------
template <int RPL, int STRIDE, bool NEEDV, bool NEEDR>
__device__ __forceinline__ void bitmap_walk(const uint32_t* s_off, int n, const uint32_t (&o1)[RPL], const uint32_t (&o2)[RPL],
                                            const uint32_t (&o3)[RPL], uint32_t (&bc)[RPL]) {
    int ch = n - 1;
    for (; ch >= 3; ch -= 4) {
        const uint32_t* p = s_off + (size_t)(ch - 3) * STRIDE;
#pragma unroll
        for (int u = 3; u >= 0; --u) {
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                uint32_t m = p[u * STRIDE + o1[r]];
                if (NEEDV) m &= p[u * STRIDE + o2[r]];
                if (NEEDR) m &= p[u * STRIDE + o3[r]];
                if (m) bc[r] = (uint32_t)(ch - 3 + u);
            }
        }
    }
    for (; ch >= 0; --ch) {
        const uint32_t* p = s_off + (size_t)ch * STRIDE;
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            uint32_t m = p[o1[r]];
            if (NEEDV) m &= p[o2[r]];
            if (NEEDR) m &= p[o3[r]];
            if (m) bc[r] = (uint32_t)ch;
        }
    }
}

template <int RPL, int STRIDE>
__global__ void __launch_bounds__(kCtaThreads, 3) k_select_bitmap(SelectArgs a, uint32_t S, uint32_t seg_chunks) {
    constexpr uint32_t kBmStride = STRIDE;
    extern __shared__ __align__(128) uint32_t s_off[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_last;
    constexpr uint32_t RPC = kCtaThreads * RPL;
    const uint32_t tile = blockIdx.x / S, seg = blockIdx.x - tile * S;
    pdl_wait();     // k_pod_prep (counts, order, rw, pos) has completed
    pdl_trigger();  // a dependent (the peer fence) may be scheduled; it waits for this grid to complete
    const TileInfo t = decode_tile(a, tile, RPC);
    if (!t.valid) return;
    const uint32_t total_chunks = (a.G + 31) / 32;
    const uint32_t c0 = seg * seg_chunks;
    const uint32_t n = c0 < total_chunks ? min(seg_chunks, total_chunks - c0) : 0;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&s_bar, n * kBmStride * 4u);
        if (n) bulk_g2s(s_off, a.view[t.cloud].bitmap + (size_t)c0 * kBmStride, n * kBmStride * 4u, &s_bar);
    }
    uint32_t o1[RPL], o2[RPL], o3[RPL], bc[RPL];
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const uint32_t slot = (uint32_t)r * kCtaThreads + threadIdx.x;
        const uint32_t w = slot < t.nrows ? a.rw[tile_row(a, t, slot)] : 0u;
        o1[r] = w & 0xFFu; o2[r] = (w >> 8) & 0xFFu; o3[r] = (w >> 16) & 0xFFu;
        bc[r] = kNone;
    }
    mbar_wait(&s_bar, 0);
    // the tile's rows all share the two "column constrains" flags: masks of a non-constraining column are never read
    switch (t.group & 3u) {
        case 0: bitmap_walk<RPL, STRIDE, false, false>(s_off, (int)n, o1, o2, o3, bc); break;
        case 1: bitmap_walk<RPL, STRIDE, false, true>(s_off, (int)n, o1, o2, o3, bc); break;
        case 2: bitmap_walk<RPL, STRIDE, true, false>(s_off, (int)n, o1, o2, o3, bc); break;
        default: bitmap_walk<RPL, STRIDE, true, true>(s_off, (int)n, o1, o2, o3, bc); break;
    }
#pragma unroll
    for (int r = 0; r < RPL; ++r) {
        const uint32_t slot = (uint32_t)r * kCtaThreads + threadIdx.x;
        if (slot < t.nrows && bc[r] != kNone) {
            const uint32_t* p = s_off + (size_t)bc[r] * kBmStride;
            const uint32_t m = p[o1[r]] & p[o2[r]] & p[o3[r]];  // a non-constraining column's word is all-available: harmless here
            atomicMin(&a.pos[tile_row(a, t, slot)], (c0 + bc[r]) * 32u + (uint32_t)__ffs(m) - 1u);
        }
    }
    finish_tile(a, t, tile, S, &s_last);
}

// ---------------------------------------------------------------------------------------------------------
// K1 (bit-sliced, fused, small batches): one launch does what k_pod_prep + k_select_bitmap + the epilogue do,
// for batches too small to fill the GPU anyway (micro-batches of the streaming path, BASELINE config 2).  One
// lane per pod row in arrival order, no grouping, no atomics, no counters to clear: the CTA walks every segment
// of the view(s) its rows use, so each row's result is final when its loop ends.
// ---------------------------------------------------------------------------------------------------------
template <int STRIDE>
__global__ void __launch_bounds__(kCtaThreads) k_select_fused(SelectArgs a, uint32_t seg_chunks) {
    extern __shared__ __align__(128) uint32_t s_off[];
    __shared__ __align__(8) uint64_t s_bar;
    const uint32_t p = blockIdx.x * kCtaThreads + threadIdx.x;
    const bool valid = p < a.P;
    uint32_t cls = 2;
    uint32_t o1[1] = {0}, o2[1] = {0}, o3[1] = {0};
    bool needv = false, needr = false;
    if (valid) {
        const uint8_t c = a.cloud ? a.cloud[p] : (uint8_t)RPK_CLOUD_SECURE;
        cls = c <= 1 ? c : 2;
        const uint32_t tm = lower_bound_i32(a.distinct[0], a.D[0], a.req_mem[p]) + 1;
        const uint32_t tv = lower_bound_i32(a.distinct[1], a.D[1], a.req_vcpu ? a.req_vcpu[p] : 0);
        const uint32_t tr = lower_bound_i32(a.distinct[2], a.D[2], a.req_ram ? a.req_ram[p] : 0);
        o1[0] = tm - 1; o2[0] = a.pk.bm_off_vcpu + tv; o3[0] = a.pk.bm_off_ram + tr;
        needv = tv != 0; needr = tr != 0;
        a.rw[p] = o1[0] | (o2[0] << 8) | (o3[0] << 16);  // k_select_top5_bitmap reads it
        if (cls == 2 && a.top5) for (int k = 0; k < RPK_TOPK; ++k) a.top5[(size_t)p * RPK_TOPK + k] = -1;
    }
    const bool wv = __any_sync(0xFFFFFFFFu, needv), wr = __any_sync(0xFFFFFFFFu, needr);  // warp-uniform
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    const uint32_t total_chunks = (a.G + 31) / 32;
    const uint32_t S = total_chunks ? (total_chunks + seg_chunks - 1) / seg_chunks : 0;
    uint32_t gpos = kNone, loads = 0;
    for (uint32_t c = 0; c < 2; ++c) {
        if (!__syncthreads_or(cls == c)) continue;  // no row of this CTA uses view c
        for (int seg = (int)S - 1; seg >= 0; --seg) {  // descending price: the last hit is the cheapest
            const uint32_t c0 = (uint32_t)seg * seg_chunks;
            const uint32_t n = min(seg_chunks, total_chunks - c0);
            if (threadIdx.x == 0) {
                mbar_expect_tx(&s_bar, n * STRIDE * 4u);
                bulk_g2s(s_off, a.view[c].bitmap + (size_t)c0 * STRIDE, n * STRIDE * 4u, &s_bar);
            }
            mbar_wait(&s_bar, loads & 1);
            ++loads;
            if (cls == c) {
                uint32_t bc[1] = {kNone};
                switch ((wv ? 2 : 0) | (wr ? 1 : 0)) {
                    case 0: bitmap_walk<1, STRIDE, false, false>(s_off, (int)n, o1, o2, o3, bc); break;
                    case 1: bitmap_walk<1, STRIDE, false, true>(s_off, (int)n, o1, o2, o3, bc); break;
                    case 2: bitmap_walk<1, STRIDE, true, false>(s_off, (int)n, o1, o2, o3, bc); break;
                    default: bitmap_walk<1, STRIDE, true, true>(s_off, (int)n, o1, o2, o3, bc); break;
                }
                if (bc[0] != kNone) {
                    const uint32_t* q = s_off + (size_t)bc[0] * STRIDE;
                    const uint32_t m = q[o1[0]] & q[o2[0]] & q[o3[0]];
                    gpos = (c0 + bc[0]) * 32u + (uint32_t)__ffs(m) - 1u;
                }
            }
            __syncthreads();  // everyone is done with the segment before the next bulk copy overwrites it
        }
    }
    if (!valid) return;
    int32_t b = -1;
    if (cls < 2 && gpos != kNone) {  // winner's price < maxPrice (strict, runpod_client.go:478); the bound is a prefix of the order
        const double pr = a.view[cls].price[gpos];
        const double mx = a.max_price ? a.max_price[p] : RPK_DEFAULT_MAX_PRICE;
        if (pr < mx) b = a.view[cls].perm[gpos];
    }
    store_best(a, p, b);
}

// ---------------------------------------------------------------------------------------------------------
// K1 (generic): full-range int32 columns, one int4 per offer, 4 instructions per offer-score
// ---------------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(kCtaThreads, 2) k_select_wide(SelectArgs a, uint32_t S, uint32_t seg_len) {
    extern __shared__ __align__(128) uint32_t s_off[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_last;
    constexpr uint32_t RPC = kWarpsPerCta * R;
    const uint32_t tile = blockIdx.x / S, seg = blockIdx.x - tile * S;
    pdl_wait();     // k_pod_prep (counts, order, rw, pos) has completed
    pdl_trigger();  // a dependent (the peer fence) may be scheduled; it waits for this grid to complete
    const TileInfo t = decode_tile(a, tile, RPC);
    if (!t.valid) return;
    const uint32_t Gc = ((a.G + kChunk - 1) / kChunk) * kChunk;
    const uint32_t g0 = seg * seg_len;
    const uint32_t len = g0 < Gc ? min(seg_len, Gc - g0) : 0;
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&s_bar, len * 16u);
        if (len) bulk_g2s(s_off, a.view[t.cloud].wide + g0, len * 16u, &s_bar);
    }
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int32_t qm[R], qv[R], qr[R];
    uint32_t best[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t slot = warp * R + r;
        const uint32_t row = slot < t.nrows ? tile_row(a, t, slot) : 0u;
        const bool ok = slot < t.nrows;
        qm[r] = ok ? a.req_mem[row] : INT32_MAX;
        qv[r] = ok && a.req_vcpu ? a.req_vcpu[row] : 0;
        qr[r] = ok && a.req_ram ? a.req_ram[row] : 0;
        best[r] = kNone;
    }
    mbar_wait(&s_bar, 0);
    const int4* s4 = reinterpret_cast<const int4*>(s_off);
    for (int ch = (int)(len / 32u) - 1; ch >= 0; --ch) {
        const int4 o = s4[ch * 32 + lane];
        const uint32_t j = g0 + (uint32_t)ch * 32u + lane;
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (o.x >= qm[r] && o.y >= qv[r] && o.z >= qr[r]) best[r] = j;  // mem/vcpu/ram >= request (:478, non-strict)
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint32_t m = __reduce_min_sync(0xFFFFFFFFu, best[r]);
        const uint32_t slot = warp * R + r;
        if (lane == 0 && slot < t.nrows && m != kNone) atomicMin(&a.pos[tile_row(a, t, slot)], m);
    }
    finish_tile(a, t, tile, S, &s_last);
}

// ---------------------------------------------------------------------------------------------------------
// top-5: "Take up to 5 GPUs" (runpod_client.go:502-509).  One warp per pod row walks the sorted view in
// ascending price order, ballots feasibility over 32 positions at a time and stops after five hits or at
// the first position whose price is not < maxPrice (everything after it is at least as expensive).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_select_top5(SelectArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t Gc = ((a.G + kChunk - 1) / kChunk) * kChunk;
    for (uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; p < a.P; p += nwarps) {
        const uint8_t c = a.cloud ? a.cloud[p] : (uint8_t)RPK_CLOUD_SECURE;
        if (c > 1) continue;  // filled by k_pod_prep
        const int32_t qm = a.req_mem[p], qv = a.req_vcpu ? a.req_vcpu[p] : 0, qr = a.req_ram ? a.req_ram[p] : 0;
        const double mx = a.max_price ? a.max_price[p] : RPK_DEFAULT_MAX_PRICE;
        const int4* __restrict__ wide = a.view[c].wide;
        const double* __restrict__ price = a.view[c].price;
        int cnt = 0;
        for (uint32_t base = 0; base < Gc && cnt < RPK_TOPK; base += 32) {
            const int4 o = wide[base + lane];
            const bool price_ok = price[base + lane] < mx;  // NaN (unavailable / padding) fails
            const bool f = price_ok && o.x >= qm && o.y >= qv && o.z >= qr;
            uint32_t mask = __ballot_sync(0xFFFFFFFFu, f);
            const bool stop = __ballot_sync(0xFFFFFFFFu, !price_ok) != 0;
            while (mask && cnt < RPK_TOPK) {
                const int b = __ffs(mask) - 1;
                const int idx = __shfl_sync(0xFFFFFFFFu, o.w, b);
                if (lane == 0) a.top5[(size_t)p * RPK_TOPK + cnt] = idx;
                ++cnt;
                mask &= mask - 1;
            }
            if (stop) break;
        }
        if (lane == 0) for (int k = cnt; k < RPK_TOPK; ++k) a.top5[(size_t)p * RPK_TOPK + k] = -1;
    }
}

// top-5 on the bit-sliced view: one lane per pod row.  The price bound is a prefix of the sorted view (binary
// search for its length), inside it the row's three masks give 32 positions per step; the walk stops at five hits.
template <int STRIDE>
__global__ void __launch_bounds__(256) k_select_top5_bitmap(SelectArgs a) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.P) return;
    const uint8_t c = a.cloud ? a.cloud[p] : (uint8_t)RPK_CLOUD_SECURE;
    if (c > 1) return;  // filled by k_pod_prep
    const uint32_t w = a.rw[p];
    const uint32_t o1 = w & 0xFFu, o2 = (w >> 8) & 0xFFu, o3 = (w >> 16) & 0xFFu;
    const double mx = a.max_price ? a.max_price[p] : RPK_DEFAULT_MAX_PRICE;
    const double* __restrict__ price = a.view[c].price;
    const uint32_t* __restrict__ bm = a.view[c].bitmap;
    const int32_t* __restrict__ perm = a.view[c].perm;
    uint32_t lo = 0, hi = a.G;  // number of positions with price < mx (NaN = unavailable sorts last and fails)
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (price[mid] < mx) lo = mid + 1; else hi = mid; }
    const uint32_t cut = lo;
    int cnt = 0;
    int32_t* out = a.top5 + (size_t)p * RPK_TOPK;
    for (uint32_t base = 0; base < cut && cnt < RPK_TOPK; base += 32) {
        const uint32_t* row = bm + (size_t)(base >> 5) * STRIDE;
        uint32_t m = row[o1] & row[o2] & row[o3];
        if (cut - base < 32) m &= (1u << (cut - base)) - 1u;
        while (m && cnt < RPK_TOPK) {
            out[cnt++] = perm[base + (uint32_t)__ffs(m) - 1u];
            m &= m - 1;
        }
    }
    for (; cnt < RPK_TOPK; ++cnt) out[cnt] = -1;
}

// ---------------------------------------------------------------------------------------------------------
// the all-gather: push this shard's finished slice [row0, row0 + P) of the local vector into every peer's vector
// with 16-byte stores over NVLink (grid = CTAs per peer x peers).  Launched behind the select kernels with PDL.
// ---------------------------------------------------------------------------------------------------------
struct PushArgs { const int32_t* src; int32_t* dst[RPK_MAX_GPUS]; uint32_t count; };

__global__ void __launch_bounds__(256) k_gather_push(PushArgs a) {
    pdl_wait();     // the select kernels have completed: the local slice is final
    pdl_trigger();  // the peer fence may be scheduled behind this grid
    const int32_t* src = a.src;  // read with ld.global.cg: written by the grid this kernel overlaps with (never the .nc path)
    int32_t* dst = a.dst[blockIdx.y];
    // all vectors share the slice offset and 256-byte aligned bases, so one head/body/tail split fits source and peers
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(src) & 15u) >> 2);
    const uint32_t head = min(a.count, (4u - mis) & 3u);
    const uint32_t nvec = (a.count - head) >> 2;
    const uint32_t tail0 = head + (nvec << 2);
    if (blockIdx.x == 0) {
        if (threadIdx.x < head) dst[threadIdx.x] = __ldcg(src + threadIdx.x);
        if (threadIdx.x < a.count - tail0) dst[tail0 + threadIdx.x] = __ldcg(src + tail0 + threadIdx.x);
    }
    const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    // four 16-byte loads in flight per thread before the first store: the copy is latency-bound, not bandwidth-bound
    constexpr int kU = 4;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += stride * kU) {
        uint4 v[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) { const uint32_t i = i0 + (uint32_t)u * stride; if (i < nvec) v[u] = __ldcg(s4 + i); }
#pragma unroll
        for (int u = 0; u < kU; ++u) { const uint32_t i = i0 + (uint32_t)u * stride; if (i < nvec) d4[i] = v[u]; }
    }
    __threadfence_system();  // this thread's peer stores are performed system-wide before the grid completes
}

// ---------------------------------------------------------------------------------------------------------
// cross-GPU fence after the fused gather: signal every peer, then wait for every peer (one warp)
// ---------------------------------------------------------------------------------------------------------
constexpr int kFenceCounterWord = 32;  // word of a rank's own flag array that counts its fences (epoch = 0 mode)

__global__ void k_peer_fence(PeerFenceArgs a) {
    const int r = (int)threadIdx.x;
    pdl_wait();  // launched early behind the select grid (PDL): its peer stores are complete from here on
    uint32_t epoch = a.epoch;
    if (epoch == 0) {  // self-counting: the launch carries no per-call value, so it can be captured in a CUDA graph and replayed
        uint32_t e = 0;
        if (r == 0) { e = a.flags[a.my_rank][kFenceCounterWord] + 1u; a.flags[a.my_rank][kFenceCounterWord] = e; }
        epoch = __shfl_sync(0xFFFFFFFFu, e, 0);
    }
    if (r < a.n) {
        __threadfence_system();  // everything this stream did before (the epilogue's NVLink stores) is ordered before the flag
        *reinterpret_cast<volatile uint32_t*>(a.flags[r] + a.my_rank) = epoch;
        const volatile uint32_t* mine = a.flags[a.my_rank] + r;
        while ((int32_t)(*mine - epoch) < 0) __nanosleep(64);
        __threadfence_system();
    }
}


// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
// RPK_TUNE="rpl=<1|2|4>,order=natural,seg=full,segmul=<1..8>,pdl=off" -- measurement hooks for tools/k1_tune.py (read per call, so one
// process can sweep them); unset = the defaults chosen from those measurements.
struct Tune { int rpl = 0; int segmul = 1; bool natural_order = false; bool full_segments = false; bool pdl = true; bool grid_kernel = false; };
// The environment is read ONCE per process (this sits on the 38 us latency path of production callers); measurement
// tools that sweep settings inside one process set RPK_TUNE_RELOAD=1 before the first call.
static bool tune_reload() { static const bool r = getenv("RPK_TUNE_RELOAD") != nullptr; return r; }
static Tune parse_tune();
static Tune read_tune() {
    if (tune_reload()) return parse_tune();
    static const Tune cached = parse_tune();
    return cached;
}
static Tune parse_tune() {
    Tune t;
    const char* e = getenv("RPK_TUNE");
    if (!e) return t;
    if (const char* p = strstr(e, "rpl=")) t.rpl = atoi(p + 4);
    t.natural_order = strstr(e, "order=natural") != nullptr;
    t.full_segments = strstr(e, "seg=full") != nullptr;
    t.pdl = strstr(e, "pdl=off") == nullptr;
    t.grid_kernel = strstr(e, "k1=grid") != nullptr;  // the first-generation (tile x segment) bit-sliced kernel, for A/B runs
    if (const char* p = strstr(e, "segmul=")) { t.segmul = atoi(p + 7); if (t.segmul < 1 || t.segmul > 8) t.segmul = 1; }
    return t;
}

uint32_t select_tiles_max(uint32_t P, int rows_per_warp) {
    const uint32_t rpc = (uint32_t)(kWarpsPerCta * rows_per_warp);  // bit-sliced kernel: rows_per_warp = 32 * RPL
    return (P + rpc - 1) / rpc + kGroups;  // every row group can end in one partial tile
}

int pick_rows_per_lane(uint32_t P, uint32_t G, int sm_count) {  // bit-sliced kernel
    // CTAs = row tiles x offer segments.  Measured on B200 (tools/k1_tune.py, G = 100k, us per select):
    //   P = 125k: 90.8 / 93.2 / 97.6   P = 250k: 160.6 / 162.5 / 169.0   P = 1M: 581.5 / 579.7 / 594.0   (1 / 2 / 4 rows per lane)
    // -- many small CTAs beat few large ones until the grid is tens of waves deep (the tail of the last wave
    // outweighs the extra segment loads), and 4 rows per lane never pays.
    const int forced = read_tune().rpl;
    if (forced == 1 || forced == 2 || forced == 4) return forced;
    const uint64_t want = (uint64_t)sm_count * 3 * 24;  // >= 24 waves of 3 CTAs/SM at 2 rows per lane
    const uint64_t total_chunks = (G + 31) / 32;
    const uint64_t kBmSegChunks = kBmSegBytes / (32 * 4);
    uint64_t S = (total_chunks + kBmSegChunks - 1) / kBmSegChunks;
    if (S == 0) S = 1;
    if ((P + (uint64_t)kCtaThreads * 2 - 1) / ((uint64_t)kCtaThreads * 2) * S >= want) return 2;
    return 1;
}

template <int RPL, int STRIDE>
static void launch_bitmap(const SelectArgs& a, cudaStream_t st) {
    const uint32_t tiles = select_tiles_max(a.P, 32 * RPL);
    const uint32_t total_chunks = (a.G + 31) / 32;
    constexpr uint32_t kBmSegChunks = kBmSegBytes / (STRIDE * 4);
    const Tune tune = read_tune();
    uint32_t S = (total_chunks + kBmSegChunks - 1) / kBmSegChunks * (uint32_t)tune.segmul;
    if (S == 0) S = 1;
    // equal segments (ceil(chunks / S) each) instead of S-1 full ones and a short one: every CTA of a row group
    // then costs the same, and the stage is no larger than the table needs
    uint32_t seg_chunks = tune.full_segments ? kBmSegChunks : (total_chunks + S - 1) / S;
    if (seg_chunks == 0) seg_chunks = 1;
    const size_t smem = (size_t)seg_chunks * STRIDE * 4;
    static thread_local int attr_dev = -1;  // the attribute is per device: set it once per (thread, device)
    int dev = 0;
    RPK_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        RPK_CUDA(cudaFuncSetAttribute(k_select_bitmap<RPL, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBmSegBytes));
        attr_dev = dev;
    }
    launch_pdl(k_select_bitmap<RPL, STRIDE>, dim3(tiles * S), dim3(kCtaThreads), smem, st, tune.pdl, a, S, seg_chunks);
}

static void seg_plan(uint32_t G, uint32_t seg_cap, uint32_t* S, uint32_t* seg_len) {
    const uint32_t Gc = ((G + kChunk - 1) / kChunk) * kChunk;
    uint32_t s = (Gc + seg_cap - 1) / seg_cap;
    if (s == 0) s = 1;
    uint32_t len = ((Gc + s - 1) / s + kChunk - 1) / kChunk * kChunk;
    *S = s; *seg_len = len;
}

int pick_rows_per_warp(uint32_t P, int sm_count) {
    // enough CTAs for ~4 waves of 3 CTAs/SM, else fewer rows per warp
    const uint64_t want = (uint64_t)sm_count * 3 * 4;
    for (int r = 16; r > 1; r >>= 1) {
        const uint64_t tiles = (P + (uint64_t)kWarpsPerCta * r - 1) / ((uint64_t)kWarpsPerCta * r);
        if (tiles >= want) return r;
    }
    return 1;
}

// Launch with the programmatic-serialization attribute (see pdl_wait): the kernel may start before its stream
// predecessor has finished and synchronises with griddepcontrol.wait itself.
template <typename... KArgs, typename... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    RPK_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device): set it once per (thread, device), so a
// warmed-up launch path issues nothing but launches (stream capture)
template <typename K>
static void allow_smem(K kernel, int bytes, int* attr_dev) {
    int dev = 0;
    RPK_CUDA(cudaGetDevice(&dev));
    if (*attr_dev == dev) return;
    RPK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    *attr_dev = dev;
}

template <int R>
static void launch_grid(const SelectArgs& a, cudaStream_t st) {
    static thread_local int dev_pos = -1, dev_packed = -1, dev_wide = -1;
    const bool pdl = read_tune().pdl;
    uint32_t S, seg_len;
    const uint32_t tiles = select_tiles_max(a.P, R);
    if (a.pk.bits && a.pk.pos_bits) {
        // fixed 2^kPosBits-offer segments: the embedded positions are segment-local
        const uint32_t Gc = ((a.G + kChunk - 1) / kChunk) * kChunk;
        S = (Gc + kSegPacked - 1) / kSegPacked; if (S == 0) S = 1;
        seg_len = kSegPacked;
        allow_smem(k_select_packed<R, true>, (int)(kSegPacked * 4), &dev_pos);
        launch_pdl(k_select_packed<R, true>, dim3(tiles * S), dim3(kCtaThreads), (size_t)kSegPacked * 4, st, pdl, a, S, seg_len);
    } else if (a.pk.bits) {
        seg_plan(a.G, kSegPacked, &S, &seg_len);
        const size_t smem = (size_t)seg_len * 4;
        allow_smem(k_select_packed<R, false>, (int)(kSegPacked * 4), &dev_packed);
        launch_pdl(k_select_packed<R, false>, dim3(tiles * S), dim3(kCtaThreads), smem, st, pdl, a, S, seg_len);
    } else {
        seg_plan(a.G, kSegWide, &S, &seg_len);
        const size_t smem = (size_t)seg_len * 16;
        allow_smem(k_select_wide<R>, (int)(kSegWide * 16), &dev_wide);
        launch_pdl(k_select_wide<R>, dim3(tiles * S), dim3(kCtaThreads), smem, st, pdl, a, S, seg_len);
    }
}

constexpr uint32_t kFusedMaxRows = kFusedRowsMax;  // below this a batch cannot fill the GPU: one fused launch beats three

template <int STRIDE>
static void launch_fused(const SelectArgs& a, cudaStream_t st) {
    constexpr uint32_t kSegChunks = kBmSegBytes / (STRIDE * 4);
    static thread_local int attr_dev = -1;
    int dev = 0;
    RPK_CUDA(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        RPK_CUDA(cudaFuncSetAttribute(k_select_fused<STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBmSegBytes));
        attr_dev = dev;
    }
    // small tables need a small stage: size the dynamic shared memory to the table
    const uint32_t total_chunks = (a.G + 31) / 32;
    const uint32_t chunks = total_chunks < kSegChunks ? (total_chunks ? total_chunks : 1) : kSegChunks;
    k_select_fused<STRIDE><<<(a.P + kCtaThreads - 1) / kCtaThreads, kCtaThreads, (size_t)chunks * STRIDE * 4, st>>>(a, kSegChunks);
}

static int launch_select_kernels(const SelectArgs& a, int R, const PersistPlan* pl, cudaStream_t st, bool* fused_push);

int launch_select(const SelectArgs& args, int R, const PersistPlan* pl, cudaStream_t st) {
    SelectArgs a = args;
    PeerFenceArgs fa{};
    for (int r = 0; r < a.n_flags; ++r) fa.flags[r] = a.flags[r];
    fa.n = a.n_flags; fa.my_rank = a.my_rank; fa.epoch = a.inline_wait ? 1u : 0u;  // k_peer_signal: non-zero = wait here too
    if (args.P == 0) return a.n_flags > 0 ? launch_peer_signal(fa, 0, st) : 0;  // an empty shard still tells its peers it is done
    a.tune_natural_order = read_tune().natural_order ? 1u : 0u;
    if (a.n_out == 1) a.self_out = 0;
    bool fused_push = false;  // the persistent kernel pushes (and signals) itself
    int launches = launch_select_kernels(a, R, pl, st, &fused_push);
    if (fused_push) return launches;
    if (a.n_out > 1 && a.self_out >= 0) {  // forward the finished slice to the peers (the all-gather)
        PushArgs pa{};
        pa.src = a.best_out[a.self_out] + a.row0;
        int np = 0;
        for (int o = 0; o < a.n_out; ++o) if (o != a.self_out) pa.dst[np++] = a.best_out[o] + a.row0;
        pa.count = a.P;
        const uint32_t nvec = a.P / 4;
        uint32_t per_peer = (nvec + 1023) / 1024;  // one pass of four 16-byte units per thread, up to two CTAs per SM
        per_peer = per_peer < 1 ? 1 : per_peer > 296 ? 296 : per_peer;
        launch_pdl(k_gather_push, dim3(per_peer, (unsigned)np), dim3(256), 0, st, read_tune().pdl, pa);
        ++launches;
    }
    if (a.n_flags > 0) launches += launch_peer_signal(fa, 0, st);
    return launches;
}

static int launch_select_kernels(const SelectArgs& a, int R, const PersistPlan* pl, cudaStream_t st, bool* fused_push) {
    int launches = 0;
    if (a.pk.bm_words && a.P <= kFusedMaxRows && !a.pk.no_fused) {
        const bool wide_rows = a.pk.bm_stride == 64;
        if (wide_rows) launch_fused<64>(a, st); else launch_fused<32>(a, st);
        ++launches;
        if (a.top5) {
            if (wide_rows) k_select_top5_bitmap<64><<<(a.P + 255) / 256, 256, 0, st>>>(a);
            else k_select_top5_bitmap<32><<<(a.P + 255) / 256, 256, 0, st>>>(a);
            ++launches;
        }
        RPK_CUDA(cudaGetLastError());
        return launches;
    }
    if (a.pk.bm_words && pl && !read_tune().grid_kernel) {  // persistent kernel on the transposed view: sorts, selects, pushes, signals
        launches += launch_select_persist(a, *pl, st);
        *fused_push = true;
        if (a.top5) {
            if (a.pk.bm_stride == 64) k_select_top5_bitmap<64><<<(a.P + 255) / 256, 256, 0, st>>>(a);
            else k_select_top5_bitmap<32><<<(a.P + 255) / 256, 256, 0, st>>>(a);
            ++launches;
        }
        RPK_CUDA(cudaGetLastError());
        return launches;
    }
    // counts / done / tile tickets are zero here: zeroed when allocated, and every grid kernel leaves them zero
    k_pod_prep<<<(a.P + 1023) / 1024, 1024, 0, st>>>(a); ++launches;
    if (a.pk.bm_words) {  // R = 32 * rows-per-lane
        const bool wide_rows = a.pk.bm_stride == 64;
        switch (R / 32) {
            case 4: if (wide_rows) launch_bitmap<4, 64>(a, st); else launch_bitmap<4, 32>(a, st); break;
            case 2: if (wide_rows) launch_bitmap<2, 64>(a, st); else launch_bitmap<2, 32>(a, st); break;
            default: if (wide_rows) launch_bitmap<1, 64>(a, st); else launch_bitmap<1, 32>(a, st); break;
        }
        ++launches;
        if (a.top5) {
            if (wide_rows) k_select_top5_bitmap<64><<<(a.P + 255) / 256, 256, 0, st>>>(a);
            else k_select_top5_bitmap<32><<<(a.P + 255) / 256, 256, 0, st>>>(a);
            ++launches;
        }
        RPK_CUDA(cudaGetLastError());
        return launches;
    }
    switch (R) {
        case 16: launch_grid<16>(a, st); break;
        case 8: launch_grid<8>(a, st); break;
        case 4: launch_grid<4>(a, st); break;
        case 2: launch_grid<2>(a, st); break;
        default: launch_grid<1>(a, st); break;
    }
    ++launches;
    if (a.top5) {
        const uint32_t blocks = (uint32_t)min((uint64_t)(a.P + 7) / 8, (uint64_t)148 * 32);
        k_select_top5<<<blocks, 256, 0, st>>>(a); ++launches;
    }
    RPK_CUDA(cudaGetLastError());
    return launches;
}

int launch_peer_fence(const PeerFenceArgs& a, cudaStream_t st) {
    launch_pdl(k_peer_fence, dim3(1), dim3(32), 0, st, read_tune().pdl, a);
    RPK_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace rpk
