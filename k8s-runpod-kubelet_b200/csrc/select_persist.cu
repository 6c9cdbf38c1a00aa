// select_persist.cu -- the P x G grid on the bit-sliced view, second generation (kernel kind 4, batches above the
// fused kernel's size).  Same predicate and tie rule as select.cu (reference runpod_client.go:465-509); what changes
// is how the mask words reach the lanes and who owns a stage of the offer table:
//
//   k_pod_classify   row -> rank thresholds -> class key (cloud, vcpu threshold, ram threshold, mem threshold); one wave of
//                    CTAs, each looping over 512-row chunks: shared-memory histogram -> global class counts; fills
//                    pos[] = none; block 0 resets the queue state of the grid kernel.
//   k_pod_scatter    counting-sort scatter, same CTA shape: every CTA scans the class counts itself, reserves its rows
//                    per class with one atomic per non-empty class and writes ONE {row, thresholds} pair per row in
//                    class order (ord_rw[]); block 0 publishes per-cloud row / work totals.  Rows of neither cloud
//                    are final here (-1, runpod_client.go:469-475).
//   k_select_persist persistent CTAs (512 threads, as many per SM as the stage allows).  A CTA loads ONE stage of one
//                    cloud view -- `per` 64-chunk sub-ranges of the transposed view, 8 bulk async copies on 8 mbarriers,
//                    so the first items start before the whole stage has landed -- and keeps it for the whole call.
//                    Its warps pull items (64 consecutive rows of the class order x a range of sub-ranges) from the
//                    stage's queue with one atomic each.  A lane owns RPL rows and reads its rows' threshold words for
//                    FOUR chunks with one LDS.128 (the view is [threshold][chunk]-major); rows of a warp share their
//                    thresholds after the sort, so the 32 lanes read at most a few distinct 16-byte addresses and the
//                    load costs 2 LSU cycles instead of the 4 that four LDS.32 cost (tools/microbench/lds_probe.cu:
//                    LDS.128 with <= 8 distinct addresses = 2.0 cycles/warp, LDS.32 = 1.0).  Every (row, chunk) mask is
//                    still loaded and ANDed -- bit j of it IS the predicate of runpod_client.go:478 (minus the price
//                    bound) for offer j; eight chunks are OR-folded before the "any hit" test, the cheapest hit's block
//                    is re-read once at the end for the bit position.
//                    Segments merge with atomicMin; the warp that takes a row block's last ticket applies
//                    price < maxPrice to the winner (strict, :478; the bound is a prefix of the price order) and stores
//                    the offer index.
//   fused all-gather every retired row block adds its rows to one counter; a CTA that has run out of items waits until the
//                    counter reaches the row count (one polling thread per CTA), then EVERY warp of the grid copies an
//                    equal share of the slice to every peer with 16-byte NVLink stores; the CTA that finishes last
//                    signals the peers' flag words (rpk_peer_bind) and, with rpk_peer_inline_wait, waits for theirs.
//                    No drain, no copy kernel, no signal kernel, no wait kernel.
#include <math_constants.h>

#include <cstdlib>
#include <cstring>

#include "rpk_device.cuh"
#include "rpk_internal.cuh"

namespace rpk {
using namespace dev;

constexpr int kPThreads = 512;
constexpr int kPWarps = kPThreads / 32;
constexpr int kCopyGroups = 8;
constexpr uint32_t kSubWords = kSubStride;

// ---------------------------------------------------------------------------------------------------------
// row classes
// ---------------------------------------------------------------------------------------------------------
struct ClassDims { uint32_t Dm1, Dv1, Dr1, use_mem, C; };
__host__ __device__ __forceinline__ ClassDims class_dims(const uint32_t (&D)[3]) {
    ClassDims cd;
    cd.Dm1 = D[0] + 1; cd.Dv1 = D[1] + 1; cd.Dr1 = D[2] + 1;
    // bm_words = Dm1 + Dv1 + Dr1 <= 64, so 2 * Dv1 * Dr1 <= 1922 always fits; the mem threshold joins the key when it fits too
    cd.use_mem = 2u * cd.Dm1 * cd.Dv1 * cd.Dr1 <= kMaxClasses ? 1u : 0u;
    cd.C = 2u * cd.Dv1 * cd.Dr1 * (cd.use_mem ? cd.Dm1 : 1u);
    return cd;
}
// Class order = processing order: cloud, then vcpu threshold DEscending, ram threshold descending (rows whose requests
// constrain nothing -- one mask word per chunk -- come last, so the warps that finish the grid are the light ones),
// then the mem threshold.
__device__ __forceinline__ uint32_t class_key(const ClassDims& cd, uint32_t c, uint32_t tm1, uint32_t tv, uint32_t tr) {
    uint32_t k = (c * cd.Dv1 + (cd.Dv1 - 1 - tv)) * cd.Dr1 + (cd.Dr1 - 1 - tr);
    return cd.use_mem ? k * cd.Dm1 + tm1 : k;
}
__device__ __forceinline__ uint32_t class_weight(const ClassDims& cd, uint32_t key) {  // mask words per (row, chunk)
    const uint32_t k = cd.use_mem ? key / cd.Dm1 : key;
    const uint32_t tr_rev = k % cd.Dr1, tv_rev = (k / cd.Dr1) % cd.Dv1;
    return 1u + (tv_rev != cd.Dv1 - 1 ? 1u : 0u) + (tr_rev != cd.Dr1 - 1 ? 1u : 0u);
}

__device__ __forceinline__ uint32_t lb_smem(const int32_t* a, uint32_t n, int32_t x) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}

// ---------------------------------------------------------------------------------------------------------
// fused all-gather: push-block accounting.  Called by whole warps; `row` / `valid` per lane.
// ---------------------------------------------------------------------------------------------------------
// copy this warp's share of the local slice [row0, row0 + P) into every peer's vector.  The slice's 16-byte body is
// split evenly over ALL warps of the grid (a 125k-row slice is ~7 units per warp: one load, then one store per peer),
// so the copy has no hand-out counter and its latency is one L2 load + one NVLink store per lane.
__device__ __forceinline__ void push_share(const SelectArgs& a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t wpc = blockDim.x >> 5, gw = blockIdx.x * wpc + (threadIdx.x >> 5), nw = gridDim.x * wpc;
    const int32_t* src = a.best_out[a.self_out];
    const uint32_t g_lo = a.row0, n = a.P;
    // all vectors share the slice's 16-byte phase (checked on the host), so one head / body / tail split fits all
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(src + g_lo) & 15u) >> 2);
    const uint32_t head = min(n, (4u - mis) & 3u);
    const uint32_t nvec = (n - head) >> 2;
    const uint32_t tail0 = head + (nvec << 2);
    if (gw == 0) {  // the <= 3 + 3 elements off the 16-byte grid
        int32_t hv = 0, tv = 0;
        if (lane < head) hv = __ldcg(src + g_lo + lane);
        if (lane < n - tail0) tv = __ldcg(src + g_lo + tail0 + lane);
        for (int o = 0; o < a.n_out; ++o) {
            if (o == a.self_out) continue;
            if (lane < head) a.best_out[o][g_lo + lane] = hv;
            if (lane < n - tail0) a.best_out[o][g_lo + tail0 + lane] = tv;
        }
    }
    const uint32_t per = (nvec + nw - 1) / nw;
    const uint32_t v_lo = min(nvec, gw * per), v_hi = min(nvec, v_lo + per);
    const uint4* s4 = reinterpret_cast<const uint4*>(src + g_lo + head);
    for (uint32_t i0 = v_lo; i0 < v_hi; i0 += 32 * 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const uint32_t i = i0 + (uint32_t)u * 32 + lane; if (i < v_hi) v[u] = __ldcg(s4 + i); }
        for (int o = 0; o < a.n_out; ++o) {
            if (o == a.self_out) continue;
            uint4* d4 = reinterpret_cast<uint4*>(a.best_out[o] + g_lo + head);
#pragma unroll
            for (int u = 0; u < 4; ++u) { const uint32_t i = i0 + (uint32_t)u * 32 + lane; if (i < v_hi) d4[i] = v[u]; }
        }
    }
}

// first warp of the CTA that pushed last: every peer store of this grid is performed system-wide (each pusher fenced
// before its CTA counted itself), so the flags may go out
__device__ __forceinline__ void signal_peers(const SelectArgs& a, uint32_t word0, uint32_t counter_word) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t e = 0;
    if (lane == 0) { e = a.flags[a.my_rank][counter_word] + 1u; a.flags[a.my_rank][counter_word] = e; }
    e = __shfl_sync(0xFFFFFFFFu, e, 0);
    __threadfence_system();
    if ((int)lane < a.n_flags) {
        *reinterpret_cast<volatile uint32_t*>(a.flags[lane] + word0 + a.my_rank) = e;
        if (a.inline_wait) {  // the other half of the fence, without a kernel of its own: this warp is the grid's last worker anyway
            const volatile uint32_t* mine = a.flags[a.my_rank] + word0 + lane;
            while ((int32_t)(*mine - e) < 0) __nanosleep(32);
            __threadfence_system();
        }
    }
}

constexpr uint32_t kFlagSelect = 0;       // words [0, 8): select epochs per source rank
constexpr uint32_t kFlagStatus = 8;       // words [8, 16): status-sweep epochs per source rank
constexpr uint32_t kFlagSelectCtr = 32;   // own select epoch counter
constexpr uint32_t kFlagStatusCtr = 33;   // own status epoch counter

// Fused all-gather, the way it works: natural-order push blocks only become complete near the END of the grid kernel
// (the rows of any 1024-row block are spread over all classes, and the classes are processed one after the other), so
// pushing "a block as soon as it is complete" degenerates into a serial tail on the few warps that finalise the last
// rows (measured: +225 us on 500k rows).  Instead every retired row block adds its rows to one counter; warps that run
// out of work wait for the counter to reach P -- the slowest warp is at most one item behind -- and then EVERY warp of
// the grid copies an equal share of the slice to the peers (push_share); the CTA that finishes last signals.
// No copy kernel, no drain, no signal kernel, no hand-out counter; the copy itself is spread over the whole grid.
__device__ __forceinline__ bool fused_gather(const SelectArgs& a) { return a.n_out > 1 && a.self_out >= 0; }

// whole-warp: `n` rows of this warp are final in the local vector
__device__ __forceinline__ void count_rows_done(const SelectArgs& a, uint32_t n) {
    if (!fused_gather(a)) return;
    __threadfence();  // this lane's store into the local vector is visible before the rows count as done
    __syncwarp();
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&a.hdr[kHdrRowsDone], n);
}

// whole CTA, after its warps have run out of items: one thread waits until every row is final (one poller per CTA, on
// a cache line nothing else touches while the grid runs), then every warp of the grid copies its share of the slice to
// the peers; the CTA that finishes last signals (and, with inline_wait, waits for the peers' signals)
__device__ __forceinline__ void push_tail(const SelectArgs& a) {
    if (!fused_gather(a)) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        const volatile uint32_t* done = a.hdr + kHdrRowsDone;
        while (*done < a.P) __nanosleep(100);
    }
    __syncthreads();
    __threadfence();  // every row counted done is visible here
    push_share(a);
    __threadfence_system();  // this lane's peer stores are performed system-wide before its CTA counts as pushed
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t last = 0;
        if (threadIdx.x == 0) last = atomicAdd(&a.hdr[kHdrPushed], 1u) + 1u == gridDim.x ? 1u : 0u;
        last = __shfl_sync(0xFFFFFFFFu, last, 0);
        if (last && a.n_flags > 0) signal_peers(a, kFlagSelect, kFlagSelectCtr);
    }
}

__device__ __forceinline__ void store_best_local(const SelectArgs& a, uint32_t row, int32_t b) {
    if (a.self_out >= 0) { a.best_out[a.self_out][a.row0 + row] = b; return; }
    for (int o = 0; o < a.n_out; ++o) a.best_out[o][a.row0 + row] = b;
}

// ---------------------------------------------------------------------------------------------------------
// K0a: classify.  No block waits for another one: the class counts are only added to here (the grid kernel of the
// previous call zeroed them once the scatter kernel was done with them); block 0 resets the grid kernel's queue cursors.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPThreads) k_pod_classify(SelectArgs a, uint32_t n_cursors) {
    pdl_trigger();  // k_pod_scatter may be scheduled; it waits before it reads anything written here
    __shared__ int32_t s_dist[3][64];
    __shared__ uint32_t s_hist[kMaxClasses];
    const uint32_t tid = threadIdx.x;
    const ClassDims cd = class_dims(a.D);
    uint32_t* hist = a.hist;
    for (uint32_t i = tid; i < 192; i += kPThreads) {
        const uint32_t d = i >> 6, k = i & 63;
        s_dist[d][k] = k < a.D[d] ? __ldg(a.distinct[d] + k) : INT32_MAX;
    }
    for (uint32_t i = tid; i < cd.C; i += kPThreads) s_hist[i] = 0u;
    if (blockIdx.x == 0) {  // the previous call's kernels have completed (this is not a programmatic launch): its queue state is free
        for (uint32_t i = tid; i < n_cursors; i += kPThreads) a.hdr[kHdrCursors + i] = 0u;
        if (tid == 0) { a.hdr[kHdrPushed] = 0u; a.hdr[kHdrRowsDone] = 0u; }
    }
    __syncthreads();
    // a CTA takes 512-row chunks b, b + grid, ...: the 4096-bin histogram (zeroing, flush) is paid once per CTA, not per chunk
    for (uint32_t p = blockIdx.x * kPThreads + tid; p < a.P; p += gridDim.x * kPThreads) {
        const uint8_t c = a.cloud ? a.cloud[p] : (uint8_t)RPK_CLOUD_SECURE;
        const uint32_t tm1 = lb_smem(s_dist[0], a.D[0], a.req_mem[p]);
        const uint32_t tv = lb_smem(s_dist[1], a.D[1], a.req_vcpu ? a.req_vcpu[p] : 0);
        const uint32_t tr = lb_smem(s_dist[2], a.D[2], a.req_ram ? a.req_ram[p] : 0);
        a.rw[p] = tm1 | ((a.pk.bm_off_vcpu + tv) << 8) | ((a.pk.bm_off_ram + tr) << 16);
        uint32_t key = 0xFFFFu;  // neither SECURE nor COMMUNITY: nothing is feasible (runpod_client.go:469-475)
        if (c <= 1) { key = class_key(cd, c, tm1, tv, tr); atomicAdd(&s_hist[key], 1u); }
        a.key[p] = (uint16_t)key;
        a.pos[p] = kNone;  // "no feasible offer yet" for sorted position p (positions [0, valid rows) are used: a coalesced fill)
    }
    __syncthreads();
    for (uint32_t i = tid; i < cd.C; i += kPThreads) { const uint32_t v = s_hist[i]; if (v) atomicAdd(&hist[i], v); }
}

// ---------------------------------------------------------------------------------------------------------
// K0b: scatter into class order.  A CTA takes 512-row chunks b, b + grid, ... (the same rows in both passes below).
// Pass 1 counts its rows per class in shared memory while the class totals arrive from L2; every CTA scans the totals
// itself (<= 4096 words: cheaper than a serial "last block" phase in the kernel before) and adds its counts to the
// global class cursors with one atomic per non-empty class, which turns its shared table into "next free position per
// class"; pass 2 re-reads the keys (L2 hits) and hands every row its position.  The per-CTA cost of the 4096-bin
// tables is paid once per CTA, not once per 512 rows.  Block 0 also publishes the per-cloud row / work totals the grid
// kernel shares its CTAs out by.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPThreads) k_pod_scatter(SelectArgs a) {
    __shared__ uint32_t s_cnt[kMaxClasses], s_base[kMaxClasses];
    __shared__ uint32_t s_warp[kPWarps];
    __shared__ unsigned long long s_work[2];
    __shared__ uint32_t s_rows[2];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ClassDims cd = class_dims(a.D);
    for (uint32_t i = tid; i < cd.C; i += kPThreads) s_cnt[i] = 0u;
    if (tid < 2) { s_work[tid] = 0ull; s_rows[tid] = 0u; }
    __syncthreads();
    pdl_wait();     // k_pod_classify has completed: keys, rw, class counts
    pdl_trigger();  // the grid kernel may be scheduled; it waits before it reads anything written here
    const uint32_t* hist = a.hist;
    uint32_t* cursor = a.cursor;
    const uint32_t stride = gridDim.x * kPThreads;
    // the class totals first (independent of pass 1: the loads overlap it); thread t owns classes [8t, 8t + 8)
    constexpr uint32_t kPer = kMaxClasses / kPThreads;
    uint32_t v[kPer], sum = 0;
#pragma unroll
    for (uint32_t k = 0; k < kPer; ++k) {
        const uint32_t i = tid * kPer + k;
        v[k] = i < cd.C ? __ldcg(hist + i) : 0u;
    }
    // pass 1: this CTA's rows per class
    for (uint32_t p = blockIdx.x * kPThreads + tid; p < a.P; p += stride) {
        const uint32_t key = (uint32_t)a.key[p];
        if (key != 0xFFFFu) atomicAdd(&s_cnt[key], 1u);
    }
    unsigned long long work[2] = {0ull, 0ull};
    uint32_t rows[2] = {0u, 0u};
#pragma unroll
    for (uint32_t k = 0; k < kPer; ++k) {
        const uint32_t i = tid * kPer + k;
        sum += v[k];
        if (blockIdx.x == 0 && v[k]) { const uint32_t c = i >= cd.C / 2 ? 1u : 0u; rows[c] += v[k]; work[c] += (unsigned long long)v[k] * class_weight(cd, i); }
    }
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, d); if ((int)lane >= d) inc += n; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();  // s_warp, and every row of pass 1 is counted
    uint32_t run = inc - sum;
    for (uint32_t w = 0; w < warp; ++w) run += s_warp[w];
    if (blockIdx.x == 0) {
        for (int c = 0; c < 2; ++c) if (rows[c]) { atomicAdd(&s_rows[c], rows[c]); atomicAdd(&s_work[c], work[c]); }
    }
    // class start + this CTA's share of the class = the first position this CTA may use
#pragma unroll
    for (uint32_t k = 0; k < kPer; ++k) {
        const uint32_t i = tid * kPer + k;
        if (i < cd.C) { const uint32_t c = s_cnt[i]; s_base[i] = run + (c ? atomicAdd(&cursor[i], c) : 0u); }
        run += v[k];
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) {
        a.hdr[kHdrRows0] = s_rows[0]; a.hdr[kHdrRows1] = s_rows[1];
        a.hdr[kHdrWork0] = (uint32_t)min(s_work[0] >> 2, 0xFFFFFFFFull); a.hdr[kHdrWork1] = (uint32_t)min(s_work[1] >> 2, 0xFFFFFFFFull);
    }
    // pass 2: positions (the order inside a class is arbitrary -- results go back by row id)
    uint32_t n_neither = 0;
    for (uint32_t p0 = blockIdx.x * kPThreads; p0 < a.P; p0 += stride) {  // whole warps: the ballot below
        const uint32_t p = p0 + tid;
        const uint32_t key = p < a.P ? (uint32_t)a.key[p] : 0u;
        const bool neither = p < a.P && key == 0xFFFFu;
        if (p < a.P && !neither) {
            const uint32_t dst = atomicAdd(&s_base[key], 1u);
            a.ord_rw[dst] = make_uint2(p, a.rw[p]);  // the one scattered store of the sort
        }
        if (neither) {
            store_best_local(a, p, -1);
            if (a.top5) for (int k = 0; k < RPK_TOPK; ++k) a.top5[(size_t)p * RPK_TOPK + k] = -1;
        }
        n_neither += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, neither));
    }
    if (n_neither) count_rows_done(a, n_neither);
}

// ---------------------------------------------------------------------------------------------------------
// K1: persistent grid kernel
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 lds4(const uint32_t* p) { return *reinterpret_cast<const uint4*>(p); }

template <int RPL, bool NV, bool NR>
__device__ __forceinline__ void walk(const uint32_t* __restrict__ stage, uint32_t w68, int sub_lo, int sub_hi, const uint32_t (&o1)[RPL],
                                     const uint32_t (&o2)[RPL], const uint32_t (&o3)[RPL], uint32_t (&bb)[RPL]) {
    for (int sub = sub_hi - 1; sub >= sub_lo; --sub) {  // descending price: the last hit is the cheapest
        const uint32_t* sp = stage + (size_t)sub * w68;
#pragma unroll
        for (int b8 = 7; b8 >= 0; --b8) {
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const uint4 a0 = lds4(sp + o1[r] + b8 * 8), a1 = lds4(sp + o1[r] + b8 * 8 + 4);
                uint32_t any;
                if (!NV && !NR) {
                    any = (a0.x | a0.y | a0.z) | (a0.w | a1.x | a1.y) | (a1.z | a1.w);
                } else if (NV != NR) {
                    const uint32_t ob = NV ? o2[r] : o3[r];
                    const uint4 b0 = lds4(sp + ob + b8 * 8), b1 = lds4(sp + ob + b8 * 8 + 4);
                    any = (a0.x & b0.x) | (a0.y & b0.y);
                    any |= (a0.z & b0.z); any |= (a0.w & b0.w);
                    any |= (a1.x & b1.x); any |= (a1.y & b1.y);
                    any |= (a1.z & b1.z); any |= (a1.w & b1.w);
                } else {
                    const uint4 b0 = lds4(sp + o2[r] + b8 * 8), b1 = lds4(sp + o2[r] + b8 * 8 + 4);
                    const uint4 c0 = lds4(sp + o3[r] + b8 * 8), c1 = lds4(sp + o3[r] + b8 * 8 + 4);
                    any = (a0.x & b0.x & c0.x) | (a0.y & b0.y & c0.y) | (a0.z & b0.z & c0.z);
                    any |= (a0.w & b0.w & c0.w) | (a1.x & b1.x & c1.x);
                    any |= (a1.y & b1.y & c1.y) | (a1.z & b1.z & c1.z);
                    any |= (a1.w & b1.w & c1.w);
                }
                if (any) bb[r] = (uint32_t)(sub * 8 + b8);
            }
        }
    }
}

struct PersistArgs {
    uint32_t S, per, qspan, n_groups, tickets_per_block, stage_cap_subs;
};

template <int RPL, int MINB>
__device__ __forceinline__ void select_persist_body(const SelectArgs& a, const PersistArgs& pa) {
    extern __shared__ __align__(128) uint32_t s_stage[];
    __shared__ __align__(8) uint64_t s_bar[kCopyGroups];
    const uint32_t tid = threadIdx.x, lane = tid & 31;
    constexpr uint32_t RPI = 32 * RPL;  // rows per item
    if (tid == 0) {
        for (int g = 0; g < kCopyGroups; ++g) mbar_init(&s_bar[g], 1);
        mbar_fence_init();
    }
    pdl_wait();     // k_pod_scatter (and k_pod_classify before it) has completed
    pdl_trigger();  // a dependent (the peer wait) may be scheduled; it waits for this grid to complete
    // the scatter kernel is done with the class counts and cursors: zero them for the next call (all kMaxClasses words --
    // the table, and with it the class count, may change between calls)
    if (blockIdx.x == 0) for (uint32_t i = tid; i < kMaxClasses; i += kPThreads) { a.hist[i] = 0u; a.cursor[i] = 0u; }
    __syncthreads();
    const uint32_t rows[2] = {__ldcg(a.hdr + kHdrRows0), __ldcg(a.hdr + kHdrRows1)};
    const uint32_t work[2] = {__ldcg(a.hdr + kHdrWork0), __ldcg(a.hdr + kHdrWork1)};
    const uint32_t S = pa.S, W = a.pk.bm_words, w68 = W * kSubWords;
    const uint32_t have0 = rows[0] ? 1u : 0u, have1 = rows[1] ? 1u : 0u;
    const uint32_t npairs = S * (have0 + have1);  // 0: every row was final in the scatter kernel (neither cloud); only the push is left
    // which (cloud, segment) stages this CTA serves: with at least as many CTAs as stages the CTAs are shared out in
    // proportion to the clouds' mask-word work and every CTA keeps one stage for the whole call; otherwise CTA b serves
    // stages b, b + grid, ... one after the other
    uint32_t first, step, ctas0 = 0;
    const bool shared_out = npairs <= gridDim.x;
    if (npairs == 0) {
        first = 0; step = 1;
    } else if (shared_out) {
        const unsigned long long tot = (unsigned long long)work[0] + work[1];
        ctas0 = tot ? (uint32_t)(((unsigned long long)gridDim.x * work[0] + tot / 2) / tot) : gridDim.x / 2;
        ctas0 = max(ctas0, S * have0);
        ctas0 = min(ctas0, gridDim.x - S * have1);
        if (!have0) ctas0 = 0;
        first = blockIdx.x < ctas0 ? blockIdx.x % S : S * have0 + (blockIdx.x - ctas0) % S;
        step = npairs;  // one stage only
    } else {
        first = blockIdx.x; step = gridDim.x;
    }
    const uint32_t nblk0 = (rows[0] + RPI - 1) / RPI;
    uint32_t it = 0, phase_bits = 0;  // bit g: parity of copy group g's next completion (groups are not all used by every stage)
    for (uint32_t pair = first; pair < npairs; pair += step, ++it) {
        const uint32_t c = have0 ? pair / S : 1u, s = pair % S;
        const uint32_t sub0 = s * pa.per, nsubs = min(pa.per, a.nsub - sub0);
        const uint32_t gs = (nsubs + pa.n_groups - 1) / pa.n_groups;  // subs per copy group
        if (it) __syncthreads();  // every warp is done with the previous stage
        uint32_t used = 0;
        for (uint32_t g = 0; g * gs < nsubs; ++g) {
            used |= 1u << g;
            if (tid == 0) {
                const uint32_t g0 = g * gs, gn = min(gs, nsubs - g0);
                mbar_expect_tx(&s_bar[g], gn * w68 * 4u);
                bulk_g2s(s_stage + (size_t)g0 * w68, a.view[c].bitmapT + ((size_t)sub0 + g0) * w68, gn * w68 * 4u, &s_bar[g]);
            }
        }
        const uint32_t parity = phase_bits;
        phase_bits ^= used;
        uint32_t ready = 0;  // copy groups this warp has seen complete
        const uint32_t n_c = rows[c], c_start = c ? rows[0] : 0u;
        const uint32_t nblk = (n_c + RPI - 1) / RPI;
        const uint32_t Qi = (nsubs + pa.qspan - 1) / pa.qspan;
        const uint32_t n_items = nblk * Qi;
        uint32_t* cursor = a.hdr + kHdrCursors + c * S + s;
        uint32_t* tickets = a.tile_ctr + (c ? nblk0 : 0u);
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(cursor, 1u);
        item = __shfl_sync(0xFFFFFFFFu, item, 0);
        const uint32_t w_none = (a.pk.bm_off_vcpu << 8) | (a.pk.bm_off_ram << 16);  // padding rows: threshold 0 everywhere
        uint32_t w_cur[RPL];
#pragma unroll
        for (int r = 0; r < RPL; ++r) {
            const uint32_t local = (item / Qi) * RPI + (uint32_t)r * 32 + lane;
            w_cur[r] = item < n_items && local < n_c ? __ldcg(a.ord_rw + c_start + local).y : w_none;
        }
        uint32_t pend1_blk = kNone, pend2_blk = kNone, pend2_ticket = 0;  // merge pipeline (see below)
        // the row block whose ticket was taken one iteration ago: if that was its last ticket, every segment of its rows
        // has been merged -- price < maxPrice on the winner (strict, runpod_client.go:478), position -> offer index
        auto retire = [&](uint32_t rblk, uint32_t ticket) {
            if (rblk == kNone) return;
            if (__shfl_sync(0xFFFFFFFFu, ticket, 0) != pa.tickets_per_block - 1) return;
            __threadfence();
            if (lane == 0) tickets[rblk] = 0u;  // self-cleaning
            uint32_t n_final = 0;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const uint32_t local = rblk * RPI + (uint32_t)r * 32 + lane;
                const bool ok = local < n_c;
                n_final += (uint32_t)__popc(__ballot_sync(0xFFFFFFFFu, ok));
                uint32_t row = 0;
                if (ok) {
                    row = a.ord_rw[c_start + local].x;
                    const uint32_t p = __ldcg(a.pos + c_start + local);
                    int32_t b = -1;
                    if (p != kNone) {
                        const double pr = a.view[c].price[p];
                        const double mx = a.max_price ? a.max_price[row] : RPK_DEFAULT_MAX_PRICE;
                        if (pr < mx) b = a.view[c].perm[p];
                    }
                    store_best_local(a, row, b);
                }
            }
            count_rows_done(a, n_final);
            if (a.n_out > 1 && a.self_out < 0) __threadfence_system();  // direct peer stores: performed before the grid completes
        };
        while (item < n_items) {
            uint32_t next = 0;
            if (lane == 0) next = atomicAdd(cursor, 1u);  // in flight while this item is walked
            const uint32_t blk = item / Qi, qi = item - blk * Qi;
            const int sub_lo = (int)(qi * pa.qspan), sub_hi = (int)min(nsubs, (qi + 1) * pa.qspan);
            uint32_t idx[RPL], o1[RPL], o2[RPL], o3[RPL], bb[RPL];
            bool valid[RPL];
            bool nv = false, nr = false;
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const uint32_t local = blk * RPI + (uint32_t)r * 32 + lane;
                valid[r] = local < n_c;
                idx[r] = c_start + local;
                const uint32_t w = w_cur[r];
                const uint32_t i1 = w & 0xFFu, i2 = (w >> 8) & 0xFFu, i3 = (w >> 16) & 0xFFu;
                o1[r] = i1 * kSubWords; o2[r] = i2 * kSubWords; o3[r] = i3 * kSubWords;
                nv |= i2 != a.pk.bm_off_vcpu; nr |= i3 != a.pk.bm_off_ram;
                bb[r] = kNone;
            }
            nv = __any_sync(0xFFFFFFFFu, nv); nr = __any_sync(0xFFFFFFFFu, nr);
            for (uint32_t g = (uint32_t)sub_lo / gs; g * gs < (uint32_t)sub_hi; ++g)
                if (!(ready >> g & 1u)) { mbar_wait(&s_bar[g], parity >> g & 1u); ready |= 1u << g; }
            // rows of an item share the "column constrains" flags except at class boundaries; a non-constraining column's
            // threshold-0 word is all-available, so reading it for a row that does not need it is harmless
            if (!nv && !nr) walk<RPL, false, false>(s_stage, w68, sub_lo, sub_hi, o1, o2, o3, bb);
            else if (nv && !nr) walk<RPL, true, false>(s_stage, w68, sub_lo, sub_hi, o1, o2, o3, bb);
            else if (!nv && nr) walk<RPL, false, true>(s_stage, w68, sub_lo, sub_hi, o1, o2, o3, bb);
            else walk<RPL, true, true>(s_stage, w68, sub_lo, sub_hi, o1, o2, o3, bb);
            // The merge of an item costs two device-wide round trips (its atomicMin's must be performed before its ticket is
            // taken, and the ticket's return value says whether these rows are final), so it is pipelined over the
            // NEXT items' walks: here the item walked one iteration ago takes its ticket (its atomicMin's were issued a
            // whole walk ago: the fence returns at once), and the item before that reads its ticket and, if it was the
            // last one of its row block, runs the epilogue.
            retire(pend2_blk, pend2_ticket);
            pend2_blk = pend1_blk; pend2_ticket = 0;
            if (pend1_blk != kNone) {
                __threadfence();
                if (lane == 0) pend2_ticket = atomicAdd(&tickets[pend1_blk], 1u);  // consumed after the next walk
            }
            // the next item's thresholds are fetched under this item's resolve and the next walk's start
            const uint32_t item_next = __shfl_sync(0xFFFFFFFFu, next, 0);
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                const uint32_t local = (item_next / Qi) * RPI + (uint32_t)r * 32 + lane;
                w_cur[r] = item_next < n_items && local < n_c ? __ldcg(a.ord_rw + c_start + local).y : w_none;
            }
#pragma unroll
            for (int r = 0; r < RPL; ++r) {
                if (valid[r] && bb[r] != kNone) {  // re-read the cheapest block with a hit: first chunk, first bit
                    const uint32_t sub = bb[r] >> 3, b8 = bb[r] & 7u;
                    const uint32_t* sp = s_stage + (size_t)sub * w68 + b8 * 8;
                    uint32_t j = 0, m = 0;
#pragma unroll
                    for (int k = 7; k >= 0; --k) {
                        const uint32_t mk = sp[o1[r] + k] & sp[o2[r] + k] & sp[o3[r] + k];
                        if (mk) { j = (uint32_t)k; m = mk; }
                    }
                    atomicMin(&a.pos[idx[r]], (((sub0 + sub) * kSubChunks + b8 * 8 + j) << 5) + (uint32_t)__ffs(m) - 1u);
                }
            }
            pend1_blk = blk;
            item = item_next;
        }
        // drain the pipeline: the last two items of this stage
        retire(pend2_blk, pend2_ticket);
        pend2_blk = pend1_blk; pend2_ticket = 0; pend1_blk = kNone;
        if (pend2_blk != kNone) {
            __threadfence();
            if (lane == 0) pend2_ticket = atomicAdd(&tickets[pend2_blk], 1u);
        }
        retire(pend2_blk, pend2_ticket);
        pend2_blk = kNone;
    }
    push_tail(a);
}

// Register budgets.  <RPL, 2>: two CTAs per SM, 64 registers.  <RPL, 1>: one CTA per SM; 128 registers give the deepest
// load pipelining, but then the CTA owns the SM's whole register file and a status sweep enqueued beside the select
// (other stream) cannot start on that SM until the select is over -- or, if it got there first, delays it.  With 96
// registers (k_select_persist96) 16K registers stay free: exactly one 256-thread CTA of the sweep kernel, so the two
// kernels -- one bound by shared memory, the other by HBM and integer multiplies -- really run side by side.
template <int RPL, int MINB>
__global__ void __launch_bounds__(kPThreads, MINB) k_select_persist(SelectArgs a, PersistArgs pa) { select_persist_body<RPL, MINB>(a, pa); }
template <int RPL>
__global__ void __maxnreg__(96) k_select_persist96(SelectArgs a, PersistArgs pa) { select_persist_body<RPL, 1>(a, pa); }

// ---------------------------------------------------------------------------------------------------------
// peer signal / wait (flags bound with rpk_peer_bind).  The persistent kernel signals from its last pusher; the
// other select paths launch k_peer_signal behind their copy kernel; k_peer_wait is what remains of the fence.
// ---------------------------------------------------------------------------------------------------------
__global__ void k_peer_signal(PeerFenceArgs a, uint32_t word0, uint32_t counter_word) {
    pdl_wait();  // everything this stream did before (peer stores included) has completed
    const uint32_t lane = threadIdx.x;
    uint32_t e = 0;
    if (lane == 0) { e = a.flags[a.my_rank][counter_word] + 1u; a.flags[a.my_rank][counter_word] = e; }
    e = __shfl_sync(0xFFFFFFFFu, e, 0);
    __threadfence_system();
    if ((int)lane < a.n) {
        *reinterpret_cast<volatile uint32_t*>(a.flags[lane] + word0 + a.my_rank) = e;
        if (a.epoch) {  // epoch != 0 doubles as "wait here too" for this kernel (inline-wait mode)
            const volatile uint32_t* mine = a.flags[a.my_rank] + word0 + lane;
            while ((int32_t)(*mine - e) < 0) __nanosleep(32);
            __threadfence_system();
        }
    }
}

__global__ void k_peer_wait(PeerFenceArgs a, uint32_t what) {
    pdl_wait();  // the kernels that signalled for this rank have completed: the own epoch counters are final
    const uint32_t lane = threadIdx.x;
    const volatile uint32_t* mine = a.flags[a.my_rank];
    if ((int)lane < a.n) {
        if (what & 1u) { const uint32_t e = mine[kFlagSelectCtr]; while ((int32_t)(mine[kFlagSelect + lane] - e) < 0) __nanosleep(32); }
        if (what & 2u) { const uint32_t e = mine[kFlagStatusCtr]; while ((int32_t)(mine[kFlagStatus + lane] - e) < 0) __nanosleep(32); }
        __threadfence_system();
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
struct PTune { int rpl = 0; int stage_kb = 0; int items_per_warp = 0; int ctas = 0; int minb = 0; int regs = 0; int k0ctas = 592; bool pdl = true; };
static PTune parse_ptune();
static PTune read_ptune() {  // read once per process unless RPK_TUNE_RELOAD is set (see select.cu)
    static const bool reload = getenv("RPK_TUNE_RELOAD") != nullptr;
    if (reload) return parse_ptune();
    static const PTune cached = parse_ptune();
    return cached;
}
static PTune parse_ptune() {
    PTune t;
    const char* e = getenv("RPK_TUNE");
    if (!e) return t;
    if (const char* p = strstr(e, "prpl=")) t.rpl = atoi(p + 5);
    if (const char* p = strstr(e, "stagekb=")) t.stage_kb = atoi(p + 8);
    if (const char* p = strstr(e, "ipw=")) t.items_per_warp = atoi(p + 4);
    if (const char* p = strstr(e, "pctas=")) t.ctas = atoi(p + 6);
    if (const char* p = strstr(e, "minb=")) t.minb = atoi(p + 5);
    if (const char* p = strstr(e, "regs=")) t.regs = atoi(p + 5);
    if (const char* p = strstr(e, "k0ctas=")) { const int v = atoi(p + 7); if (v > 0) t.k0ctas = v; }
    t.pdl = strstr(e, "pdl=off") == nullptr;
    return t;
}

uint32_t persist_hdr_words(uint32_t G, uint32_t bm_words) {
    (void)bm_words;
    const uint32_t nsub = ((G + 31) / 32 + kSubChunks - 1) / kSubChunks;
    return kHdrCursors + 2 * (nsub ? nsub : 1) + 8;  // S <= nsub
}

bool persist_plan(const SelectArgs& a, int sm_count, PersistPlan* pl) {
    if (!a.pk.bm_words || !a.nsub) return false;
    const PTune t = read_ptune();
    const uint32_t sub_bytes = a.pk.bm_words * kSubWords * 4u;
    // minb = 2: two CTAs of 512 threads per SM share the 227 KB (stage up to ~110 KB, 64 registers per thread);
    // minb = 1: one CTA per SM with the whole shared memory and up to 128 registers per thread (deeper load pipelining)
    pl->minb = t.minb == 2 ? 2 : 1;
    uint32_t max_stage = (uint32_t)(t.stage_kb > 0 ? t.stage_kb : (pl->minb == 1 ? 220 : 110)) * 1024u;
    if (max_stage > 220u * 1024u) max_stage = 220u * 1024u;
    uint32_t cap = max_stage / sub_bytes;
    if (cap == 0) cap = 1;
    const uint32_t S = (a.nsub + cap - 1) / cap;
    const uint32_t per = (a.nsub + S - 1) / S;
    pl->cap_subs = cap; pl->S = (a.nsub + per - 1) / per; pl->per = per;
    pl->smem_bytes = per * sub_bytes;
    uint32_t ctas = (227u * 1024u) / (pl->smem_bytes + 1024u + 128u);
    if (ctas > 2048u / kPThreads) ctas = 2048u / kPThreads;
    if (ctas == 0) ctas = 1;
    if (t.ctas > 0 && (uint32_t)t.ctas < ctas) ctas = (uint32_t)t.ctas;
    if (pl->minb == 1) ctas = 1;
    pl->grid = (uint32_t)sm_count * ctas;
    pl->rpl = t.rpl == 1 ? 1 : t.rpl == 4 && pl->minb == 1 ? 4 : 2;
    const uint64_t warps = (uint64_t)pl->grid * kPWarps;
    const uint64_t nblk = ((uint64_t)a.P + 32u * pl->rpl - 1) / (32u * pl->rpl);
    const uint64_t want = (uint64_t)(t.items_per_warp > 0 ? t.items_per_warp : 4) * warps;  // items per warp: the tail is one item long
    const uint64_t visits = nblk * pl->S > 0 ? nblk * pl->S : 1;  // (row block, stage) visits
    uint64_t Qi = (want + visits - 1) / visits;
    if (Qi < 1) Qi = 1;
    if (Qi > per) Qi = per;
    pl->qspan = (uint32_t)((per + Qi - 1) / Qi);
    pl->Qi = (per + pl->qspan - 1) / pl->qspan;
    return true;
}

template <typename... KArgs, typename... Args>
static void launch_pdl_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    RPK_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

int launch_select_persist(const SelectArgs& a, const PersistPlan& pl, cudaStream_t st) {
    const PTune t = read_ptune();
    PersistArgs pa;
    pa.S = pl.S; pa.per = pl.per; pa.qspan = pl.qspan; pa.stage_cap_subs = pl.cap_subs;
    pa.n_groups = pl.per < (uint32_t)kCopyGroups ? pl.per : (uint32_t)kCopyGroups;
    uint32_t tickets = 0;
    for (uint32_t s = 0; s < pl.S; ++s) {
        const uint32_t sub0 = s * pl.per, nsubs = a.nsub - sub0 < pl.per ? a.nsub - sub0 : pl.per;
        tickets += (nsubs + pl.qspan - 1) / pl.qspan;
    }
    pa.tickets_per_block = tickets;
    // K0: one wave of CTAs (four per SM), each looping over 512-row chunks
    const uint32_t chunks = (a.P + kPThreads - 1) / kPThreads;
    const uint32_t blocks = chunks < (uint32_t)t.k0ctas ? (chunks ? chunks : 1u) : (uint32_t)t.k0ctas;
    k_pod_classify<<<blocks, kPThreads, 0, st>>>(a, 2 * pl.S);
    launch_pdl_k(k_pod_scatter, dim3(blocks), dim3(kPThreads), 0, st, t.pdl, a);
    static thread_local int attr_dev[5] = {-1, -1, -1, -1, -1};
    int dev = 0;
    RPK_CUDA(cudaGetDevice(&dev));
    auto go = [&](auto kernel, int slot) {
        if (attr_dev[slot] != dev) { RPK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024)); attr_dev[slot] = dev; }
        launch_pdl_k(kernel, dim3(pl.grid), dim3(kPThreads), pl.smem_bytes, st, t.pdl, a, pa);
    };
    if (pl.minb == 1) {
        if (pl.rpl == 4) go(k_select_persist<4, 1>, 3);
        else if (t.regs == 128) go(k_select_persist<2, 1>, 2);
        else go(k_select_persist96<2>, 4);
    } else {
        if (pl.rpl == 1) go(k_select_persist<1, 2>, 0); else go(k_select_persist<2, 2>, 1);
    }
    RPK_CUDA(cudaGetLastError());
    return 3;
}

int launch_peer_signal(const PeerFenceArgs& a, uint32_t word0, cudaStream_t st) {
    launch_pdl_k(k_peer_signal, dim3(1), dim3(32), 0, st, read_ptune().pdl, a, word0, word0 == kFlagSelect ? kFlagSelectCtr : kFlagStatusCtr);
    RPK_CUDA(cudaGetLastError());
    return 1;
}

int launch_peer_wait(const PeerFenceArgs& a, uint32_t what, cudaStream_t st) {
    launch_pdl_k(k_peer_wait, dim3(1), dim3(32), 0, st, read_ptune().pdl, a, what);
    RPK_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace rpk
