// status.cu -- batched pod-status diff: the predicate of Provider.updateAllPodStatuses
// (reference kubelet.go:857-880) over N tracked slots in one pass.
//
// One record = exactly the two fields the reference compares (InstanceInfo.Status, .PortsExposed:
// runpod_client.go:103,108) in a fixed slot [len][status][0x00][ports][pad].  The previous state lives on
// the device as one 64-bit XXH64 per slot; "status changed || ports changed" (kubelet.go:870-873) becomes
// hash != previous hash.  Single pass, HBM-bound: every record is read once (coalesced 16-byte loads into a
// padded shared-memory tile), hashed, compared, the new hash written back, and the changed slot indices are
// emitted in ascending order through a decoupled look-back scan (no second pass over the data).
#include "rpk_internal.cuh"

namespace rpk {

using u64 = unsigned long long;

constexpr u64 P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full, P3 = 0x165667B19E3779F9ull,
              P4 = 0x85EBCA77C2B2AE63ull, P5 = 0x27D4EB2F165667C5ull;

__device__ __forceinline__ u64 rotl64(u64 x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ u64 xround(u64 acc, u64 in) { return rotl64(acc + in * P2, 31) * P1; }
__device__ __forceinline__ u64 xmerge(u64 h, u64 v) { return (h ^ xround(0, v)) * P1 + P4; }

// The slot's data starts at byte 1 (after the length byte), so every 4/8-byte lane of the hash input sits
// at word offset +1 byte: one funnel shift by 8 per 32-bit half.
__device__ __forceinline__ uint32_t rd32(const uint32_t* w, uint32_t off) {  // off % 4 == 0
    const uint32_t q = off >> 2;
    return __funnelshift_r(w[q], w[q + 1], 8);
}
__device__ __forceinline__ u64 rd64(const uint32_t* w, uint32_t off) {  // off % 8 == 0
    const uint32_t q = off >> 2;
    const uint32_t a = w[q], b = w[q + 1], c = w[q + 2];
    return (u64)__funnelshift_r(a, b, 8) | ((u64)__funnelshift_r(b, c, 8) << 32);
}
__device__ __forceinline__ uint32_t rd8(const uint32_t* w, uint32_t off) {
    const uint32_t s = off + 1;
    return (w[s >> 2] >> ((s & 3) * 8)) & 0xFFu;
}

// XXH64, seed 0, over `len` data bytes of a slot held as 32-bit words (public xxHash spec; this is what
// github.com/cespare/xxhash/v2 Sum64 computes, go.mod:60).
__device__ u64 xxh64_slot(const uint32_t* w, uint32_t len) {
    uint32_t off = 0;
    u64 h;
    if (len >= 32) {
        u64 v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0ull - P1;
        do {
            v1 = xround(v1, rd64(w, off)); v2 = xround(v2, rd64(w, off + 8));
            v3 = xround(v3, rd64(w, off + 16)); v4 = xround(v4, rd64(w, off + 24));
            off += 32;
        } while (off + 32 <= len);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xmerge(h, v1); h = xmerge(h, v2); h = xmerge(h, v3); h = xmerge(h, v4);
    } else {
        h = P5;
    }
    h += (u64)len;
    while (off + 8 <= len) { h ^= xround(0, rd64(w, off)); h = rotl64(h, 27) * P1 + P4; off += 8; }
    if (off + 4 <= len) { h ^= (u64)rd32(w, off) * P1; h = rotl64(h, 23) * P2 + P3; off += 4; }
    while (off < len) { h ^= (u64)rd8(w, off) * P5; h = rotl64(h, 11) * P1; ++off; }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

constexpr int kStThreads = 256;
#define kFlagAgg (1ull << 32)
#define kFlagPrefix (2ull << 32)
#define kFlagMask (3ull << 32)

static int items_for_stride(uint32_t stride) { return stride == 32 ? 2 : stride <= 32 ? 4 : stride <= 64 ? 2 : 1; }
uint32_t status_tiles(uint32_t N, uint32_t stride) {
    const uint32_t tile = (uint32_t)(kStThreads * items_for_stride(stride));
    return (N + tile - 1) / tile;
}

// Ordered compaction of the changed slots of one tile: per-(item, warp) counts -> exclusive scan ->
// decoupled look-back across tiles (tile ids were handed out in scheduling order, so every predecessor is
// already running) -> scatter of the slot indices, ascending.
template <int ITEMS>
__device__ __forceinline__ void emit_changed(const StatusArgs& a, uint32_t tile, uint32_t rec0, uint32_t nrec,
                                             const bool (&changed)[ITEMS], const uint32_t (&bal)[ITEMS],
                                             uint32_t* s_wcnt, uint32_t* s_excl_p) {
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __syncthreads();

    // warp 0: exclusive scan of the per-(item, warp) counts, then decoupled look-back for the tile prefix
    if (warp == 0) {
        constexpr uint32_t kCnt = ITEMS * (kStThreads / 32);  // <= 32
        uint32_t c = lane < kCnt ? s_wcnt[lane] : 0u, inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, d); if ((int)lane >= d) inc += n; }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        if (lane < kCnt) s_wcnt[lane] = inc - c;
        volatile u64* st = a.tile_state;
        uint32_t excl = 0;
        if (tile == 0) {
            if (lane == 0) st[0] = kFlagPrefix | total;
        } else {
            if (lane == 0) st[tile] = kFlagAgg | total;
            int look = (int)tile - 1;
            while (true) {
                const int idx = look - (int)lane;
                u64 v;
                do { v = kFlagPrefix; if (idx >= 0) v = st[idx]; } while (__any_sync(0xFFFFFFFFu, (v & kFlagMask) == 0));
                const uint32_t pm = __ballot_sync(0xFFFFFFFFu, (v & kFlagMask) == kFlagPrefix);
                const int first = pm ? __ffs(pm) - 1 : 31;
                uint32_t val = (int)lane <= first ? (uint32_t)v : 0u;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) val += __shfl_xor_sync(0xFFFFFFFFu, val, d);
                excl += val;
                if (pm) break;
                look -= 32;
            }
            if (lane == 0) st[tile] = kFlagPrefix | (u64)(excl + total);
        }
        if (lane == 0) {
            *s_excl_p = excl;
            if (rec0 + nrec == a.N) *a.n_changed = excl + total;  // the last tile in record order owns the count
        }
    }
    __syncthreads();
    const uint32_t excl = *s_excl_p;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (changed[k]) {
            const uint32_t posn = excl + s_wcnt[k * (kStThreads / 32) + warp] + __popc(bal[k] & ((1u << lane) - 1));
            a.changed_idx[posn] = a.idx_base + rec0 + (uint32_t)k * kStThreads + tid;
        }
    }
}

// ---- stride 32 fast path: the whole slot lives in 8 registers, no shared-memory staging -------------------
// Data byte i of the hash input is slot byte i+1, so the 8-byte lane m is words (2m, 2m+1, 2m+2) funnel-
// shifted by 8; len <= 31 keeps the input on XXH64's short path (no 32-byte stripes).
__device__ __forceinline__ u64 lane64(uint32_t a, uint32_t b, uint32_t c) {
    return (u64)__funnelshift_r(a, b, 8) | ((u64)__funnelshift_r(b, c, 8) << 32);
}
__device__ __forceinline__ u64 step8(u64 h, u64 lane) { h ^= xround(0, lane); return rotl64(h, 27) * P1 + P4; }

__device__ __forceinline__ u64 xxh64_slot32(const uint4 lo, const uint4 hi) {
    const uint32_t len = min(lo.x & 0xFFu, 31u);
    u64 h = P5 + (u64)len;
    const u64 l0 = lane64(lo.x, lo.y, lo.z), l1 = lane64(lo.z, lo.w, hi.x);
    u64 t;  // the 8-byte lane holding the <8 tail bytes
    if (len >= 16) {  // no RunPod status is this long: warp-uniformly skipped in practice
        const u64 l2 = lane64(hi.x, hi.y, hi.z), l3 = lane64(hi.z, hi.w, 0u);
        h = step8(step8(h, l0), l1);
        if (len >= 24) { h = step8(h, l2); t = l3; } else t = l2;
    } else if (len >= 8) {
        h = step8(h, l0); t = l1;
    } else {
        t = l0;
    }
    // tail: every lane of a warp has its own length, so all four sub-steps run anyway -- keep them branch-free
    {
        u64 h4 = h ^ ((u64)(uint32_t)t * P1);
        h4 = rotl64(h4, 23) * P2 + P3;
        const bool f4 = (len & 4u) != 0;
        h = f4 ? h4 : h;
        t = f4 ? (t >> 32) : t;
    }
    const uint32_t nb = len & 3u;
#pragma unroll
    for (uint32_t k = 0; k < 3; ++k) {
        u64 hb = h ^ (((t >> (8 * k)) & 0xFFull) * P5);
        hb = rotl64(hb, 11) * P1;
        h = nb > k ? hb : h;
    }
    h ^= h >> 33; h *= P2; h ^= h >> 29; h *= P3; h ^= h >> 32;
    return h;
}

// bulk async copy global -> shared with mbarrier completion (cp.async.bulk, SASS UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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

// Persistent CTAs over contiguous runs of 512-slot tiles.  Two 20 KB stages per CTA (512 slots x 32 B +
// their 512 previous hashes) are refilled by bulk async copies, so the next tile of this CTA -- and of the
// other CTA on the SM -- is in flight while a tile is hashed.  No CTA ever waits for another one: changed
// slot indices are staged in slot order inside the CTA's own chunk of `stage_idx`, and k_status_compact
// (a few microseconds) concatenates the chunks, which keeps the output ascending without a serial scan.
constexpr int kItems32 = 2;
constexpr uint32_t kTile32 = kStThreads * kItems32;  // 512 slots per tile, 20 KB per stage, 4 CTAs per SM
constexpr int kCtasPerSm32 = 4;
struct __align__(128) Stage32 {
    uint4 rec[kTile32 * 2];
    u64 prev[kTile32];
};

__device__ __forceinline__ void chunk_tiles(uint32_t n_tiles, uint32_t n_ctas, uint32_t c, uint32_t* lo, uint32_t* hi) {
    *lo = (uint32_t)((u64)n_tiles * c / n_ctas);
    *hi = (uint32_t)((u64)n_tiles * (c + 1) / n_ctas);
}

__global__ void __launch_bounds__(kStThreads, kCtasPerSm32) k_status_diff32(StatusArgs a, uint32_t n_tiles) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    Stage32* stage = reinterpret_cast<Stage32*>(s_raw);
    __shared__ __align__(8) uint64_t s_full[2];
    __shared__ uint32_t s_wcnt[2][kItems32 * (kStThreads / 32)];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_full = a.N / kTile32;  // tiles holding exactly kTile32 slots go through the bulk copies
    uint32_t t_lo, t_hi;
    chunk_tiles(n_tiles, gridDim.x, blockIdx.x, &t_lo, &t_hi);

    auto fill = [&](int s, uint32_t t) {  // thread 0 only
        if (t < t_hi && t < n_full) {
            mbar_expect_tx(&s_full[s], (uint32_t)sizeof(Stage32));
            bulk_g2s(stage[s].rec, a.records + (size_t)t * kTile32 * 32, kTile32 * 32, &s_full[s]);
            bulk_g2s(stage[s].prev, a.hash_prev + (size_t)t * kTile32, kTile32 * 8, &s_full[s]);
        } else {
            mbar_expect_tx(&s_full[s], 0);  // ragged last tile (direct loads) or past the chunk: nothing to wait for
        }
    };
    if (tid == 0) {
        mbar_init(&s_full[0], 1);
        mbar_init(&s_full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fill(0, t_lo);
        fill(1, t_lo + 1);
    }
    __syncthreads();

    uint32_t running = 0;  // changed slots of this chunk so far (same value in every thread)
    const uint32_t stage_base = t_lo * kTile32;
    for (uint32_t tile = t_lo, it = 0; tile < t_hi; ++tile, ++it) {
        const int s = (int)(it & 1);
        mbar_wait(&s_full[s], (it >> 1) & 1);
        const uint32_t rec0 = tile * kTile32;
        const uint32_t nrec = min(kTile32, a.N - rec0);
        uint4 lo[kItems32], hi[kItems32];
        u64 prev[kItems32];
        if (tile < n_full) {
#pragma unroll
            for (int k = 0; k < kItems32; ++k) {
                const uint32_t lr = (uint32_t)k * kStThreads + tid;
                lo[k] = stage[s].rec[lr * 2]; hi[k] = stage[s].rec[lr * 2 + 1]; prev[k] = stage[s].prev[lr];
            }
        } else {
            const uint4* __restrict__ src = reinterpret_cast<const uint4*>(a.records) + (size_t)rec0 * 2;
#pragma unroll
            for (int k = 0; k < kItems32; ++k) {
                const uint32_t lr = (uint32_t)k * kStThreads + tid;
                if (lr < nrec) { lo[k] = __ldcs(src + (size_t)lr * 2); hi[k] = __ldcs(src + (size_t)lr * 2 + 1); prev[k] = a.hash_prev[rec0 + lr]; }
                else { lo[k] = make_uint4(0, 0, 0, 0); hi[k] = lo[k]; prev[k] = 0; }
            }
        }
        bool changed[kItems32];
        uint32_t bal[kItems32];
        uint32_t* wcnt = s_wcnt[it & 1];  // double-buffered: a warp that runs ahead writes the other buffer
#pragma unroll
        for (int k = 0; k < kItems32; ++k) {
            const uint32_t lr = (uint32_t)k * kStThreads + tid;
            changed[k] = false;
            if (lr < nrec) {
                const u64 h = xxh64_slot32(lo[k], hi[k]);
                changed[k] = (prev[k] == 0ull) || (h != prev[k]);  // 0 = never seen
                if (changed[k]) a.hash_prev[rec0 + lr] = h;        // kubelet.go:875-880
                if (a.hash_out) a.hash_out[rec0 + lr] = h;
            }
            bal[k] = __ballot_sync(0xFFFFFFFFu, changed[k]);
            if (lane == 0) wcnt[k * (kStThreads / 32) + warp] = __popc(bal[k]);
        }
        __syncthreads();                     // the only CTA barrier per tile: counts are visible AND the stage is consumed
        if (tid == 0) fill(s, tile + 2);     // refill the stage (the other one is already in flight)
        if (a.stage_idx == nullptr) continue;  // seed: state only
        // every warp scans the (item, warp) counts itself: no second barrier
        constexpr uint32_t kCnt = kItems32 * (kStThreads / 32);
        const uint32_t c = lane < kCnt ? wcnt[lane] : 0u;
        uint32_t inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t n = __shfl_up_sync(0xFFFFFFFFu, inc, d); if ((int)lane >= d) inc += n; }
        const uint32_t excl = inc - c, total = __shfl_sync(0xFFFFFFFFu, inc, 31);
#pragma unroll
        for (int k = 0; k < kItems32; ++k) {
            const uint32_t base = __shfl_sync(0xFFFFFFFFu, excl, k * (kStThreads / 32) + warp);
            if (changed[k])
                a.stage_idx[stage_base + running + base + __popc(bal[k] & ((1u << lane) - 1))] = a.idx_base + rec0 + (uint32_t)k * kStThreads + tid;
        }
        running += total;
    }
    if (tid == 0 && a.cta_count) a.cta_count[blockIdx.x] = running;
}

// Concatenate the per-CTA index segments (each already ascending, chunks in slot order).
__global__ void __launch_bounds__(256) k_status_compact(StatusArgs a, uint32_t n_tiles, uint32_t n_ctas) {
    __shared__ uint32_t s_excl;
    const uint32_t c = blockIdx.x, tid = threadIdx.x;
    if (tid < 32) {
        uint32_t excl = 0, total = 0;
        for (uint32_t base = 0; base < n_ctas; base += 32) {
            const uint32_t i = base + tid;
            const uint32_t v = i < n_ctas ? a.cta_count[i] : 0u;
            const uint32_t before = i < c ? v : 0u;
            uint32_t sb = before, sv = v;
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) { sb += __shfl_xor_sync(0xFFFFFFFFu, sb, d); sv += __shfl_xor_sync(0xFFFFFFFFu, sv, d); }
            excl += sb; total += sv;
        }
        if (tid == 0) { s_excl = excl; if (c == 0) *a.n_changed = total; }
    }
    __syncthreads();
    uint32_t t_lo, t_hi;
    chunk_tiles(n_tiles, n_ctas, c, &t_lo, &t_hi);
    const uint32_t cnt = a.cta_count[c];
    const uint32_t* __restrict__ src = a.stage_idx + (size_t)t_lo * kTile32;
    uint32_t* __restrict__ dst = a.changed_idx + s_excl;
    for (uint32_t i = tid; i < cnt; i += blockDim.x) dst[i] = src[i];
}

template <int ITEMS>
__global__ void __launch_bounds__(kStThreads) k_status_diff(StatusArgs a) {
    extern __shared__ __align__(16) uint32_t s_rec[];  // tile_recs rows of (stride/4 + 1) words
    __shared__ uint32_t s_tile, s_excl;
    __shared__ uint32_t s_wcnt[ITEMS * (kStThreads / 32)];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(a.tile_counter, 1u);  // tile ids in scheduling order: look-back cannot starve
    __syncthreads();
    const uint32_t tile = s_tile;
    constexpr uint32_t kTileRecs = kStThreads * ITEMS;
    const uint32_t rec0 = tile * kTileRecs;
    const uint32_t nrec = min(kTileRecs, a.N - rec0);
    const uint32_t wps = a.stride >> 2, row = wps + 1, u16 = a.stride >> 4;  // words / padded row / 16B units per slot

    // coalesced copy of the tile into padded rows (bank-conflict-free per-thread row walks)
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(a.records + (size_t)rec0 * a.stride);
    const uint32_t units = nrec * u16;
    for (uint32_t u = tid; u < units; u += kStThreads) {
        const uint4 v = __ldcs(src + u);  // streamed once
        const uint32_t r = u / u16, part = u - r * u16;
        uint32_t* d = s_rec + r * row + part * 4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();

    bool changed[ITEMS];
    uint32_t bal[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t lr = (uint32_t)k * kStThreads + tid;
        changed[k] = false;
        if (lr < nrec) {
            const uint32_t* w = s_rec + lr * row;
            uint32_t len = w[0] & 0xFFu;
            len = min(len, a.stride - 1);
            const u64 h = xxh64_slot(w, len);
            const u64 prev = a.hash_prev[rec0 + lr];
            changed[k] = (prev == 0ull) || (h != prev);  // 0 = never seen (after reset)
            if (changed[k]) a.hash_prev[rec0 + lr] = h;  // kubelet.go:875-880: state replaced only on change
            if (a.hash_out) a.hash_out[rec0 + lr] = h;
        }
        bal[k] = __ballot_sync(0xFFFFFFFFu, changed[k]);
        if (lane == 0) s_wcnt[k * (kStThreads / 32) + warp] = __popc(bal[k]);
    }
    if (a.changed_idx != nullptr) emit_changed<ITEMS>(a, tile, rec0, nrec, changed, bal, s_wcnt, &s_excl);
}

int launch_status_diff(const StatusArgs& a, cudaStream_t st) {
    if (a.N == 0) {
        if (a.n_changed) RPK_CUDA(cudaMemsetAsync(a.n_changed, 0, sizeof(uint32_t), st));
        return 0;
    }
    const int items = items_for_stride(a.stride);
    const uint32_t tiles = status_tiles(a.N, a.stride);
    if (a.stride != 32) {
        RPK_CUDA(cudaMemsetAsync(a.tile_state, 0, (size_t)tiles * sizeof(u64), st));
        RPK_CUDA(cudaMemsetAsync(a.tile_counter, 0, sizeof(uint32_t), st));
    }
    if (a.stride == 32) {
        int dev = 0, sms = 148;
        RPK_CUDA(cudaGetDevice(&dev));
        RPK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const uint32_t grid = tiles < (uint32_t)(kCtasPerSm32 * sms) ? tiles : (uint32_t)(kCtasPerSm32 * sms);
        static thread_local int attr_dev = -1;  // the attribute is per device: set it once per (thread, device)
        if (attr_dev != dev) {
            RPK_CUDA(cudaFuncSetAttribute(k_status_diff32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * sizeof(Stage32))));
            attr_dev = dev;
        }
        StatusArgs b = a;
        if (a.changed_idx == nullptr) b.stage_idx = nullptr;
        k_status_diff32<<<grid, kStThreads, 2 * sizeof(Stage32), st>>>(b, tiles);
        int launches = 1;
        if (a.changed_idx != nullptr) { k_status_compact<<<grid, 256, 0, st>>>(b, tiles, grid); ++launches; }
        RPK_CUDA(cudaGetLastError());
        return launches;
    }
    const size_t smem = (size_t)kStThreads * items * (a.stride + 4);
    switch (items) {
        case 4:
            RPK_CUDA(cudaFuncSetAttribute(k_status_diff<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_status_diff<4><<<tiles, kStThreads, smem, st>>>(a);
            break;
        case 2:
            RPK_CUDA(cudaFuncSetAttribute(k_status_diff<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_status_diff<2><<<tiles, kStThreads, smem, st>>>(a);
            break;
        default:
            RPK_CUDA(cudaFuncSetAttribute(k_status_diff<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_status_diff<1><<<tiles, kStThreads, smem, st>>>(a);
            break;
    }
    RPK_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace rpk
