// rpk_internal.cuh -- shared declarations of the rpk engine (not part of the C-ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "rpk.h"

namespace rpk {

constexpr uint32_t kNone = 0xFFFFFFFFu;      // "no feasible position"
constexpr uint32_t kGroups = 8;              // row groups: cloud view x (vcpu constrains) x (ram constrains)
constexpr int kWarpsPerCta = 8;              // select kernels: 256 threads
constexpr int kCtaThreads = kWarpsPerCta * 32;
constexpr uint32_t kChunk = 128;             // offers per warp-iteration (LDS.128 per lane)
constexpr uint32_t kSegPacked = 16384;       // packed words staged per CTA (64 KB); a power of two (see kPosBits)
constexpr uint32_t kPosBits = 14;            // log2(kSegPacked): position-in-segment bits of a "pos" packed word
constexpr uint32_t kSegWide = 4096;          // wide int4 views staged per CTA (64 KB)

// Layout of the packed offer word (sorted-by-price order, one u32 per offer), from bit `sh_ram` up:
//   [guard1][mem rank' : b1][guard2][vcpu rank : b2][guard3][ram rank : b3][position in segment : pos_bits]
// rank' = rank+1 (0 = offer not available in this cloud).  A pod's word holds its thresholds in the same
// fields with guards 0, so (offer - pod) keeps a field's guard iff offer_field >= pod_field.
// pos_bits = kPosBits when bits <= 32 - kPosBits: the low bits then carry the offer's position inside its
// 16384-offer segment and the kernel reduces key = (~d & guard) | (d & pos_mask) with a plain unsigned min
// (feasible keys are positions, infeasible ones are >= 2^pos_bits).  Otherwise pos_bits = 0 and the kernel
// tracks the position itself with predicated selects.
struct PackLayout {
    uint32_t guard;    // the three guard bits
    uint32_t sh_mem;   // shift of the mem field
    uint32_t sh_vcpu;  // shift of the vcpu field
    uint32_t sh_ram;   // shift of the ram field (= pos_bits)
    uint32_t pos_bits; // kPosBits or 0
    uint32_t bits;     // field + guard bits used (0 = table not packable -> generic kernel)
    // bit-sliced view (kernel kind 4): per 32-offer chunk, kBmStride words; word i < bm_words is the 32-bit
    // mask "offer j of the chunk passes threshold i" -- mem thresholds first (rank' >= t', t' = 1..D_mem+1, so
    // unavailable offers fail all of them), then vcpu thresholds 0..D_vcpu, then ram thresholds 0..D_ram.
    uint32_t bm_words; // D_mem+1 + D_vcpu+1 + D_ram+1 if that fits 64 words, else 0
    uint32_t bm_off_vcpu, bm_off_ram;
    uint32_t bm_stride; // words per chunk row: 32 (conflict-free) or 64 (up to 64 thresholds; words 32 apart share a bank)
    uint32_t no_fused;  // test hook (RPK_FORCE_KERNEL=bitmap_grouped): small batches also take the three-launch path
};
constexpr uint32_t kBmMaxStride = 64;
constexpr uint32_t kBmSegBytes = 65536;      // bit-sliced rows staged per CTA: 512 chunks (stride 32) or 256 (stride 64)
constexpr uint32_t kSmallBatch = 4096;       // rows: the host entry point's single-copy latency path

struct OfferView {       // one per cloud, all arrays in price-sorted order, length Gpad
    uint32_t* packed = nullptr;
    uint32_t* bitmap = nullptr;  // [Gpad/32][bm_stride]
    int4* wide = nullptr;     // (mem_gb, vcpu, ram_gb, offer index)
    double* price = nullptr;  // NaN when the offer is not available in this cloud / padding
    int32_t* perm = nullptr;  // offer index, -1 when unavailable / padding
};

struct SelectArgs {
    // pod columns of this shard (device); nullable ones follow the C-ABI defaults
    const int32_t* req_mem;
    const int32_t* req_vcpu;
    const int32_t* req_ram;
    const double* max_price;
    const uint8_t* cloud;
    uint32_t P;
    // offer views
    OfferView view[2];
    uint32_t G, Gpad;
    const int32_t* distinct[3];
    uint32_t D[3];
    PackLayout pk;
    // per-call scratch
    uint32_t* rw;        // [P] packed thresholds
    uint32_t* order;     // [kGroups][P] row indices per group
    uint32_t* pos;       // [P] best sorted position so far (atomicMin target)
    uint32_t* counts;    // [kGroups] rows per group          } zero between calls: the grid kernel's last CTAs
    uint32_t* done;      // [1] row tiles finished             } clear them again (finish_tile), so no memset is
    uint32_t* tile_ctr;  // [ntiles] CTAs arrived per row tile } launched per call
    // outputs: the shard's slice is written into every peer's full-length vector at row0
    int32_t* best_out[RPK_MAX_GPUS];
    int n_out;
    int self_out;        // which of best_out lives on this GPU: the select kernels store there only and k_gather_push copies
                         // the slice to the peers in 16-byte NVLink stores; -1 = unknown, the kernels store to every vector
    uint32_t row0;
    int32_t* top5;       // [P*5] local, nullable
    uint32_t tune_natural_order;  // tuning hook (RPK_TUNE=order=natural): row tiles in group order instead of heaviest first
};

struct StatusArgs {
    const uint8_t* records;
    uint32_t stride;
    uint32_t N;
    uint64_t* hash_prev;   // updated in place
    uint64_t* hash_out;    // nullable
    uint32_t* changed_idx; // nullable (seed)
    uint32_t* n_changed;   // nullable (seed)
    uint32_t idx_base;     // added to emitted indices (shard offset)
    unsigned long long* tile_state;  // [ntiles] look-back state (strides other than 32)
    uint32_t* tile_counter;
    uint32_t* stage_idx;             // [N] per-CTA ordered index segments (stride 32)
    uint32_t* cta_count;             // [<= 2 * SMs] changed slots per persistent CTA (stride 32)
};

// launchers (each returns the number of kernels it launched, or throws CudaError)
struct CudaError { cudaError_t err; const char* what; const char* file; int line; };
#define RPK_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) throw ::rpk::CudaError{e__, #x, __FILE__, __LINE__}; } while (0)

struct OfferIngest {   // raw device columns in, views out
    uint32_t G;
    const int32_t* mem; const int32_t* vcpu; const int32_t* ram;  // vcpu/ram never null here (zero-filled)
    const double* secure_price; const double* community_price; const uint8_t* flags;
};

struct DeviceState;
int launch_offer_ingest(DeviceState& ds, const OfferIngest& in, cudaStream_t st);
int launch_select(const SelectArgs& a, int rows_per_warp, cudaStream_t st);
uint32_t select_tiles_max(uint32_t P, int rows_per_warp);
int pick_rows_per_warp(uint32_t P, int sm_count);
int pick_rows_per_lane(uint32_t P, uint32_t G, int sm_count);
int launch_status_diff(const StatusArgs& a, cudaStream_t st);
struct PeerFenceArgs { uint32_t* flags[RPK_MAX_GPUS]; int n; int my_rank; uint32_t epoch; };
int launch_peer_fence(const PeerFenceArgs& a, cudaStream_t st);
uint32_t status_tiles(uint32_t N, uint32_t stride);

template <typename T>
struct DevBuf {  // grow-only device buffer
    T* p = nullptr;
    size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) RPK_CUDA(cudaFree(p));
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        RPK_CUDA(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct DeviceState {
    int dev = -1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // offer table
    uint32_t G = 0, Gpad = 0;
    bool offers_ready = false;
    DevBuf<int32_t> raw_mem, raw_vcpu, raw_ram; DevBuf<double> raw_sp, raw_cp; DevBuf<uint8_t> raw_flags;
    DevBuf<unsigned long long> sort_keys; DevBuf<uint32_t> sort_vals;
    DevBuf<uint32_t> v_packed[2]; DevBuf<uint32_t> v_bitmap[2]; DevBuf<int4> v_wide[2]; DevBuf<double> v_price[2]; DevBuf<int32_t> v_perm[2];
    DevBuf<int32_t> distinct[3]; DevBuf<uint32_t> dcount;
    uint32_t D[3] = {0, 0, 0};
    PackLayout pk = {};
    int force_kind = 0;  // RPK_FORCE_KERNEL (tests): 0 auto, 1 generic, 2 packed+select, 3 packed+pos, 4 bit-sliced
    // select scratch.  The host entry point pipelines row sub-batches over two lanes (stream + staging +
    // scratch each) so that the H2D of one sub-batch, the kernels of another and the D2H of a third overlap;
    // the device entry points use lane 0's scratch on the caller's stream.
    struct Lane {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        DevBuf<int32_t> p_req_mem, p_req_vcpu, p_req_ram; DevBuf<double> p_max_price; DevBuf<uint8_t> p_cloud;
        DevBuf<int32_t> top5;
        DevBuf<uint32_t> rw, order, pos, ctrs;
        bool ctrs_dirty = false;  // a select on this lane failed between its launches: re-zero the counters before the next one
        void release() {
            p_req_mem.release(); p_req_vcpu.release(); p_req_ram.release(); p_max_price.release(); p_cloud.release();
            top5.release(); rw.release(); order.release(); pos.release(); ctrs.release();
        }
    };
    Lane lane[2];
    DevBuf<int32_t> best_full;
    // small-batch (latency) path: one pinned staging block in, one out
    unsigned char* h_small = nullptr; DevBuf<unsigned char> d_small_in; DevBuf<int32_t> d_small_out;
    // status
    DevBuf<uint8_t> s_records; DevBuf<uint64_t> s_hash_prev, s_hash_out; DevBuf<uint32_t> s_changed, s_misc;
    DevBuf<unsigned long long> s_tile_state; DevBuf<uint32_t> s_stage_idx;
    uint32_t statusN = 0; bool status_sized = false;
};

}  // namespace rpk
