// Shared declarations of the whole-stream encoders (TAF and Event Volume): constants, the
// bucketing plan handed from the bucketing passes to the tile kernels, PTX wrappers for
// mbarrier / TMA bulk copies / named barriers, and the host-side front end.
#pragma once

#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace evrep {

constexpr int kTafThreads = 384;        // threads per tile CTA: 3 warps per SM sub-partition -> up to 168 registers
constexpr int kMaxSlots = 6;            // pixels per thread held in registers (6 x 384 = 2304 >= 2240)
constexpr int kMaxTilePixels = 2240;    // largest tile: see make_layout
constexpr int kChunkRecords = 1024;     // records per TMA bulk copy (4 KB)
constexpr int kStages = 8;              // ring depth (32 KB in flight per SM)
constexpr int kBatchBins = 16;          // bins whose offsets are staged in smem at once
constexpr int kMaxTiles = 2048;         // kLocalBins * kMaxTiles counters fit the 13-bit key of the scatter pass
constexpr int kLocalBins = 4;           // bins covered by a bucketing CTA's smem histogram
constexpr int kBucketThreads = 512;
constexpr int kBucketPerThread = 8;     // 4096 events per bucketing CTA
constexpr int kSavedThreads = 256;       // scatter pass over saved records: 256 threads x 16 records, 4 CTAs per SM
constexpr int kSavedCtasPerSm = 4;
constexpr int kScatterPerThread = 16;
constexpr int kScatterChunks = kSavedThreads * kScatterPerThread / (kBucketThreads * kBucketPerThread);   // count chunks per scatter chunk
static_assert(kScatterChunks >= 1 && kSavedThreads * kScatterPerThread % (kBucketThreads * kBucketPerThread) == 0, "scatter chunks are whole count chunks");
constexpr uint32_t kDMax = (1u << 18) - 1;
constexpr float kTafInit = -6000.0f;    // generate_taf.py:207-209

// Exact unsigned division by a runtime constant (Granlund-Montgomery).
struct FastDiv {
    uint32_t mul, sh1, sh2, d;
    static FastDiv make(uint32_t d) {
        FastDiv f;
        f.d = d;
        uint32_t l = 0;
        while ((1ull << l) < d) ++l;
        f.mul = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
        f.sh1 = l < 1 ? l : 1;
        f.sh2 = l > 0 ? l - 1 : 0;
        return f;
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const {
        uint32_t t1 = __umulhi(mul, n);
        return (t1 + ((n - t1) >> sh1)) >> sh2;
    }
};

struct Batch {            // <= kBatchBins consecutive bins of one window
    int gbin0;            // first global bin
    int nb;               // bins in this batch
    int flags;            // bit0: reset state before; bit1: emit window tensor after
    int win;              // window index (selects the output slot)
};

struct StreamPlan {       // device pointers into the scratch buffer
    const int64_t* w_begin;
    const int64_t* w_end;
    const int64_t* w_start;
    const int32_t* w_nbins;
    const int32_t* w_binbase;
    const Batch* batches;
    uint32_t* counts;     // [n_tiles][TB]  histogram, then scatter cursors
    uint32_t* bin_any;    // [TB]
    uint32_t* off_rel;    // [n_tiles][TB+1] record offsets relative to the tile's list
    uint32_t* tile_total; // [n_tiles]
    uint32_t* tile_base;  // [n_tiles+1] (multiples of 4 records: 16-byte aligned lists)
    uint32_t* records;    // [n_events + 4 n_tiles]
    uint32_t* tile_bits;  // [n_tiles][n_batches]: per batch, bit b = tile has records in bin b, bit 16+b = bin is non-empty anywhere
    uint32_t* saved_rec;  // [events] record of every event as classified by the count pass (event order)
    uint16_t* saved_key;  // [events] its shared-memory counter (local bin * n_tiles + tile), kKeyFar or kKeyDropped
    int n_windows, n_batches, TB, n_tiles, P, H, W;
    FastDiv div_abin, div_P;
    uint32_t abin;
    uint32_t tile_mul;    // pix / P == umulhi(pix, tile_mul) for every pixel of the grid, or 0 (use div_P)
};

// ---- mbarrier / TMA bulk-copy primitives (PTX) --------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {      // read-only tables: free to be scheduled
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// Shared-memory accesses by 32-bit shared address: the compiler emits LDS / STS / ATOMS with 32-bit address arithmetic.
__device__ __forceinline__ uint32_t sld_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t sld_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 sld_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 sld_v4f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sst_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sst_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sst_u8(uint32_t addr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sst_v4f(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t satom_add(uint32_t addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void sred_add(uint32_t addr, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t satom_exch(uint32_t addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
// Pulls `bytes` (a multiple of 16) of 16-byte aligned global memory into L2 without waiting for them.
__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {
    if (reinterpret_cast<uintptr_t>(src) & 15) return;      // a hint only: skipped when the range is not aligned
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- warp-specialised tile kernel: roles, ring and barriers
constexpr int kProducerThreads = 128;   // the producer warpgroup: 3 accumulate warps + 1 store warp
constexpr int kAccumThreads = 96;
constexpr int kStoreThreads = 32;
constexpr int kConsumerThreads = 384;
constexpr int kWsThreads = kProducerThreads + kConsumerThreads;
constexpr int kWsChunkRecords = 512;    // 2 KB TMA bulk copies
constexpr int kWsStages = 8;            // 16 KB ring = a flat circular buffer of 4096 records
constexpr int kWsRing = kWsChunkRecords * kWsStages;
constexpr int kWsPrefetch = 8;          // chunks (16 KB, one ring) the store warp prefetches into L2 ahead of the ring
static_assert((kWsRing & (kWsRing - 1)) == 0, "ring size must be a power of two");

enum : int { kBarFull0 = 1, kBarFull1 = 2, kBarEmpty0 = 3, kBarEmpty1 = 4, kBarProducers = 6, kBarStaged = 7, kBarStageFree = 8 };

__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct TileSmemWS {
    int ring, acc, stage, bars, feed_p, total;
    static constexpr int kFeedBytes = 2 * (kBatchBins + 1) * 4 + 2 * kBatchBins * 4 + 2 * (int)sizeof(Batch) + 8;
    __host__ __device__ TileSmemWS(int P, int K) {
        int o = 0;
        ring = o;   o += kWsStages * kWsChunkRecords * 4;
        acc = o;    o += 2 * 2 * P * (int)sizeof(uint2);            // two buffers of {n, sum d} per (pixel, polarity)
        stage = o;  o += 2 * K * P * 4;                             // [2K][P] output staging
        bars = o;   o += 96;            // 8 record chunks, STAGED, STAGE_FREE, the drained-chunks word
        feed_p = o; o += (kFeedBytes + 15) / 16 * 16;
        total = o;
    }
};

// Role-local view of the batch list: offsets / flags of the current batch in shared memory,
// the next batch prefetched into registers (see taf_tile_kernel).
struct BatchFeed {
    uint32_t* s_off;       // [2][kBatchBins + 1]
    uint32_t* s_any;       // [2][kBatchBins]
    Batch* s_meta;         // [2]
    const uint32_t* my_off;
    const StreamPlan* pl;
    int rtid, bar_id, nthreads;
    Batch nmeta, nnmeta;
    uint32_t pre_off, pre_any;

    __device__ __forceinline__ void init(unsigned char* base, const StreamPlan* plan, const uint32_t* tile_off, int role_tid,
                                         int barrier_id, int role_threads) {
        s_off = reinterpret_cast<uint32_t*>(base);
        s_any = s_off + 2 * (kBatchBins + 1);
        s_meta = reinterpret_cast<Batch*>(s_any + 2 * kBatchBins);
        my_off = tile_off; pl = plan; rtid = role_tid; bar_id = barrier_id; nthreads = role_threads;
        const Batch m0 = pl->batches[0];
        if (rtid == 0) s_meta[0] = m0;
        if (rtid <= m0.nb) s_off[rtid] = my_off[m0.gbin0 + rtid];
        if (rtid < m0.nb) s_any[rtid] = pl->bin_any[m0.gbin0 + rtid];
        nmeta = pl->batches[pl->n_batches > 1 ? 1 : 0];
        named_sync(bar_id, nthreads);
    }
    __device__ __forceinline__ Batch begin(int j) {          // start of batch j: issue the prefetches
        pre_off = 0; pre_any = 0; nnmeta = nmeta;
        if (j + 1 < pl->n_batches) {
            if (rtid <= nmeta.nb) pre_off = my_off[nmeta.gbin0 + rtid];
            if (rtid < nmeta.nb) pre_any = pl->bin_any[nmeta.gbin0 + rtid];
            if (j + 2 < pl->n_batches) nnmeta = pl->batches[j + 2];
        }
        return s_meta[j & 1];
    }
    __device__ __forceinline__ void publish(int j) {         // store batch j+1; the caller supplies the barrier
        if (j + 1 < pl->n_batches) {
            const int nb = (j & 1) ^ 1;
            if (rtid == 0) s_meta[nb] = nmeta;
            if (rtid <= nmeta.nb) s_off[nb * (kBatchBins + 1) + rtid] = pre_off;
            if (rtid < nmeta.nb) s_any[nb * kBatchBins + rtid] = pre_any;
            nmeta = nnmeta;
        }
    }
    __device__ __forceinline__ void end(int j) {             // end of batch j: publish batch j+1
        publish(j);
        named_sync(bar_id, nthreads);
    }
};

// ---- host side (bucketing.cu) ---------------------------------------------------------------
struct Layout {
    int P, n_tiles, slots;
    int64_t o_wbegin, o_wend, o_wstart, o_wnbins, o_wbinbase, o_batches, meta_bytes;
    int64_t o_counts, o_binany, o_offrel, o_tiletotal, o_tilebase, o_tilebits, o_origins, o_records, o_savedrec, o_savedkey, total;
    int n_batches_max;
};

int make_layout(int64_t n_events, int n_windows, int64_t TB, int H, int W, int n_batches, Layout& L, int tiles_per_sm = 1);
int64_t batches_upper_bound(int n_windows, int64_t TB);
// Small host table -> device memory as kernel arguments (no copy engine involved; see prepare_stream).
int upload_words(const uint32_t* host, int64_t n_words, uint32_t* dst, cudaStream_t st);

// Validates the window list, uploads the window / batch tables and runs the bucketing passes.
// On return `pl` describes the bucketed records of every (tile, bin).
int prepare_stream(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                   const evrep_taf_window* windows_host, int n_windows, int abin, int H, int W,
                   const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                   void* scratch, int64_t scratch_bytes, cudaStream_t st, StreamPlan& pl, Layout& L,
                   int tiles_per_sm = 1);

// Time-ordered input: one-pass bin-major sort in front of the same tile kernels (bucketing.cu).
int64_t binmajor_scratch_bytes(int64_t n_events, int n_windows, int64_t TB, int H, int W);
int64_t binmajor_status_offset(int64_t n_events, int n_windows, int64_t TB, int H, int W);
int prepare_stream_binmajor(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                            const evrep_taf_window* windows_host, int n_windows, int abin, int H, int W,
                            const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                            void* scratch, int64_t scratch_bytes, cudaStream_t st, StreamPlan& pl, Layout& L,
                            const uint32_t*& src, uint32_t*& status);

}  // namespace evrep
