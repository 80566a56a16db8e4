// Slice sort: the one-pass front end of the whole-stream encoders on time-ordered input.
//
// A stream that is ordered in time (src/io/psee_loader.py hands out nothing else) has every
// bin of a window as ONE contiguous index range, so the counting sort by (sensor tile, bin)
// of bucketing.cu -- count pass, global scans, scatter pass -- collapses into a purely local
// job: find the bins' index ranges by bisection, cut every bin into slices of at most 8188
// events, let one CTA sort a slice by tile in shared memory, write it back with one TMA bulk
// store and publish the slice's tile offsets in a table laid out tile-major (a tile reads its
// runs as one contiguous row).  A tile kernel then reads its records as one short run per slice
// (TMA bulk copies into a ring of stages).  Per event:
// 9 bytes read, 4 written -- no second pass over the events, no global atomics on the data path.
#pragma once

#include "stream_common.cuh"

namespace evrep {

constexpr int kSortThreads = 512;
constexpr int kSortPerThread = 16;                 // events held in registers across the rank / place phases
constexpr int kSliceCap = kSortThreads * kSortPerThread;   // 8192 event slots per slice (a 16-byte aligned window)
constexpr int kSliceMax = kSliceCap - 4;           // events per slice: the window starts at a multiple of 4 events
constexpr uint32_t kNullRecord = 0xFFFFFFFFu;      // run padding (runs are multiples of 4 records = 16 bytes)
constexpr uint32_t kPackedDMax = 32767;            // largest offset the packed (count:9 | sum:23) accumulators are safe for
constexpr int kMaxSliceTiles = 2040;               // tiles per grid (the offset table row and the sort's histogram)
constexpr int kMaxSlicePixels = 4096;              // pixels per tile: 2 * P cells fit the 14-bit cell field below the null record

// Record written by the sort:  [ d:18 | cell:14 ],  cell = p * P + local pixel,  d = t - bin start (saturated).

enum : uint32_t {
    kBinFirst = 1u << 0,         // bin 0 of its window (timestamps before the window start are clamped, not strays)
    kBinLast = 1u << 1,          // last bin of its window (timestamps past its end are clamped, not strays)
    kBinAny = 1u << 30,          // set by the sort: some event of the bin landed on the grid (any tile)
    kBinBigD = 1u << 31,         // set by the sort: some record of the bin has d > kPackedDMax
};

struct BinDesc {                 // 32 bytes, one per global bin
    uint32_t lo, hi;             // events [lo, hi)
    int64_t t0;                  // bin start time
    uint32_t first_slice;        // slices first_slice .. first_slice + ceil((hi - lo) / kSliceMax) - 1
    uint32_t flags;              // kBinFirst | kBinLast
    uint32_t dyn;                // kBinAny | kBinBigD, OR-ed in by the sort CTAs
    uint32_t win;                // after slice_layout_kernel: first record of the bin's region in the bin-major layout
};
static_assert(sizeof(BinDesc) == 32, "BinDesc layout");

__host__ __device__ __forceinline__ uint32_t slice_parts(uint32_t lo, uint32_t hi) {
    return (hi - lo + (uint32_t)kSliceMax - 1u) / (uint32_t)kSliceMax;
}

struct SlicePlan {               // device pointers into the scratch buffer
    uint32_t* status;            // [0] events outside the bin their position implies (input not time-ordered), [1] n_slices
    const int64_t* w_begin;
    const int64_t* w_end;
    const int64_t* w_start;
    const int32_t* w_nbins;
    const int32_t* w_binbase;    // [n_windows + 1]
    const int32_t* w_fresh;
    BinDesc* bins;               // [TB]
    uint32_t* slice_bin;         // [max_slices] global bin of a slice
    uint32_t* runs;              // [n_tiles][pitch]: run of tile i in slice s = records 4 * (v & 0xFFFF) .. 4 * (v >> 16) of the slice
    uint32_t* records;           // [max_slices][slice_stride]
    int n_windows, TB, n_tiles, P, H, W, pitch, slice_stride;
    int max_slices;
    uint32_t abin;
    uint32_t tile_mul;           // pix / P == umulhi(pix, tile_mul) for every pixel of the grid, or 0 (use div_P)
    FastDiv div_P;
};

struct SliceLayout {
    int P, n_tiles, pitch, slice_stride, ctas_per_sm;
    int64_t max_slices;
    int64_t o_status, o_wbegin, o_wend, o_wstart, o_wnbins, o_wbinbase, o_wfresh, meta_bytes;
    int64_t o_bins, o_slicebin, o_runs, o_records, total;
};

// Shared memory a tile CTA needs for a tile of P pixels (K slots), per kernel family.
typedef size_t (*TileSmemFn)(int P, int K);

// Picks the tile size: `ctas_per_sm` resident CTAs per SM, tiles as large as their shared memory
// allows, then shrunk so that the grid is covered in the same number of waves with equal tiles.
int choose_tile(int H, int W, int K, TileSmemFn smem_of, int ctas_per_sm, int& P, int& n_tiles);

int make_slice_layout(int64_t n_events, int n_windows, int64_t TB, int H, int W, int P, int n_tiles, SliceLayout& L);

// Validates the window list, uploads the window tables, finds the bins' event ranges, lays out the
// slices and sorts them.  On return (stream order) `sp` describes the sorted slices.
int prepare_slices(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                   const evrep_taf_window* windows_host, int n_windows, int abin, int H, int W, int P, int n_tiles,
                   const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                   void* scratch, int64_t scratch_bytes, cudaStream_t st, SlicePlan& sp, SliceLayout& L);

// Output of the sort's bin-major mode (see slice_sort_kernel): what the register-resident TAF tile kernel reads.
struct BinMajorOut {
    uint16_t* cnt16;         // [max_slices][pitch16] tile counts of every slice
    uint32_t* bin_done;      // [TB] slices of the bin that have published their counts (zero on entry)
    uint32_t* counts;        // [n_tiles][TB] padded run length of (tile, bin): input of taf_scan_rows_kernel (zero on entry)
    uint32_t* src;           // [n_tiles][TB] first record of the (tile, bin) run in `records`
    uint32_t* bin_any;       // [TB] (zero on entry)
    uint32_t* records;
    int pitch16, TB;
};
constexpr int kBinMajorMaxParts = 128;             // slices a bin may have in that mode: they must all be resident at once

// Bins by bisection, slice layout, the sort (slices, or bin-major when `bm` is given), on `st`.
int run_slice_front(const SoA& ev, const SlicePlan& sp, int64_t n_events, int sensor_h, int sensor_w, const BinMajorOut* bm,
                    cudaStream_t st);

// ---- record feed of the tile kernels ------------------------------------------------------------
// The producer warps of a CTA walk the windows and bins of the plan (all of them the same way), gather the
// tile's run of every slice of a bin with TMA bulk copies -- each warp its share of the runs -- into a ring
// of stages in shared memory and hand the workers
// one *segment* (the records of one bin, or a piece of it) per stage through a descriptor guarded by
// mbarriers: FULL[stage] (transaction bytes + the producer's arrival), EMPTY[stage] (one arrival
// per worker warp once it has read the records).
constexpr int kFeedStages = 6;
constexpr int kFeedStageRecords = 512;             // 2 KB per stage

enum : uint32_t {
    kSegBinEnd = 1u << 0,        // the segment completes its bin: age by `age_inc`, then push the active cells
    kSegWide = 1u << 1,          // the bin holds offsets beyond kPackedDMax: two-word accumulators
    kSegReset = 1u << 2,         // control: reset the state (fresh window)
    kSegEmit = 1u << 3,          // control: emit the window tensor
    kSegDone = 1u << 4,          // control: last descriptor of the launch
};

struct SegDesc {                 // 16 bytes
    uint32_t n_rec;              // records in the stage (0 for a control descriptor)
    uint32_t flags;
    uint32_t age_inc;            // ageing steps: bins that were non-empty somewhere on the grid
    uint32_t arg;                // kSegBinEnd: global bin; kSegEmit: window index
};

struct FeedSmem {                // offsets inside a CTA's dynamic shared memory (bytes)
    int ring, desc, full, empty, total;
    __host__ __device__ FeedSmem(int base) {
        int o = (base + 127) / 128 * 128;
        ring = o;  o += kFeedStages * kFeedStageRecords * 4;
        desc = o;  o += kFeedStages * (int)sizeof(SegDesc);
        full = o;  o += kFeedStages * 8;
        empty = o; o += kFeedStages * 8;
        total = o;
    }
};

__device__ __forceinline__ void mbar_add_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// 16-byte asynchronous copy global -> shared (LDGSTS) and its completion hooked to an mbarrier: the arrival fires once all
// earlier cp.async of the executing thread have landed, and counts as one of the barrier's expected arrivals (.noinc).
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
constexpr int kFeedWarps = 4;                      // producer warps (one warpgroup): a bulk copy is issued lane by lane, so the runs of a bin are dealt out to four warps
constexpr int kFeedFullCount = kFeedWarps;         // FULL[stage]: one arrival per producer warp (+ the TMA transaction bytes)

// Producer side of the feed (state of one producer warp).  `seq` counts descriptors; stage = seq % kFeedStages.
struct FeedProducer {
    uint32_t* ring;
    SegDesc* desc;
    uint64_t* full;
    uint64_t* empty;
    uint32_t seq;
    int lane, pw;                // lane, and which of the kFeedWarps producer warps this is

    __device__ __forceinline__ void init(unsigned char* smem, const FeedSmem& fs) {
        ring = reinterpret_cast<uint32_t*>(smem + fs.ring);
        desc = reinterpret_cast<SegDesc*>(smem + fs.desc);
        full = reinterpret_cast<uint64_t*>(smem + fs.full);
        empty = reinterpret_cast<uint64_t*>(smem + fs.empty);
        seq = 0;
        lane = threadIdx.x & 31;
        pw = (threadIdx.x >> 5) % kFeedWarps;
    }
    __device__ __forceinline__ int slot() const { return (int)(seq % kFeedStages); }
    // wait until the workers have drained the stage the next descriptor goes to
#ifdef EVREP_TAF_TIMING
    long long t_acquire = 0;
    __device__ __forceinline__ void acquire() {
        const long long t0 = clock64();
        mbar_wait(empty + slot(), ((seq / kFeedStages) & 1u) ^ 1u);
        t_acquire += clock64() - t0;
    }
#else
    __device__ __forceinline__ void acquire() { mbar_wait(empty + slot(), ((seq / kFeedStages) & 1u) ^ 1u); }
#endif
    // publish the descriptor of the acquired stage (all copies into it have been issued)
    __device__ __forceinline__ void publish(uint32_t n_rec, uint32_t flags, uint32_t age_inc, uint32_t arg) {
        __syncwarp();
        if (lane == 0) {
            if (pw == 0) desc[slot()] = SegDesc{n_rec, flags, age_inc, arg};
            mbar_arrive(full + slot());
        }
        ++seq;
    }
    __device__ __forceinline__ void control(uint32_t flags, uint32_t age_inc, uint32_t arg) {
        acquire();
        publish(0u, flags, age_inc, arg);
    }
    // ---- the tile's row of the run table, read through a window of 64 consecutive slices (two registers per lane) ----
    const uint32_t* row;         // runs of this tile, one entry per slice
    uint32_t win_base;           // slice index of lane 0 of `win_lo`
    uint32_t win_lo, win_hi;     // entries win_base + lane and win_base + 32 + lane (win_hi is the prefetch)
    uint32_t n_slices;

    __device__ __forceinline__ void open_row(const SlicePlan& sp, int tile) {
        row = sp.runs + (int64_t)tile * sp.pitch;
        n_slices = sp.status[1];
        rewind(0u);
    }
    // The window only slides forward; a reader that goes back (overlapping Event Volume spans) opens it again.
    __device__ __forceinline__ void rewind(uint32_t slice) {
        win_base = slice & ~31u;
        win_lo = win_base + lane < n_slices ? __ldg(row + win_base + lane) : 0u;
        win_hi = win_base + 32u + lane < n_slices ? __ldg(row + win_base + 32 + lane) : 0u;
    }
    // Entry of slice s0 + lane (0 beyond the window's reach or `count`): slices s0 .. s0 + 31 must start inside the window.
    __device__ __forceinline__ uint32_t entries(uint32_t s0, uint32_t count) {
        while (s0 >= win_base + 32u) {                   // slide: the prefetched half becomes current, the next one is requested
            win_base += 32u;
            win_lo = win_hi;
            win_hi = win_base + 32u + lane < n_slices ? __ldg(row + win_base + 32 + lane) : 0u;
        }
        const uint32_t k = s0 - win_base + lane;         // 0 .. 62
        const uint32_t from_lo = __shfl_sync(0xFFFFFFFFu, win_lo, k & 31u);
        const uint32_t from_hi = __shfl_sync(0xFFFFFFFFu, win_hi, k & 31u);
        return lane < count ? (k < 32u ? from_lo : from_hi) : 0u;
    }
    // All records of this tile in bin `gbin` -> one or more segments; returns false when the tile has none.
    // `age_inc` goes into the descriptor of the bin's last segment, or of every segment (`age_on_all`, the Event Volume
    // kernel keeps a time offset there).
    __device__ bool feed_bin(const SlicePlan& sp, uint32_t gbin, uint32_t first_slice, uint32_t parts, uint32_t dyn, uint32_t age_inc,
                             bool age_on_all = false);
};

__device__ __forceinline__ bool FeedProducer::feed_bin(const SlicePlan& sp, uint32_t gbin, uint32_t first_slice, uint32_t parts,
                                                       uint32_t dyn, uint32_t age_inc, bool age_on_all) {
    const uint32_t wide = (dyn & kBinBigD) ? kSegWide : 0u;
    // pass 1 (bins of more than 32 slices only): the bin's total, so that the last segment is known when it is built
    uint32_t total = 0;
    if (parts > 32u) {
        for (uint32_t c0 = 0; c0 < parts; c0 += 32u) {
            const uint32_t v = __ldg(row + first_slice + c0 + (c0 + lane < parts ? lane : 0u));
            total += __reduce_add_sync(0xFFFFFFFFu, c0 + lane < parts ? ((v >> 16) - (v & 0xFFFFu)) * 4u : 0u);
        }
    }
    uint32_t q0 = 0;                                    // records of the bin placed so far
    bool open = false;                                  // the stage holding record q0 is acquired and not yet published
    for (uint32_t c0 = 0; c0 < parts; c0 += 32u) {
        const uint32_t cnt = parts - c0 < 32u ? parts - c0 : 32u;
        uint32_t v;
        if (parts <= 32u) v = entries(first_slice, cnt);
        else v = c0 + lane < parts ? __ldg(row + first_slice + c0 + lane) : 0u;
        const uint32_t a = v & 0xFFFFu, b = v >> 16;
        const uint32_t len = (b - a) * 4u;
        uint32_t incl = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t chunk_total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (parts <= 32u) total = chunk_total;
        if (total == 0u) return false;
        if (chunk_total == 0u) continue;
        const uint32_t start = q0 + incl - len, end = q0 + incl;          // this lane's run, bin-relative
        const uint32_t* src = sp.records + (int64_t)(first_slice + c0 + lane) * sp.slice_stride + a * 4u;
        const uint32_t q1 = q0 + chunk_total;
        for (uint32_t sb = q0 / kFeedStageRecords * kFeedStageRecords; sb < q1; sb += kFeedStageRecords) {
            if (!open) { acquire(); open = true; }
            const uint32_t se = sb + kFeedStageRecords;
            const uint32_t lo = start > sb ? start : sb, hi = end < se ? end : se;
            if (lo < hi && (lane % kFeedWarps) == pw) {      // this warp's share of the runs
                const uint32_t bytes = (hi - lo) * 4u;
                mbar_add_tx(full + slot(), bytes);
                tma_load_1d(ring + slot() * kFeedStageRecords + (lo - sb), src + (lo - start), bytes, full + slot());
            }
            if (q1 >= se || q1 == total) {              // the stage is complete: full, or the bin ends in it
                const bool last = (se >= total);
                publish((last ? total : se) - sb, wide | (last ? kSegBinEnd : 0u), (last || age_on_all) ? age_inc : 0u, gbin);
                open = false;
            }
        }
        q0 = q1;
    }
    return true;
}

}  // namespace evrep
