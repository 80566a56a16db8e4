// Temporal Active Focus over whole streams (generate_taf.py:160-238) in two steps.
//
//  (1) bucketing: a counting sort of the time-ordered events by (sensor tile, 10 ms bin)
//      into packed 4-byte records  [ d:18 | local pixel:13 | p:1 ],  d = t - bin start.
//      Three small passes: count (shared-memory histograms per 4096-event chunk),
//      scan (per tile row, then across tiles), scatter.
//  (2) the persistent tile kernel: one CTA per sensor tile (a contiguous range of <= 2304
//      pixels, chosen so that there are <= #SM tiles when possible).  The CTA keeps the
//      tile's FIFO state -- 2K floats per pixel -- in REGISTERS for the whole stream,
//      streams its own record list through a ring of TMA bulk copies (cp.async.bulk +
//      mbarrier), accumulates (count, sum d) per cell with shared-memory atomics, applies
//      the FIFO push / ageing rule bin by bin, and writes the [2K,H,W] tensor (and the
//      state) once per window with coalesced stores.  Tiles never talk to each other:
//      the only cross-tile fact, "did any pixel see an event in this bin"
//      (generate_taf.py:40-41), is a per-bin flag produced by the bucketing pass.
//
// HBM-bound byte/float work: no tensor cores.  Sums of d are exact integers, so the
// result does not depend on the order in which records are accumulated.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace evrep {

constexpr int kTafThreads = 384;        // threads per tile CTA: 3 warps per SM sub-partition -> up to 168 registers
constexpr int kMaxSlots = 6;            // pixels per thread held in registers (6 x 384 = 2304 >= 2240)
constexpr int kChunkRecords = 1024;     // records per TMA bulk copy (4 KB)
constexpr int kStages = 8;              // ring depth (32 KB in flight per SM)
constexpr int kBatchBins = 16;          // bins whose offsets are staged in smem at once
constexpr int kMaxTiles = 2048;         // kLocalBins * kMaxTiles counters fit the 13-bit key of the scatter pass
constexpr int kLocalBins = 4;           // bins covered by a bucketing CTA's smem histogram
constexpr int kBucketThreads = 512;
constexpr int kBucketPerThread = 8;     // 4096 events per bucketing CTA
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
    int n_windows, n_batches, TB, n_tiles, P, H, W;
    FastDiv div_abin, div_P;
    uint32_t abin;
};

// Window descriptor staged in shared memory by the bucketing kernels.
struct WinInfo {
    int64_t begin, end, start;
    int nbins, binbase;
};

__device__ __forceinline__ WinInfo load_window(const StreamPlan& pl, int w) {
    WinInfo wi;
    wi.begin = pl.w_begin[w]; wi.end = pl.w_end[w]; wi.start = pl.w_start[w];
    wi.nbins = pl.w_nbins[w]; wi.binbase = pl.w_binbase[w];
    return wi;
}

// First window whose event range ends after event index i.
__device__ __forceinline__ int first_window(const StreamPlan& pl, int64_t i) {
    int lo = 0, hi = pl.n_windows;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(pl.w_end + mid) > i) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Bin of a timestamp inside window `wi` and the offset d from the bin start:
// z = clamp(floor((t - start) / abin), 0, nbins - 1) -- inclusive edges, later bin wins
// (generate_taf.py:201-202) -- and d = t - (start + z abin), saturated to 18 bits.
__device__ __forceinline__ void bin_of(const StreamPlan& pl, const WinInfo& wi, uint32_t t, uint32_t& z, uint32_t& d) {
    const int64_t dt = (int64_t)t - wi.start;
    z = 0; d = 0;
    if (dt > 0) {
        const uint32_t u = dt > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)dt;
        z = pl.div_abin.div(u);
        if (z > (uint32_t)(wi.nbins - 1)) z = wi.nbins - 1;
        const uint32_t rem = u - z * pl.abin;
        d = rem > kDMax ? kDMax : rem;
    }
}

// Block-wide exclusive scan of n <= kBucketThreads * 16 shared-memory counters.
// Returns the total.  `tmp` holds one word per warp (+1).
__device__ __forceinline__ uint32_t block_exclusive_scan(const uint32_t* in, uint32_t* out, int n, uint32_t* tmp) {
    const int per = (n + kBucketThreads - 1) / kBucketThreads;
    const int lo = threadIdx.x * per, hi = min(lo + per, n);
    uint32_t mine = 0;
    for (int i = lo; i < hi; ++i) mine += in[i];
    uint32_t incl = mine;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) tmp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < kBucketThreads / 32 ? tmp[lane] : 0u, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= o) wi += v;
        }
        if (lane < kBucketThreads / 32) tmp[lane] = wi - w;
        if (lane == kBucketThreads / 32 - 1) tmp[kBucketThreads / 32] = wi;
    }
    __syncthreads();
    uint32_t run = tmp[wid] + incl - mine;
    for (int i = lo; i < hi; ++i) { const uint32_t c = in[i]; out[i] = run; run += c; }
    const uint32_t total = tmp[kBucketThreads / 32];
    __syncthreads();                     // `out` is complete (and `tmp` reusable) for every thread
    return total;
}

// Per-chunk prologue of the bucketing passes, computed once by a tiny kernel: the window of
// the chunk's first event, the first global bin the chunk can touch and whether the whole
// chunk lies inside that window (the fast path).
struct ChunkOrigin {
    WinInfo win;
    int w0, gb0, single, pad;
};

__global__ void __launch_bounds__(256)
taf_chunk_origin_kernel(SoA ev, StreamPlan pl, int64_t ev_first, int64_t ev_last, int n_chunks, ChunkOrigin* __restrict__ origins) {
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk >= n_chunks) return;
    const int64_t c0 = ev_first + (int64_t)chunk * (kBucketThreads * kBucketPerThread);
    const int64_t c1 = min(c0 + kBucketThreads * kBucketPerThread, ev_last);
    ChunkOrigin o;
    o.w0 = first_window(pl, c0);
    o.gb0 = 0; o.single = 0; o.pad = 0;
    o.win.begin = o.win.end = o.win.start = 0; o.win.nbins = 0; o.win.binbase = 0;
    if (o.w0 < pl.n_windows) {
        const WinInfo wi = load_window(pl, o.w0);
        o.win = wi;
        o.single = (c0 >= wi.begin && c1 <= wi.end && wi.nbins > 0) ? 1 : 0;
        const int64_t i = c0 > wi.begin ? c0 : wi.begin;
        o.gb0 = wi.binbase;
        if (i < c1 && i < wi.end && wi.nbins > 0) {
            uint32_t z, d;
            bin_of(pl, wi, ev.t[i], z, d);
            o.gb0 += (int)z;
        }
    }
    origins[chunk] = o;
}

// Shared-memory carve-up of the bucketing kernels.
struct BucketSmem {
    int lutx, luty, hist, loff, gbase, sorted, skey, total;
    __host__ __device__ BucketSmem(int lut_w, int lut_h, int nh, bool scatter) {
        int o = 0;
        lutx = o;  o += (lut_w * 2 + 15) / 16 * 16;
        luty = o;  o += (lut_h * 2 + 15) / 16 * 16;
        hist = o;  o += nh * 4;
        loff = o;  o += scatter ? nh * 4 : 0;
        gbase = o; o += scatter ? nh * 4 : 0;
        sorted = o; o += scatter ? kBucketThreads * kBucketPerThread * 4 : 0;
        skey = o;  o += scatter ? kBucketThreads * kBucketPerThread * 2 : 0;
        total = (o + 15) / 16 * 16;
    }
};

// Bucketing passes.  Persistent CTAs walk 4096-event chunks of the time-ordered stream.
//  count   (kScatter = false): per-chunk shared-memory histogram over (local bin, tile),
//          flushed with one global atomic per non-empty counter; sets the per-bin flags.
//  scatter (kScatter = true):  the same histogram with ranks, a block scan, one global
//          reservation per non-empty counter, then the chunk's records are ordered in shared
//          memory so that each (tile, bin) run is written with consecutive addresses.
template <bool kScatter>
__global__ void __launch_bounds__(kBucketThreads, 2)
taf_bucket_kernel(SoA ev, StreamPlan pl, int64_t ev_first, int64_t ev_last, int n_chunks, int lut_w, int lut_h,
                  const ChunkOrigin* __restrict__ origins, int vec_ok) {
    extern __shared__ __align__(16) unsigned char bsm[];
    const int nh = kLocalBins * pl.n_tiles;
    const bool use_lut = ev.xmap != nullptr && ev.ymap != nullptr;
    const BucketSmem lay(use_lut ? lut_w : 0, use_lut ? lut_h : 0, nh, kScatter);
    uint16_t* s_lutx = reinterpret_cast<uint16_t*>(bsm + lay.lutx);
    uint16_t* s_luty = reinterpret_cast<uint16_t*>(bsm + lay.luty);
    uint32_t* hist = reinterpret_cast<uint32_t*>(bsm + lay.hist);
    uint32_t* loff = reinterpret_cast<uint32_t*>(bsm + lay.loff);
    uint32_t* gbase = reinterpret_cast<uint32_t*>(bsm + lay.gbase);
    uint32_t* sorted = reinterpret_cast<uint32_t*>(bsm + lay.sorted);
    uint16_t* skey = reinterpret_cast<uint16_t*>(bsm + lay.skey);
    __shared__ uint32_t s_tmp[kBucketThreads / 32 + 1];

    if (use_lut) {
        for (int i = threadIdx.x; i < lut_w; i += kBucketThreads) s_lutx[i] = ev.xmap[i];
        for (int i = threadIdx.x; i < lut_h; i += kBucketThreads) s_luty[i] = ev.ymap[i];
    }
    const uint32_t W = pl.W, H = pl.H;

    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int64_t c0 = ev_first + (int64_t)chunk * (kBucketThreads * kBucketPerThread);
        const int64_t c1 = min(c0 + kBucketThreads * kBucketPerThread, ev_last);
        const ChunkOrigin org = origins[chunk];            // same address for every thread: one broadcast load
        __syncthreads();                                   // previous chunk is done with smem
        for (int i = threadIdx.x; i < nh; i += kBucketThreads) hist[i] = 0;
        const int gb0 = org.gb0;
        const bool single = org.single != 0;
        int w = org.w0;
        WinInfo wi = org.win;

        // all global loads of the chunk are issued before any of them is used
        const bool fast = single && (c1 - c0) == kBucketThreads * kBucketPerThread &&
                          wi.start >= 0 && wi.start <= 0xFFFFFFFFll;
        // per event: timestamp and x | y << 14 | p << 28 (the .dat word), kBad when out of range
        constexpr uint32_t kBad = 0xFFFFFFFFu;
        auto pack = [](uint32_t x, uint32_t y, uint32_t p) -> uint32_t {
            return (((x | y) >> 14) | (p >> 1)) ? kBad : (x | (y << 14) | (p << 28));
        };
        uint32_t tt[kBucketPerThread], xyp[kBucketPerThread];
        if (fast && vec_ok) {
            // 4 consecutive events per 128/64/64/32-bit load (c0 is a multiple of 4 events)
            static_assert(kBucketPerThread % 4 == 0, "vector path loads events in groups of 4");
#pragma unroll
            for (int g = 0; g < kBucketPerThread / 4; ++g) {
                const int64_t base = c0 + ((int64_t)g * kBucketThreads + threadIdx.x) * 4;
                const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(ev.t + base));
                const uint2 x4 = __ldg(reinterpret_cast<const uint2*>(ev.x + base));
                const uint2 y4 = __ldg(reinterpret_cast<const uint2*>(ev.y + base));
                const uint32_t p4 = __ldg(reinterpret_cast<const uint32_t*>(ev.p + base));
                tt[4 * g + 0] = t4.x; tt[4 * g + 1] = t4.y; tt[4 * g + 2] = t4.z; tt[4 * g + 3] = t4.w;
                xyp[4 * g + 0] = pack(x4.x & 0xFFFFu, y4.x & 0xFFFFu, p4 & 0xFFu);
                xyp[4 * g + 1] = pack(x4.x >> 16, y4.x >> 16, (p4 >> 8) & 0xFFu);
                xyp[4 * g + 2] = pack(x4.y & 0xFFFFu, y4.y & 0xFFFFu, (p4 >> 16) & 0xFFu);
                xyp[4 * g + 3] = pack(x4.y >> 16, y4.y >> 16, p4 >> 24);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kBucketPerThread; ++k) {
                const int64_t i = c0 + k * kBucketThreads + threadIdx.x;
                xyp[k] = kBad;
                if (i < c1) { tt[k] = __ldg(ev.t + i); xyp[k] = pack(__ldg(ev.x + i), __ldg(ev.y + i), __ldg(ev.p + i)); }
            }
        }

        __syncthreads();                                   // histogram is zeroed
        // per event: (smem counter << 12) | rank inside the chunk, or kNone when dropped
        constexpr uint32_t kNone = 0xFFFFFFFFu;
        uint32_t slot[kBucketPerThread], rec[kBucketPerThread] = {};

        // count / rank one classified event
        auto deposit = [&](int k, uint32_t tile, int gbin) {
            const uint32_t lb = (uint32_t)(gbin - gb0);
            if (lb < (uint32_t)kLocalBins) {
                const uint32_t key = lb * (uint32_t)pl.n_tiles + tile;
                if (kScatter) slot[k] = (key << 12) | atomicAdd(&hist[key], 1u);
                else atomicAdd(&hist[key], 1u);
            } else {                          // unsorted input or a very sparse stream: go straight to global
                uint32_t* cursor = pl.counts + (int64_t)tile * pl.TB + gbin;
                if (kScatter)
                    pl.records[pl.tile_base[tile] + pl.off_rel[(int64_t)tile * (pl.TB + 1) + gbin] + atomicAdd(cursor, 1u)] = rec[k];
                else { atomicAdd(cursor, 1u); pl.bin_any[gbin] = 1u; }
            }
        };
        // map raw coordinates to the grid; false when the event is to be dropped
        auto locate = [&](int k, uint32_t& pix) -> bool {
            uint32_t xm = xyp[k] & 0x3FFFu, ym = (xyp[k] >> 14) & 0x3FFFu;
            bool ok = xyp[k] != kBad;
            if (use_lut) {
                ok = ok && xm < (uint32_t)lut_w && ym < (uint32_t)lut_h;
                xm = s_lutx[min(xm, (uint32_t)lut_w - 1u)];
                ym = s_luty[min(ym, (uint32_t)lut_h - 1u)];
            }
            pix = ym * W + xm;
            return ok && xm < W && ym < H;
        };

        if (fast) {
            // the whole chunk lies in one window: 32-bit time arithmetic, no bounds checks
            const uint32_t start32 = (uint32_t)wi.start, zmax = (uint32_t)(wi.nbins - 1);
#pragma unroll
            for (int k = 0; k < kBucketPerThread; ++k) {
                slot[k] = kNone;
                uint32_t pix;
                if (!locate(k, pix)) continue;
                const uint32_t u = tt[k] >= start32 ? tt[k] - start32 : 0u;
                const uint32_t z = min(pl.div_abin.div(u), zmax);
                const uint32_t tile = pl.div_P.div(pix);
                if (kScatter) rec[k] = (min(u - z * pl.abin, kDMax) << 14) | ((pix - tile * pl.P) << 1) | (xyp[k] >> 28);
                deposit(k, tile, wi.binbase + (int)z);
            }
        } else {
#pragma unroll
            for (int k = 0; k < kBucketPerThread; ++k) {
                const int64_t i = c0 + k * kBucketThreads + threadIdx.x;
                slot[k] = kNone;
                if (i >= c1) continue;
                if (!single) {                  // chunk straddles a window boundary or a gap
                    while (w < pl.n_windows && i >= __ldg(pl.w_end + w)) ++w;
                    if (w >= pl.n_windows) continue;
                    wi = load_window(pl, w);
                    if (i < wi.begin || wi.nbins <= 0) continue;
                }
                uint32_t pix;
                if (!locate(k, pix)) continue;
                uint32_t z, d;
                bin_of(pl, wi, tt[k], z, d);
                const uint32_t tile = pl.div_P.div(pix);
                rec[k] = (d << 14) | ((pix - tile * pl.P) << 1) | (xyp[k] >> 28);
                deposit(k, tile, wi.binbase + (int)z);
            }
        }
        __syncthreads();
        if (!kScatter) {
            for (int i = threadIdx.x; i < nh; i += kBucketThreads) {
                const uint32_t c = hist[i];
                if (!c) continue;
                const int lb = i / pl.n_tiles, tile = i - lb * pl.n_tiles, gbin = gb0 + lb;
                atomicAdd(pl.counts + (int64_t)tile * pl.TB + gbin, c);
                pl.bin_any[gbin] = 1u;
            }
            continue;
        }
        const uint32_t n_valid = block_exclusive_scan(hist, loff, nh, s_tmp);
        for (int i = threadIdx.x; i < nh; i += kBucketThreads) {
            const uint32_t c = hist[i];
            if (!c) continue;
            const int lb = i / pl.n_tiles, tile = i - lb * pl.n_tiles, gbin = gb0 + lb;
            gbase[i] = pl.tile_base[tile] + pl.off_rel[(int64_t)tile * (pl.TB + 1) + gbin] +
                       atomicAdd(pl.counts + (int64_t)tile * pl.TB + gbin, c) - loff[i];
        }
#pragma unroll
        for (int k = 0; k < kBucketPerThread; ++k) {
            if (slot[k] == kNone) continue;
            const uint32_t key = slot[k] >> 12;
            const uint32_t pos = loff[key] + (slot[k] & 0xFFFu);
            sorted[pos] = rec[k];
            skey[pos] = (uint16_t)key;
        }
        __syncthreads();
        for (uint32_t pos = threadIdx.x; pos < n_valid; pos += kBucketThreads)
            pl.records[gbase[skey[pos]] + pos] = sorted[pos];
    }
}

// Exclusive scan of one tile's per-bin counts -> relative offsets; counts are zeroed so
// that the scatter pass can reuse them as cursors.
__global__ void __launch_bounds__(256)
taf_scan_rows_kernel(StreamPlan pl) {
    __shared__ uint32_t warp_sum[8];
    __shared__ uint32_t s_carry;
    const int tile = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t* cnt = pl.counts + (int64_t)tile * pl.TB;
    uint32_t* off = pl.off_rel + (int64_t)tile * (pl.TB + 1);
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < pl.TB; base += 256 * 4) {
        const int i0 = base + threadIdx.x * 4;
        uint32_t c[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) c[k] = (i0 + k < pl.TB) ? cnt[i0 + k] : 0u;
        const uint32_t mine = c[0] + c[1] + c[2] + c[3];
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        uint32_t before = s_carry;
        for (int k = 0; k < wid; ++k) before += warp_sum[k];
        uint32_t run = before + incl - mine;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i0 + k < pl.TB) { off[i0 + k] = run; cnt[i0 + k] = 0u; }
            run += c[k];
        }
        __syncthreads();
        if (threadIdx.x == 255) s_carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) { off[pl.TB] = s_carry; pl.tile_total[tile] = s_carry; }
}

__global__ void __launch_bounds__(1024)
taf_scan_tiles_kernel(StreamPlan pl) {         // n_tiles <= kMaxTiles = 2 * 1024
    __shared__ uint32_t warp_sum[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t c[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i = threadIdx.x * 2 + k;
        c[k] = i < pl.n_tiles ? ((pl.tile_total[i] + 3u) & ~3u) : 0u;
    }
    const uint32_t mine = c[0] + c[1];
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (int k = 0; k < wid; ++k) before += warp_sum[k];
    uint32_t run = before + incl - mine;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i = threadIdx.x * 2 + k;
        if (i < pl.n_tiles) pl.tile_base[i] = run;
        run += c[k];
        if (i == pl.n_tiles - 1) pl.tile_base[pl.n_tiles] = run;
    }
}

// Per (tile, batch) summary for the consumer warps of the tile kernel: which bins of the batch
// have records of this tile (bit b) and which are non-empty anywhere (bit 16 + b).
static_assert(kBatchBins <= 16, "two 16-bit masks per batch");
__global__ void __launch_bounds__(256)
taf_tile_bits_kernel(StreamPlan pl) {
    const int tile = blockIdx.x;
    const uint32_t* off = pl.off_rel + (int64_t)tile * (pl.TB + 1);
    for (int j = threadIdx.x; j < pl.n_batches; j += blockDim.x) {
        const Batch m = pl.batches[j];
        uint32_t bits = 0;
        for (int b = 0; b < m.nb; ++b) {
            if (off[m.gbin0 + b + 1] > off[m.gbin0 + b]) bits |= 1u << b;
            if (pl.bin_any[m.gbin0 + b]) bits |= 1u << (16 + b);
        }
        pl.tile_bits[(int64_t)tile * pl.n_batches + j] = bits;
    }
}

// ---- mbarrier / TMA bulk-copy primitives (PTX) --------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct TileParams {
    StreamPlan pl;
    float* state;          // [H,W,2,K]
    float* out;            // window w at out + w * out_stride
    int64_t out_stride;
    int emit_state;        // write the state after every window (always after the last)
    int n_emits;           // number of windows (batches that end a window)
    int bulk_out;          // out rows are 16-byte aligned: emit through smem + TMA bulk stores
    float span;            // f32(abin + 1e-8)
};

// Shared-memory carve-up of the tile kernel (all offsets multiples of 128 bytes).
struct TileSmem {
    int ring, acc, stage, bars, off, any, meta, total;
    __host__ __device__ TileSmem(int P, int K) {
        int o = 0;
        ring = o;  o += kStages * kChunkRecords * 4;
        acc = o;   o += 2 * P * (int)sizeof(uint2);                 // {n, sum d} per (pixel, polarity)
        stage = o; o += 2 * K * P * 4;                              // [2K][P] output staging
        bars = o;  o += 128;
        off = o;   o += 2 * (kBatchBins + 1) * 4; o = (o + 127) / 128 * 128;
        any = o;   o += 2 * kBatchBins * 4;
        meta = o;  o += 2 * (int)sizeof(Batch); o = (o + 127) / 128 * 128;
        total = o;
    }
};

template <int K, int SLOTS>
__global__ void __launch_bounds__(kTafThreads, 1)
taf_tile_kernel(TileParams tp) {
    const StreamPlan& pl = tp.pl;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmem lay(pl.P, K);
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw + lay.ring);     // [kStages][kChunkRecords]
    uint2* acc = reinterpret_cast<uint2*>(smem_raw + lay.acc);             // [2P]
    float* stage = reinterpret_cast<float*>(smem_raw + lay.stage);         // [2K][P]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);
    uint32_t* s_off = reinterpret_cast<uint32_t*>(smem_raw + lay.off);     // [2][kBatchBins+1]
    uint32_t* s_any = reinterpret_cast<uint32_t*>(smem_raw + lay.any);     // [2][kBatchBins]
    Batch* s_meta = reinterpret_cast<Batch*>(smem_raw + lay.meta);         // [2]

    const int tid = threadIdx.x, tile = blockIdx.x;
    const int64_t HW = (int64_t)pl.H * pl.W;
    const int64_t pix0 = (int64_t)tile * pl.P;
    const int npix = (int)min((int64_t)pl.P, HW - pix0);
    const uint32_t* my_records = pl.records + pl.tile_base[tile];
    const uint32_t list_len = (pl.tile_total[tile] + 3u) & ~3u;
    const int n_chunks = (int)((list_len + kChunkRecords - 1) / kChunkRecords);
    const uint32_t* my_off = pl.off_rel + (int64_t)tile * (pl.TB + 1);

    auto issue = [&](int c) {           // thread 0 only
        const uint32_t first = (uint32_t)c * kChunkRecords;
        const uint32_t bytes = min((uint32_t)kChunkRecords, list_len - first) * 4u;
        uint64_t* bar = full + (c % kStages);
        mbar_expect_tx(bar, bytes);
        tma_load_1d(ring + (c % kStages) * kChunkRecords, my_records + first, bytes, bar);
    };

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 2 * pl.P; i += kTafThreads) acc[i] = make_uint2(0u, 0u);
    if (tid == 0) s_meta[0] = pl.batches[0];
    __syncthreads();
    if (tid == 0)
        for (int c = 0; c < n_chunks && c < kStages; ++c) issue(c);
    {
        const Batch m0 = s_meta[0];
        if (tid <= m0.nb) s_off[tid] = my_off[m0.gbin0 + tid];
        if (tid < m0.nb) s_any[tid] = pl.bin_any[m0.gbin0 + tid];
    }

    // FIFO state of this thread's pixels, as float2 pairs for the packed f32x2 adds of
    // sm_100: element k of (slot, polarity) is v[s][p][k / 2].{x,y}; k = K-1 is the newest.
    static_assert(K % 4 == 0, "K must be a multiple of 4");
    float2 v[SLOTS][2][K / 2];
    const bool first_fresh = (pl.batches[0].flags & 1) != 0;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int lp = s * kTafThreads + tid;
        if (lp < npix && !first_fresh) {
            const float4* src = reinterpret_cast<const float4*>(tp.state + (pix0 + lp) * 2 * K);
#pragma unroll
            for (int q = 0; q < 2 * K / 4; ++q) {
                const float4 f = src[q];
                v[s][(q * 4) / K][((q * 4) % K) / 2 + 0] = make_float2(f.x, f.y);
                v[s][(q * 4) / K][((q * 4) % K) / 2 + 1] = make_float2(f.z, f.w);
            }
        } else {
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int k = 0; k < K / 2; ++k) v[s][p][k] = make_float2(kTafInit, kTafInit);
        }
    }
    __syncthreads();

    int ready_chunk = -1;      // highest chunk this thread has observed complete
    int next_refill = kStages; // next chunk to load; its stage is free once chunk (next_refill - kStages) is drained
    bool staged_once = false;
    // batch descriptors are fetched two batches ahead, their offsets / flags one batch ahead
    Batch nmeta = pl.batches[pl.n_batches > 1 ? 1 : 0];
    for (int j = 0; j < pl.n_batches; ++j) {
        const int buf = j & 1;
        const Batch meta = s_meta[buf];
        uint32_t pre_off = 0, pre_any = 0;
        Batch nnmeta = nmeta;
        if (j + 1 < pl.n_batches) {
            if (tid <= nmeta.nb) pre_off = my_off[nmeta.gbin0 + tid];
            if (tid < nmeta.nb) pre_any = pl.bin_any[nmeta.gbin0 + tid];
            if (j + 2 < pl.n_batches) nnmeta = pl.batches[j + 2];
        }
        if (meta.flags & 1) {
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int k = 0; k < K / 2; ++k) v[s][p][k] = make_float2(kTafInit, kTafInit);
        }
        for (int b = 0; b < meta.nb; ++b) {
            const uint32_t o0 = s_off[buf * (kBatchBins + 1) + b], o1 = s_off[buf * (kBatchBins + 1) + b + 1];
            if (!s_any[buf * kBatchBins + b]) continue;          // nobody saw an event: no ageing
            const bool have = o1 > o0;
            if (have) {
                uint32_t cur = o0;
                while (cur < o1) {
                    const int c = (int)(cur / kChunkRecords);
                    const uint32_t chunk_end = (uint32_t)(c + 1) * kChunkRecords;
                    const uint32_t seg_end = o1 < chunk_end ? o1 : chunk_end;
                    if (c >= next_refill) {
                        // a single bin longer than the whole ring: recycle drained stages now
                        __syncthreads();
                        if (tid == 0)
                            for (int r = next_refill; r <= c && r < n_chunks; ++r) issue(r);
                        next_refill = c + 1;
                    }
                    if (c > ready_chunk) { mbar_wait(full + (c % kStages), (uint32_t)(c / kStages) & 1u); ready_chunk = c; }
                    const uint32_t* chunk = ring + (c % kStages) * kChunkRecords;
                    for (uint32_t r = cur + tid; r < seg_end; r += kTafThreads) {
                        const uint32_t rec = chunk[r & (kChunkRecords - 1)];
                        uint2* cell = acc + (rec & 0x3FFFu);         // 2 * local pixel + p
                        atomicAdd(&cell->x, 1u);
                        atomicAdd(&cell->y, rec >> 14);
                    }
                    cur = seg_end;
                }
                __syncthreads();                                  // all records of the bin are in `acc`
                // every chunk that ends at or before o1 is drained: refill those ring stages
                const int drained = (int)(o1 / kChunkRecords);    // chunks [0, drained) fully consumed
                if (tid == 0)
                    for (int r = next_refill; r < drained + kStages && r < n_chunks; ++r) issue(r);
                if (drained + kStages > next_refill) next_refill = drained + kStages;
            }
            const float2 minus1 = make_float2(-1.0f, -1.0f);
            if (!have) {
                // the tile saw nothing in this bin, but some other tile did: everything ages
#pragma unroll
                for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                    for (int p = 0; p < 2; ++p)
#pragma unroll
                        for (int k = 0; k < K / 2; ++k) v[s][p][k] = __fadd2_rn(v[s][p][k], minus1);
            } else {
                // read and clear this thread's accumulators, then release `acc` for the next bin
                // BEFORE the arithmetic: the long update phase runs without a barrier behind it
                uint4 a[SLOTS];
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kTafThreads + tid;
                    a[s] = make_uint4(0u, 0u, 0u, 0u);
                    // every slot but the last lies inside the tile's accumulator array
                    if (s < SLOTS - 1 || lp < pl.P) {
                        a[s] = *reinterpret_cast<uint4*>(acc + 2 * lp);           // {n0, S0, n1, S1}
                        if (a[s].x | a[s].z) *reinterpret_cast<uint4*>(acc + 2 * lp) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                __syncthreads();
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const uint32_t nn[2] = {a[s].x, a[s].z}, ss[2] = {a[s].y, a[s].w};
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        // mean(t_norm) - 1 = S / (n span) - 1 (generate_taf.py:23-27); for n == 0
                        // the value is NaN and is never selected
                        const bool active = nn[p] != 0u;
                        float r;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)nn[p] * tp.span));
                        const float mean = fmaf((float)ss[p], r, -1.0f);
                        float2 aged[K / 2];
#pragma unroll
                        for (int k = 0; k < K / 2; ++k) aged[k] = __fadd2_rn(v[s][p][k], minus1);
#pragma unroll
                        for (int k = 0; k < K / 2; ++k) {
                            const float next = (k + 1 < K / 2) ? aged[k + 1].x : mean;
                            v[s][p][k].x = active ? aged[k].y : aged[k].x;
                            v[s][p][k].y = active ? next : aged[k].y;
                        }
                    }
                }
            }
        }
        if (meta.flags & 2) {
            const bool write_state = tp.emit_state || (j == pl.n_batches - 1);
            float* o = tp.out + (int64_t)meta.win * tp.out_stride + pix0;
            if (tp.bulk_out) {
                // [2K][npix] staging tile -> one TMA bulk store per channel row
                if (staged_once) {
                    if (tid < 2 * K) bulk_wait_read();             // previous window's rows have left smem
                    __syncthreads();
                }
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kTafThreads + tid;
                    if (s == SLOTS - 1 && lp >= pl.P) continue;     // columns >= npix are staged but never stored
#pragma unroll
                    for (int k = 0; k < K / 2; ++k)
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            stage[(4 * k + p) * pl.P + lp] = v[s][p][k].x;
                            stage[(4 * k + 2 + p) * pl.P + lp] = v[s][p][k].y;
                        }
                }
                fence_async_smem();
                __syncthreads();
                if (tid < 2 * K) {
                    bulk_store_1d(o + (int64_t)tid * HW, stage + tid * pl.P, (uint32_t)npix * 4u);
                    bulk_commit();
                }
                staged_once = true;
            } else {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kTafThreads + tid;
                    if (lp >= npix) continue;
#pragma unroll
                    for (int k = 0; k < K / 2; ++k)
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            __stcs(o + (int64_t)(4 * k + p) * HW + lp, v[s][p][k].x);
                            __stcs(o + (int64_t)(4 * k + 2 + p) * HW + lp, v[s][p][k].y);
                        }
                }
            }
            if (write_state) {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kTafThreads + tid;
                    if (lp >= npix) continue;
                    float4* dst = reinterpret_cast<float4*>(tp.state + (pix0 + lp) * 2 * K);
#pragma unroll
                    for (int q = 0; q < 2 * K / 4; ++q) {
                        const float2 lo = v[s][(q * 4) / K][((q * 4) % K) / 2], hi = v[s][(q * 4) / K][((q * 4) % K) / 2 + 1];
                        dst[q] = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
                }
            }
        }
        if (j + 1 < pl.n_batches) {
            const int nb = buf ^ 1;
            if (tid == 0) s_meta[nb] = nmeta;
            if (tid <= nmeta.nb) s_off[nb * (kBatchBins + 1) + tid] = pre_off;
            if (tid < nmeta.nb) s_any[nb * kBatchBins + tid] = pre_any;
            nmeta = nnmeta;
        }
        __syncthreads();
    }
    if (tp.bulk_out && tid < 2 * K) bulk_wait_all();               // smem must outlive the bulk reads
}

// ---- warp-specialised tile kernel -------------------------------------------------------------
// Same algorithm as taf_tile_kernel, split into two roles so that the latency-bound record
// bookkeeping runs ahead of, and concurrently with, the arithmetic:
//   * producer warpgroup (warps 0-3, 56 registers after setmaxnreg.dec).  Three accumulate warps
//     walk the bins, wait for the TMA ring, accumulate (n, sum d) of the next bins into one of two
//     accumulator buffers and refill the ring; the fourth, the store warp, sends every staged
//     window tensor with TMA bulk stores;
//   * consumer warpgroups (warps 4-15, 152 registers after setmaxnreg.inc): own the FIFO state of
//     the tile (6 pixels per thread), read + clear the accumulator of a bin, apply the update, and
//     copy the window tensor into the staging tile.  They never synchronise among themselves.
// Hand-over uses named barriers (bar.arrive / bar.sync): FULL[buf] accumulate -> consumers,
// EMPTY[buf] consumers -> accumulate, STAGED consumers -> store warp, STAGE_FREE store warp ->
// consumers.  4 warps per SM sub-partition: 1 producer-group warp + 3 consumers.
constexpr int kProducerThreads = 128;   // the producer warpgroup: 3 accumulate warps + 1 store warp
constexpr int kAccumThreads = 96;
constexpr int kStoreThreads = 32;
constexpr int kConsumerThreads = 384;
constexpr int kWsThreads = kProducerThreads + kConsumerThreads;
constexpr int kWsChunkRecords = 512;    // 2 KB TMA bulk copies
constexpr int kWsStages = 8;            // 16 KB ring = a flat circular buffer of 4096 records
constexpr int kWsRing = kWsChunkRecords * kWsStages;
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
        bars = o;   o += 64;
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

template <int K, int SLOTS>
__global__ void __launch_bounds__(kWsThreads, 1)
taf_tile_ws_kernel(TileParams tp) {
    const StreamPlan& pl = tp.pl;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmemWS lay(pl.P, K);
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw + lay.ring);     // [kWsStages][kWsChunkRecords]
    uint2* acc = reinterpret_cast<uint2*>(smem_raw + lay.acc);             // [2][2P]
    float* stage = reinterpret_cast<float*>(smem_raw + lay.stage);         // [2K][P]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);

    const int tid = threadIdx.x, tile = blockIdx.x;
    const int64_t HW = (int64_t)pl.H * pl.W;
    const int64_t pix0 = (int64_t)tile * pl.P;
    const int npix = (int)min((int64_t)pl.P, HW - pix0);
    const uint32_t* my_off = pl.off_rel + (int64_t)tile * (pl.TB + 1);

    if (tid == 0) {
        for (int s = 0; s < kWsStages; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 4 * pl.P; i += kWsThreads) acc[i] = make_uint2(0u, 0u);
    __syncthreads();

    if (tid < kProducerThreads) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (tid >= kAccumThreads) {
            // ================================== store warp ==================================
            // Waits for the consumers to stage a window tensor, sends its rows with TMA bulk
            // stores and tells the consumers when the staging tile may be overwritten.
            if (!tp.bulk_out) return;
            const int lane = tid - kAccumThreads;
            int emitted = 0;
            for (int j = 0; j < pl.n_batches; ++j) {
                const Batch m = pl.batches[j];
                if (!(m.flags & 2)) continue;
                named_sync(kBarStaged, kStoreThreads + kConsumerThreads);
                if (lane < 2 * K) {
                    float* o = tp.out + (int64_t)m.win * tp.out_stride + pix0;
                    bulk_store_1d(o + (int64_t)lane * HW, stage + lane * pl.P, (uint32_t)npix * 4u);
                    bulk_commit();
                    bulk_wait_read();                                // the rows have left shared memory
                }
                __syncwarp();
                if (++emitted < tp.n_emits) named_arrive(kBarStageFree, kStoreThreads + kConsumerThreads);
            }
            if (lane < 2 * K) bulk_wait_all();
            return;
        }
        // ================================ accumulate warps ================================
        const uint32_t* my_records = pl.records + pl.tile_base[tile];
        const uint32_t list_len = (pl.tile_total[tile] + 3u) & ~3u;
        const int n_chunks = (int)((list_len + kWsChunkRecords - 1) / kWsChunkRecords);
        auto issue = [&](int c) {           // thread 0 only
            const uint32_t first = (uint32_t)c * kWsChunkRecords;
            const uint32_t bytes = min((uint32_t)kWsChunkRecords, list_len - first) * 4u;
            uint64_t* bar = full + (c % kWsStages);
            mbar_expect_tx(bar, bytes);
            tma_load_1d(ring + (c % kWsStages) * kWsChunkRecords, my_records + first, bytes, bar);
        };
        if (tid == 0)
            for (int c = 0; c < n_chunks && c < kWsStages; ++c) issue(c);
        BatchFeed feed;
        feed.init(smem_raw + lay.feed_p, &pl, my_off, tid, kBarProducers, kAccumThreads);
        int ready_chunk = -1, next_refill = kWsStages, buf = 0;
        int uses0 = 0, uses1 = 0;
        constexpr int kHandOver = kAccumThreads + kConsumerThreads;
        for (int j = 0; j < pl.n_batches; ++j) {
            const Batch meta = feed.begin(j);
            const int jb = j & 1;
            for (int b = 0; b < meta.nb; ++b) {
                const uint32_t o0 = feed.s_off[jb * (kBatchBins + 1) + b], o1 = feed.s_off[jb * (kBatchBins + 1) + b + 1];
                if (!feed.s_any[jb * kBatchBins + b] || o1 <= o0) continue;
                uint2* my_acc = acc + buf * 2 * pl.P;
                const int last_c = (int)((o1 - 1) / kWsChunkRecords);
                const bool had_use = (buf ? uses1 : uses0) > 0;
                if (last_c < next_refill) {
                    // common case: every chunk of the bin is already in flight
                    while (ready_chunk < last_c) {
                        ++ready_chunk;
                        mbar_wait(full + (ready_chunk % kWsStages), (uint32_t)(ready_chunk / kWsStages) & 1u);
                    }
                    // records are pulled into registers BEFORE waiting for the accumulator buffer
                    constexpr int kPre = 8;
                    constexpr uint32_t kNoRec = 0xFFFFFFFFu;        // d = 2^18-1, pixel 8191: never produced for P <= 2560
                    uint32_t pre[kPre];
#pragma unroll
                    for (int i = 0; i < kPre; ++i) {
                        const uint32_t r = o0 + tid + i * kAccumThreads;
                        pre[i] = r < o1 ? ring[r & (kWsRing - 1)] : kNoRec;
                    }
                    const bool all_pre = (o1 - o0) <= (uint32_t)(kPre * kAccumThreads);
                    // the consumers must have drained this buffer (its first use needs no wait); the
                    // barrier also tells that every accumulate thread is done with all earlier bins
                    if (had_use) named_sync(kBarEmpty0 + buf, kHandOver);
                    else named_sync(kBarProducers, kAccumThreads);
                    {
                        // ring stages whose chunk ends before the first record still to be read are free
                        const int drained = (int)((all_pre ? o1 : o0) / kWsChunkRecords);
                        if (tid == 0)
                            for (int r = next_refill; r < drained + kWsStages && r < n_chunks; ++r) issue(r);
                        if (drained + kWsStages > next_refill) next_refill = drained + kWsStages;
                    }
#pragma unroll
                    for (int i = 0; i < kPre; ++i) {
                        if (pre[i] != kNoRec) {
                            uint2* cell = my_acc + (pre[i] & 0x3FFFu);   // 2 * local pixel + p
                            atomicAdd(&cell->x, 1u);
                            atomicAdd(&cell->y, pre[i] >> 14);
                        }
                    }
                    for (uint32_t r = o0 + tid + kPre * kAccumThreads; r < o1; r += kAccumThreads) {
                        const uint32_t rec = ring[r & (kWsRing - 1)];
                        uint2* cell = my_acc + (rec & 0x3FFFu);
                        atomicAdd(&cell->x, 1u);
                        atomicAdd(&cell->y, rec >> 14);
                    }
                } else {
                    // a single bin longer than the ring: go chunk by chunk, recycling drained stages
                    if (had_use) named_sync(kBarEmpty0 + buf, kHandOver);
                    uint32_t cur = o0;
                    while (cur < o1) {
                        const int c = (int)(cur / kWsChunkRecords);
                        const uint32_t chunk_end = (uint32_t)(c + 1) * kWsChunkRecords;
                        const uint32_t seg_end = o1 < chunk_end ? o1 : chunk_end;
                        if (c >= next_refill) {
                            named_sync(kBarProducers, kAccumThreads);
                            if (tid == 0)
                                for (int r = next_refill; r <= c && r < n_chunks; ++r) issue(r);
                            next_refill = c + 1;
                        }
                        while (ready_chunk < c) {
                            ++ready_chunk;
                            mbar_wait(full + (ready_chunk % kWsStages), (uint32_t)(ready_chunk / kWsStages) & 1u);
                        }
                        for (uint32_t r = cur + tid; r < seg_end; r += kAccumThreads) {
                            const uint32_t rec = ring[r & (kWsRing - 1)];
                            uint2* cell = my_acc + (rec & 0x3FFFu);
                            atomicAdd(&cell->x, 1u);
                            atomicAdd(&cell->y, rec >> 14);
                        }
                        cur = seg_end;
                    }
                }
                named_arrive(kBarFull0 + buf, kHandOver);            // hand the accumulator to the consumers
                if (buf) ++uses1; else ++uses0;
                buf ^= 1;
            }
            feed.end(j);
        }
        // match the consumers' last EMPTY arrivals so that no barrier phase is left open
        if (uses0 > 0) named_sync(kBarEmpty0, kHandOver);
        if (uses1 > 0) named_sync(kBarEmpty1, kHandOver);
        return;
    }

    // =================================== consumer warpgroups ===================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int ctid = tid - kProducerThreads;
    constexpr int kHandOver = kAccumThreads + kConsumerThreads;
    static_assert(K % 4 == 0, "K must be a multiple of 4");
    float2 v[SLOTS][2][K / 2];
    const bool first_fresh = (pl.batches[0].flags & 1) != 0;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int lp = s * kConsumerThreads + ctid;
        if (lp < npix && !first_fresh) {
            const float4* src = reinterpret_cast<const float4*>(tp.state + (pix0 + lp) * 2 * K);
#pragma unroll
            for (int q = 0; q < 2 * K / 4; ++q) {
                const float4 f = src[q];
                v[s][(q * 4) / K][((q * 4) % K) / 2 + 0] = make_float2(f.x, f.y);
                v[s][(q * 4) / K][((q * 4) % K) / 2 + 1] = make_float2(f.z, f.w);
            }
        } else {
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int k = 0; k < K / 2; ++k) v[s][p][k] = make_float2(kTafInit, kTafInit);
        }
    }
    // batch descriptors and the tile's per-batch bit masks come straight from global memory,
    // one batch ahead: the consumers never synchronise among themselves
    const uint32_t* my_bits = pl.tile_bits + (int64_t)tile * pl.n_batches;
    Batch meta = pl.batches[0];
    uint32_t bits = my_bits[0];
    int buf = 0, emitted = 0;
    const float2 minus1 = make_float2(-1.0f, -1.0f);
    for (int j = 0; j < pl.n_batches; ++j) {
        Batch nmeta = meta;
        uint32_t nbits = 0;
        if (j + 1 < pl.n_batches) { nmeta = pl.batches[j + 1]; nbits = my_bits[j + 1]; }
        if (meta.flags & 1) {
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int k = 0; k < K / 2; ++k) v[s][p][k] = make_float2(kTafInit, kTafInit);
        }
        for (int b = 0; b < meta.nb; ++b) {
            if (!((bits >> (16 + b)) & 1u)) continue;               // nobody saw an event: no ageing
            if (!((bits >> b) & 1u)) {
                // the tile saw nothing in this bin, but some other tile did: everything ages
#pragma unroll
                for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                    for (int p = 0; p < 2; ++p)
#pragma unroll
                        for (int k = 0; k < K / 2; ++k) v[s][p][k] = __fadd2_rn(v[s][p][k], minus1);
                continue;
            }
            named_sync(kBarFull0 + buf, kHandOver);                  // the producers filled this accumulator
            uint2* my_acc = acc + buf * 2 * pl.P;
            uint4 a[SLOTS];
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int lp = s * kConsumerThreads + ctid;
                a[s] = make_uint4(0u, 0u, 0u, 0u);
                if (s < SLOTS - 1 || lp < pl.P) {                    // every slot but the last lies inside the array
                    a[s] = *reinterpret_cast<uint4*>(my_acc + 2 * lp);               // {n0, S0, n1, S1}
                    if (a[s].x | a[s].z) *reinterpret_cast<uint4*>(my_acc + 2 * lp) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            named_arrive(kBarEmpty0 + buf, kHandOver);               // clean again: give it back
            buf ^= 1;
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const uint32_t nn[2] = {a[s].x, a[s].z}, ss[2] = {a[s].y, a[s].w};
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    // mean(t_norm) - 1 = S / (n span) - 1 (generate_taf.py:23-27); NaN for n == 0, never selected
                    const bool active = nn[p] != 0u;
                    float r;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)nn[p] * tp.span));
                    const float mean = fmaf((float)ss[p], r, -1.0f);
                    float2 aged[K / 2];
#pragma unroll
                    for (int k = 0; k < K / 2; ++k) aged[k] = __fadd2_rn(v[s][p][k], minus1);
#pragma unroll
                    for (int k = 0; k < K / 2; ++k) {
                        const float next = (k + 1 < K / 2) ? aged[k + 1].x : mean;
                        v[s][p][k].x = active ? aged[k].y : aged[k].x;
                        v[s][p][k].y = active ? next : aged[k].y;
                    }
                }
            }
        }
        if (meta.flags & 2) {
            const bool write_state = tp.emit_state || (j == pl.n_batches - 1);
            if (tp.bulk_out) {
                // stage the [2K][P] tile; the store warp sends it.  No consumer waits for another:
                // each warp streams its columns, fences, signals and moves on to the next bin.
                if (emitted > 0) named_sync(kBarStageFree, kStoreThreads + kConsumerThreads);
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kConsumerThreads + ctid;
                    if (s == SLOTS - 1 && lp >= pl.P) continue;     // columns >= npix are staged but never stored
#pragma unroll
                    for (int k = 0; k < K / 2; ++k)
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            stage[(4 * k + p) * pl.P + lp] = v[s][p][k].x;
                            stage[(4 * k + 2 + p) * pl.P + lp] = v[s][p][k].y;
                        }
                }
                fence_async_smem();
                named_arrive(kBarStaged, kStoreThreads + kConsumerThreads);
                ++emitted;
            } else {
                float* o = tp.out + (int64_t)meta.win * tp.out_stride + pix0;
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kConsumerThreads + ctid;
                    if (lp >= npix) continue;
#pragma unroll
                    for (int k = 0; k < K / 2; ++k)
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            __stcs(o + (int64_t)(4 * k + p) * HW + lp, v[s][p][k].x);
                            __stcs(o + (int64_t)(4 * k + 2 + p) * HW + lp, v[s][p][k].y);
                        }
                }
            }
            if (write_state) {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kConsumerThreads + ctid;
                    if (lp >= npix) continue;
                    float4* dst = reinterpret_cast<float4*>(tp.state + (pix0 + lp) * 2 * K);
#pragma unroll
                    for (int q = 0; q < 2 * K / 4; ++q) {
                        const float2 lo = v[s][(q * 4) / K][((q * 4) % K) / 2], hi = v[s][(q * 4) / K][((q * 4) % K) / 2 + 1];
                        dst[q] = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
                }
            }
        }
        meta = nmeta;
        bits = nbits;
    }
}

// ---- Event Volume over whole streams ----------------------------------------------------------
// generate_eventvolume.py:15-42 for a list of non-overlapping windows: the same bucketing (one
// "bin" per window, d = t - t0) feeds one CTA per sensor tile.  The tile's [2K][P] float
// accumulator lives in shared memory: splat (shared-memory float atomics), then one pass that
// reads, clears, scales by /5*255 and stores 16 bytes per thread to the tensor (rows are contiguous).
// Two CTAs per SM, so one tile's output pass overlaps the other tile's splat.  Records arrive through the same ring of TMA bulk copies as in the TAF kernel.
constexpr int kEvThreads = 512;
constexpr int kEvTilesPerSm = 2;        // two CTAs per SM: one splats while the other's bulk store drains

// v / 5 * 255 (generate_eventvolume.py:37) without the IEEE-division subroutine: one Newton
// correction of v * RN(1/5) is the correctly rounded quotient for every finite v away from the
// denormal range, so the two roundings of the reference are reproduced.
__device__ __forceinline__ float div5_mul255(float v) {
    const float q = v * 0.2f;
    const float r = fmaf(-q, 5.0f, v);
    return fmaf(r, 0.2f, q) * 255.0f;
}

struct EvTileParams {
    StreamPlan pl;
    float* out;
    int64_t out_stride;
    double tw;             // window length: t_norm = d / tw in float64 (generate_eventvolume.py:141)
    int K;
    int bulk_out;
};

struct EvTileSmem {
    int ring, acc, bars, feed, total;
    __host__ __device__ EvTileSmem(int P, int K) {
        int o = 0;
        ring = o; o += kWsRing * 4;
        acc = o;  o += 2 * K * P * 4;
        bars = o; o += 64;
        feed = o; o += (TileSmemWS::kFeedBytes + 15) / 16 * 16;
        total = o;
    }
};

__global__ void __launch_bounds__(kEvThreads, kEvTilesPerSm)
ev_tile_kernel(EvTileParams tp) {
    const StreamPlan& pl = tp.pl;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const EvTileSmem lay(pl.P, tp.K);
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw + lay.ring);
    float* acc = reinterpret_cast<float*>(smem_raw + lay.acc);             // [2K][P]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);

    const int tid = threadIdx.x, tile = blockIdx.x;
    const int K = tp.K, rows = 2 * K;
    const int64_t HW = (int64_t)pl.H * pl.W;
    const int64_t pix0 = (int64_t)tile * pl.P;
    const int npix = (int)min((int64_t)pl.P, HW - pix0);
    const uint32_t* my_off = pl.off_rel + (int64_t)tile * (pl.TB + 1);
    const uint32_t* my_records = pl.records + pl.tile_base[tile];
    const uint32_t list_len = (pl.tile_total[tile] + 3u) & ~3u;
    const int n_chunks = (int)((list_len + kWsChunkRecords - 1) / kWsChunkRecords);
    auto issue = [&](int c) {           // thread 0 only
        const uint32_t first = (uint32_t)c * kWsChunkRecords;
        const uint32_t bytes = min((uint32_t)kWsChunkRecords, list_len - first) * 4u;
        uint64_t* bar = full + (c % kWsStages);
        mbar_expect_tx(bar, bytes);
        tma_load_1d(ring + (c % kWsStages) * kWsChunkRecords, my_records + first, bytes, bar);
    };
    if (tid == 0) {
        for (int s = 0; s < kWsStages; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int c = 0; c < n_chunks && c < kWsStages; ++c) issue(c);
    BatchFeed feed;
    feed.init(smem_raw + lay.feed, &pl, my_off, tid, 0, kEvThreads);       // barrier 0 = the whole CTA

    int ready_chunk = -1, next_refill = kWsStages;
    const float Kf = (float)K;
    const double inv_tw = 1.0 / tp.tw;
    const int n4 = rows * pl.P / 4;
    const int p4 = pl.P / 4;                                               // float4 columns per row
    const int row_first = tid / p4, c4_first = tid - row_first * p4;
    const int row_step = kEvThreads / p4, c4_step = kEvThreads - row_step * p4;
    for (int i = tid; i < rows * pl.P; i += kEvThreads) acc[i] = 0.0f;     // afterwards the output pass keeps it clean
    __syncthreads();
    for (int j = 0; j < pl.n_batches; ++j) {
        const Batch meta = feed.begin(j);
        const int jb = j & 1;
        // every window is one bin; a zero-bin window still emits an all-zero tensor
        const uint32_t o0 = meta.nb > 0 ? feed.s_off[jb * (kBatchBins + 1)] : 0u;
        const uint32_t o1 = meta.nb > 0 ? feed.s_off[jb * (kBatchBins + 1) + meta.nb] : 0u;
        uint32_t cur = o0;
        while (cur < o1) {
            const uint32_t avail = (uint32_t)next_refill * kWsChunkRecords;     // records requested so far
            const uint32_t limit = o1 < avail ? o1 : avail;
            const int last_c = (int)((limit - 1) / kWsChunkRecords);
            while (ready_chunk < last_c) {
                ++ready_chunk;
                mbar_wait(full + (ready_chunk % kWsStages), (uint32_t)(ready_chunk / kWsStages) & 1u);
            }
            for (uint32_t r = cur + tid; r < limit; r += kEvThreads) {
                const uint32_t rec = ring[r & (kWsRing - 1)];
                const uint32_t lp = (rec >> 1) & 0x1FFFu, pol = rec & 1u;
                // (t - t0) / tw in float64 (:141) then .float() (:23): reciprocal + one Newton step
                const double dd = (double)(rec >> 14);
                const double q0 = dd * inv_tw;
                const float tn = (float)fma(fma(-q0, tp.tw, dd), inv_tw, q0);
                const float ts = Kf * tn;                                        // t* = K * t
                const int c0 = (int)floorf(ts);
#pragma unroll
                for (int d = 0; d < 2; ++d) {                                     // centres c0, c0 + 1 (1..K)
                    const int c = c0 + d;
                    if (c < 1 || c > K) continue;
                    const float w = 1.0f - fabsf((float)c - ts);
                    if (w > 0.0f) atomicAdd(acc + (2 * (c - 1) + (1 - (int)pol)) * pl.P + lp, w);
                }
            }
            cur = limit;
            if (cur < o1) {                                        // the window outgrew the ring: recycle stages
                __syncthreads();
                const int drained = (int)(cur / kWsChunkRecords);
                if (tid == 0)
                    for (int r = next_refill; r < drained + kWsStages && r < n_chunks; ++r) issue(r);
                next_refill = drained + kWsStages;
            }
        }
        __syncthreads();
        {
            const int drained = (int)(o1 / kWsChunkRecords);
            if (drained + kWsStages > next_refill) {
                if (tid == 0)
                    for (int r = next_refill; r < drained + kWsStages && r < n_chunks; ++r) issue(r);
                next_refill = drained + kWsStages;
            }
        }
        if (meta.flags & 2) {
            // read, clear and scale the accumulator (:37  / 5 * 255), 16 bytes per thread straight
            // to global memory: each row is contiguous, so every warp store is 512 contiguous bytes
            float* o = tp.out + (int64_t)meta.win * tp.out_stride + pix0;
            if (tp.bulk_out) {
                // (row, column) of float4 number i = tid + k * kEvThreads, advanced without divisions
                int row = row_first, c4 = c4_first;
                for (int i = tid; i < n4; i += kEvThreads) {
                    const int lp = c4 * 4;
                    float4 v = reinterpret_cast<float4*>(acc)[i];
                    reinterpret_cast<float4*>(acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (lp < npix) {
                        v.x = div5_mul255(v.x); v.y = div5_mul255(v.y); v.z = div5_mul255(v.z); v.w = div5_mul255(v.w);
                        __stcs(reinterpret_cast<float4*>(o + (int64_t)row * HW + lp), v);
                    }
                    row += row_step; c4 += c4_step;
                    if (c4 >= p4) { c4 -= p4; ++row; }
                }
            } else {
                for (int i = tid; i < rows * pl.P; i += kEvThreads) {
                    const int row = i / pl.P, lp = i - row * pl.P;
                    const float v = acc[i];
                    acc[i] = 0.0f;
                    if (lp < npix) __stcs(o + (int64_t)row * HW + lp, div5_mul255(v));
                }
            }
        }
        feed.end(j);                                               // also orders the clears before the next splat
    }
}

// ---- host side -------------------------------------------------------------------------
struct Layout {
    int P, n_tiles, slots;
    int64_t o_wbegin, o_wend, o_wstart, o_wnbins, o_wbinbase, o_batches, meta_bytes;
    int64_t o_counts, o_binany, o_offrel, o_tiletotal, o_tilebase, o_tilebits, o_origins, o_records, total;
    int n_batches_max;
};

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static int make_layout(int64_t n_events, int n_windows, int64_t TB, int H, int W, int n_batches, Layout& L,
                       int tiles_per_sm = 1) {
    const int64_t HW = (int64_t)H * W;
    const int sms = sm_count() * tiles_per_sm;
    int64_t P = (HW + sms - 1) / sms;
    P = (P + 31) / 32 * 32;
    if (P > kTafThreads * kMaxSlots) P = kTafThreads * kMaxSlots;
    if (P < 32) P = 32;
    L.P = (int)P;
    L.n_tiles = (int)((HW + P - 1) / P);
    L.slots = (int)((P + kTafThreads - 1) / kTafThreads);
    if (L.n_tiles > kMaxTiles) return EVREP_ERR_RANGE;
    if (n_events >= (1ll << 31) || TB >= (1ll << 24) || (int64_t)L.n_tiles * (TB + 1) >= (1ll << 31)) return EVREP_ERR_RANGE;
    L.n_batches_max = n_batches;
    int64_t o = 0;
    L.o_wbegin = o;   o += align_up(8ll * n_windows, 16);
    L.o_wend = o;     o += align_up(8ll * n_windows, 16);
    L.o_wstart = o;   o += align_up(8ll * n_windows, 16);
    L.o_wnbins = o;   o += align_up(4ll * n_windows, 16);
    L.o_wbinbase = o; o += align_up(4ll * (n_windows + 1), 16);
    L.o_batches = o;  o += align_up(16ll * n_batches, 16);
    L.meta_bytes = o;
    o = align_up(o, 256);
    L.o_counts = o;   o += align_up(4ll * L.n_tiles * TB, 16);
    L.o_binany = o;   o += align_up(4ll * TB, 16);
    L.o_offrel = o;   o += align_up(4ll * L.n_tiles * (TB + 1), 16);
    L.o_tiletotal = o; o += align_up(4ll * L.n_tiles, 16);
    L.o_tilebase = o; o += align_up(4ll * (L.n_tiles + 1), 16);
    L.o_tilebits = o; o += align_up(4ll * L.n_tiles * (n_batches > 0 ? n_batches : 1), 16);
    o = align_up(o, 256);
    L.o_origins = o;  o += align_up((int64_t)sizeof(ChunkOrigin) * (n_events / (kBucketThreads * kBucketPerThread) + 2), 256);
    L.o_records = o;  o += align_up(4ll * (n_events + 4ll * L.n_tiles), 256);
    L.total = o;
    return EVREP_OK;
}

static inline int64_t batches_upper_bound(int n_windows, int64_t TB) {
    return (int64_t)n_windows + TB / kBatchBins + 1;
}

template <int K>
static int launch_tiles_ws(const TileParams& tp, cudaStream_t st) {
    const int slots = (tp.pl.P + kConsumerThreads - 1) / kConsumerThreads;
    const size_t smem = (size_t)TileSmemWS(tp.pl.P, K).total;
#define EVREP_TILE_WS(S)                                                                                  \
    case S:                                                                                               \
        EVREP_CUDA(cudaFuncSetAttribute(taf_tile_ws_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        taf_tile_ws_kernel<K, S><<<tp.pl.n_tiles, kWsThreads, smem, st>>>(tp);                            \
        break;
    switch (slots) {
        EVREP_TILE_WS(1) EVREP_TILE_WS(2) EVREP_TILE_WS(3) EVREP_TILE_WS(4) EVREP_TILE_WS(5) EVREP_TILE_WS(6)
        default: return EVREP_ERR_RANGE;
    }
#undef EVREP_TILE_WS
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

template <int K>
static int launch_tiles(const TileParams& tp, int slots, size_t smem, cudaStream_t st) {
#define EVREP_TILE(S)                                                                                     \
    case S:                                                                                               \
        EVREP_CUDA(cudaFuncSetAttribute(taf_tile_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        taf_tile_kernel<K, S><<<tp.pl.n_tiles, kTafThreads, smem, st>>>(tp);                              \
        break;
    switch (slots) {
        EVREP_TILE(1) EVREP_TILE(2) EVREP_TILE(3) EVREP_TILE(4) EVREP_TILE(5) EVREP_TILE(6)
        default: return EVREP_ERR_RANGE;
    }
#undef EVREP_TILE
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

// Shared front end of the stream entry points: validates the window list, uploads the
// window / batch tables and runs the bucketing passes.  On return `pl` describes the bucketed
// records of every (tile, bin).
static int prepare_stream(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                          const evrep_taf_window* windows_host, int n_windows, int abin, int H, int W,
                          const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                          void* scratch, int64_t scratch_bytes, cudaStream_t st, StreamPlan& pl, Layout& L,
                          int tiles_per_sm = 1) {
    if (xmap && ymap && (sensor_h <= 0 || sensor_w <= 0 || sensor_h > EVREP_COORD_LUT_LEN || sensor_w > EVREP_COORD_LUT_LEN))
        return EVREP_ERR_ARG;
    if ((uint32_t)abin > kDMax) return EVREP_ERR_RANGE;
    if (!windows_host || (n_events > 0 && (!t || !x || !y || !p))) return EVREP_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(scratch) & 255) return EVREP_ERR_ARG;

    // windows -> bins -> batches (host, O(n_windows + bins / 16))
    int64_t TB = 0;
    int64_t prev_end = 0;
    for (int w = 0; w < n_windows; ++w) {
        const evrep_taf_window& win = windows_host[w];
        if (win.ev_begin < prev_end || win.ev_end < win.ev_begin || win.ev_end > n_events || win.n_bins < 0) return EVREP_ERR_ARG;
        if ((int64_t)win.n_bins * abin >= (1ll << 32)) return EVREP_ERR_RANGE;
        prev_end = win.ev_end;
        TB += win.n_bins;
    }
    std::vector<Batch> batches;
    batches.reserve((size_t)batches_upper_bound(n_windows, TB));
    {
        int gbin = 0;
        for (int w = 0; w < n_windows; ++w) {
            const int nb = windows_host[w].n_bins;
            int done = 0;
            do {
                Batch b;
                b.gbin0 = gbin + done;
                b.nb = nb - done < kBatchBins ? nb - done : kBatchBins;
                b.flags = (done == 0 && windows_host[w].fresh ? 1 : 0) | (done + b.nb >= nb ? 2 : 0);
                b.win = w;
                batches.push_back(b);
                done += b.nb;
            } while (done < nb);
            gbin += nb;
        }
    }
    int rc = make_layout(n_events, n_windows, TB, H, W, (int)batches.size(), L, tiles_per_sm);
    if (rc) return rc;
    if (scratch_bytes < L.total) return EVREP_ERR_SCRATCH;
    // pack and upload the metadata
    std::vector<unsigned char> meta((size_t)L.meta_bytes, 0);
    int64_t* hb = reinterpret_cast<int64_t*>(meta.data() + L.o_wbegin);
    int64_t* he = reinterpret_cast<int64_t*>(meta.data() + L.o_wend);
    int64_t* hs = reinterpret_cast<int64_t*>(meta.data() + L.o_wstart);
    int32_t* hn = reinterpret_cast<int32_t*>(meta.data() + L.o_wnbins);
    int32_t* hbb = reinterpret_cast<int32_t*>(meta.data() + L.o_wbinbase);
    int32_t base = 0;
    for (int w = 0; w < n_windows; ++w) {
        hb[w] = windows_host[w].ev_begin; he[w] = windows_host[w].ev_end; hs[w] = windows_host[w].start_time;
        hn[w] = windows_host[w].n_bins; hbb[w] = base;
        base += windows_host[w].n_bins;
    }
    hbb[n_windows] = base;
    memcpy(meta.data() + L.o_batches, batches.data(), batches.size() * sizeof(Batch));
    char* s = reinterpret_cast<char*>(scratch);
    EVREP_CUDA(cudaMemcpyAsync(s, meta.data(), (size_t)L.meta_bytes, cudaMemcpyHostToDevice, st));
    // pageable source: the copy has been staged when the call returns, `meta` may die

    pl.w_begin = reinterpret_cast<const int64_t*>(s + L.o_wbegin);
    pl.w_end = reinterpret_cast<const int64_t*>(s + L.o_wend);
    pl.w_start = reinterpret_cast<const int64_t*>(s + L.o_wstart);
    pl.w_nbins = reinterpret_cast<const int32_t*>(s + L.o_wnbins);
    pl.w_binbase = reinterpret_cast<const int32_t*>(s + L.o_wbinbase);
    pl.batches = reinterpret_cast<const Batch*>(s + L.o_batches);
    pl.counts = reinterpret_cast<uint32_t*>(s + L.o_counts);
    pl.bin_any = reinterpret_cast<uint32_t*>(s + L.o_binany);
    pl.off_rel = reinterpret_cast<uint32_t*>(s + L.o_offrel);
    pl.tile_total = reinterpret_cast<uint32_t*>(s + L.o_tiletotal);
    pl.tile_base = reinterpret_cast<uint32_t*>(s + L.o_tilebase);
    pl.records = reinterpret_cast<uint32_t*>(s + L.o_records);
    pl.tile_bits = reinterpret_cast<uint32_t*>(s + L.o_tilebits);
    pl.n_windows = n_windows; pl.n_batches = (int)batches.size(); pl.TB = (int)TB;
    pl.n_tiles = L.n_tiles; pl.P = L.P; pl.H = H; pl.W = W;
    pl.div_abin = FastDiv::make((uint32_t)abin);
    pl.div_P = FastDiv::make((uint32_t)L.P);
    pl.abin = (uint32_t)abin;

    if (TB > 0) {
        EVREP_CUDA(cudaMemsetAsync(s + L.o_counts, 0, (size_t)(L.o_offrel - L.o_counts), st));   // counts + bin_any
        // chunks start on a multiple of 4 events so that full chunks can use vector loads
        const int64_t ev_first = windows_host[0].ev_begin & ~3ll, ev_last = windows_host[n_windows - 1].ev_end;
        const int vec_ok = !((reinterpret_cast<uintptr_t>(t) & 15) | (reinterpret_cast<uintptr_t>(x) & 7) |
                             (reinterpret_cast<uintptr_t>(y) & 7) | (reinterpret_cast<uintptr_t>(p) & 3));
        const int64_t per_cta = kBucketThreads * kBucketPerThread;
        const int64_t n_chunks = (ev_last - ev_first + per_cta - 1) / per_cta;
        if (n_chunks >= (1ll << 31)) return EVREP_ERR_RANGE;
        const int grid = (int)(n_chunks < 2ll * sm_count() ? n_chunks : 2ll * sm_count());
        const bool use_lut = xmap && ymap;
        const int nh = kLocalBins * L.n_tiles;
        const size_t smem_count = (size_t)BucketSmem(use_lut ? sensor_w : 0, use_lut ? sensor_h : 0, nh, false).total;
        const size_t smem_scatter = (size_t)BucketSmem(use_lut ? sensor_w : 0, use_lut ? sensor_h : 0, nh, true).total;
        EVREP_CUDA(cudaFuncSetAttribute(taf_bucket_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count));
        EVREP_CUDA(cudaFuncSetAttribute(taf_bucket_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scatter));
        SoA ev{t, x, y, p, xmap, ymap};
        ChunkOrigin* origins = reinterpret_cast<ChunkOrigin*>(s + L.o_origins);
        if (grid > 0) {
            taf_chunk_origin_kernel<<<(int)((n_chunks + 255) / 256), 256, 0, st>>>(ev, pl, ev_first, ev_last, (int)n_chunks, origins);
            EVREP_LAUNCH_CHECK();
            taf_bucket_kernel<false><<<grid, kBucketThreads, smem_count, st>>>(ev, pl, ev_first, ev_last, (int)n_chunks, sensor_w, sensor_h, origins, vec_ok);
            EVREP_LAUNCH_CHECK();
        }
        taf_scan_rows_kernel<<<L.n_tiles, 256, 0, st>>>(pl);
        EVREP_LAUNCH_CHECK();
        taf_scan_tiles_kernel<<<1, 1024, 0, st>>>(pl);
        EVREP_LAUNCH_CHECK();
        taf_tile_bits_kernel<<<L.n_tiles, 256, 0, st>>>(pl);
        EVREP_LAUNCH_CHECK();
        if (grid > 0) {
            taf_bucket_kernel<true><<<grid, kBucketThreads, smem_scatter, st>>>(ev, pl, ev_first, ev_last, (int)n_chunks, sensor_w, sensor_h, origins, vec_ok);
            EVREP_LAUNCH_CHECK();
        }
    } else {
        EVREP_CUDA(cudaMemsetAsync(s + L.o_tiletotal, 0, (size_t)(L.o_origins - L.o_tiletotal), st));   // totals, bases, tile bits
    }

    return EVREP_OK;
}

}  // namespace evrep

using namespace evrep;

extern "C" {

int64_t evrep_taf_stream_scratch_bytes(int64_t n_events, int n_windows, int64_t total_bins, int H, int W) {
    if (n_events < 0 || n_windows < 0 || total_bins < 0 || H <= 0 || W <= 0) return EVREP_ERR_ARG;
    Layout L;
    int rc = make_layout(n_events, n_windows, total_bins, H, W, (int)batches_upper_bound(n_windows, total_bins), L);
    if (rc) return rc;
    return L.total;
}

int evrep_taf_stream(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                     const evrep_taf_window* windows_host, int n_windows, int abin, int H, int W, int K,
                     const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                     float* state_inout, int emit_state_every_window,
                     float* out, int64_t out_stride, void* scratch, int64_t scratch_bytes,
                     void* ev_tiles_begin, void* ev_tiles_end, evrep_stream_t stream) {
    if (n_events < 0 || n_windows < 0 || H <= 0 || W <= 0 || abin <= 0 || !state_inout || !scratch) return EVREP_ERR_ARG;
    if (K != 4 && K != 8) return EVREP_ERR_ARG;
    if (n_windows == 0) return EVREP_OK;
    if (!out || (reinterpret_cast<uintptr_t>(state_inout) & 15)) return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    StreamPlan pl;
    Layout L;
    int rc = prepare_stream(t, x, y, p, n_events, windows_host, n_windows, abin, H, W, xmap, ymap, sensor_h, sensor_w,
                            scratch, scratch_bytes, st, pl, L);
    if (rc) return rc;

    TileParams tp;
    tp.pl = pl; tp.state = state_inout; tp.out = out; tp.out_stride = out_stride;
    tp.emit_state = emit_state_every_window;
    tp.n_emits = n_windows;
    tp.span = (float)((double)abin + 1e-8);
    tp.bulk_out = (((int64_t)H * W) % 4 == 0 && out_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;

    const size_t smem = (size_t)TileSmem(L.P, K).total;
    if (ev_tiles_begin) EVREP_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev_tiles_begin), st));
    const char* legacy = getenv("EVREP_TAF_TILE_KERNEL");            // "single" = the non-specialised kernel (A/B runs)
    const bool ws = !(legacy && strcmp(legacy, "single") == 0) && (size_t)TileSmemWS(L.P, K).total <= 232448;
    if (ws) rc = K == 8 ? launch_tiles_ws<8>(tp, st) : launch_tiles_ws<4>(tp, st);
    else rc = K == 8 ? launch_tiles<8>(tp, L.slots, smem, st) : launch_tiles<4>(tp, L.slots, smem, st);
    if (rc) return rc;
    if (ev_tiles_end) EVREP_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev_tiles_end), st));
    return EVREP_OK;
}

int64_t evrep_event_volume_stream_scratch_bytes(int64_t n_events, int n_windows, int H, int W) {
    if (n_events < 0 || n_windows < 0 || H <= 0 || W <= 0) return EVREP_ERR_ARG;
    Layout L;
    int rc = make_layout(n_events, n_windows, n_windows, H, W, (int)batches_upper_bound(n_windows, n_windows), L, kEvTilesPerSm);
    if (rc) return rc;
    return L.total;
}

int evrep_event_volume_stream(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                              const evrep_ev_window* windows_host, int n_windows, int64_t tw, int H, int W, int K,
                              const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                              float* out, int64_t out_stride, void* scratch, int64_t scratch_bytes,
                              evrep_stream_t stream) {
    if (n_events < 0 || n_windows < 0 || H <= 0 || W <= 0 || tw <= 0 || K < 1 || !scratch) return EVREP_ERR_ARG;
    if (tw > (int64_t)kDMax) return EVREP_ERR_RANGE;
    if (n_windows == 0) return EVREP_OK;
    if (!out || !windows_host) return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    std::vector<evrep_taf_window> wins((size_t)n_windows);
    for (int w = 0; w < n_windows; ++w) {
        wins[w].ev_begin = windows_host[w].ev_begin; wins[w].ev_end = windows_host[w].ev_end;
        wins[w].start_time = windows_host[w].t0; wins[w].n_bins = 1; wins[w].fresh = 0;
    }
    StreamPlan pl;
    Layout L;
    int rc = prepare_stream(t, x, y, p, n_events, wins.data(), n_windows, (int)tw, H, W, xmap, ymap, sensor_h, sensor_w,
                            scratch, scratch_bytes, st, pl, L, kEvTilesPerSm);
    if (rc) return rc;
    const size_t smem = (size_t)EvTileSmem(L.P, K).total;
    if (smem > 232448) return EVREP_ERR_RANGE;
    EvTileParams tp;
    tp.pl = pl; tp.out = out; tp.out_stride = out_stride; tp.tw = (double)tw; tp.K = K;
    tp.bulk_out = (((int64_t)H * W) % 4 == 0 && out_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                   L.P % 4 == 0) ? 1 : 0;
    EVREP_CUDA(cudaFuncSetAttribute(ev_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ev_tile_kernel<<<L.n_tiles, kEvThreads, smem, st>>>(tp);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

}  // extern "C"
