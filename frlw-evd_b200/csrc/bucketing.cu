// Bucketing passes of the whole-stream encoders: a counting sort of the time-ordered events
// by (sensor tile, bin) into packed 4-byte records  [ d:18 | local pixel:13 | p:1 ],
// d = t - bin start.  Count (shared-memory histograms per 4096-event chunk), scan (per tile
// row, then across tiles), scatter (records ordered in shared memory, runs written with
// consecutive addresses).  Also the host-side front end shared by the stream entry points.
#include "slices.cuh"

namespace evrep {

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

// Block-wide exclusive scan of n shared-memory counters by kThreads threads.
// Returns the total.  `tmp` holds one word per warp (+1).
template <int kThreads = kBucketThreads>
__device__ __forceinline__ uint32_t block_exclusive_scan(const uint32_t* in, uint32_t* out, int n, uint32_t* tmp) {
    const int per = (n + kThreads - 1) / kThreads;
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
        uint32_t w = lane < kThreads / 32 ? tmp[lane] : 0u, wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= o) wi += v;
        }
        if (lane < kThreads / 32) tmp[lane] = wi - w;
        if (lane == kThreads / 32 - 1) tmp[kThreads / 32] = wi;
    }
    __syncthreads();
    uint32_t run = tmp[wid] + incl - mine;
    for (int i = lo; i < hi; ++i) { const uint32_t c = in[i]; out[i] = run; run += c; }
    const uint32_t total = tmp[kThreads / 32];
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
taf_chunk_origin_kernel(SoA ev, StreamPlan pl, int64_t ev_first, int64_t ev_last, int n_chunks, int chunk_events,
                        ChunkOrigin* __restrict__ origins) {
    const int chunk = blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk >= n_chunks) return;
    const int64_t c0 = ev_first + (int64_t)chunk * chunk_events;
    const int64_t c1 = min(c0 + chunk_events, ev_last);
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
        lutx = o;  o += (lut_w * 4 + 15) / 16 * 16;      // column of a raw x (u32), kOffGrid when dropped
        luty = o;  o += (lut_h * 4 + 15) / 16 * 16;      // first pixel of the row of a raw y, kOffGrid when dropped
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
//  count + save (kMode = 2): the count pass that also stores every event's record and 16-bit
//          histogram key, so that taf_scatter_saved_kernel need not classify the events again.
enum : int { kModeCount = 0, kModeScatter = 1, kModeCountSave = 2 };
constexpr uint32_t kKeyFar = 0xFFFEu, kKeyDropped = 0xFFFFu;     // saved keys that are not shared-memory counters

template <int kMode>
__global__ void __launch_bounds__(kBucketThreads, 2)
taf_bucket_kernel(SoA ev, StreamPlan pl, int64_t ev_first, int64_t ev_last, int n_chunks, int lut_w, int lut_h,
                  const ChunkOrigin* __restrict__ origins, int vec_ok, int stage_ok) {
    constexpr bool kScatter = kMode == kModeScatter;
    constexpr bool kSave = kMode == kModeCountSave;
    extern __shared__ __align__(128) unsigned char bsm[];
    const int nh = kLocalBins * pl.n_tiles;
    const BucketSmem lay(lut_w, lut_h, nh, kScatter);
    // Count + save pass: a CTA's NEXT full chunk is copied into shared memory by TMA bulk copies while the current one is
    // classified (stage_ok: the event arrays are 16-byte aligned), so that no thread waits for DRAM: a quarter of this
    // kernel's stall samples were the first use of the chunk's global loads.
    constexpr uint32_t kChunkEvents = kBucketThreads * kBucketPerThread;
    unsigned char* stage = bsm + (lay.total + 127) / 128 * 128;          // t (4 B), x (2 B), y (2 B), p (1 B) per event
    const uint32_t* st_t = reinterpret_cast<const uint32_t*>(stage);
    const uint16_t* st_x = reinterpret_cast<const uint16_t*>(stage + kChunkEvents * 4);
    const uint16_t* st_y = reinterpret_cast<const uint16_t*>(stage + kChunkEvents * 6);
    const uint8_t* st_p = stage + kChunkEvents * 8;
    __shared__ __align__(8) uint64_t stage_bar;
    const bool staging = kSave && stage_ok != 0;
    auto stage_issue = [&](int chunk) {                                   // thread 0: full chunks only
        const int64_t b0 = ev_first + (int64_t)chunk * kChunkEvents;
        if (chunk >= n_chunks || b0 + kChunkEvents > ev_last) return;
        mbar_expect_tx(&stage_bar, kChunkEvents * 9u);
        tma_load_1d(stage, ev.t + b0, kChunkEvents * 4u, &stage_bar);
        tma_load_1d(stage + kChunkEvents * 4, ev.x + b0, kChunkEvents * 2u, &stage_bar);
        tma_load_1d(stage + kChunkEvents * 6, ev.y + b0, kChunkEvents * 2u, &stage_bar);
        tma_load_1d(stage + kChunkEvents * 8, ev.p + b0, kChunkEvents, &stage_bar);
    };
    if (staging && threadIdx.x == 0) {
        mbar_init(&stage_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stage_issue((int)blockIdx.x);
    }
    uint32_t n_staged = 0;                                                // staged chunks consumed so far (barrier phase)
    uint32_t* s_col = reinterpret_cast<uint32_t*>(bsm + lay.lutx);
    uint32_t* s_row = reinterpret_cast<uint32_t*>(bsm + lay.luty);
    uint32_t* hist = reinterpret_cast<uint32_t*>(bsm + lay.hist);
    uint32_t* loff = reinterpret_cast<uint32_t*>(bsm + lay.loff);
    uint32_t* gbase = reinterpret_cast<uint32_t*>(bsm + lay.gbase);
    uint32_t* sorted = reinterpret_cast<uint32_t*>(bsm + lay.sorted);
    uint16_t* skey = reinterpret_cast<uint16_t*>(bsm + lay.skey);
    __shared__ uint32_t s_tmp[kBucketThreads / 32 + 1];

    // Coordinate tables: raw x -> grid column, raw y -> first pixel of the grid row; entries that
    // leave the grid hold kOffGrid, so that "row + column < H W" is the only range test per event.
    // Without user LUTs the tables are the identity over the grid.
    const uint32_t W = pl.W, H = pl.H, HW = W * H;
    constexpr uint32_t kOffGrid = 0x40000000u;
    for (int i = threadIdx.x; i < lut_w; i += kBucketThreads) {
        const uint32_t xm = ev.xmap ? ev.xmap[i] : (uint32_t)i;
        s_col[i] = xm < W ? xm : kOffGrid;
    }
    for (int i = threadIdx.x; i < lut_h; i += kBucketThreads) {
        const uint32_t ym = ev.ymap ? ev.ymap[i] : (uint32_t)i;
        s_row[i] = ym < H ? ym * W : kOffGrid;
    }
    uint32_t col_addr = smem_u32(s_col), row_addr = smem_u32(s_row);
    uint32_t n_cols = (uint32_t)lut_w, n_rows = (uint32_t)lut_h;
    uint32_t tile_mul = pl.tile_mul, P = (uint32_t)pl.P, n_tiles = (uint32_t)pl.n_tiles;
    // Per-event constants pinned in registers: left to itself the compiler re-derives them for every event (the shared
    // window base through S2UR / ULEA, the kernel parameters through LDCU): a fifth of the pass's instructions.
    uint32_t k_mul = pl.div_abin.mul, k_sh1 = pl.div_abin.sh1, k_sh2 = pl.div_abin.sh2, k_abin = pl.abin, k_hw = W * H;
    asm volatile("" : "+r"(col_addr), "+r"(row_addr), "+r"(n_cols), "+r"(n_rows));
    asm volatile("" : "+r"(tile_mul), "+r"(P), "+r"(n_tiles), "+r"(k_hw));
    asm volatile("" : "+r"(k_mul), "+r"(k_sh1), "+r"(k_sh2), "+r"(k_abin));
    auto div_abin = [&](uint32_t n) -> uint32_t {      // FastDiv::div with the constants above
        const uint32_t t1 = __umulhi(k_mul, n);
        return (t1 + ((n - t1) >> k_sh1)) >> k_sh2;
    };

    for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int64_t c0 = ev_first + (int64_t)chunk * (kBucketThreads * kBucketPerThread);
        const int64_t c1 = min(c0 + kBucketThreads * kBucketPerThread, ev_last);
        // same address for every thread: one broadcast load.  In save mode the local-bin window is
        // that of the enclosing scatter chunk, whose keys the scatter pass shares.
        const ChunkOrigin org = origins[kSave ? chunk / kScatterChunks : chunk];
        __syncthreads();                                   // previous chunk is done with smem
        for (int i = threadIdx.x; i < nh; i += kBucketThreads) hist[i] = 0;
        const int gb0 = org.gb0;
        const bool single = org.single != 0;
        int w = org.w0;
        WinInfo wi = org.win;

        // all global loads of the chunk are issued before any of them is used
        const bool fast = single && (c1 - c0) == kBucketThreads * kBucketPerThread &&
                          wi.start >= 0 && wi.start <= 0xFFFFFFFFll;
        // per event: timestamp, x | y << 16, and the polarity bytes four to a word (0xFF = no event)
        uint32_t tt[kBucketPerThread], xy[kBucketPerThread], pw[kBucketPerThread / 4];
        static_assert(kBucketPerThread % 4 == 0, "events are handled in groups of 4");
        const bool staged = staging && (c1 - c0) == kBucketThreads * kBucketPerThread;
        if (staged) {
            mbar_wait(&stage_bar, n_staged & 1u);
            ++n_staged;
            if (fast) {
#pragma unroll
                for (int g = 0; g < kBucketPerThread / 4; ++g) {
                    const uint32_t base = ((uint32_t)g * kBucketThreads + threadIdx.x) * 4u;
                    const uint4 t4 = *reinterpret_cast<const uint4*>(st_t + base);
                    const uint2 x4 = *reinterpret_cast<const uint2*>(st_x + base);
                    const uint2 y4 = *reinterpret_cast<const uint2*>(st_y + base);
                    pw[g] = *reinterpret_cast<const uint32_t*>(st_p + base);
                    tt[4 * g + 0] = t4.x; tt[4 * g + 1] = t4.y; tt[4 * g + 2] = t4.z; tt[4 * g + 3] = t4.w;
                    xy[4 * g + 0] = __byte_perm(x4.x, y4.x, 0x5410); xy[4 * g + 1] = __byte_perm(x4.x, y4.x, 0x7632);
                    xy[4 * g + 2] = __byte_perm(x4.y, y4.y, 0x5410); xy[4 * g + 3] = __byte_perm(x4.y, y4.y, 0x7632);
                }
            } else {
#pragma unroll
                for (int g = 0; g < kBucketPerThread / 4; ++g) {
                    uint32_t pol4 = 0;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int k = 4 * g + e;
                        const uint32_t i = (uint32_t)k * kBucketThreads + threadIdx.x;
                        tt[k] = st_t[i]; xy[k] = st_x[i] | ((uint32_t)st_y[i] << 16);
                        pol4 |= (uint32_t)st_p[i] << (8 * e);
                    }
                    pw[g] = pol4;
                }
            }
        } else if (fast && vec_ok) {
            // 4 consecutive events per 128/64/64/32-bit load (c0 is a multiple of 4 events)
#pragma unroll
            for (int g = 0; g < kBucketPerThread / 4; ++g) {
                const int64_t base = c0 + ((int64_t)g * kBucketThreads + threadIdx.x) * 4;
                const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(ev.t + base));
                const uint2 x4 = __ldg(reinterpret_cast<const uint2*>(ev.x + base));
                const uint2 y4 = __ldg(reinterpret_cast<const uint2*>(ev.y + base));
                pw[g] = __ldg(reinterpret_cast<const uint32_t*>(ev.p + base));
                tt[4 * g + 0] = t4.x; tt[4 * g + 1] = t4.y; tt[4 * g + 2] = t4.z; tt[4 * g + 3] = t4.w;
                xy[4 * g + 0] = __byte_perm(x4.x, y4.x, 0x5410); xy[4 * g + 1] = __byte_perm(x4.x, y4.x, 0x7632);
                xy[4 * g + 2] = __byte_perm(x4.y, y4.y, 0x5410); xy[4 * g + 3] = __byte_perm(x4.y, y4.y, 0x7632);
            }
        } else {
#pragma unroll
            for (int g = 0; g < kBucketPerThread / 4; ++g) {
                uint32_t pol4 = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = 4 * g + e;
                    const int64_t i = c0 + k * kBucketThreads + threadIdx.x;
                    uint32_t pv = 0xFFu;
                    tt[k] = 0; xy[k] = 0;
                    if (i < c1) { tt[k] = __ldg(ev.t + i); xy[k] = __ldg(ev.x + i) | ((uint32_t)__ldg(ev.y + i) << 16); pv = __ldg(ev.p + i); }
                    pol4 |= pv << (8 * e);
                }
                pw[g] = pol4;
            }
        }

        __syncthreads();                                   // histogram is zeroed; the staged chunk is in registers
        if (staging && threadIdx.x == 0) stage_issue(chunk + (int)gridDim.x);
        // per event: (smem counter << 12) | rank inside the chunk, or kNone when dropped
        constexpr uint32_t kNone = 0xFFFFFFFFu;
        uint32_t slot[kScatter ? kBucketPerThread : 1], rec[kBucketPerThread] = {};
        uint32_t key16[kSave ? kBucketPerThread : 1];

        // count / rank one classified event
        auto deposit = [&](int k, uint32_t tile, int gbin) {
            const uint32_t lb = (uint32_t)(gbin - gb0);
            if (lb < (uint32_t)kLocalBins) {
                const uint32_t key = lb * n_tiles + tile;
                if (kScatter) slot[k] = (key << 12) | atomicAdd(&hist[key], 1u);
                else atomicAdd(&hist[key], 1u);
                if (kSave) key16[k] = key;
            } else {                          // unsorted input or a very sparse stream: go straight to global
                uint32_t* cursor = pl.counts + (int64_t)tile * pl.TB + gbin;
                if (kScatter)
                    pl.records[pl.tile_base[tile] + pl.off_rel[(int64_t)tile * (pl.TB + 1) + gbin] + atomicAdd(cursor, 1u)] = rec[k];
                else { atomicAdd(cursor, 1u); pl.bin_any[gbin] = 1u; }
                if (kSave) key16[k] = kKeyFar;
            }
        };
        // grid pixel and polarity of event k; false when the event is to be dropped
        auto locate = [&](int k, uint32_t& pix, uint32_t& pol) -> bool {
            const uint32_t xv = xy[k] & 0xFFFFu, yv = xy[k] >> 16;
            pol = (pw[k >> 2] >> ((k & 3) * 8)) & 0xFFu;
            if (xv >= n_cols || yv >= n_rows || pol > 1u) return false;
            pix = lds_u32(col_addr + xv * 4u) + lds_u32(row_addr + yv * 4u);
            return pix < k_hw;
        };
        // tile of a pixel: one multiply when the host proved the magic number exact for this grid
        auto tile_of = [&](uint32_t pix) -> uint32_t { return tile_mul ? __umulhi(pix, tile_mul) : pl.div_P.div(pix); };

        if (fast) {
            // the whole chunk lies in one window: 32-bit time arithmetic, no bounds checks
            const uint32_t start32 = (uint32_t)wi.start, zmax = (uint32_t)(wi.nbins - 1);
#pragma unroll
            for (int k = 0; k < kBucketPerThread; ++k) {
                if (kScatter) slot[k] = kNone;
                if (kSave) key16[k] = kKeyDropped;
                uint32_t pix, pol;
                if (locate(k, pix, pol)) {
                    const uint32_t u = max(tt[k], start32) - start32;
                    const uint32_t z = min(div_abin(u), zmax);
                    const uint32_t tile = tile_of(pix);
                    if (kScatter || kSave) rec[k] = (min(u - z * k_abin, kDMax) << 14) | ((pix - tile * P) << 1) | pol;
                    deposit(k, tile, wi.binbase + (int)z);
                }
                if (kSave && vec_ok && (k & 3) == 3) {
                    // events 4g .. 4g+3 of a thread are consecutive in the stream: one 16-byte and one 8-byte store
                    const int64_t at = c0 - ev_first + ((int64_t)(k >> 2) * kBucketThreads + threadIdx.x) * 4;
                    *reinterpret_cast<uint4*>(pl.saved_rec + at) = make_uint4(rec[k - 3], rec[k - 2], rec[k - 1], rec[k]);
                    *reinterpret_cast<uint2*>(pl.saved_key + at) = make_uint2(key16[k - 3] | (key16[k - 2] << 16), key16[k - 1] | (key16[k] << 16));
                }
            }
            if (kSave && !vec_ok) {
#pragma unroll
                for (int k = 0; k < kBucketPerThread; ++k) {
                    const int64_t at = c0 - ev_first + k * kBucketThreads + threadIdx.x;
                    pl.saved_rec[at] = rec[k]; pl.saved_key[at] = (uint16_t)key16[k];
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < kBucketPerThread; ++k) {
                const int64_t i = c0 + k * kBucketThreads + threadIdx.x;
                if (kScatter) slot[k] = kNone;
                if (i >= c1) continue;
                bool in_window = true;
                if (!single) {                  // chunk straddles a window boundary or a gap
                    while (w < pl.n_windows && i >= __ldg(pl.w_end + w)) ++w;
                    in_window = w < pl.n_windows;
                    if (in_window) {
                        wi = load_window(pl, w);
                        in_window = i >= wi.begin && wi.nbins > 0;
                    }
                }
                uint32_t pix, pol, key = kKeyDropped;
                if (in_window && locate(k, pix, pol)) {
                    uint32_t z, d;
                    bin_of(pl, wi, tt[k], z, d);
                    const uint32_t tile = tile_of(pix);
                    rec[k] = (d << 14) | ((pix - tile * P) << 1) | pol;
                    deposit(k, tile, wi.binbase + (int)z);
                    if (kSave) key = key16[k];
                }
                if (kSave) { pl.saved_rec[i - ev_first] = rec[k]; pl.saved_key[i - ev_first] = (uint16_t)key; }
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

// Tile, global bin and record of event i, from global memory only (the rare events whose bin lies
// outside the local-bin window of their chunk: unsorted input, very sparse streams).
__device__ __noinline__ bool classify_general(const SoA& ev, const StreamPlan& pl, int lut_w, int lut_h, int64_t i,
                                              uint32_t& tile, int& gbin, uint32_t& rec) {
    const int w = first_window(pl, i);
    if (w >= pl.n_windows) return false;
    const WinInfo wi = load_window(pl, w);
    if (i < wi.begin || wi.nbins <= 0) return false;
    uint32_t xm = ev.x[i], ym = ev.y[i];
    const uint32_t pol = ev.p[i];
    if (xm >= (uint32_t)lut_w || ym >= (uint32_t)lut_h || pol > 1u) return false;
    if (ev.xmap) xm = ev.xmap[xm];
    if (ev.ymap) ym = ev.ymap[ym];
    if (xm >= (uint32_t)pl.W || ym >= (uint32_t)pl.H) return false;
    uint32_t z, d;
    bin_of(pl, wi, ev.t[i], z, d);
    const uint32_t pix = ym * pl.W + xm;
    tile = pl.div_P.div(pix);
    gbin = wi.binbase + (int)z;
    rec = (d << 14) | ((pix - tile * pl.P) << 1) | pol;
    return true;
}

// Scatter pass over the records and keys saved by the count pass: no classification, 16 events per
// thread (8192 per CTA, so the runs written per (tile, bin) are twice as long).  Ranks come from
// returning shared-memory atomics, then block scan, one global reservation per non-empty counter,
// records ordered in shared memory and written run by run.
__global__ void __launch_bounds__(kSavedThreads, kSavedCtasPerSm)
taf_scatter_saved_kernel(SoA ev, StreamPlan pl, int64_t ev_first, int64_t ev_last, int n_super, int lut_w, int lut_h,
                         const ChunkOrigin* __restrict__ origins) {
    extern __shared__ __align__(16) unsigned char bsm[];
    constexpr int kPer = kScatterPerThread;
    constexpr int64_t kSuper = (int64_t)kSavedThreads * kPer;
    const int nh = kLocalBins * pl.n_tiles;
    uint32_t* hist = reinterpret_cast<uint32_t*>(bsm);
    uint32_t* loff = hist + nh;
    uint32_t* gbase = loff + nh;
    uint32_t* sorted = gbase + nh;
    uint16_t* skey = reinterpret_cast<uint16_t*>(sorted + kSuper);
    __shared__ uint32_t s_tmp[kSavedThreads / 32 + 1];
    const uint32_t n_tiles = (uint32_t)pl.n_tiles;
    constexpr uint32_t kNone = 0xFFFFFFFFu;

    for (int sc = blockIdx.x; sc < n_super; sc += gridDim.x) {
        const int64_t c0 = ev_first + (int64_t)sc * kSuper;
        const int64_t c1 = min(c0 + kSuper, ev_last);
        const int gb0 = origins[sc].gb0;
        __syncthreads();                                   // previous chunk is done with smem
        for (int i = threadIdx.x; i < nh; i += kSavedThreads) hist[i] = 0;
        uint32_t rec[kPer], key[kPer];
        if (c1 - c0 == kSuper) {
#pragma unroll
            for (int g = 0; g < kPer / 4; ++g) {
                const int64_t at = c0 - ev_first + ((int64_t)g * kSavedThreads + threadIdx.x) * 4;
                const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(pl.saved_rec + at));
                const uint2 k4 = __ldg(reinterpret_cast<const uint2*>(pl.saved_key + at));
                rec[4 * g + 0] = r4.x; rec[4 * g + 1] = r4.y; rec[4 * g + 2] = r4.z; rec[4 * g + 3] = r4.w;
                key[4 * g + 0] = k4.x & 0xFFFFu; key[4 * g + 1] = k4.x >> 16; key[4 * g + 2] = k4.y & 0xFFFFu; key[4 * g + 3] = k4.y >> 16;
            }
        } else {
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                const int64_t i = c0 + ((int64_t)(k >> 2) * kSavedThreads + threadIdx.x) * 4 + (k & 3);
                key[k] = kKeyDropped; rec[k] = 0;
                if (i < c1) { rec[k] = __ldg(pl.saved_rec + (i - ev_first)); key[k] = __ldg(pl.saved_key + (i - ev_first)); }
            }
        }
        __syncthreads();                                   // histogram is zeroed
        // key[k] becomes (counter << 13) | rank inside the chunk, kNone (dropped) or kFarMark
        constexpr uint32_t kFarMark = 0xFFFFFFFEu;
        bool any_far = false;
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            if (key[k] < kKeyFar) key[k] = (key[k] << 13) | atomicAdd(&hist[key[k]], 1u);
            else { any_far |= key[k] == kKeyFar; key[k] = key[k] == kKeyFar ? kFarMark : kNone; }
        }
        if (any_far) {                                     // rare: classify from the raw event, write straight to its run
#pragma unroll
            for (int k = 0; k < kPer; ++k) {
                if (key[k] != kFarMark) continue;
                const int64_t i = c0 + ((int64_t)(k >> 2) * kSavedThreads + threadIdx.x) * 4 + (k & 3);
                uint32_t tile, r;
                int gbin;
                if (classify_general(ev, pl, lut_w, lut_h, i, tile, gbin, r))
                    pl.records[pl.tile_base[tile] + pl.off_rel[(int64_t)tile * (pl.TB + 1) + gbin] +
                               atomicAdd(pl.counts + (int64_t)tile * pl.TB + gbin, 1u)] = r;
            }
        }
        __syncthreads();
        const uint32_t n_valid = block_exclusive_scan<kSavedThreads>(hist, loff, nh, s_tmp);
        for (int i = threadIdx.x; i < nh; i += kSavedThreads) {
            const uint32_t c = hist[i];
            if (!c) continue;
            const uint32_t lb = (uint32_t)i / n_tiles, tile = (uint32_t)i - lb * n_tiles;
            const int gbin = gb0 + (int)lb;
            gbase[i] = pl.tile_base[tile] + pl.off_rel[(int64_t)tile * (pl.TB + 1) + gbin] +
                       atomicAdd(pl.counts + (int64_t)tile * pl.TB + gbin, c) - loff[i];
        }
#pragma unroll
        for (int k = 0; k < kPer; ++k) {
            if (key[k] >= kFarMark) continue;
            const uint32_t kk = key[k] >> 13;
            const uint32_t pos = loff[kk] + (key[k] & 0x1FFFu);
            sorted[pos] = rec[k];
            skey[pos] = (uint16_t)kk;
        }
        __syncthreads();
        for (uint32_t pos = threadIdx.x; pos < n_valid; pos += kSavedThreads)
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

// Window / batch tables are handed to the device as by-value kernel arguments (see prepare_stream).
struct MetaPiece { uint32_t words[960]; };          // 3840 bytes, inside the 4 KB argument limit
__global__ void __launch_bounds__(256)
meta_upload_kernel(const __grid_constant__ MetaPiece piece, uint32_t* __restrict__ dst, int n_words) {
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) dst[i] = piece.words[i];
}

int upload_words(const uint32_t* host, int64_t n_words, uint32_t* dst, cudaStream_t st) {
    constexpr int64_t kPiece = (int64_t)(sizeof(MetaPiece) / 4);
    for (int64_t done = 0; done < n_words; done += kPiece) {
        MetaPiece piece;
        const int64_t n = n_words - done < kPiece ? n_words - done : kPiece;
        memcpy(piece.words, host + done, (size_t)n * 4);
        meta_upload_kernel<<<1, 256, 0, st>>>(piece, dst + done, (int)n);
        EVREP_LAUNCH_CHECK();
    }
    return EVREP_OK;
}

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

int make_layout(int64_t n_events, int n_windows, int64_t TB, int H, int W, int n_batches, Layout& L, int tiles_per_sm) {
    const int64_t HW = (int64_t)H * W;
    const int sms = sm_count() * tiles_per_sm;
    int64_t P = (HW + sms - 1) / sms;
    P = (P + 31) / 32 * 32;
    // at most 6 pixels per consumer thread (2304), and small enough for the warp-specialised tile kernel's
    // shared memory at K = 8 (ring + two accumulators + staging tile <= 227 KB  =>  P <= 2240)
    if (P > kMaxTilePixels) P = kMaxTilePixels;
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
    {   // records / keys saved by the count pass, padded to whole scatter chunks
        const int64_t super = (int64_t)kSavedThreads * kScatterPerThread;
        const int64_t slots = (n_events / super + 2) * super;
        L.o_savedrec = o; o += align_up(4ll * slots, 256);
        L.o_savedkey = o; o += align_up(2ll * slots, 256);
    }
    L.total = o;
    return EVREP_OK;
}

int64_t batches_upper_bound(int n_windows, int64_t TB) {
    return (int64_t)n_windows + TB / kBatchBins + 1;
}

// Shared front end of the stream entry points: validates the window list, uploads the
// window / batch tables and runs the bucketing passes.  On return `pl` describes the bucketed
// records of every (tile, bin).
int prepare_stream(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                          const evrep_taf_window* windows_host, int n_windows, int abin, int H, int W,
                          const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                          void* scratch, int64_t scratch_bytes, cudaStream_t st, StreamPlan& pl, Layout& L,
                   int tiles_per_sm) {
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
    // The tables travel as kernel arguments, not as a host->device copy: a copy would queue on
    // the copy engine behind whatever bulk transfer the caller has in flight on another stream
    // (the event payload of the next chunk in generate_taf.HostPipeline) and stall this stream.
    rc = upload_words(reinterpret_cast<const uint32_t*>(meta.data()), L.meta_bytes / 4, reinterpret_cast<uint32_t*>(s), st);
    if (rc) return rc;

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
    pl.saved_rec = reinterpret_cast<uint32_t*>(s + L.o_savedrec);
    pl.saved_key = reinterpret_cast<uint16_t*>(s + L.o_savedkey);
    pl.n_windows = n_windows; pl.n_batches = (int)batches.size(); pl.TB = (int)TB;
    pl.n_tiles = L.n_tiles; pl.P = L.P; pl.H = H; pl.W = W;
    pl.div_abin = FastDiv::make((uint32_t)abin);
    pl.div_P = FastDiv::make((uint32_t)L.P);
    {   // pix / P as one multiply-high: exact while pix * (mul * P - 2^32) < 2^32
        const uint64_t mul = (1ull << 32) / (uint64_t)L.P + 1;
        const uint64_t err = mul * (uint64_t)L.P - (1ull << 32);
        pl.tile_mul = (mul < (1ull << 32) && (uint64_t)H * W * err < (1ull << 32)) ? (uint32_t)mul : 0u;
    }
    pl.abin = (uint32_t)abin;

    if (TB > 0) {
        EVREP_CUDA(cudaMemsetAsync(s + L.o_counts, 0, (size_t)(L.o_offrel - L.o_counts), st));   // counts + bin_any
        // chunks start on a multiple of 4 events so that full chunks can use vector loads
        // (a multiple of 16 events: the count pass stages whole chunks with 16-byte aligned TMA bulk copies)
        const int64_t ev_first = windows_host[0].ev_begin & ~15ll, ev_last = windows_host[n_windows - 1].ev_end;
        const int vec_ok = !((reinterpret_cast<uintptr_t>(t) & 15) | (reinterpret_cast<uintptr_t>(x) & 7) |
                             (reinterpret_cast<uintptr_t>(y) & 7) | (reinterpret_cast<uintptr_t>(p) & 3));
        const int stage_ok = !((reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(x) |
                                reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(p)) & 15);
        const int64_t per_cta = kBucketThreads * kBucketPerThread;
        const int64_t n_chunks = (ev_last - ev_first + per_cta - 1) / per_cta;
        if (n_chunks >= (1ll << 31)) return EVREP_ERR_RANGE;
        const int grid = (int)(n_chunks < 2ll * sm_count() ? n_chunks : 2ll * sm_count());
        // coordinate tables of the bucketing kernels: the user LUTs over the sensor, or the identity over the grid
        const bool use_lut = xmap && ymap;
        const int lut_w = use_lut ? sensor_w : W, lut_h = use_lut ? sensor_h : H;
        const int nh = kLocalBins * L.n_tiles;
        const size_t smem_count = (size_t)BucketSmem(lut_w, lut_h, nh, false).total;
        const size_t smem_save = (smem_count + 127) / 128 * 128 + (size_t)kBucketThreads * kBucketPerThread * 9;   // + the staged chunk
        const size_t smem_scatter = (size_t)BucketSmem(lut_w, lut_h, nh, true).total;
        if (smem_scatter > 100 * 1024) return EVREP_ERR_RANGE;          // two CTAs per SM
        SoA ev{t, x, y, p, xmap, ymap};
        ChunkOrigin* origins = reinterpret_cast<ChunkOrigin*>(s + L.o_origins);
        // Default: the count pass saves every event's record and key and the scatter pass sorts those
        // (no second classification).  EVREP_BUCKETING=reclassify selects the older scatter pass that
        // reads and classifies the raw events again (kept for A/B runs).
        const char* mode_env = getenv("EVREP_BUCKETING");
        const bool reclassify = mode_env && strcmp(mode_env, "reclassify") == 0;
        if (reclassify) {
            EVREP_CUDA(cudaFuncSetAttribute(taf_bucket_kernel<kModeCount>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_count));
            EVREP_CUDA(cudaFuncSetAttribute(taf_bucket_kernel<kModeScatter>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_scatter));
            if (grid > 0) {
                taf_chunk_origin_kernel<<<(int)((n_chunks + 255) / 256), 256, 0, st>>>(ev, pl, ev_first, ev_last, (int)n_chunks, (int)per_cta, origins);
                EVREP_LAUNCH_CHECK();
                taf_bucket_kernel<kModeCount><<<grid, kBucketThreads, smem_count, st>>>(ev, pl, ev_first, ev_last, (int)n_chunks, lut_w, lut_h, origins, vec_ok, 0);
                EVREP_LAUNCH_CHECK();
            }
        } else {
            EVREP_CUDA(cudaFuncSetAttribute(taf_bucket_kernel<kModeCountSave>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_save));
            if (grid > 0) {
                const int64_t n_super = (n_chunks + kScatterChunks - 1) / kScatterChunks;
                taf_chunk_origin_kernel<<<(int)((n_super + 255) / 256), 256, 0, st>>>(ev, pl, ev_first, ev_last, (int)n_super,
                                                                                    (int)(per_cta * kScatterChunks), origins);
                EVREP_LAUNCH_CHECK();
                taf_bucket_kernel<kModeCountSave><<<grid, kBucketThreads, smem_save, st>>>(ev, pl, ev_first, ev_last, (int)n_chunks, lut_w, lut_h, origins, vec_ok, stage_ok);
                EVREP_LAUNCH_CHECK();
            }
        }
        taf_scan_rows_kernel<<<L.n_tiles, 256, 0, st>>>(pl);
        EVREP_LAUNCH_CHECK();
        taf_scan_tiles_kernel<<<1, 1024, 0, st>>>(pl);
        EVREP_LAUNCH_CHECK();
        taf_tile_bits_kernel<<<L.n_tiles, 256, 0, st>>>(pl);
        EVREP_LAUNCH_CHECK();
        if (grid > 0 && reclassify) {
            taf_bucket_kernel<kModeScatter><<<grid, kBucketThreads, smem_scatter, st>>>(ev, pl, ev_first, ev_last, (int)n_chunks, lut_w, lut_h, origins, vec_ok, 0);
            EVREP_LAUNCH_CHECK();
        } else if (grid > 0) {
            const int64_t n_super = (n_chunks + kScatterChunks - 1) / kScatterChunks;
            const size_t smem_saved = (size_t)3 * nh * 4 + (size_t)kSavedThreads * kScatterPerThread * 6;
            const int64_t resident = (int64_t)kSavedCtasPerSm * sm_count();
            const int grid_saved = (int)(n_super < resident ? n_super : resident);
            EVREP_CUDA(cudaFuncSetAttribute(taf_scatter_saved_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_saved));
            taf_scatter_saved_kernel<<<grid_saved, kSavedThreads, smem_saved, st>>>(ev, pl, ev_first, ev_last, (int)n_super, lut_w, lut_h, origins);
            EVREP_LAUNCH_CHECK();
        }
    } else {
        EVREP_CUDA(cudaMemsetAsync(s + L.o_tiletotal, 0, (size_t)(L.o_origins - L.o_tiletotal), st));   // totals, bases, tile bits
    }

    return EVREP_OK;
}

// ---- time-ordered input: one-pass bin-major sort (slices.cu) in front of the same tile kernels ----------------------
// Same tables as prepare_stream (counts -> off_rel / tile_total, bin_any, tile_bits, batches), but the records come from
// the slice sort's bin-major mode: every (bin, tile) run contiguous at src[tile][bin], padded to 16 bytes with null
// records; off_rel counts the padded records.  Windows of more than kBinMajorMaxParts slices are refused (their bins'
// slices might not all be resident at once): the caller uses prepare_stream for those.
struct BinMajorLayout {
    Layout L;
    int64_t o_status, o_wfresh, o_src, o_bins, o_slicebin, o_cnt16, o_bindone, total;
    int64_t max_slices;
    int pitch16;
};

static int make_binmajor_layout(int64_t n_events, int n_windows, int64_t TB, int H, int W, int n_batches, BinMajorLayout& B) {
    Layout& L = B.L;
    int rc = make_layout(0, n_windows, TB, H, W, n_batches, L);          // tile geometry and the table offsets; records re-sized below
    if (rc) return rc;
    if (n_events >= (1ll << 31) || n_events + (4ll * L.n_tiles + 3) * TB >= (1ll << 32)) return EVREP_ERR_RANGE;
    B.max_slices = n_events / kSliceMax + TB + 1;
    B.pitch16 = (L.n_tiles + 7) / 8 * 8;
    int64_t o = L.o_origins;                                              // the two-pass scratch (origins, saved records) is not needed
    L.o_records = o;  o += align_up(4ll * (n_events + (4ll * L.n_tiles + 3) * TB + 64), 256);   // per bin: events rounded up to 4 + 4 per tile
    B.o_status = o;   o += 256;
    B.o_wfresh = o;   o += align_up(4ll * (n_windows + 1), 256);
    B.o_src = o;      o += align_up(4ll * L.n_tiles * (TB > 0 ? TB : 1), 256);
    B.o_bins = o;     o += align_up((int64_t)sizeof(BinDesc) * (TB > 0 ? TB : 1), 256);
    B.o_slicebin = o; o += align_up(4ll * B.max_slices, 256);
    B.o_cnt16 = o;    o += align_up(2ll * B.max_slices * B.pitch16, 256);
    B.o_bindone = o;  o += align_up(4ll * (TB > 0 ? TB : 1), 256);
    B.total = o;
    L.total = o;
    return EVREP_OK;
}

int64_t binmajor_status_offset(int64_t n_events, int n_windows, int64_t TB, int H, int W) {
    BinMajorLayout B;
    int rc = make_binmajor_layout(n_events, n_windows, TB, H, W, (int)batches_upper_bound(n_windows, TB), B);
    return rc ? rc : B.o_status;
}

int64_t binmajor_scratch_bytes(int64_t n_events, int n_windows, int64_t TB, int H, int W) {
    BinMajorLayout B;
    int rc = make_binmajor_layout(n_events, n_windows, TB, H, W, (int)batches_upper_bound(n_windows, TB), B);
    return rc ? rc : B.total;
}

int prepare_stream_binmajor(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                            const evrep_taf_window* windows_host, int n_windows, int abin, int H, int W,
                            const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                            void* scratch, int64_t scratch_bytes, cudaStream_t st, StreamPlan& pl, Layout& L_out,
                            const uint32_t*& src_out, uint32_t*& status_out) {
    if (xmap && ymap && (sensor_h <= 0 || sensor_w <= 0 || sensor_h > EVREP_COORD_LUT_LEN || sensor_w > EVREP_COORD_LUT_LEN))
        return EVREP_ERR_ARG;
    if (abin <= 0 || (uint32_t)abin > kDMax) return EVREP_ERR_RANGE;
    if (!windows_host || (n_events > 0 && (!t || !x || !y || !p))) return EVREP_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(scratch) & 255) return EVREP_ERR_ARG;
    int64_t TB = 0, prev_end = 0;
    for (int w = 0; w < n_windows; ++w) {
        const evrep_taf_window& win = windows_host[w];
        if (win.ev_begin < prev_end || win.ev_end < win.ev_begin || win.ev_end > n_events || win.n_bins < 0) return EVREP_ERR_ARG;
        if ((int64_t)win.n_bins * abin >= (1ll << 32)) return EVREP_ERR_RANGE;
        if (win.ev_end - win.ev_begin > (int64_t)kBinMajorMaxParts * kSliceMax) return EVREP_ERR_RANGE;   // a bin could exceed the resident CTAs
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
    BinMajorLayout B;      // laid out for the upper bound of the batch count, like the scratch size and the status offset
    int rc = make_binmajor_layout(n_events, n_windows, TB, H, W, (int)batches_upper_bound(n_windows, TB), B);
    if (rc) return rc;
    if (scratch_bytes < B.total) return EVREP_ERR_SCRATCH;
    const Layout& L = B.L;
    L_out = L;
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
    rc = upload_words(reinterpret_cast<const uint32_t*>(meta.data()), L.meta_bytes / 4, reinterpret_cast<uint32_t*>(s), st);
    if (rc) return rc;

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
    pl.saved_rec = nullptr; pl.saved_key = nullptr;
    pl.n_windows = n_windows; pl.n_batches = (int)batches.size(); pl.TB = (int)TB;
    pl.n_tiles = L.n_tiles; pl.P = L.P; pl.H = H; pl.W = W;
    pl.div_abin = FastDiv::make((uint32_t)abin);
    pl.div_P = FastDiv::make((uint32_t)L.P);
    pl.tile_mul = 0; pl.abin = (uint32_t)abin;
    src_out = reinterpret_cast<const uint32_t*>(s + B.o_src);
    status_out = reinterpret_cast<uint32_t*>(s + B.o_status);

    if (TB == 0) {
        EVREP_CUDA(cudaMemsetAsync(s + L.o_tiletotal, 0, (size_t)(L.o_origins - L.o_tiletotal), st));   // totals, bases, tile bits
        EVREP_CUDA(cudaMemsetAsync(s + B.o_status, 0, 16, st));
        return EVREP_OK;
    }
    EVREP_CUDA(cudaMemsetAsync(s + L.o_counts, 0, (size_t)(L.o_offrel - L.o_counts), st));   // counts + bin_any
    EVREP_CUDA(cudaMemsetAsync(s + B.o_bindone, 0, (size_t)(4 * TB), st));
    SlicePlan sp;
    sp.status = status_out;
    sp.w_begin = pl.w_begin; sp.w_end = pl.w_end; sp.w_start = pl.w_start; sp.w_nbins = pl.w_nbins; sp.w_binbase = pl.w_binbase;
    sp.w_fresh = nullptr;
    sp.bins = reinterpret_cast<BinDesc*>(s + B.o_bins);
    sp.slice_bin = reinterpret_cast<uint32_t*>(s + B.o_slicebin);
    sp.runs = nullptr; sp.records = nullptr;
    sp.n_windows = n_windows; sp.TB = (int)TB; sp.n_tiles = L.n_tiles; sp.P = L.P; sp.H = H; sp.W = W;
    sp.pitch = 0; sp.slice_stride = 0; sp.max_slices = (int)B.max_slices;
    sp.abin = (uint32_t)abin;
    sp.div_P = pl.div_P;
    {
        const uint64_t mul = (1ull << 32) / (uint64_t)L.P + 1;
        const uint64_t err = mul * (uint64_t)L.P - (1ull << 32);
        sp.tile_mul = (mul < (1ull << 32) && (uint64_t)H * W * err < (1ull << 32)) ? (uint32_t)mul : 0u;
    }
    BinMajorOut bm;
    bm.cnt16 = reinterpret_cast<uint16_t*>(s + B.o_cnt16);
    bm.bin_done = reinterpret_cast<uint32_t*>(s + B.o_bindone);
    bm.counts = pl.counts; bm.src = reinterpret_cast<uint32_t*>(s + B.o_src); bm.bin_any = pl.bin_any; bm.records = pl.records;
    bm.pitch16 = B.pitch16; bm.TB = (int)TB;
    SoA ev{t, x, y, p, xmap, ymap};
    rc = run_slice_front(ev, sp, n_events, sensor_h, sensor_w, &bm, st);
    if (rc) return rc;
    taf_scan_rows_kernel<<<L.n_tiles, 256, 0, st>>>(pl);
    EVREP_LAUNCH_CHECK();
    taf_scan_tiles_kernel<<<1, 1024, 0, st>>>(pl);
    EVREP_LAUNCH_CHECK();
    taf_tile_bits_kernel<<<L.n_tiles, 256, 0, st>>>(pl);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

}  // namespace evrep
