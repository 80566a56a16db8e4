// Slice sort (see slices.cuh): bin ranges by bisection over the time-ordered stream, slice layout,
// and the one-pass sort of every slice by sensor tile.  Host front end: prepare_slices.
#include "slices.cuh"

namespace evrep {

// ---- bins: event range of every bin of every window ---------------------------------------------
// First index i in [lo, hi) with t[i] >= T (t is non-decreasing over the range).
__device__ __forceinline__ uint32_t lower_bound_time(const uint32_t* __restrict__ t, uint32_t lo, uint32_t hi, int64_t T) {
    if (T <= 0) return lo;
    if (T > 0xFFFFFFFFll) return hi;
    const uint32_t key = (uint32_t)T;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(t + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(128)
slice_bins_kernel(const uint32_t* __restrict__ t, SlicePlan sp) {
    const int gbin = blockIdx.x * blockDim.x + threadIdx.x;
    if (gbin == 0 && threadIdx.x == 0) { sp.status[0] = 0u; sp.status[1] = 0u; sp.status[2] = 0u; sp.status[3] = 0u; }
    if (gbin >= sp.TB) return;
    // window of the bin: the last w with binbase[w] <= gbin (windows without bins share their base with the next one)
    int lo = 0, hi = sp.n_windows;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(sp.w_binbase + mid) <= gbin) lo = mid; else hi = mid;
    }
    while (__ldg(sp.w_nbins + lo) == 0) --lo;            // never below 0: gbin < TB lies in a window that has bins
    const int w = lo;
    const int z = gbin - sp.w_binbase[w], nb = sp.w_nbins[w];
    const uint32_t begin = (uint32_t)sp.w_begin[w], end = (uint32_t)sp.w_end[w];
    const int64_t t0 = sp.w_start[w] + (int64_t)z * sp.abin;
    // inclusive edges, the later bin wins (generate_taf.py:201-202); events before the first edge
    // belong to bin 0, events past the last one to the last bin (the clamp of evrep_taf_stream)
    BinDesc bd;
    bd.lo = z == 0 ? begin : lower_bound_time(t, begin, end, t0);
    bd.hi = z == nb - 1 ? end : lower_bound_time(t, begin, end, t0 + sp.abin);
    if (bd.hi < bd.lo) bd.hi = bd.lo;
    bd.t0 = t0;
    bd.first_slice = 0;
    bd.flags = (z == 0 ? kBinFirst : 0u) | (z == nb - 1 ? kBinLast : 0u);
    bd.dyn = 0;
    bd.win = (uint32_t)w;
    sp.bins[gbin] = bd;
}

// One CTA: exclusive scan of the bins' slice counts (and of their record capacities: events + 4 per tile of padding,
// the bin's region in the bin-major layout), then the slices' bin indices.
__global__ void __launch_bounds__(1024)
slice_layout_kernel(SlicePlan sp) {
    __shared__ uint32_t warp_sum[32], warp_cap[32];
    __shared__ uint32_t s_carry, s_carry_cap;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) { s_carry = 0; s_carry_cap = 0; }
    __syncthreads();
    for (int base = 0; base < sp.TB; base += 1024) {
        const int b = base + threadIdx.x;
        uint32_t parts = 0, cap = 0;
        if (b < sp.TB) {
            parts = slice_parts(sp.bins[b].lo, sp.bins[b].hi);
            cap = parts ? ((sp.bins[b].hi - sp.bins[b].lo + 3u) & ~3u) + 4u * (uint32_t)sp.n_tiles : 0u;   // a multiple of 4 records
        }
        uint32_t incl = parts, incl_cap = cap;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o), c = __shfl_up_sync(0xFFFFFFFFu, incl_cap, o);
            if (lane >= o) { incl += v; incl_cap += c; }
        }
        if (lane == 31) { warp_sum[wid] = incl; warp_cap[wid] = incl_cap; }
        __syncthreads();
        uint32_t before = s_carry, before_cap = s_carry_cap;
        for (int k = 0; k < wid; ++k) { before += warp_sum[k]; before_cap += warp_cap[k]; }
        const uint32_t first = before + incl - parts;
        if (b < sp.TB) {
            sp.bins[b].first_slice = first;
            sp.bins[b].win = before_cap + incl_cap - cap;      // first record of the bin's region (bin-major layout)
            for (uint32_t j = 0; j < parts; ++j)
                if (first + j < (uint32_t)sp.max_slices) sp.slice_bin[first + j] = (uint32_t)b;
        }
        __syncthreads();
        if (threadIdx.x == 1023) { s_carry = before + incl; s_carry_cap = before_cap + incl_cap; }
        __syncthreads();
    }
    if (threadIdx.x == 0) sp.status[1] = s_carry < (uint32_t)sp.max_slices ? s_carry : (uint32_t)sp.max_slices;
}

// ---- the sort ----------------------------------------------------------------------------------
struct SortSmem {
    int lutx, luty, hist, off, sorted, total;
    __host__ __device__ SortSmem(int lut_w, int lut_h, int n_tiles) {
        int o = 0;
        lutx = o;   o += (lut_w * 4 + 15) / 16 * 16;     // column of a raw x (u32), kOffGrid when dropped
        luty = o;   o += (lut_h * 4 + 15) / 16 * 16;     // first pixel of the row of a raw y, kOffGrid when dropped
        hist = o;   o += (n_tiles + 1) * 4; o = (o + 15) / 16 * 16;
        off = o;    o += (n_tiles + 1) * 4; o = (o + 127) / 128 * 128;
        sorted = o; o += (kSliceCap + 4 * n_tiles) * 4;
        total = o;
    }
};

// Persistent CTAs, one slice at a time: load (vectorised), classify (coordinate tables in shared
// memory, tile by one multiply-high), rank inside the tile's run with a returning shared-memory
// atomic, scan the tile histogram (runs padded to 4 records), place the records, TMA bulk store.
// kBinMajor: instead of one bulk store per slice, the CTAs that sort the slices of one bin exchange their tile counts
// (a few KB through L2; a spin on the bin's counter -- every slice of a bin sits in a different resident CTA) and every
// CTA writes its part of each (bin, tile) run where it belongs: all records of a tile in a bin end up contiguous, the
// layout the register-resident TAF tile kernel (taf_tile.cu) streams with one copy per chunk.  Records then carry
// [ d:18 | 2 * local pixel + p ] like the two-pass bucketing's.
template <bool kBinMajor>
__global__ void __launch_bounds__(kSortThreads, 2)
slice_sort_kernel(SoA ev, SlicePlan sp, int64_t n_events, int lut_w, int lut_h, int vec_ok, BinMajorOut bm) {
    extern __shared__ __align__(128) unsigned char ssm[];
    const SortSmem lay(lut_w, lut_h, sp.n_tiles);
    uint32_t* s_col = reinterpret_cast<uint32_t*>(ssm + lay.lutx);
    uint32_t* s_row = reinterpret_cast<uint32_t*>(ssm + lay.luty);
    uint32_t* hist = reinterpret_cast<uint32_t*>(ssm + lay.hist);
    uint32_t* off = reinterpret_cast<uint32_t*>(ssm + lay.off);
    uint32_t* sorted = reinterpret_cast<uint32_t*>(ssm + lay.sorted);
    __shared__ uint32_t s_tmp[kSortThreads / 32 + 1];
    __shared__ uint32_t s_flags;                          // bit 0: some record with d > kPackedDMax

    const uint32_t W = sp.W, H = sp.H, HW = W * H;
    constexpr uint32_t kOffGrid = 0x40000000u;
    for (int i = threadIdx.x; i < lut_w; i += kSortThreads) {
        const uint32_t xm = ev.xmap ? ev.xmap[i] : (uint32_t)i;
        s_col[i] = xm < W ? xm : kOffGrid;
    }
    for (int i = threadIdx.x; i < lut_h; i += kSortThreads) {
        const uint32_t ym = ev.ymap ? ev.ymap[i] : (uint32_t)i;
        s_row[i] = ym < H ? ym * W : kOffGrid;
    }
    uint32_t col_addr = smem_u32(s_col), row_addr = smem_u32(s_row), hist_addr = smem_u32(hist);
    uint32_t n_cols = (uint32_t)lut_w, n_rows = (uint32_t)lut_h;
    uint32_t tile_mul = sp.tile_mul, P = (uint32_t)sp.P, k_hw = HW;
    // per-event constants pinned in registers (see taf_bucket_kernel): the compiler otherwise re-derives the shared
    // window base and re-loads the kernel parameters for every event
    asm volatile("" : "+r"(col_addr), "+r"(row_addr), "+r"(hist_addr), "+r"(n_cols), "+r"(n_rows));
    asm volatile("" : "+r"(tile_mul), "+r"(P), "+r"(k_hw));
    const int n_tiles = sp.n_tiles;
    const uint32_t n_slices = sp.status[1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    bool store_pending = false;

    // the header of the next slice (its bin, then the bin's descriptor) is fetched while the current slice is sorted,
    // and its events are pulled into L2 ahead of time
    uint32_t gbin = 0, gbin_next = 0;
    BinDesc bd, bd_next;
    if (blockIdx.x < n_slices) { gbin = sp.slice_bin[blockIdx.x]; bd = sp.bins[gbin]; }
    for (uint32_t s = blockIdx.x; s < n_slices; s += gridDim.x) {
        const uint32_t s_next = s + gridDim.x;
        if (s_next < n_slices) gbin_next = __ldg(sp.slice_bin + s_next);
        const uint32_t part = s - bd.first_slice;
        const uint32_t lo = bd.lo + part * (uint32_t)kSliceMax;
        const uint32_t hi = min(bd.hi, lo + (uint32_t)kSliceMax);
        const uint32_t base = lo & ~3u;

        // the previous slice's bulk store must have read `sorted` before it is overwritten
        if (store_pending && threadIdx.x == 0) bulk_wait_read();
        __syncthreads();
        for (int i = threadIdx.x; i <= n_tiles; i += kSortThreads) hist[i] = 0;
        if (threadIdx.x == 0) s_flags = 0;
        if (!kBinMajor) {   // padding: every slot a run does not fill holds the null record
            uint4* s4 = reinterpret_cast<uint4*>(sorted);
            const int n4 = (kSliceCap + 4 * n_tiles) / 4;
            for (int i = threadIdx.x; i < n4; i += kSortThreads) s4[i] = make_uint4(kNullRecord, kNullRecord, kNullRecord, kNullRecord);
        }

        // all global loads of the slice are issued before any of them is used
        uint32_t tt[kSortPerThread], xy[kSortPerThread], pw[kSortPerThread / 4];
#pragma unroll
        for (int g = 0; g < kSortPerThread / 4; ++g) {
            const uint32_t i4 = base + (uint32_t)(g * kSortThreads + threadIdx.x) * 4u;
            pw[g] = 0xFFFFFFFFu;
#pragma unroll
            for (int e = 0; e < 4; ++e) { tt[4 * g + e] = 0; xy[4 * g + e] = 0; }
            if (i4 >= hi) continue;
            if (vec_ok && (int64_t)i4 + 4 <= n_events) {
                const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(ev.t + i4));
                const uint2 x4 = __ldg(reinterpret_cast<const uint2*>(ev.x + i4));
                const uint2 y4 = __ldg(reinterpret_cast<const uint2*>(ev.y + i4));
                pw[g] = __ldg(reinterpret_cast<const uint32_t*>(ev.p + i4));
                tt[4 * g + 0] = t4.x; tt[4 * g + 1] = t4.y; tt[4 * g + 2] = t4.z; tt[4 * g + 3] = t4.w;
                xy[4 * g + 0] = __byte_perm(x4.x, y4.x, 0x5410); xy[4 * g + 1] = __byte_perm(x4.x, y4.x, 0x7632);
                xy[4 * g + 2] = __byte_perm(x4.y, y4.y, 0x5410); xy[4 * g + 3] = __byte_perm(x4.y, y4.y, 0x7632);
            } else {
                uint32_t pol4 = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int64_t i = (int64_t)i4 + e;
                    uint32_t pv = 0xFFu;
                    if (i < n_events) {
                        tt[4 * g + e] = __ldg(ev.t + i);
                        xy[4 * g + e] = __ldg(ev.x + i) | ((uint32_t)__ldg(ev.y + i) << 16);
                        pv = __ldg(ev.p + i);
                    }
                    pol4 |= pv << (8 * e);
                }
                pw[g] = pol4;
            }
        }
        __syncthreads();                                   // histogram zeroed, padding written
        if (s_next < n_slices) {
            bd_next = sp.bins[gbin_next];
            if (threadIdx.x == 0 && vec_ok) {
                const uint32_t nlo = (bd_next.lo + (s_next - bd_next.first_slice) * (uint32_t)kSliceMax) & ~15u;
                const uint32_t nhi = min(bd_next.hi, nlo + (uint32_t)kSliceCap);
                if (nhi > nlo && (int64_t)nlo + kSliceCap <= n_events) {
                    const uint32_t n16 = (nhi - nlo + 15u) & ~15u;
                    l2_prefetch(ev.t + nlo, n16 * 4u); l2_prefetch(ev.x + nlo, n16 * 2u);
                    l2_prefetch(ev.y + nlo, n16 * 2u); l2_prefetch(ev.p + nlo, n16);
                }
            }
        }

        // classify and rank: slot = (tile << 14) | rank inside the tile's run, or kNone
        constexpr uint32_t kNone = 0xFFFFFFFFu;
        uint32_t slot[kSortPerThread], rec[kSortPerThread];
        const bool first_bin = (bd.flags & kBinFirst) != 0, last_bin = (bd.flags & kBinLast) != 0;
        const bool t0_32 = bd.t0 >= 0 && bd.t0 <= 0xFFFFFFFFll;
        const uint32_t t0u = (uint32_t)bd.t0, abin = sp.abin;
        uint32_t strays = 0, big = 0;
#pragma unroll
        for (int k = 0; k < kSortPerThread; ++k) {
            slot[k] = kNone; rec[k] = kNullRecord;
            const uint32_t i = base + (uint32_t)((k >> 2) * kSortThreads + threadIdx.x) * 4u + (uint32_t)(k & 3);
            const uint32_t xv = xy[k] & 0xFFFFu, yv = xy[k] >> 16;
            const uint32_t pol = (pw[k >> 2] >> ((k & 3) * 8)) & 0xFFu;
            if (i < lo || i >= hi || xv >= n_cols || yv >= n_rows || pol > 1u) continue;
            const uint32_t pix = lds_u32(col_addr + xv * 4u) + lds_u32(row_addr + yv * 4u);
            if (pix >= k_hw) continue;
            // d = t - bin start; an event whose time lies outside its bin's edges is either clamped (first / last
            // bin of a window) or evidence that the input is not ordered in time
            uint32_t d;
            bool before, after;
            if (t0_32) {
                before = tt[k] < t0u;
                d = before ? 0u : tt[k] - t0u;
                after = d >= abin;
            } else {
                const int64_t dd = (int64_t)tt[k] - bd.t0;
                before = dd < 0;
                d = before ? 0u : (dd > (int64_t)kDMax ? kDMax : (uint32_t)dd);
                after = dd >= (int64_t)abin;
            }
            strays += ((before && !first_bin) || (after && !last_bin)) ? 1u : 0u;
            d = min(d, kDMax);
            big |= d > kPackedDMax ? 1u : 0u;
            const uint32_t tile = tile_mul ? __umulhi(pix, tile_mul) : sp.div_P.div(pix);
            rec[k] = kBinMajor ? (d << 14) | ((pix - tile * P) << 1) | pol : (d << 14) | (pol * P + (pix - tile * P));
            slot[k] = (tile << 14) | satom_add(hist_addr + tile * 4u, 1u);
        }
        if (strays) atomicAdd(sp.status, strays);
        if (big) s_flags = 1u;
        __syncthreads();

        // exclusive scan of the run lengths, each rounded up to 4 records
        uint32_t n_valid;
        {
            const int per = (n_tiles + kSortThreads - 1) / kSortThreads;
            const int b0 = threadIdx.x * per, b1 = min(b0 + per, n_tiles);
            uint32_t mine = 0, raw = 0;
            for (int i = b0; i < b1; ++i) { const uint32_t c = hist[i]; raw += c; mine += kBinMajor ? c : (c + 3u) & ~3u; }
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            raw = __reduce_add_sync(0xFFFFFFFFu, raw);
            if (lane == 31) s_tmp[wid] = incl;
            if (lane == 0 && raw) atomicAdd(&hist[n_tiles], raw);      // hist[n_tiles] = events kept
            __syncthreads();
            if (wid == 0) {
                uint32_t w = lane < kSortThreads / 32 ? s_tmp[lane] : 0u, wi = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, o);
                    if (lane >= o) wi += v;
                }
                if (lane < kSortThreads / 32) s_tmp[lane] = wi - w;
                if (lane == kSortThreads / 32 - 1) s_tmp[kSortThreads / 32] = wi;
            }
            __syncthreads();
            uint32_t run = s_tmp[wid] + incl - mine;
            for (int i = b0; i < b1; ++i) { off[i] = run; run += kBinMajor ? hist[i] : (hist[i] + 3u) & ~3u; }
            if (threadIdx.x == kSortThreads - 1 || b1 == n_tiles) off[n_tiles] = s_tmp[kSortThreads / 32];
            n_valid = hist[n_tiles];
            __syncthreads();
        }
        const uint32_t total = off[n_tiles];               // (padded) records of the slice
        if (!kBinMajor) {
            // the run table is tile-major: a tile kernel streams its row; neighbouring slices fill the same sectors in L2
            for (int i = threadIdx.x; i < n_tiles; i += kSortThreads)
                sp.runs[(int64_t)i * sp.pitch + s] = (off[i] >> 2) | ((off[i + 1] >> 2) << 16);
#pragma unroll
            for (int k = 0; k < kSortPerThread; ++k)
                if (slot[k] != kNone) sorted[off[slot[k] >> 14] + (slot[k] & 0x3FFFu)] = rec[k];
            fence_async_smem();
            __syncthreads();
            if (threadIdx.x == 0) {
                if (total) {
                    bulk_store_1d(sp.records + (int64_t)s * sp.slice_stride, sorted, total * 4u);
                    bulk_commit();
                }
                const uint32_t dyn = (n_valid ? kBinAny : 0u) | (s_flags ? kBinBigD : 0u);
                if (dyn) atomicOr(&sp.bins[gbin].dyn, dyn);
            }
            store_pending = true;
        } else {
            // (1) publish this slice's tile counts, then wait until every slice of the bin has
            const uint32_t parts = slice_parts(bd.lo, bd.hi);
            uint16_t* my_cnt = bm.cnt16 + (int64_t)s * bm.pitch16;
            for (int i = threadIdx.x; i < n_tiles; i += kSortThreads) my_cnt[i] = (uint16_t)hist[i];
#pragma unroll
            for (int k = 0; k < kSortPerThread; ++k)
                if (slot[k] != kNone) sorted[off[slot[k] >> 14] + (slot[k] & 0x3FFFu)] = rec[k];
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) {
                atomicAdd(bm.bin_done + gbin, 1u);
                // bounded: a CTA that is never scheduled next to its siblings must not hang the GPU
                uint32_t polls = 0;
                while (*reinterpret_cast<volatile uint32_t*>(bm.bin_done + gbin) < parts && ++polls < (1u << 24)) __nanosleep(64);
                if (polls >= (1u << 24)) atomicAdd(sp.status + 2, 1u);
                __threadfence();
            }
            __syncthreads();
            // (2) per tile: records of the whole bin (every slice) and of the slices before this one
            uint32_t* x_prev = sorted + kSliceCap;         // the pad area of the slice layout (4 * n_tiles words) is free here
            uint32_t* x_tot = x_prev + n_tiles;
            uint32_t* x_dest = x_tot + n_tiles;
            for (int i = threadIdx.x; i < n_tiles; i += kSortThreads) {
                const uint16_t* col = bm.cnt16 + (int64_t)bd.first_slice * bm.pitch16 + i;
                uint32_t all = 0, prev = 0;
                for (uint32_t j = 0; j < parts; ++j) {
                    const uint32_t c = __ldcg(col + (int64_t)j * bm.pitch16);
                    all += c;
                    prev += j < part ? c : 0u;
                }
                x_prev[i] = prev;
                x_tot[i] = all;
            }
            __syncthreads();
            // (3) exclusive scan over the tiles of the bin's run lengths, each rounded up to 4 records
            {
                const int per = (n_tiles + kSortThreads - 1) / kSortThreads;
                const int b0 = threadIdx.x * per, b1 = min(b0 + per, n_tiles);
                uint32_t mine = 0;
                for (int i = b0; i < b1; ++i) mine += (x_tot[i] + 3u) & ~3u;
                uint32_t incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (lane >= o) incl += v;
                }
                if (lane == 31) s_tmp[wid] = incl;
                __syncthreads();
                if (wid == 0) {
                    uint32_t w = lane < kSortThreads / 32 ? s_tmp[lane] : 0u, wi = w;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, o);
                        if (lane >= o) wi += v;
                    }
                    if (lane < kSortThreads / 32) s_tmp[lane] = wi - w;
                    if (lane == kSortThreads / 32 - 1) s_tmp[kSortThreads / 32] = wi;
                }
                __syncthreads();
                uint32_t run = bd.win + s_tmp[wid] + incl - mine;       // bd.win = first record of the bin's region
                for (int i = b0; i < b1; ++i) {
                    const uint32_t padded = (x_tot[i] + 3u) & ~3u;
                    if (part == 0u) {                       // the bin's first slice writes the bin's tables and the padding
                        bm.counts[(int64_t)i * bm.TB + gbin] = padded;
                        bm.src[(int64_t)i * bm.TB + gbin] = run;
                        for (uint32_t q = x_tot[i]; q < padded; ++q) bm.records[run + q] = kNullRecord;
                    }
                    x_dest[i] = run + x_prev[i];
                    run += padded;
                }
                if (threadIdx.x == 0 && part == 0u) bm.bin_any[gbin] = s_tmp[kSortThreads / 32] ? 1u : 0u;
                __syncthreads();
            }
            // (4) every warp copies whole runs: sorted[local offset ...] -> records[destination ...]
            for (int i = wid; i < n_tiles; i += kSortThreads / 32) {
                const uint32_t cnt = hist[i], from = off[i], to = x_dest[i];
                for (uint32_t q = lane; q < cnt; q += 32u) bm.records[to + q] = sorted[from + q];
            }
        }
        gbin = gbin_next; bd = bd_next;
    }
    if (store_pending && threadIdx.x == 0) bulk_wait_all();
}

// Number of adjacent pairs t[i] > t[i + 1] (0 = the stream is ordered in time).
__global__ void __launch_bounds__(256)
order_check_kernel(const uint32_t* __restrict__ t, int64_t n, uint32_t* __restrict__ violations) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i + 1 < n; i += stride)
        bad += __ldg(t + i) > __ldg(t + i + 1) ? 1u : 0u;
    bad = __reduce_add_sync(0xFFFFFFFFu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(violations, bad);
}

// ---- host side ---------------------------------------------------------------------------------
static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

int choose_tile(int H, int W, int K, TileSmemFn smem_of, int ctas_per_sm, int& P, int& n_tiles) {
    const int64_t HW = (int64_t)H * W;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    // largest tile (multiple of 8 pixels) whose shared memory lets `ctas_per_sm` CTAs share an SM
    const size_t per_cta = (size_t)(233472 / ctas_per_sm) - 1024;
    int p_max = 8;
    while (p_max + 8 <= kMaxSlicePixels && smem_of(p_max + 8, K) <= per_cta) p_max += 8;
    if (smem_of(p_max, K) > per_cta) return EVREP_ERR_RANGE;
    const int64_t resident = (int64_t)sm_count() * ctas_per_sm;
    const int64_t waves = (HW + (int64_t)p_max * resident - 1) / ((int64_t)p_max * resident);
    int64_t p = (HW + waves * resident - 1) / (waves * resident);
    p = (p + 7) / 8 * 8;
    if (p > p_max) p = p_max;
    if (p < 8) p = 8;
    P = (int)p;
    n_tiles = (int)((HW + p - 1) / p);
    return n_tiles <= kMaxSliceTiles ? EVREP_OK : EVREP_ERR_RANGE;
}

int make_slice_layout(int64_t n_events, int n_windows, int64_t TB, int H, int W, int P, int n_tiles, SliceLayout& L) {
    if (n_events >= (1ll << 31) || TB >= (1ll << 24)) return EVREP_ERR_RANGE;
    L.P = P; L.n_tiles = n_tiles;
    L.slice_stride = kSliceCap + 4 * n_tiles;
    L.max_slices = n_events / kSliceMax + TB + 1;
    if (L.max_slices >= (1ll << 31) - 64) return EVREP_ERR_RANGE;
    L.pitch = (int)((L.max_slices + 63) / 32 * 32);         // the feed reads 32-slice windows, one ahead
    int64_t o = 0;
    L.o_status = o;   o += 16;
    L.o_wbegin = o;   o += align_up(8ll * n_windows, 16);
    L.o_wend = o;     o += align_up(8ll * n_windows, 16);
    L.o_wstart = o;   o += align_up(8ll * n_windows, 16);
    L.o_wnbins = o;   o += align_up(4ll * n_windows, 16);
    L.o_wbinbase = o; o += align_up(4ll * (n_windows + 1), 16);
    L.o_wfresh = o;   o += align_up(4ll * n_windows, 16);
    L.meta_bytes = o;
    o = align_up(o, 256);
    L.o_bins = o;     o += align_up((int64_t)sizeof(BinDesc) * TB, 256);
    L.o_slicebin = o; o += align_up(4ll * L.max_slices, 256);
    L.o_runs = o;     o += align_up(4ll * n_tiles * L.pitch, 256);
    L.o_records = o;  o += align_up(4ll * L.max_slices * L.slice_stride, 256);
    L.total = o;
    return EVREP_OK;
}

int prepare_slices(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                   const evrep_taf_window* windows_host, int n_windows, int abin, int H, int W, int P, int n_tiles,
                   const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                   void* scratch, int64_t scratch_bytes, cudaStream_t st, SlicePlan& sp, SliceLayout& L) {
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
        prev_end = win.ev_end;
        TB += win.n_bins;
    }
    int rc = make_slice_layout(n_events, n_windows, TB, H, W, P, n_tiles, L);
    if (rc) return rc;
    if (scratch_bytes < L.total) return EVREP_ERR_SCRATCH;

    std::vector<unsigned char> meta((size_t)L.meta_bytes, 0);
    int64_t* hb = reinterpret_cast<int64_t*>(meta.data() + L.o_wbegin);
    int64_t* he = reinterpret_cast<int64_t*>(meta.data() + L.o_wend);
    int64_t* hs = reinterpret_cast<int64_t*>(meta.data() + L.o_wstart);
    int32_t* hn = reinterpret_cast<int32_t*>(meta.data() + L.o_wnbins);
    int32_t* hbb = reinterpret_cast<int32_t*>(meta.data() + L.o_wbinbase);
    int32_t* hf = reinterpret_cast<int32_t*>(meta.data() + L.o_wfresh);
    int32_t base = 0;
    for (int w = 0; w < n_windows; ++w) {
        hb[w] = windows_host[w].ev_begin; he[w] = windows_host[w].ev_end; hs[w] = windows_host[w].start_time;
        hn[w] = windows_host[w].n_bins; hbb[w] = base; hf[w] = windows_host[w].fresh ? 1 : 0;
        base += windows_host[w].n_bins;
    }
    hbb[n_windows] = base;
    char* s = reinterpret_cast<char*>(scratch);
    // the tables travel as kernel arguments (see upload_words in bucketing.cu)
    rc = upload_words(reinterpret_cast<const uint32_t*>(meta.data()), L.meta_bytes / 4, reinterpret_cast<uint32_t*>(s), st);
    if (rc) return rc;

    sp.status = reinterpret_cast<uint32_t*>(s + L.o_status);
    sp.w_begin = reinterpret_cast<const int64_t*>(s + L.o_wbegin);
    sp.w_end = reinterpret_cast<const int64_t*>(s + L.o_wend);
    sp.w_start = reinterpret_cast<const int64_t*>(s + L.o_wstart);
    sp.w_nbins = reinterpret_cast<const int32_t*>(s + L.o_wnbins);
    sp.w_binbase = reinterpret_cast<const int32_t*>(s + L.o_wbinbase);
    sp.w_fresh = reinterpret_cast<const int32_t*>(s + L.o_wfresh);
    sp.bins = reinterpret_cast<BinDesc*>(s + L.o_bins);
    sp.slice_bin = reinterpret_cast<uint32_t*>(s + L.o_slicebin);
    sp.runs = reinterpret_cast<uint32_t*>(s + L.o_runs);
    sp.records = reinterpret_cast<uint32_t*>(s + L.o_records);
    sp.n_windows = n_windows; sp.TB = (int)TB; sp.n_tiles = n_tiles; sp.P = P; sp.H = H; sp.W = W;
    sp.pitch = L.pitch; sp.slice_stride = L.slice_stride; sp.max_slices = (int)L.max_slices;
    sp.abin = (uint32_t)abin;
    sp.div_P = FastDiv::make((uint32_t)P);
    {   // pix / P as one multiply-high: exact while pix * (mul * P - 2^32) < 2^32
        const uint64_t mul = (1ull << 32) / (uint64_t)P + 1;
        const uint64_t err = mul * (uint64_t)P - (1ull << 32);
        sp.tile_mul = (mul < (1ull << 32) && (uint64_t)H * W * err < (1ull << 32)) ? (uint32_t)mul : 0u;
    }

    SoA ev{t, x, y, p, xmap, ymap};
    return run_slice_front(ev, sp, n_events, sensor_h, sensor_w, nullptr, st);
}

int run_slice_front(const SoA& ev, const SlicePlan& sp, int64_t n_events, int sensor_h, int sensor_w, const BinMajorOut* bm,
                    cudaStream_t st) {
    // status words are reset by the bins kernel; with no bins at all there is nothing to sort
    if (sp.TB == 0) {
        EVREP_CUDA(cudaMemsetAsync(sp.status, 0, 16, st));
        return EVREP_OK;
    }
    slice_bins_kernel<<<(sp.TB + 127) / 128, 128, 0, st>>>(ev.t, sp);
    EVREP_LAUNCH_CHECK();
    slice_layout_kernel<<<1, 1024, 0, st>>>(sp);
    EVREP_LAUNCH_CHECK();
    const bool use_lut = ev.xmap && ev.ymap;
    const int lut_w = use_lut ? sensor_w : sp.W, lut_h = use_lut ? sensor_h : sp.H;
    const size_t smem = (size_t)SortSmem(lut_w, lut_h, sp.n_tiles).total;
    if (smem > 110 * 1024) return EVREP_ERR_RANGE;                        // two CTAs per SM
    const int vec_ok = !((reinterpret_cast<uintptr_t>(ev.t) & 15) | (reinterpret_cast<uintptr_t>(ev.x) & 7) |
                         (reinterpret_cast<uintptr_t>(ev.y) & 7) | (reinterpret_cast<uintptr_t>(ev.p) & 3));
    const int64_t resident = 2ll * sm_count();
    const int grid = (int)(sp.max_slices < resident ? sp.max_slices : resident);
    if (bm) {
        EVREP_CUDA(cudaFuncSetAttribute(slice_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        slice_sort_kernel<true><<<grid, kSortThreads, smem, st>>>(ev, sp, n_events, lut_w, lut_h, vec_ok, *bm);
    } else {
        EVREP_CUDA(cudaFuncSetAttribute(slice_sort_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        slice_sort_kernel<false><<<grid, kSortThreads, smem, st>>>(ev, sp, n_events, lut_w, lut_h, vec_ok, BinMajorOut{});
    }
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

}  // namespace evrep

extern "C" int evrep_events_order_check(const uint32_t* t, int64_t n_events, uint32_t* violations_dev, evrep_stream_t stream) {
    using namespace evrep;
    if (n_events < 0 || !violations_dev || (n_events > 0 && !t)) return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    EVREP_CUDA(cudaMemsetAsync(violations_dev, 0, 4, st));
    if (n_events < 2) return EVREP_OK;
    order_check_kernel<<<grid_for(n_events, 4), 256, 0, st>>>(t, n_events, violations_dev);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}
