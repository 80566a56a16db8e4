// Temporal Active Focus over whole time-ordered streams (generate_taf.py:160-238): the tile
// kernel that consumes the slice sort (slices.cu).
//
// One CTA per sensor tile (two or three CTAs per SM).  The tile's FIFO state lives in SHARED
// memory for the whole launch as a circular buffer per cell -- K values `u` and a head index --
// with LAZY AGEING: the reference subtracts 1 from every value of every pixel whenever a bin saw
// an event anywhere (generate_taf.py:44-46); here a value is stored as u = v + A, A = the number
// of such bins since the last rebase, so a bin costs work only for the cells that received
// events: one packed shared-memory atomic per record ((1 << 23) + d into a count:9 | sum:23 word;
// the first toucher of a cell appends it to the bin's active list), then one push per active
// cell (mean, one store into the circular buffer, head + 1).  v = u - A is materialised when a
// window tensor is emitted (and the buffer rebased -- also every 8 ageing steps inside a long
// window -- so |u| stays below 8 and the rounding of u = mean + A stays below 2.4e-7).  Records arrive through a ring of TMA bulk copies driven by a
// producer warp (slices.cuh); the eight worker warps synchronise once per bin.
//
// Cells that collect 256 or more events in one bin, or bins with offsets beyond 32767 us, cannot
// use the packed word: they are detected (returning atomics / a flag from the sort) and the bin is
// accumulated again with two-word accumulators -- exact, slower, rare.
//
// HBM-bound byte/float work: no tensor cores.  Sums of d are exact integers, so the result does
// not depend on the order in which records are accumulated: run-to-run bit-identical.
#include "slices.cuh"

namespace evrep {

constexpr int kTsWorkers = 256;                    // worker threads (8 warps)
constexpr int kTsThreads = kTsWorkers + 32 * kFeedWarps;   // + the producer warps
constexpr int kTsWarps = kTsWorkers / 32;
constexpr int kBarWorkers = 1;                     // named barrier of the worker warps
constexpr uint32_t kCountShift = 23;               // packed accumulator: count in bits 23..31, sum of d below
constexpr uint32_t kSumMask = (1u << kCountShift) - 1u;
constexpr uint32_t kHotCount = 256;                // a cell that reaches this count in one bin forces the exact path
constexpr int kRebaseAlways = 8;                   // A never exceeds this: rebase inside a window when it gets there
constexpr int kHeldSegments = 2;                   // segments of a bin whose records a thread keeps in registers for the push
constexpr int kPerSeg = kFeedStageRecords / kTsWorkers;   // records per thread per segment
static_assert(kFeedStageRecords % kTsWorkers == 0, "a stage is a whole number of records per worker thread");

struct TafSliceParams {
    SlicePlan sp;
    float* state;          // [H,W,2,K]
    float* out;            // window w at out + w * out_stride (may be null when out_u8 is set)
    int64_t out_stride;
    uint8_t* out_u8;       // optional: u8 [n_windows][K,2,H,W], leaky transform + slot flip (generate_taf.py:226-235)
    int64_t out_u8_stride;
    int emit_state;        // write the state after every window (always after the last)
    int vec_out;           // out / out_u8 rows allow 16-byte / 4-byte vector stores (4 pixels per thread)
    float span;            // f32(abin + 1e-8)
};

struct TafSliceSmem {
    int state, acc, head, ctrl, feed_base, total;
    __host__ __device__ TafSliceSmem(int P, int K) {
        int o = 0;
        state = o; o += K * 2 * P * 4;                  // u[K][2P]
        acc = o;   o += 2 * 2 * P * 4;                  // two buffers of packed {count | sum d} per cell
        head = o;  o += 2 * P; o = (o + 15) / 16 * 16;  // next slot to overwrite, per cell
        ctrl = o;  o += 16;                             // hot counters [2] (monotonic)
        feed_base = o;
        total = FeedSmem(o).total;
    }
};

static size_t taf_slice_smem(int P, int K) { return (size_t)TafSliceSmem(P, K).total; }

// generate_taf.py:69-76 on one value, 255 * max(0, 1 - log1p(-v) / 8.7) truncated to uint8, with the logarithm
// from MUFU.LG2 (v <= 0, so 1 - v >= 1): within 1e-5 of the float32 reference before truncation
__device__ __forceinline__ uint32_t leaky_u8_fast(float v) {
    const float r = fmaf(__log2f(1.0f - v), -0.6931471805599453f / 8.7f, 1.0f);
    return (uint32_t)(int)(fmaxf(r, 0.0f) * 255.0f);
}

#ifdef EVREP_TAF_TIMING
// Diagnostic build only: cycles spent per tile CTA in [0] producer total, [1] producer waiting for a free stage,
// [2] worker total, [3] worker waiting for a full stage, [4] worker barriers at bin ends, [5] sweeps, [6] accumulate, [7] push.
__device__ unsigned long long g_taf_timing[kMaxSliceTiles][8];
#define TAF_T0(v) const long long v = clock64()
#define TAF_ADD(slot, v) do { t_acc[slot] += clock64() - (v); } while (0)
#else
#define TAF_T0(v) do {} while (0)
#define TAF_ADD(slot, v) do {} while (0)
#endif

template <int K>
__global__ void __launch_bounds__(kTsThreads, 2)
taf_slice_tile_kernel(const __grid_constant__ TafSliceParams tp) {
    const SlicePlan& sp = tp.sp;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = sp.P;
    const TafSliceSmem lay(P, K);
    const FeedSmem fs(lay.feed_base);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + fs.full);
    uint64_t* empty = reinterpret_cast<uint64_t*>(smem_raw + fs.empty);

    const int tid = threadIdx.x, tile = blockIdx.x, lane = tid & 31;
    const int64_t HW = (int64_t)sp.H * sp.W;
    const int64_t pix0 = (int64_t)tile * P;
    const int npix = (int)min((int64_t)P, HW - pix0);
    const int C = 2 * P;                                 // cells of the tile: cell = p * P + pixel

    if (tid == 0) {
        for (int s = 0; s < kFeedStages; ++s) { mbar_init(full + s, kFeedFullCount); mbar_init(empty + s, kTsWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= kTsWorkers) {
        // ================================== producer warps =================================
        FeedProducer fp;
        fp.init(smem_raw, fs);
        fp.open_row(sp, tile);
#ifdef EVREP_TAF_TIMING
        const long long p_start = clock64();
#endif
        for (int w = 0; w < sp.n_windows; ++w) {
            if (w > 0 && sp.w_fresh[w]) fp.control(kSegReset, 0u, (uint32_t)w);
            const int g0 = sp.w_binbase[w], g1 = sp.w_binbase[w + 1];
            uint32_t pend_age = 0;                       // bins that age the tile without bringing it records
            for (int gb = g0; gb < g1; gb += 32) {
                // descriptors of 32 bins at a time, one per lane
                uint32_t first = 0, parts = 0, dyn = 0;
                if (gb + lane < g1) {
                    const BinDesc* bd = sp.bins + gb + lane;
                    first = bd->first_slice; parts = slice_parts(bd->lo, bd->hi); dyn = bd->dyn;
                }
                const int nb = min(32, g1 - gb);
                for (int k = 0; k < nb; ++k) {
                    const uint32_t k_dyn = __shfl_sync(0xFFFFFFFFu, dyn, k);
                    if (!(k_dyn & kBinAny)) continue;    // nobody saw an event: no ageing (generate_taf.py:40-41)
                    const uint32_t k_first = __shfl_sync(0xFFFFFFFFu, first, k), k_parts = __shfl_sync(0xFFFFFFFFu, parts, k);
                    if (fp.feed_bin(sp, (uint32_t)(gb + k), k_first, k_parts, k_dyn, pend_age + 1u)) pend_age = 0;
                    else ++pend_age;
                }
            }
            fp.control(kSegEmit | (w == sp.n_windows - 1 ? kSegDone : 0u), pend_age, (uint32_t)w);
        }
#ifdef EVREP_TAF_TIMING
        if (tid == kTsWorkers) { g_taf_timing[tile][0] = clock64() - p_start; g_taf_timing[tile][1] = fp.t_acquire; }
#endif
        return;
    }

    // ==================================== worker warps ====================================
    auto worker_sync = [] { named_sync(kBarWorkers, kTsWorkers); };
    // everything below addresses shared memory by 32-bit shared addresses (LDS / STS / ATOMS)
    const uint32_t a_u = smem_u32(smem_raw + lay.state);          // f32 [K][2P]
    const uint32_t a_acc = smem_u32(smem_raw + lay.acc);          // u32 [2][2P]
    const uint32_t a_head = smem_u32(smem_raw + lay.head);        // u8 [2P]
    const uint32_t a_hot = smem_u32(smem_raw + lay.ctrl);         // u32 [2]
    const uint32_t a_ring = smem_u32(smem_raw + fs.ring);
    const uint32_t a_desc = smem_u32(smem_raw + fs.desc);
    const uint32_t row_bytes = (uint32_t)C * 4u;                  // one slot row of the state

    const bool first_fresh = sp.w_fresh[0] != 0;
    if (tid < 2) sst_u32(a_hot + tid * 4u, 0u);
    // state in: u = v (A = 0), heads at 0 -- slot e of a cell holds FIFO position e (K - 1 = newest)
    for (int c = tid; c < C; c += kTsWorkers) {
        const int p = c >= P ? 1 : 0, pix = c - p * P;
        sst_u8(a_head + c, 0u);
        sst_u32(a_acc + c * 4u, 0u); sst_u32(a_acc + (C + c) * 4u, 0u);
        if (pix < npix && !first_fresh) {
            const float4* src = reinterpret_cast<const float4*>(tp.state + ((pix0 + pix) * 2 + p) * K);
#pragma unroll
            for (int q = 0; q < K / 4; ++q) {
                const float4 f = __ldg(src + q);
                sst_f32(a_u + (4 * q + 0) * row_bytes + c * 4u, f.x); sst_f32(a_u + (4 * q + 1) * row_bytes + c * 4u, f.y);
                sst_f32(a_u + (4 * q + 2) * row_bytes + c * 4u, f.z); sst_f32(a_u + (4 * q + 3) * row_bytes + c * 4u, f.w);
            }
        } else {
#pragma unroll
            for (int e = 0; e < K; ++e) sst_f32(a_u + e * row_bytes + c * 4u, kTafInit);
        }
    }
    worker_sync();

    int A = 0;                                           // ageing steps not yet materialised in `u`
    uint32_t acc_cur = a_acc, acc_alt = a_acc + row_bytes;    // accumulator of the open bin / of the previous one
    uint32_t hot_cur = a_hot, hot_alt = a_hot + 4u;
    uint32_t hot_seen_cur = 0u, hot_seen_alt = 0u;       // last values seen of the two monotonic hot counters
    uint32_t held[kHeldSegments * kPerSeg];              // this thread's records of the open bin (the push visits them again)
#pragma unroll
    for (int i = 0; i < kHeldSegments * kPerSeg; ++i) held[i] = kNullRecord;
    int seg_in_bin = 0;                                  // segments of the open bin seen so far
    bool dirty = false;                                  // pushes since the last worker barrier

    // one value into the circular buffer of a cell: mean(t_norm) - 1 = S / (n span) - 1 (generate_taf.py:23-27), stored as u = v + A
    auto push_cell = [&](uint32_t cell, uint32_t n, uint32_t s, float fa) {
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)n * tp.span));
        const uint32_t h = sld_u8(a_head + cell);
        sst_f32(a_u + h * row_bytes + cell * 4u, fmaf((float)s, r, fa));
        sst_u8(a_head + cell, (h + 1u) & (uint32_t)(K - 1));
    };
    // every record of this tile in bin `gbin`, straight from global memory (the exact path)
    auto for_bin_records = [&](uint32_t gbin, auto&& fn) {
        const BinDesc bd = sp.bins[gbin];
        const uint32_t parts = slice_parts(bd.lo, bd.hi);
        for (uint32_t c = tid >> 5; c < parts; c += kTsWarps) {
            const int64_t s = (int64_t)bd.first_slice + c;
            const uint32_t v = sp.runs[(int64_t)tile * sp.pitch + s];
            const uint32_t* src = sp.records + s * sp.slice_stride + (v & 0xFFFFu) * 4u;
            const uint32_t n = ((v >> 16) - (v & 0xFFFFu)) * 4u;
            for (uint32_t i = lane; i < n; i += 32u) {
                const uint32_t rec = __ldg(src + i);
                if (rec != kNullRecord) fn(rec);
            }
        }
    };

    // Emission.  v = u - A for every cell, written back in FIFO order with the head reset to slot 0 (rebase + unrotate in
    // place): the state array [K][2][P] then IS the tile's slice of the [2K,H,W] window tensor, channel 2e + p = row (e, p),
    // and one thread sends it with 2K TMA bulk stores straight out of the state.  The rows must stay untouched until the
    // copies have read them: `store_pending` makes the next push / sweep wait (thread 0, before the workers' barrier).
    // The uint8 file bytes and the [H,W,2,K] state tensor are written by the threads themselves.
    bool store_pending = false;
    auto drain_stores = [&] {
        if (store_pending) {
            if (tid == 0) bulk_wait_read();
            store_pending = false;
        }
    };
    auto sweep = [&](float* o, uint8_t* o8, bool write_state) {
        const float fa = (float)A;
        if (tp.vec_out) {
            const int quads = (npix + 3) >> 2;           // npix is a multiple of 4 here
            for (int q = tid; q < 2 * quads; q += kTsWorkers) {
                const int p = q >= quads ? 1 : 0, pix = (q - p * quads) * 4, c = p * P + pix;
                const uint32_t h4 = sld_u32(a_head + c);
                float4 r[K];
                const uint32_t addr = a_u + (uint32_t)c * 4u;
#pragma unroll
                for (int e = 0; e < K; ++e) r[e] = sld_v4f(addr + e * row_bytes);     // all loads in flight before the first use
                // rotate each pixel's slots into FIFO order: position e = slot (head + e) mod K
                float v[4][K];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t h = (h4 >> (8 * j)) & 0xFFu;
                    float a[K];
#pragma unroll
                    for (int e = 0; e < K; ++e) a[e] = (j == 0 ? r[e].x : j == 1 ? r[e].y : j == 2 ? r[e].z : r[e].w) - fa;
#pragma unroll
                    for (int sh = 1; sh < K; sh <<= 1) {
                        const bool on = (h & (uint32_t)sh) != 0u;
                        float t[K];
#pragma unroll
                        for (int e = 0; e < K; ++e) t[e] = on ? a[(e + sh) & (K - 1)] : a[e];
#pragma unroll
                        for (int e = 0; e < K; ++e) a[e] = t[e];
                    }
#pragma unroll
                    for (int e = 0; e < K; ++e) v[j][e] = a[e];
                }
                if (h4) sst_u32(a_head + c, 0u);
#pragma unroll
                for (int e = 0; e < K; ++e) sst_v4f(addr + e * row_bytes, make_float4(v[0][e], v[1][e], v[2][e], v[3][e]));
                if (o8) {
                    // [K,2,H,W] with slot 0 = newest bin (np.flip of the [K,2,H,W] view)
                    uint8_t* row = o8 + ((int64_t)(K - 1) * 2 + p) * HW + pix;
#pragma unroll
                    for (int e = 0; e < K; ++e, row -= 2 * HW) {
                        const uint32_t b4 = leaky_u8_fast(v[0][e]) | (leaky_u8_fast(v[1][e]) << 8) | (leaky_u8_fast(v[2][e]) << 16) |
                                            (leaky_u8_fast(v[3][e]) << 24);
                        *reinterpret_cast<uint32_t*>(row) = b4;
                    }
                }
                if (write_state) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4* dst = reinterpret_cast<float4*>(tp.state + ((pix0 + pix + j) * 2 + p) * K);
#pragma unroll
                        for (int qq = 0; qq < K / 4; ++qq) dst[qq] = make_float4(v[j][4 * qq], v[j][4 * qq + 1], v[j][4 * qq + 2], v[j][4 * qq + 3]);
                    }
                }
            }
            if (o) {
                fence_async_smem();                          // the rewritten rows, before the bulk copies read them
                worker_sync();
                if (tid == 0) {
#pragma unroll
                    for (int e = 0; e < K; ++e)
#pragma unroll
                        for (int p = 0; p < 2; ++p)
                            bulk_store_1d(o + (int64_t)(2 * e + p) * HW, smem_raw + lay.state + (size_t)(e * C + p * P) * 4, (uint32_t)npix * 4u);
                    bulk_commit();
                }
                store_pending = true;
            }
            return;
        }
        for (int c = tid; c < C; c += kTsWorkers) {
            const int p = c >= P ? 1 : 0, pix = c - p * P;
            if (pix >= npix) continue;
            const uint32_t h = sld_u8(a_head + c);
            float v[K];
#pragma unroll
            for (int e = 0; e < K; ++e) v[e] = __uint_as_float(sld_u32(a_u + ((h + e) & (K - 1)) * row_bytes + (uint32_t)c * 4u)) - fa;
            sst_u8(a_head + c, 0u);
#pragma unroll
            for (int e = 0; e < K; ++e) sst_f32(a_u + e * row_bytes + (uint32_t)c * 4u, v[e]);
            if (o) {
#pragma unroll
                for (int e = 0; e < K; ++e) __stcs(o + (int64_t)(2 * e + p) * HW + pix, v[e]);
            }
            if (o8) {
#pragma unroll
                for (int e = 0; e < K; ++e) o8[((int64_t)(K - 1 - e) * 2 + p) * HW + pix] = (uint8_t)leaky_u8_fast(v[e]);
            }
            if (write_state) {
                float4* dst = reinterpret_cast<float4*>(tp.state + ((pix0 + pix) * 2 + p) * K);
#pragma unroll
                for (int q = 0; q < K / 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
        }
    };

#ifdef EVREP_TAF_TIMING
    long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long w_start = clock64();
#endif
    // The push of a bin is software-pipelined behind the accumulation of the next one: after the barrier that closes bin b
    // the threads first fire the atomics of bin b + 1 (other accumulator buffer), then visit the cells of bin b again.
    bool push_pending = false;
    bool pend_scan = false;                              // the pending bin had more segments than a thread holds: scan instead
    float pend_fa = 0.0f;
    uint32_t pend[kHeldSegments * kPerSeg];
#pragma unroll
    for (int i = 0; i < kHeldSegments * kPerSeg; ++i) pend[i] = kNullRecord;
    auto flush_push = [&] {
        if (!push_pending) return;
        TAF_T0(tp_);
        // acc_alt is the pending bin's buffer (the buffers were swapped when it closed)
        if (!pend_scan) {
            // whoever exchanges a cell's word first gets its content and pushes; later visitors read zero
            uint32_t xw[kHeldSegments * kPerSeg];
#pragma unroll
            for (int i = 0; i < kHeldSegments * kPerSeg; ++i)
                xw[i] = pend[i] != kNullRecord ? satom_exch(acc_alt + (pend[i] & 0x3FFFu) * 4u, 0u) : 0u;
#pragma unroll
            for (int i = 0; i < kHeldSegments * kPerSeg; ++i)
                if (xw[i]) push_cell(pend[i] & 0x3FFFu, xw[i] >> kCountShift, xw[i] & kSumMask, pend_fa);
        } else {
            for (int c = tid; c < C; c += kTsWorkers) {
                const uint32_t xw = sld_u32(acc_alt + c * 4u);
                if (xw) { sst_u32(acc_alt + c * 4u, 0u); push_cell((uint32_t)c, xw >> kCountShift, xw & kSumMask, pend_fa); }
            }
        }
        push_pending = false;
        dirty = true;
        TAF_ADD(7, tp_);
    };
    // everything that sweeps the state first completes the pending push and waits for everybody's pushes and bulk stores
    auto quiesce = [&] {
        flush_push();
        if (store_pending) { drain_stores(); dirty = true; }
        if (dirty) { worker_sync(); dirty = false; }
    };

    for (uint32_t seq = 0;; ++seq) {
        const uint32_t slot = seq % kFeedStages;
        TAF_T0(tw);
        mbar_wait(full + slot, (seq / kFeedStages) & 1u);
        TAF_ADD(3, tw);
        const uint4 dd = sld_v4(a_desc + slot * 16u);
        const uint32_t n_rec = dd.x, flags = dd.y, age_inc = dd.z, arg = dd.w;
        if (flags & kSegReset) {
            quiesce();
            for (int c = tid; c < C; c += kTsWorkers) {
                sst_u8(a_head + c, 0u);
#pragma unroll
                for (int e = 0; e < K; ++e) sst_f32(a_u + e * row_bytes + c * 4u, kTafInit);
            }
            A = 0;
        }
        const bool wide = (flags & kSegWide) != 0;         // offsets beyond the packed word: the bin takes the exact path
        TAF_T0(ta);
        if (n_rec && !wide) {
            // one packed atomic per record; the records stay in registers so that the push can visit their cells again
            const uint32_t recs = a_ring + slot * (kFeedStageRecords * 4u) + (uint32_t)tid * 4u;
            uint32_t mine[kPerSeg];
#pragma unroll
            for (int j = 0; j < kPerSeg; ++j)
                mine[j] = (uint32_t)(tid + j * kTsWorkers) < n_rec ? sld_u32(recs + j * (kTsWorkers * 4u)) : kNullRecord;
#pragma unroll
            for (int j = 0; j < kPerSeg; ++j) {
                if (mine[j] != kNullRecord) {
                    const uint32_t old = satom_add(acc_cur + (mine[j] & 0x3FFFu) * 4u, (1u << kCountShift) + (mine[j] >> 14));
                    if ((old >> kCountShift) >= kHotCount) sred_add(hot_cur, 1u);
                }
            }
#pragma unroll
            for (int g = 0; g < kHeldSegments; ++g)
                if (seg_in_bin == g) {
#pragma unroll
                    for (int j = 0; j < kPerSeg; ++j) held[g * kPerSeg + j] = mine[j];
                }
            ++seg_in_bin;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);           // this warp is done with the stage
        TAF_ADD(6, ta);
        flush_push();                                        // the previous bin, behind this segment's atomics
        A += (int)age_inc;
        if (flags & kSegBinEnd) {
            TAF_T0(tb);
            drain_stores();                                  // the window tensor has left the state rows
            worker_sync();                                   // the bin is accumulated; earlier pushes are complete
            dirty = false;
            TAF_ADD(4, tb);
            const float fa = (float)(A - 1);
            const uint32_t hot_now = sld_u32(hot_cur);
            const bool exact = wide || hot_now != hot_seen_cur;
            hot_seen_cur = hot_now;
            if (!exact) {
                push_pending = true;
                pend_scan = seg_in_bin > kHeldSegments;
                pend_fa = fa;
#pragma unroll
                for (int i = 0; i < kHeldSegments * kPerSeg; ++i) pend[i] = held[i];
                { const uint32_t t = acc_cur; acc_cur = acc_alt; acc_alt = t; }
                { const uint32_t t = hot_cur; hot_cur = hot_alt; hot_alt = t; }
                { const uint32_t t = hot_seen_cur; hot_seen_cur = hot_seen_alt; hot_seen_alt = t; }
            } else {
                // Exact path (a cell gathered >= 256 events, or offsets beyond 32767): the bin is read again from global
                // memory, twice -- counts, parked in the slot the mean will go to, then sums of (d + 1) -- so one
                // accumulator word per cell suffices.  Rare; five extra barriers.
                for (int c = tid; c < C; c += kTsWorkers) sst_u32(acc_cur + c * 4u, 0u);
                worker_sync();
                for_bin_records(arg, [&](uint32_t rec) { sred_add(acc_cur + (rec & 0x3FFFu) * 4u, 1u); });
                worker_sync();
                for (int c = tid; c < C; c += kTsWorkers) {
                    const uint32_t n = sld_u32(acc_cur + c * 4u);
                    if (n) { sst_u32(a_u + sld_u8(a_head + c) * row_bytes + c * 4u, n); sst_u32(acc_cur + c * 4u, 0u); }
                }
                worker_sync();
                for_bin_records(arg, [&](uint32_t rec) { sred_add(acc_cur + (rec & 0x3FFFu) * 4u, (rec >> 14) + 1u); });
                worker_sync();
                for (int c = tid; c < C; c += kTsWorkers) {
                    const uint32_t s1 = sld_u32(acc_cur + c * 4u);
                    if (s1) {
                        const uint32_t n = sld_u32(a_u + sld_u8(a_head + c) * row_bytes + c * 4u);
                        sst_u32(acc_cur + c * 4u, 0u);
                        push_cell((uint32_t)c, n, s1 - n, fa);
                    }
                }
                worker_sync();
            }
#pragma unroll
            for (int i = 0; i < kHeldSegments * kPerSeg; ++i) held[i] = kNullRecord;
            seg_in_bin = 0;
        }
        if (flags & kSegEmit) {
            quiesce();
            const bool last = (flags & kSegDone) != 0;
            // every emission rebases (u = v, A = 0): the state after a window does not depend on how the
            // windows are split over launches, so split launches == one launch, bit for bit
            float* o = tp.out ? tp.out + (int64_t)arg * tp.out_stride + pix0 : nullptr;
            uint8_t* o8 = tp.out_u8 ? tp.out_u8 + (int64_t)arg * tp.out_u8_stride + pix0 : nullptr;
            TAF_T0(ts);
            sweep(o, o8, tp.emit_state || last);
            TAF_ADD(5, ts);
            A = 0;
        } else if (A >= kRebaseAlways) {
            quiesce();
            sweep(nullptr, nullptr, false);
            A = 0;
        }
        if (flags & kSegDone) break;
    }
    if (tid == 0) bulk_wait_all();                           // shared memory must outlive the bulk reads
#ifdef EVREP_TAF_TIMING
    if (tid == 0) {
        g_taf_timing[tile][2] = clock64() - w_start;
        for (int i = 3; i < 8; ++i) g_taf_timing[tile][i] = t_acc[i];
    }
#endif
}

template <int K>
static int launch_taf_slices(const TafSliceParams& tp, cudaStream_t st) {
    const size_t smem = taf_slice_smem(tp.sp.P, K);
    EVREP_CUDA(cudaFuncSetAttribute(taf_slice_tile_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    taf_slice_tile_kernel<K><<<tp.sp.n_tiles, kTsThreads, smem, st>>>(tp);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

static int taf_ordered_ctas_per_sm() {
    const char* e = getenv("EVREP_TAF_CTAS_PER_SM");     // tuning knob: resident tile CTAs per SM (1..4)
    const int v = e ? atoi(e) : 2;
    return v < 1 ? 1 : (v > 4 ? 4 : v);
}

}  // namespace evrep

using namespace evrep;

extern "C" {

int64_t evrep_taf_stream_sliced_scratch_bytes(int64_t n_events, int n_windows, int64_t total_bins, int H, int W, int K) {
    if (n_events < 0 || n_windows < 0 || total_bins < 0 || H <= 0 || W <= 0 || (K != 4 && K != 8)) return EVREP_ERR_ARG;
    int P, n_tiles;
    int rc = choose_tile(H, W, K, taf_slice_smem, taf_ordered_ctas_per_sm(), P, n_tiles);
    if (rc) return rc;
    SliceLayout L;
    rc = make_slice_layout(n_events, n_windows, total_bins, H, W, P, n_tiles, L);
    if (rc) return rc;
    return L.total;
}

int evrep_taf_stream_sliced(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                             const evrep_taf_window* windows_host, int n_windows, int abin, int H, int W, int K,
                             const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                             float* state_inout, int emit_state_every_window,
                             float* out, int64_t out_stride, uint8_t* out_u8, int64_t out_u8_stride,
                             void* scratch, int64_t scratch_bytes,
                             void* ev_tiles_begin, void* ev_tiles_end, evrep_stream_t stream) {
    if (n_events < 0 || n_windows < 0 || H <= 0 || W <= 0 || abin <= 0 || !state_inout || !scratch) return EVREP_ERR_ARG;
    if (K != 4 && K != 8) return EVREP_ERR_ARG;
    if (n_windows == 0) return EVREP_OK;
    if ((!out && !out_u8) || (reinterpret_cast<uintptr_t>(state_inout) & 15)) return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    int P, n_tiles;
    int rc = choose_tile(H, W, K, taf_slice_smem, taf_ordered_ctas_per_sm(), P, n_tiles);
    if (rc) return rc;
    TafSliceParams tp;
    SliceLayout L;
    rc = prepare_slices(t, x, y, p, n_events, windows_host, n_windows, abin, H, W, P, n_tiles, xmap, ymap, sensor_h, sensor_w,
                        scratch, scratch_bytes, st, tp.sp, L);
    if (rc) return rc;
    tp.state = state_inout; tp.out = out; tp.out_stride = out_stride;
    tp.out_u8 = out_u8; tp.out_u8_stride = out_u8_stride;
    tp.emit_state = emit_state_every_window;
    tp.span = (float)((double)abin + 1e-8);
    // four pixels per thread in the emission sweep: every tile starts on a multiple of 4 pixels (P is a multiple of 8),
    // so rows are 16-byte (float) / 4-byte (uint8) aligned when the planes and window strides are
    tp.vec_out = (((int64_t)H * W) % 4 == 0 && (!out || (out_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)) &&
                  (!out_u8 || (out_u8_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out_u8) & 3) == 0))) ? 1 : 0;
    if (ev_tiles_begin) EVREP_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev_tiles_begin), st));
    rc = K == 8 ? launch_taf_slices<8>(tp, st) : launch_taf_slices<4>(tp, st);
    if (rc) return rc;
    if (ev_tiles_end) EVREP_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev_tiles_end), st));
    return EVREP_OK;
}

int evrep_stream_order_violations(const void* scratch, uint32_t* host_out, evrep_stream_t stream) {
    if (!scratch || !host_out) return EVREP_ERR_ARG;
    EVREP_CUDA(cudaMemcpyAsync(host_out, scratch, 4, cudaMemcpyDeviceToHost, as_stream(stream)));
    EVREP_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return EVREP_OK;
}

#ifdef EVREP_TAF_TIMING
int evrep_debug_taf_timing(unsigned long long* host, int n_tiles) {
    EVREP_CUDA(cudaDeviceSynchronize());
    EVREP_CUDA(cudaMemcpyFromSymbol(host, g_taf_timing, sizeof(unsigned long long) * 8 * (size_t)n_tiles));
    return EVREP_OK;
}
#endif

}  // extern "C"
