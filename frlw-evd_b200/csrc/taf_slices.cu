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
constexpr int kTsThreads = kTsWorkers + 32;        // + the producer warp
constexpr int kTsWorkerWarps = kTsWorkers / 32;
constexpr int kBarWorkers = 1;                     // named barrier of the worker warps
constexpr uint32_t kCountShift = 23;               // packed accumulator: count in bits 23..31, sum of d below
constexpr uint32_t kHotCount = 256;                // a cell that reaches this count in one bin forces the exact path
constexpr int kRebaseAlways = 8;                   // A never exceeds this: rebase inside a window when it gets there

struct TafSliceParams {
    SlicePlan sp;
    float* state;          // [H,W,2,K]
    float* out;            // window w at out + w * out_stride (may be null when out_u8 is set)
    int64_t out_stride;
    uint8_t* out_u8;       // optional: u8 [n_windows][K,2,H,W], leaky transform + slot flip (generate_taf.py:226-235)
    int64_t out_u8_stride;
    int emit_state;        // write the state after every window (always after the last)
    float span;            // f32(abin + 1e-8)
};

struct TafSliceSmem {
    int state, acc, list, head, ctrl, feed_base, total;
    __host__ __device__ TafSliceSmem(int P, int K) {
        int o = 0;
        state = o; o += K * 2 * P * 4;                  // u[K][2P]
        acc = o;   o += 2 * 2 * P * 4;                  // two buffers of packed {count | sum d} per cell
        list = o;  o += 2 * 2 * P * 2;                  // two active lists (u16 cells)
        head = o;  o += 2 * P; o = (o + 15) / 16 * 16;  // next slot to overwrite, per cell
        ctrl = o;  o += 32;                             // list counters [2], hot flags [2]
        feed_base = o;
        total = FeedSmem(o).total;
    }
};

static size_t taf_slice_smem(int P, int K) { return (size_t)TafSliceSmem(P, K).total; }

__device__ __forceinline__ void worker_sync() { named_sync(kBarWorkers, kTsWorkers); }

// generate_taf.py:69-76 on one value: 255 * max(0, 1 - log1p(-v) / 8.7), truncated to uint8
__device__ __forceinline__ uint32_t leaky_u8(float v) {
    const float r = 1.0f - __fdiv_rn(log1pf(-v), 8.7f);          // same arithmetic as taf_leaky_u8_kernel
    return (uint32_t)(int)((r < 0.0f ? 0.0f : r) * 255.0f);
}

template <int K>
__global__ void __launch_bounds__(kTsThreads)
taf_slice_tile_kernel(const __grid_constant__ TafSliceParams tp) {
    const SlicePlan& sp = tp.sp;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = sp.P;
    const TafSliceSmem lay(P, K);
    const FeedSmem fs(lay.feed_base);
    float* u = reinterpret_cast<float*>(smem_raw + lay.state);            // [K][2P]
    uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw + lay.acc);      // [2][2P]
    uint16_t* list = reinterpret_cast<uint16_t*>(smem_raw + lay.list);    // [2][2P]
    uint8_t* head = smem_raw + lay.head;                                  // [2P]
    uint32_t* list_count = reinterpret_cast<uint32_t*>(smem_raw + lay.ctrl);   // [2], monotonic
    uint32_t* hot = list_count + 2;                                       // [2]
    const uint32_t* ring = reinterpret_cast<const uint32_t*>(smem_raw + fs.ring);
    const SegDesc* desc = reinterpret_cast<const SegDesc*>(smem_raw + fs.desc);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + fs.full);
    uint64_t* empty = reinterpret_cast<uint64_t*>(smem_raw + fs.empty);

    const int tid = threadIdx.x, tile = blockIdx.x, lane = tid & 31;
    const int64_t HW = (int64_t)sp.H * sp.W;
    const int64_t pix0 = (int64_t)tile * P;
    const int npix = (int)min((int64_t)P, HW - pix0);
    const int C = 2 * P;                                 // cells of the tile: cell = p * P + pixel

    if (tid == 0) {
        for (int s = 0; s < kFeedStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kTsWorkerWarps); }
        list_count[0] = list_count[1] = 0u;
        hot[0] = hot[1] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= kTsWorkers) {
        // ================================== producer warp ==================================
        FeedProducer fp;
        fp.init(smem_raw, fs);
        for (int w = 0; w < sp.n_windows; ++w) {
            if (w > 0 && sp.w_fresh[w]) fp.control(kSegReset, 0u, (uint32_t)w);
            const int g0 = sp.w_binbase[w], g1 = sp.w_binbase[w + 1];
            uint32_t pend_age = 0;                       // bins that age the tile without bringing it records
            for (int gb = g0; gb < g1; gb += 32) {
                // descriptors of 32 bins at a time, one per lane
                BinDesc mine;
                mine.lo = mine.hi = 0; mine.t0 = 0; mine.first_slice = 0; mine.flags = 0; mine.dyn = 0; mine.win = 0;
                if (gb + lane < g1) mine = sp.bins[gb + lane];
                const int nb = min(32, g1 - gb);
                for (int k = 0; k < nb; ++k) {
                    BinDesc bd;
                    bd.lo = __shfl_sync(0xFFFFFFFFu, mine.lo, k);
                    bd.hi = __shfl_sync(0xFFFFFFFFu, mine.hi, k);
                    bd.first_slice = __shfl_sync(0xFFFFFFFFu, mine.first_slice, k);
                    bd.dyn = __shfl_sync(0xFFFFFFFFu, mine.dyn, k);
                    bd.t0 = 0; bd.flags = 0; bd.win = 0;
                    if (!(bd.dyn & kBinAny)) continue;   // nobody saw an event: no ageing (generate_taf.py:40-41)
                    if (fp.feed_bin(sp, tile, (uint32_t)(gb + k), bd, pend_age + 1u)) pend_age = 0;
                    else ++pend_age;
                }
            }
            fp.control(kSegEmit | (w == sp.n_windows - 1 ? kSegDone : 0u), pend_age, (uint32_t)w);
        }
        return;
    }

    // ==================================== worker warps ====================================
    const int wid = tid >> 5;
    const bool first_fresh = sp.w_fresh[0] != 0;
    // state in: u = v (A = 0), heads at 0 -- slot e of a cell holds FIFO position e (K - 1 = newest)
    for (int c = tid; c < C; c += kTsWorkers) {
        const int p = c >= P ? 1 : 0, pix = c - p * P;
        head[c] = 0;
        acc[c] = 0u; acc[C + c] = 0u;
        if (pix < npix && !first_fresh) {
            const float4* src = reinterpret_cast<const float4*>(tp.state + ((pix0 + pix) * 2 + p) * K);
#pragma unroll
            for (int q = 0; q < K / 4; ++q) {
                const float4 f = __ldg(src + q);
                u[(4 * q + 0) * C + c] = f.x; u[(4 * q + 1) * C + c] = f.y;
                u[(4 * q + 2) * C + c] = f.z; u[(4 * q + 3) * C + c] = f.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < K; ++e) u[e * C + c] = kTafInit;
        }
    }
    worker_sync();

    int A = 0;                                           // ageing steps not yet materialised in `u`
    int buf = 0;
    uint32_t list_base[2] = {0u, 0u};
    bool dirty = false;                                  // pushes since the last worker barrier
    bool wide_open = false;                              // inside a bin that uses the two-word accumulators

    // add one record to the packed accumulator of `b`; returns through `first` whether the cell was untouched
    auto append = [&](int b, bool first, uint32_t cell) {
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, first);
        if (mask) {
            uint32_t base = 0;
            const int leader = __ffs(mask) - 1;
            if (lane == leader) base = atomicAdd(&list_count[b], (uint32_t)__popc(mask));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (first) list[b * C + (base - list_base[b]) + __popc(mask & ((1u << lane) - 1u))] = (uint16_t)cell;
        }
    };
    auto accumulate_packed = [&](int b, uint32_t rec) {
        bool first = false;
        const uint32_t cell = rec & 0x3FFFu;
        if (rec != kNullRecord) {
            const uint32_t old = atomicAdd(&acc[b * C + cell], (1u << kCountShift) + (rec >> 14));
            first = old == 0u;
            if ((old >> kCountShift) >= kHotCount) hot[b] = 1u;
        }
        append(b, first, cell);
    };
    // two-word accumulators: counts in buffer b, sums of d in the other one
    auto accumulate_wide = [&](int b, uint32_t rec) {
        bool first = false;
        const uint32_t cell = rec & 0x3FFFu;
        if (rec != kNullRecord) {
            first = atomicAdd(&acc[b * C + cell], 1u) == 0u;
            atomicAdd(&acc[(b ^ 1) * C + cell], rec >> 14);
        }
        append(b, first, cell);
    };
    // push the mean of every active cell: one store into its circular buffer
    auto push = [&](int b, bool wide) {
        const uint32_t count = list_count[b] - list_base[b];
        const float fa = (float)(A - 1);
        for (uint32_t j = tid; j < count; j += kTsWorkers) {
            const uint32_t cell = list[b * C + j];
            uint32_t n, s;
            if (wide) {
                n = acc[b * C + cell]; s = acc[(b ^ 1) * C + cell];
                acc[b * C + cell] = 0u; acc[(b ^ 1) * C + cell] = 0u;
            } else {
                const uint32_t xw = acc[b * C + cell];
                acc[b * C + cell] = 0u;
                n = xw >> kCountShift; s = xw & ((1u << kCountShift) - 1u);
            }
            // mean(t_norm) - 1 = S / (n span) - 1 (generate_taf.py:23-27), stored as u = v + A
            float r;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)n * tp.span));
            const uint32_t h = head[cell];
            u[h * C + cell] = fmaf((float)s, r, fa);
            head[cell] = (uint8_t)((h + 1u) & (uint32_t)(K - 1));
        }
        list_base[b] += count;
    };
    // v = u - A for every cell (FIFO order restored); optionally written back (rebase)
    auto sweep = [&](float* o, uint8_t* o8, bool write_state, bool rebase) {
        const float fa = (float)A;
        for (int c = tid; c < C; c += kTsWorkers) {
            const int p = c >= P ? 1 : 0, pix = c - p * P;
            if (pix >= npix) continue;
            const uint32_t h = head[c];
            float v[K];
#pragma unroll
            for (int e = 0; e < K; ++e) v[e] = u[((h + e) & (K - 1)) * C + c] - fa;
            if (rebase) {
#pragma unroll
                for (int e = 0; e < K; ++e) u[((h + e) & (K - 1)) * C + c] = v[e];
            }
            if (o) {
#pragma unroll
                for (int e = 0; e < K; ++e) __stcs(o + (int64_t)(2 * e + p) * HW + pix, v[e]);
            }
            if (o8) {
                // [K,2,H,W] with slot 0 = newest bin (np.flip of the [K,2,H,W] view)
#pragma unroll
                for (int e = 0; e < K; ++e) o8[((int64_t)(K - 1 - e) * 2 + p) * HW + pix] = (uint8_t)leaky_u8(v[e]);
            }
            if (write_state) {
                float4* dst = reinterpret_cast<float4*>(tp.state + ((pix0 + pix) * 2 + p) * K);
#pragma unroll
                for (int q = 0; q < K / 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
        }
    };

    for (uint32_t seq = 0;; ++seq) {
        const int slot = (int)(seq % kFeedStages);
        mbar_wait(full + slot, (seq / kFeedStages) & 1u);
        const SegDesc d = desc[slot];
        if (d.flags & kSegReset) {
            if (dirty) { worker_sync(); dirty = false; }
            for (int c = tid; c < C; c += kTsWorkers) {
                head[c] = 0;
#pragma unroll
                for (int e = 0; e < K; ++e) u[e * C + c] = kTafInit;
            }
            A = 0;
        }
        const bool wide = (d.flags & kSegWide) != 0;
        if (d.n_rec) {
            if (wide && !wide_open) {
                // the other accumulator buffer must be idle: everybody has finished the previous push
                worker_sync(); dirty = false; wide_open = true;
            }
            const uint32_t* recs = ring + slot * kFeedStageRecords;
            for (uint32_t i0 = (uint32_t)wid * 32u; i0 < d.n_rec; i0 += kTsWorkers) {
                const uint32_t i = i0 + lane;
                const uint32_t rec = i < d.n_rec ? recs[i] : kNullRecord;
                if (wide) accumulate_wide(buf, rec); else accumulate_packed(buf, rec);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);           // this warp is done with the stage
        A += (int)d.age_inc;
        if (d.flags & kSegBinEnd) {
            worker_sync();                                   // the bin is accumulated; earlier pushes are complete
            bool exact = wide;
            if (!wide && hot[buf]) {
                // some cell overflowed the packed word: clear what the bin touched and accumulate it again from
                // global memory with two-word accumulators
                const uint32_t count = list_count[buf] - list_base[buf];
                for (uint32_t j = tid; j < count; j += kTsWorkers) acc[buf * C + list[buf * C + j]] = 0u;
                list_base[buf] += count;
                worker_sync();
                if (tid == 0) hot[buf] = 0u;
                const BinDesc bd = sp.bins[d.arg];
                const uint32_t parts = slice_parts(bd.lo, bd.hi);
                for (uint32_t c = wid; c < parts; c += kTsWorkerWarps) {
                    const int64_t s = (int64_t)bd.first_slice + c;
                    const uint32_t a = sp.off16[s * sp.pitch + tile], b = sp.off16[s * sp.pitch + tile + 1];
                    const uint32_t* src = sp.records + s * sp.slice_stride + a * 4u;
                    const uint32_t n = (b - a) * 4u;
                    for (uint32_t i0 = 0; i0 < n; i0 += 32u) {
                        const uint32_t i = i0 + lane;
                        accumulate_wide(buf, i < n ? __ldg(src + i) : kNullRecord);
                    }
                }
                worker_sync();
                exact = true;
            }
            push(buf, exact);
            if (exact) { worker_sync(); dirty = false; wide_open = false; }    // both buffers are clean again
            else { dirty = true; buf ^= 1; }
        }
        if (d.flags & kSegEmit) {
            if (dirty) { worker_sync(); dirty = false; }
            const bool last = (d.flags & kSegDone) != 0;
            // every emission rebases (u = v, A = 0): the state after a window does not depend on how the
            // windows are split over launches, so split launches == one launch, bit for bit
            float* o = tp.out ? tp.out + (int64_t)d.arg * tp.out_stride + pix0 : nullptr;
            uint8_t* o8 = tp.out_u8 ? tp.out_u8 + (int64_t)d.arg * tp.out_u8_stride + pix0 : nullptr;
            sweep(o, o8, tp.emit_state || last, true);
            A = 0;
        } else if (A >= kRebaseAlways) {
            if (dirty) { worker_sync(); dirty = false; }
            sweep(nullptr, nullptr, false, true);
            A = 0;
        }
        if (d.flags & kSegDone) break;
    }
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

int64_t evrep_taf_stream_ordered_scratch_bytes(int64_t n_events, int n_windows, int64_t total_bins, int H, int W, int K) {
    if (n_events < 0 || n_windows < 0 || total_bins < 0 || H <= 0 || W <= 0 || (K != 4 && K != 8)) return EVREP_ERR_ARG;
    int P, n_tiles;
    int rc = choose_tile(H, W, K, taf_slice_smem, taf_ordered_ctas_per_sm(), P, n_tiles);
    if (rc) return rc;
    SliceLayout L;
    rc = make_slice_layout(n_events, n_windows, total_bins, H, W, P, n_tiles, L);
    if (rc) return rc;
    return L.total;
}

int evrep_taf_stream_ordered(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
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

}  // extern "C"
