#include "stream_common.cuh"

namespace evrep {

// ---- Event Volume over whole streams ----------------------------------------------------------
// generate_eventvolume.py:15-42 for a list of non-overlapping windows: the same bucketing (one
// "bin" per window, d = t - t0) feeds one CTA per sensor tile.  The tile's [2K][P] accumulator
// lives in shared memory: splat, then one pass that reads, clears, scales by /5*255 and stores
// 16 bytes per thread to the tensor (rows are contiguous).  Two CTAs per SM, so one tile's
// output pass overlaps the other tile's splat.  Records arrive through the same ring of TMA
// bulk copies as in the TAF kernel.
//
// Accumulator format.  Shared-memory float atomics are compare-and-swap loops and the events of
// a moving edge pile onto a handful of cells, so the weights (f32, computed exactly as the
// reference does) are summed in fixed point with native u32 atomics instead: a cell is a u32
// `lo` in units of 2^-27 (wraps at 32) plus a u8 count of wraps, value = 32 hi + lo 2^-27.
// A weight converts to within 2^-28 (absolute); the sum itself is exact and order independent,
// so the output is deterministic.  The reference's own f32 running sum rounds by up to 2^-24 of
// the partial sum per add, which is the larger error once a cell holds more than one event.
// The wrap count saturates at 255, i.e. a cell sum of 8160; the reference's on-disk format
// already saturates at a sum of 5 (uint8 of sum / 5 * 255, generate_eventvolume.py:160).
constexpr int kEvThreads = 512;
constexpr int kEvTilesPerSm = 2;        // two CTAs per SM: one splats while the other's stores drain
constexpr float kEvUnit = 134217728.0f; // 2^27
constexpr float kEvWrap = 32.0f;        // 2^32 / 2^27

struct EvTileParams {
    StreamPlan pl;
    float* out;
    int64_t out_stride;
    double tw;             // window length: t_norm = d / tw in float64 (generate_eventvolume.py:141)
    int K;
    int vec_out;
};

struct EvTileSmem {
    int ring, acc, hi, bars, total;
    __host__ __device__ EvTileSmem(int P, int K) {
        int o = 0;
        ring = o; o += kWsRing * 4;
        acc = o;  o += 2 * K * P * 4;      // lo words
        hi = o;   o += 2 * K * P;          // wrap counts, one byte per cell
        bars = o; o += 64;
        total = o;
    }
};

__global__ void __launch_bounds__(kEvThreads, kEvTilesPerSm)
ev_tile_kernel(EvTileParams tp) {
    const StreamPlan& pl = tp.pl;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const EvTileSmem lay(pl.P, tp.K);
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw + lay.ring);
    uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw + lay.acc);       // [2K][P] lo words
    uint32_t* acc_hi = reinterpret_cast<uint32_t*>(smem_raw + lay.hi);     // [2K][P] bytes, 4 cells per word
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);

    const int tid = threadIdx.x, tile = blockIdx.x, lane = tid & 31;
    const int K = tp.K, rows = 2 * K;
    const uint32_t HW = (uint32_t)(pl.H * pl.W), P = (uint32_t)pl.P;
    const uint32_t pix0 = (uint32_t)tile * P;
    const uint32_t npix = min(P, HW - pix0);
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

    // Window w is bin w of the plan (one bin per window).  Its record range [off[w], off[w+1])
    // comes from registers: lane l of every warp holds the bounds of window g*32 + l, and the
    // next group of 32 is loaded one group ahead.
    const int n_windows = pl.n_windows;
    auto load_bounds = [&](int first, uint32_t& lo, uint32_t& hi) {
        const int w = min(first + lane, n_windows - 1);
        lo = __ldg(my_off + w); hi = __ldg(my_off + w + 1);
    };
    uint32_t b_lo, b_hi, nb_lo = 0, nb_hi = 0;
    load_bounds(0, b_lo, b_hi);

    int ready_chunk = -1, next_refill = kWsStages;
    const float Kf = (float)K;
    const double inv_tw = 1.0 / tp.tw;
    // output pass: thread walks the float4 cells i = tid + k * kEvThreads of the [2K][P] tile
    const uint32_t n4 = (uint32_t)rows * P / 4u, p4 = P / 4u;
    const uint32_t row_first = (uint32_t)tid / p4, c4_first = (uint32_t)tid - row_first * p4;
    const uint32_t row_step = kEvThreads / p4, c4_step = kEvThreads - row_step * p4;
    const uint32_t g_first = row_first * HW + c4_first * 4u;              // element offset inside the window tensor
    const uint32_t g_step = row_step * HW + c4_step * 4u, g_wrap = HW - P;
    for (uint32_t i = tid; i < (uint32_t)rows * P; i += kEvThreads) acc[i] = 0u;     // afterwards the output pass keeps it clean
    for (uint32_t i = tid; i < n4; i += kEvThreads) acc_hi[i] = 0u;
    __syncthreads();
    for (int w = 0; w < n_windows; ++w) {
        if ((w & 31) == 0) {
            if (w) { b_lo = nb_lo; b_hi = nb_hi; }
            if (w + 32 < n_windows) load_bounds(w + 32, nb_lo, nb_hi);
        }
        const uint32_t o0 = __shfl_sync(0xFFFFFFFFu, b_lo, w & 31), o1 = __shfl_sync(0xFFFFFFFFu, b_hi, w & 31);
        uint32_t cur = o0;
        while (cur < o1) {
            const uint32_t avail = (uint32_t)next_refill * kWsChunkRecords;     // records requested so far
            const uint32_t limit = o1 < avail ? o1 : avail;
            const int last_c = (int)((limit - 1) / kWsChunkRecords);
            while (ready_chunk < last_c) {
                ++ready_chunk;
                mbar_wait(full + (ready_chunk % kWsStages), (uint32_t)(ready_chunk / kWsStages) & 1u);
            }
            // A warp does not take 32 consecutive records (a burst on one pixel would serialise on
            // one cell): lane l walks records l * stride + warp + 16 k; stride is odd, so the ring
            // reads stay free of bank conflicts.
            const uint32_t span = limit - cur;
            const uint32_t stride = ((span + 31u) >> 5) | 1u;
            const uint32_t lane_base = (uint32_t)lane * stride;
            for (uint32_t q = (uint32_t)(tid >> 5); q < stride; q += kEvThreads / 32) {
                const uint32_t idx = lane_base + q;
                if (idx >= span) continue;
                const uint32_t rec = ring[(cur + idx) & (kWsRing - 1)];
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
                    const float wgt = 1.0f - fabsf((float)c - ts);
                    if (wgt > 0.0f) {
                        const uint32_t cell = (uint32_t)(2 * (c - 1) + (1 - (int)pol)) * P + lp;
                        const uint32_t fx = __float2uint_rn(wgt * kEvUnit);       // wgt <= 1
                        const uint32_t old = atomicAdd(acc + cell, fx);
                        if (old + fx < old) {                                     // lo wrapped: count it
                            const uint32_t sh = (cell & 3u) * 8u;
                            const uint32_t before = atomicAdd(acc_hi + (cell >> 2), 1u << sh);
                            if (((before >> sh) & 0xFFu) == 0xFFu) atomicSub(acc_hi + (cell >> 2), 1u << sh);   // saturate
                        }
                    }
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
        // read, clear and scale the accumulator (:37  / 5 * 255); each row of the tile is
        // contiguous in the tensor, so every warp store is 512 contiguous bytes
        float* o = tp.out + (int64_t)w * tp.out_stride + pix0;
        auto value = [](uint32_t lo, uint32_t wraps) -> float {
            return div5_mul255(fmaf((float)wraps, kEvWrap, (float)lo * (1.0f / kEvUnit)));
        };
        if (tp.vec_out) {
            uint32_t c4 = c4_first, g = g_first;
            for (uint32_t i = tid; i < n4; i += kEvThreads) {
                const uint4 lo = reinterpret_cast<uint4*>(acc)[i];
                const uint32_t wr = acc_hi[i];
                reinterpret_cast<uint4*>(acc)[i] = make_uint4(0u, 0u, 0u, 0u);
                if (wr) acc_hi[i] = 0u;
                if (c4 * 4u < npix) {
                    float4 v;
                    v.x = value(lo.x, wr & 0xFFu); v.y = value(lo.y, (wr >> 8) & 0xFFu);
                    v.z = value(lo.z, (wr >> 16) & 0xFFu); v.w = value(lo.w, wr >> 24);
                    __stcs(reinterpret_cast<float4*>(o + g), v);
                }
                c4 += c4_step; g += g_step;
                if (c4 >= p4) { c4 -= p4; g += g_wrap; }
            }
        } else {
            for (uint32_t i = tid; i < (uint32_t)rows * P; i += kEvThreads) {
                const uint32_t row = i / P, lp = i - row * P;
                const uint32_t lo = acc[i];
                const uint32_t wr = (acc_hi[i >> 2] >> ((i & 3u) * 8u)) & 0xFFu;
                acc[i] = 0u;
                if (lp < npix) __stcs(o + (int64_t)row * HW + lp, value(lo, wr));
            }
            __syncthreads();                                       // the byte counters share words
            for (uint32_t i = tid; i < n4; i += kEvThreads) acc_hi[i] = 0u;
        }
        __syncthreads();                                           // clears are ordered before the next splat
    }
}

}  // namespace evrep

using namespace evrep;

extern "C" {

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
    tp.vec_out = (((int64_t)H * W) % 4 == 0 && out_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                  L.P % 4 == 0) ? 1 : 0;
    EVREP_CUDA(cudaFuncSetAttribute(ev_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ev_tile_kernel<<<L.n_tiles, kEvThreads, smem, st>>>(tp);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

}  // extern "C"
