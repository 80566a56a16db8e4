#include "stream_common.cuh"

namespace evrep {

// ---- Event Volume over whole streams ----------------------------------------------------------
// generate_eventvolume.py:15-42 for a list of non-overlapping windows: the same bucketing (one
// "bin" per window, d = t - t0) feeds one CTA per sensor tile.  The tile's [2K][P] float
// accumulator lives in shared memory: splat (shared-memory float atomics), then one pass that
// reads, clears, scales by /5*255 and stores 16 bytes per thread to the tensor (rows are contiguous).
// Two CTAs per SM, so one tile's output pass overlaps the other tile's splat.  Records arrive through the same ring of TMA bulk copies as in the TAF kernel.
constexpr int kEvThreads = 512;
constexpr int kEvTilesPerSm = 2;        // two CTAs per SM: one splats while the other's bulk store drains


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
    tp.bulk_out = (((int64_t)H * W) % 4 == 0 && out_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                   L.P % 4 == 0) ? 1 : 0;
    EVREP_CUDA(cudaFuncSetAttribute(ev_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ev_tile_kernel<<<L.n_tiles, kEvThreads, smem, st>>>(tp);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

}  // extern "C"
