// Event Volume over time-ordered streams for windows that overlap and nest (the driver of
// generate_eventvolume.py:118-169 encodes, for every label, the last 250 / 500 / 1000 ms), on top
// of the slice sort (slices.cu).
//
// The caller cuts the stream at every window boundary into consecutive *segments* (each shorter
// than 262144 us, so that a record's 18-bit offset from the segment start suffices) and describes
// a window as a run of segments plus its own time origin t0 and length tw ("span").  The events
// are sorted ONCE, by (segment, sensor tile); every span then reads the records of its segments
// again -- 4 bytes per event and span instead of 9 -- and splats them with its own normalisation
// t_norm = (segment start - t0 + d) / tw.  One CTA per sensor tile, two per SM: four producer
// warps feed a ring of TMA bulk copies, eight worker warps accumulate the tile's [2K][P] slice
// in shared memory in fixed point (same format and arithmetic as ev_tile.cu: u32 in units of
// 2^-27 plus a byte of wraps, exact and order independent) and write it out once per span, as
// float32 (/ 5 * 255, generate_eventvolume.py:37) and / or as the uint8 file bytes (clamped at
// 255 and truncated, :158-160).  HBM-bound byte/float work: no tensor cores.
#include "slices.cuh"

namespace evrep {

constexpr int kEvsWorkers = 512;
constexpr int kEvsThreads = kEvsWorkers + 32 * kFeedWarps;
constexpr int kEvsWarps = kEvsWorkers / 32;
constexpr int kEvsBar = 1;
constexpr float kEvsUnit = 134217728.0f;   // 2^27
constexpr float kEvsWrap = 32.0f;          // 2^32 / 2^27

struct EvSpanDev {               // 24 bytes
    int32_t first_segment, last_segment;
    int64_t t0, tw;
};

struct EvSliceParams {
    SlicePlan sp;
    const EvSpanDev* spans;
    const int64_t* seg_start;    // [n_segments] = sp.w_start
    int n_spans;
    float* out;                  // f32 [n_spans][2K,H,W] or null
    int64_t out_stride;
    uint8_t* out_u8;             // u8 [n_spans][2K,H,W] or null
    int64_t out_u8_stride;
    int K;
    int vec_out;
};

struct EvSliceSmem {
    int acc, hi, feed_base, total;
    __host__ __device__ EvSliceSmem(int P, int K) {
        int o = 0;
        acc = o; o += 2 * K * P * 4;       // lo words
        hi = o;  o += 2 * K * P;           // wrap counts, one byte per cell
        o = (o + 15) / 16 * 16;
        feed_base = o;
        total = FeedSmem(o).total;
    }
};

static size_t ev_slice_smem(int P, int K) { return (size_t)EvSliceSmem(P, K).total; }

__global__ void __launch_bounds__(kEvsThreads, 2)
ev_slice_tile_kernel(const __grid_constant__ EvSliceParams tp) {
    const SlicePlan& sp = tp.sp;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int P = sp.P, K = tp.K, rows = 2 * K;
    const EvSliceSmem lay(P, K);
    const FeedSmem fs(lay.feed_base);
    uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw + lay.acc);       // [2K][P] lo words
    uint32_t* acc_hi = reinterpret_cast<uint32_t*>(smem_raw + lay.hi);     // [2K][P] bytes, 4 cells per word
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + fs.full);
    uint64_t* empty = reinterpret_cast<uint64_t*>(smem_raw + fs.empty);
    const int tid = threadIdx.x, tile = blockIdx.x, lane = tid & 31;
    const uint32_t HW = (uint32_t)(sp.H * sp.W);
    const uint32_t pix0 = (uint32_t)tile * (uint32_t)P;
    const uint32_t npix = min((uint32_t)P, HW - pix0);

    if (tid == 0) {
        for (int s = 0; s < kFeedStages; ++s) { mbar_init(full + s, kFeedFullCount); mbar_init(empty + s, kEvsWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= kEvsWorkers) {
        // ================================== producer warps =================================
        FeedProducer fp;
        fp.init(smem_raw, fs);
        fp.open_row(sp, tile);
        for (int s = 0; s < tp.n_spans; ++s) {
            const EvSpanDev span = tp.spans[s];
            for (int g0 = span.first_segment; g0 <= span.last_segment; g0 += 32) {
                uint32_t first = 0, parts = 0, dyn = 0, off = 0;
                if (g0 + lane <= span.last_segment) {
                    const BinDesc* bd = sp.bins + g0 + lane;
                    first = bd->first_slice; parts = slice_parts(bd->lo, bd->hi); dyn = bd->dyn;
                    off = (uint32_t)(int32_t)(bd->t0 - span.t0);           // segment start relative to the span's origin
                }
                const int nb = min(32, span.last_segment - g0 + 1);
                for (int k = 0; k < nb; ++k) {
                    const uint32_t k_dyn = __shfl_sync(0xFFFFFFFFu, dyn, k);
                    if (!(k_dyn & kBinAny)) continue;
                    const uint32_t k_first = __shfl_sync(0xFFFFFFFFu, first, k), k_parts = __shfl_sync(0xFFFFFFFFu, parts, k);
                    const uint32_t k_off = __shfl_sync(0xFFFFFFFFu, off, k);
                    // the row window of the run table only moves forward; a span that starts before the previous one
                    // ended (nested and overlapping windows) opens it again
                    if (k_first < fp.win_base) fp.rewind(k_first);
                    fp.feed_bin(sp, (uint32_t)s, k_first, k_parts, 0u, k_off, true);
                }
            }
            fp.control(kSegEmit | (s == tp.n_spans - 1 ? kSegDone : 0u), 0u, (uint32_t)s);
        }
        return;
    }

    // ==================================== worker warps ====================================
    const uint32_t* ring = reinterpret_cast<const uint32_t*>(smem_raw + fs.ring);
    const SegDesc* desc = reinterpret_cast<const SegDesc*>(smem_raw + fs.desc);
    auto worker_sync = [] { named_sync(kEvsBar, kEvsWorkers); };
    const uint32_t Pu = (uint32_t)P;
    const uint32_t n4 = (uint32_t)rows * Pu / 4u, p4 = Pu / 4u;
    for (uint32_t i = tid; i < (uint32_t)rows * Pu; i += kEvsWorkers) acc[i] = 0u;      // afterwards the output pass keeps it clean
    for (uint32_t i = tid; i < ((uint32_t)rows * Pu + 3u) / 4u; i += kEvsWorkers) acc_hi[i] = 0u;
    worker_sync();

    const uint32_t row_first = (uint32_t)tid / p4, c4_first = (uint32_t)tid - row_first * p4;
    const uint32_t row_step = kEvsWorkers / p4, c4_step = kEvsWorkers - row_step * p4;
    const float Kf = (float)K;
    uint32_t cur_span = 0xFFFFFFFFu;
    double tw = 1.0, inv_tw = 1.0;
    for (uint32_t seq = 0;; ++seq) {
        const int slot = (int)(seq % kFeedStages);
        mbar_wait(full + slot, (seq / kFeedStages) & 1u);
        const SegDesc d = desc[slot];
        if (d.arg != cur_span) {
            cur_span = d.arg;
            tw = (double)tp.spans[cur_span].tw;
            inv_tw = 1.0 / tw;
        }
        if (d.n_rec) {
            // A thread does not take every 256th record but `stride` consecutive ones (a burst on one pixel would
            // otherwise serialise the atomics of one warp on one cell); stride is odd: no bank conflicts on the ring.
            const uint32_t* recs = ring + slot * kFeedStageRecords;
            const uint32_t stride = ((d.n_rec + kEvsWorkers - 1) / kEvsWorkers) | 1u;
            const int32_t off = (int32_t)d.age_inc;
            for (uint32_t q = 0; q < stride; ++q) {
                const uint32_t idx = (uint32_t)tid * stride + q;
                if (idx >= d.n_rec) break;
                const uint32_t rec = recs[idx];
                if (rec == kNullRecord) continue;
                const uint32_t cell_px = rec & 0x3FFFu;                       // p * P + pixel
                const uint32_t pol = cell_px >= Pu ? 1u : 0u, lp = cell_px - pol * Pu;
                const int32_t rel = off + (int32_t)(rec >> 14);
                // (t - t0) / tw in float64 (:141) then .float() (:23): reciprocal + one Newton step
                const double dd = __int2double_rn(rel < 0 ? 0 : rel);
                const double q0 = dd * inv_tw;
                const float tn = (float)fma(fma(-q0, tw, dd), inv_tw, q0);
                const float ts = Kf * tn;                                        // t* = K * t
                const int c0 = (int)floorf(ts);
#pragma unroll
                for (int dc = 0; dc < 2; ++dc) {                                  // centres c0, c0 + 1 (1..K)
                    const int c = c0 + dc;
                    if (c < 1 || c > K) continue;
                    const float wgt = 1.0f - fabsf((float)c - ts);
                    if (wgt > 0.0f) {
                        const uint32_t cell = (uint32_t)(2 * (c - 1) + (1 - (int)pol)) * Pu + lp;
                        const uint32_t fx = __float2uint_rn(wgt * kEvsUnit);      // wgt <= 1
                        const uint32_t old = atomicAdd(acc + cell, fx);
                        if (old + fx < old) {                                     // lo wrapped: count it
                            const uint32_t sh = (cell & 3u) * 8u;
                            const uint32_t before = atomicAdd(acc_hi + (cell >> 2), 1u << sh);
                            if (((before >> sh) & 0xFFu) == 0xFFu) atomicSub(acc_hi + (cell >> 2), 1u << sh);   // saturate
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + slot);
        if (d.flags & kSegEmit) {
            worker_sync();
            // read, clear and scale the accumulator (:37  / 5 * 255); each row of the tile is contiguous in the tensor
            float* o = tp.out ? tp.out + (int64_t)d.arg * tp.out_stride + pix0 : nullptr;
            uint8_t* o8 = tp.out_u8 ? tp.out_u8 + (int64_t)d.arg * tp.out_u8_stride + pix0 : nullptr;
            auto value = [](uint32_t lo, uint32_t wraps) -> float {
                return div5_mul255(fmaf((float)wraps, kEvsWrap, (float)lo * (1.0f / kEvsUnit)));
            };
            auto to_byte = [](float v) -> uint32_t { return (uint32_t)(int)fminf(v, 255.0f); };   // clamp, then astype(uint8)
            if (tp.vec_out) {
                // thread walks the float4 cells i = tid + k * kEvsWorkers of the [2K][P] tile: (row, c4) advance without a division
                uint32_t row = row_first, c4 = c4_first;
                for (uint32_t i = tid; i < n4; i += kEvsWorkers, row += row_step, c4 += c4_step) {
                    if (c4 >= p4) { c4 -= p4; ++row; }
                    const uint4 lo = reinterpret_cast<uint4*>(acc)[i];
                    const uint32_t wr = acc_hi[i];
                    reinterpret_cast<uint4*>(acc)[i] = make_uint4(0u, 0u, 0u, 0u);
                    if (wr) acc_hi[i] = 0u;
                    if (c4 * 4u < npix) {
                        float4 v;
                        if (wr == 0u) {
                            // no wrap in these four cells (the usual case).  Scaling by 2^-27 commutes with every rounding
                            // of div5_mul255, so it is folded into the last factor: same bits, one multiply less
                            // (packed f32x2 arithmetic: the kernel is bound by instruction issue, each lane rounds like FMUL / FFMA)
                            auto fast2 = [](uint32_t w0, uint32_t w1) -> float2 {
                                const float2 a = make_float2((float)w0, (float)w1), fifth = make_float2(0.2f, 0.2f);
                                const float2 q = __fmul2_rn(a, fifth);
                                const float2 r = __ffma2_rn(q, make_float2(-5.0f, -5.0f), a);
                                return __fmul2_rn(__ffma2_rn(r, fifth, q), make_float2(255.0f / kEvsUnit, 255.0f / kEvsUnit));
                            };
                            const float2 v01 = fast2(lo.x, lo.y), v23 = fast2(lo.z, lo.w);
                            v.x = v01.x; v.y = v01.y; v.z = v23.x; v.w = v23.y;
                        } else {
                            v.x = value(lo.x, wr & 0xFFu); v.y = value(lo.y, (wr >> 8) & 0xFFu);
                            v.z = value(lo.z, (wr >> 16) & 0xFFu); v.w = value(lo.w, wr >> 24);
                        }
                        const uint32_t g = row * HW + c4 * 4u;
                        if (o) __stcs(reinterpret_cast<float4*>(o + g), v);
                        if (o8) *reinterpret_cast<uint32_t*>(o8 + g) = to_byte(v.x) | (to_byte(v.y) << 8) | (to_byte(v.z) << 16) | (to_byte(v.w) << 24);
                    }
                }
            } else {
                for (uint32_t i = tid; i < (uint32_t)rows * Pu; i += kEvsWorkers) {
                    const uint32_t row = i / Pu, lp = i - row * Pu;
                    const uint32_t lo = acc[i];
                    const uint32_t wr = (acc_hi[i >> 2] >> ((i & 3u) * 8u)) & 0xFFu;
                    acc[i] = 0u;
                    if (lp < npix) {
                        const float v = value(lo, wr);
                        if (o) __stcs(o + (int64_t)row * HW + lp, v);
                        if (o8) o8[(int64_t)row * HW + lp] = (uint8_t)to_byte(v);
                    }
                }
                worker_sync();                                     // the byte counters share words
                for (uint32_t i = tid; i < ((uint32_t)rows * Pu + 3u) / 4u; i += kEvsWorkers) acc_hi[i] = 0u;
            }
            worker_sync();                                         // clears are ordered before the next splat
        }
        if (d.flags & kSegDone) break;
    }
}

// f32 [n][C,H,W] -> u8 [n][C,Ht,Wt]: optional nearest resize (legacy index maps), clamp at 255, truncation
// (generate_eventvolume.py:152-160 for a batch of windows).
__global__ void __launch_bounds__(256)
ev_u8_batch_kernel(const float* __restrict__ vol, int64_t vol_stride, int n, int C, int H, int W, int Ht, int Wt,
                   const int32_t* __restrict__ ysrc, const int32_t* __restrict__ xsrc, uint8_t* __restrict__ out) {
    const int64_t per = (int64_t)C * Ht * Wt, total = per * n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t w = i / per, j = i - w * per;
        const int X = (int)(j % Wt);
        const int64_t r = j / Wt;
        const int Y = (int)(r % Ht), ch = (int)(r / Ht);
        const int ys = ysrc ? ysrc[Y] : Y, xs = xsrc ? xsrc[X] : X;
        const float v = __ldcs(vol + w * vol_stride + ((int64_t)ch * H + ys) * W + xs);
        out[i] = (uint8_t)(int)fminf(v, 255.0f);
    }
}

}  // namespace evrep

using namespace evrep;

extern "C" {

static int ev_spans_tile(int H, int W, int K, int& P, int& n_tiles) {
    return choose_tile(H, W, K, ev_slice_smem, 2, P, n_tiles);
}

int64_t evrep_event_volume_spans_scratch_bytes(int64_t n_events, int n_segments, int n_spans, int H, int W, int K) {
    if (n_events < 0 || n_segments < 0 || n_spans < 0 || H <= 0 || W <= 0 || K < 1) return EVREP_ERR_ARG;
    int P, n_tiles;
    int rc = ev_spans_tile(H, W, K, P, n_tiles);
    if (rc) return rc;
    SliceLayout L;
    rc = make_slice_layout(n_events, n_segments, n_segments, H, W, P, n_tiles, L);
    if (rc) return rc;
    return L.total + 256 + (int64_t)sizeof(EvSpanDev) * n_spans + 4096;
}

int evrep_event_volume_spans(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                             const evrep_ev_segment* segments_host, int n_segments,
                             const evrep_ev_span* spans_host, int n_spans, int H, int W, int K,
                             const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                             float* out, int64_t out_stride, uint8_t* out_u8, int64_t out_u8_stride,
                             void* scratch, int64_t scratch_bytes, void* ev_tiles_begin, void* ev_tiles_end,
                             evrep_stream_t stream) {
    if (n_events < 0 || n_segments < 0 || n_spans < 0 || H <= 0 || W <= 0 || K < 1 || !scratch) return EVREP_ERR_ARG;
    if (n_spans == 0) return EVREP_OK;
    if ((!out && !out_u8) || !spans_host || (n_segments > 0 && !segments_host)) return EVREP_ERR_ARG;
    for (int s = 0; s < n_spans; ++s) {
        const evrep_ev_span& sp = spans_host[s];
        if (sp.tw <= 0 || sp.first_segment < 0 || sp.last_segment >= n_segments) return EVREP_ERR_ARG;
        if (sp.tw >= (1ll << 31)) return EVREP_ERR_RANGE;
    }
    cudaStream_t st = as_stream(stream);
    int P, n_tiles;
    int rc = ev_spans_tile(H, W, K, P, n_tiles);
    if (rc) return rc;
    // one "window" of one bin per segment: the slice sort then sorts every segment by tile, offsets relative to its start
    std::vector<evrep_taf_window> wins((size_t)n_segments);
    for (int g = 0; g < n_segments; ++g) {
        wins[g].ev_begin = segments_host[g].ev_begin; wins[g].ev_end = segments_host[g].ev_end;
        wins[g].start_time = segments_host[g].start_time; wins[g].n_bins = 1; wins[g].fresh = 0;
    }
    EvSliceParams tp;
    SliceLayout L;
    static const evrep_taf_window no_windows[1] = {};
    rc = prepare_slices(t, x, y, p, n_events, n_segments ? wins.data() : no_windows, n_segments, (int)kDMax, H, W, P, n_tiles,
                        xmap, ymap, sensor_h, sensor_w, scratch, scratch_bytes, st, tp.sp, L);
    if (rc) return rc;
    const int64_t o_spans = (L.total + 255) / 256 * 256;
    if (scratch_bytes < o_spans + (int64_t)sizeof(EvSpanDev) * n_spans) return EVREP_ERR_SCRATCH;
    std::vector<EvSpanDev> spans((size_t)n_spans + 1);
    for (int s = 0; s < n_spans; ++s) {
        spans[s].first_segment = spans_host[s].first_segment; spans[s].last_segment = spans_host[s].last_segment;
        spans[s].t0 = spans_host[s].t0; spans[s].tw = spans_host[s].tw;
    }
    EvSpanDev* spans_dev = reinterpret_cast<EvSpanDev*>(reinterpret_cast<char*>(scratch) + o_spans);
    rc = upload_words(reinterpret_cast<const uint32_t*>(spans.data()), (int64_t)sizeof(EvSpanDev) / 4 * n_spans,
                      reinterpret_cast<uint32_t*>(spans_dev), st);
    if (rc) return rc;
    tp.spans = spans_dev; tp.seg_start = tp.sp.w_start; tp.n_spans = n_spans;
    tp.out = out; tp.out_stride = out_stride; tp.out_u8 = out_u8; tp.out_u8_stride = out_u8_stride; tp.K = K;
    tp.vec_out = (((int64_t)H * W) % 4 == 0 && P % 4 == 0 &&
                  (!out || (out_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)) &&
                  (!out_u8 || (out_u8_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out_u8) & 3) == 0))) ? 1 : 0;
    const size_t smem = ev_slice_smem(P, K);
    if (ev_tiles_begin) EVREP_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev_tiles_begin), st));
    EVREP_CUDA(cudaFuncSetAttribute(ev_slice_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ev_slice_tile_kernel<<<n_tiles, kEvsThreads, smem, st>>>(tp);
    EVREP_LAUNCH_CHECK();
    if (ev_tiles_end) EVREP_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev_tiles_end), st));
    return EVREP_OK;
}

int evrep_event_volume_u8_batch(const float* volumes, int64_t volume_stride, int n, int C, int H, int W, int Ht, int Wt,
                                const int32_t* ysrc, const int32_t* xsrc, uint8_t* out, evrep_stream_t stream) {
    if (!volumes || !out || n < 0 || C <= 0 || H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0) return EVREP_ERR_ARG;
    if ((!ysrc || !xsrc) && (Ht != H || Wt != W)) return EVREP_ERR_ARG;
    if (n == 0) return EVREP_OK;
    ev_u8_batch_kernel<<<grid_for((int64_t)n * C * Ht * Wt, 4), 256, 0, as_stream(stream)>>>(volumes, volume_stride, n, C, H, W, Ht, Wt,
                                                                                       ysrc, xsrc, out);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

}  // extern "C"
