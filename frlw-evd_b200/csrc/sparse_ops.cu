// Batched "online" encoders of the reference's plugin surface (data/sparse_ops.py, S1-S6 of
// SURVEY.md 8a) and the event_queue_tensor extension (N1).  Dormant callers in the
// reference (data/fetcher.py:53), kept for API completeness: plain scatter kernels into
// L2-resident accumulators, float32 atomics (sum order differs from the sequential CPU
// reference by rounding only).
#include "common.cuh"

namespace evrep {

// events: float64 [N,5] (b, x, y, t, p).  mode 0: t* = (K t) / window            (sparse_ops.py:12)
//                                        mode 1: t* = ((t - iter) + infer) / window * K   (:15)
//                                        mode 2: t* = ((K - 1) t) / window          (:56)
// weight 1 - |c - t*| >= 0 for centres c = 0..C-1 into acc[(HW b + x + W y), c, 1 - p].
__global__ void __launch_bounds__(kBlock)
sparse_splat_kernel(const double* __restrict__ ev, int64_t n, int B, int H, int W, int C, int mode,
                    float Kf, float window, float iter, float infer, float* __restrict__ acc) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t HW = (int64_t)H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double* r = ev + 5 * i;
        const long long b = (long long)r[0], x = (long long)r[1], y = (long long)r[2], p = (long long)r[4];
        if (b < 0 || b >= B || x < 0 || x >= W || y < 0 || y >= H || p < 0 || p > 1) continue;
        const float t = (float)r[3];
        float ts;
        if (mode == 0) ts = __fdiv_rn(Kf * t, window);
        else if (mode == 1) ts = __fdiv_rn((t - iter) + infer, window) * Kf;
        else ts = __fdiv_rn((Kf - 1.0f) * t, window);
        if (!(ts > -1.0f) || ts >= (float)C) continue;
        const int c0 = (int)floorf(ts);
        float* cell = acc + ((HW * b + x + (int64_t)W * y) * C) * 2 + (1 - p);
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const int c = c0 + d;
            if (c < 0 || c >= C) continue;
            const float w = 1.0f - fabsf((float)c - ts);
            if (w > 0.0f) atomicAdd(cell + 2 * c, w);
        }
    }
}

// [B*HW, C2] pixel-major -> [B, C2, H*W] planar (sparse_ops.py:34, permute(0,3,1,2,4).contiguous()).
__global__ void __launch_bounds__(kBlock)
pixel_major_to_planar_kernel(const float* __restrict__ in, int B, int64_t HW, int C2, float* __restrict__ out) {
    const int64_t total = (int64_t)B * HW * C2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t pix = i % HW;
        const int64_t r = i / HW;
        const int c = (int)(r % C2);
        const int64_t b = r / C2;
        out[i] = in[(b * HW + pix) * C2 + c];
    }
}

// Incremental agile volume (sparse_ops.py:25-32): the newest bin of `past` receives the first fresh
// bin IN PLACE, the oldest bin is dropped, the second fresh bin is appended.
__global__ void __launch_bounds__(kBlock)
agile_shift_kernel(float* __restrict__ past, const float* __restrict__ fresh, int64_t BHW, int K, float* __restrict__ out) {
    const int64_t total = BHW * 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t pix = i >> 1;
        const int q = (int)(i & 1);
        float* pv = past + pix * K * 2 + q;
        const float merged = pv[(K - 1) * 2] + fresh[pix * 4 + q];
        pv[(K - 1) * 2] = merged;
        float* o = out + pix * K * 2 + q;
        for (int k = 0; k + 2 < K; ++k) o[k * 2] = pv[(k + 1) * 2];
        if (K >= 2) o[(K - 2) * 2] = merged;
        o[(K - 1) * 2] = fresh[pix * 4 + 2 + q];
    }
}

// sparse_ops.py:72-85.  events float64 [N,7] (b, x, y, t, c, p, feature) -> out [B, C, H, W, 2].
__global__ void __launch_bounds__(kBlock)
sparse_taf_scatter_kernel(const double* __restrict__ ev, int64_t n, int B, int H, int W, int C, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double* r = ev + 7 * i;
        const long long b = (long long)r[0], x = (long long)r[1], y = (long long)r[2], c = (long long)r[4], p = (long long)r[5];
        if (b < 0 || b >= B || x < 0 || x >= W || y < 0 || y >= H || c < 0 || c >= C || p < 0 || p > 1) continue;
        atomicAdd(out + ((((int64_t)b * C + c) * H + y) * W + x) * 2 + p, (float)r[6]);
    }
}
__global__ void __launch_bounds__(kBlock)
sparse_taf_fix_kernel(float* __restrict__ out, int64_t cells) {          // plane 1: 0 -> -1e8, else +1  (:84)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
        const float v = out[2 * i + 1];
        out[2 * i + 1] = (v == 0.0f) ? -1e8f : v + 1.0f;
    }
}

// sparse_ops.py:88-107: polarity-agnostic occupancy, both output channels identical.
__global__ void __launch_bounds__(kBlock)
event_frame_mark_kernel(const double* __restrict__ ev, int64_t n, int B, int H, int W, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t HW = (int64_t)H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double* r = ev + 5 * i;
        const long long b = (long long)r[0], x = (long long)r[1], y = (long long)r[2];
        if (b < 0 || b >= B || x < 0 || x >= W || y < 0 || y >= H) continue;
        const int64_t pix = x + (int64_t)W * y;
        out[(b * 2 + 0) * HW + pix] = 255.0f;
        out[(b * 2 + 1) * HW + pix] = 255.0f;
    }
}

// sparse_ops.py:109-121: scatter-add feature rows at (b, y, x) -> [B, H, W, C].
__global__ void __launch_bounds__(kBlock)
sparse_to_dense_kernel(const int64_t* __restrict__ loc, const float* __restrict__ feat, int64_t n, int B, int H, int W,
                       int C, float* __restrict__ out) {
    const int64_t total = n * C;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t e = i / C;
        const int c = (int)(i - e * C);
        const int64_t b = loc[3 * e], y = loc[3 * e + 1], x = loc[3 * e + 2];
        if (b < 0 || b >= B || x < 0 || x >= W || y < 0 || y >= H) continue;
        atomicAdd(out + (((int64_t)b * H + y) * W + x) * C + c, feat[i]);
    }
}

// event_queue_tensor.cpp:42-75 as it behaves: every event adds
// 1 - (start[b] + abin (z + 1) - t) / abin (float32) to its (p, b, h, w) cell.
__global__ void __launch_bounds__(kBlock)
queue_accumulate_kernel(const float* __restrict__ ev, int64_t n, int B, int H, int W, const int32_t* __restrict__ start,
                        int abin, float* __restrict__ totals) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t BHW = (int64_t)B * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float* r = ev + 6 * i;
        const int b = (int)r[0], w = (int)r[1], h = (int)r[2], p = (int)r[4], z = (int)r[5];
        if (b < 0 || b >= B || w < 0 || w >= W || h < 0 || h >= H || p < 0 || p > 1) continue;
        const float edge = (float)(start[b] + abin * (z + 1));
        const float v = 1.0f - __fdiv_rn(edge - r[3], (float)abin);
        atomicAdd(totals + BHW * p + ((int64_t)H * W) * b + (int64_t)W * h + w, v);
    }
}
// :79-116: cells with a positive total occupy queue slot Q-1 of plane 0; plane 1 is -1 everywhere.
__global__ void __launch_bounds__(kBlock)
queue_emit_kernel(float* __restrict__ totals, int64_t cells, int Q, double* __restrict__ out) {
    const int64_t total = cells * Q;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t cell = i % cells;
        const int k = (int)(i / cells);
        double v = 0.0;
        if (k == Q - 1) { const float t = totals[cell]; if (t > 0.0f) v = (double)t; }
        out[i] = v;
        out[total + i] = -1.0;
    }
}
__global__ void __launch_bounds__(kBlock)
zero_f32_kernel(float* __restrict__ p, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = 0.0f;
}


// ---- stable row compaction (denseToSparse, the raw-event memory of generate_event_volume_cuda) -------------
// Three small passes: flags + per-block counts, a scan of the block counts, then each block ranks its rows and
// copies the kept ones -- the output keeps the input order, like torch.nonzero / boolean indexing do.
constexpr int kCompactBlock = 256;
constexpr int kCompactPerThread = 8;
constexpr int kCompactRows = kCompactBlock * kCompactPerThread;

// kind 0: a row of `cols` floats is kept when any element is non-zero (|.|-sum != 0, sparse_ops.py:131)
// kind 1: a row of `cols` doubles is kept when column 3 >= thr                          (sparse_ops.py:42)
template <int kKind>
__device__ __forceinline__ bool keep_row(const void* src, int64_t row, int cols, double thr) {
    if (kKind == 0) {
        const float* r = reinterpret_cast<const float*>(src) + row * cols;
        bool any = false;
        for (int c = 0; c < cols; ++c) any |= (r[c] != 0.0f);
        return any;
    }
    return reinterpret_cast<const double*>(src)[row * cols + 3] >= thr;
}

template <int kKind>
__global__ void __launch_bounds__(kCompactBlock)
compact_count_kernel(const void* __restrict__ src, int64_t n, int cols, double thr, uint32_t* __restrict__ block_counts) {
    const int64_t base = (int64_t)blockIdx.x * kCompactRows;
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < kCompactPerThread; ++k) {
        const int64_t row = base + (int64_t)threadIdx.x * kCompactPerThread + k;
        if (row < n && keep_row<kKind>(src, row, cols, thr)) ++mine;
    }
    mine = __reduce_add_sync(0xFFFFFFFFu, mine);
    __shared__ uint32_t warp_sum[kCompactBlock / 32];
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (int w = 0; w < kCompactBlock / 32; ++w) total += warp_sum[w];
        block_counts[blockIdx.x] = total;
    }
}

// exclusive scan of the block counts in place (one CTA); total -> block_counts[n_blocks]
__global__ void __launch_bounds__(1024)
compact_scan_kernel(uint32_t* __restrict__ block_counts, int n_blocks) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n_blocks ? block_counts[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sum[wid] = incl;
        __syncthreads();
        uint32_t before = carry;
        for (int k = 0; k < wid; ++k) before += warp_sum[k];
        if (i < n_blocks) block_counts[i] = before + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_counts[n_blocks] = carry;
}

// copies the kept rows (row_bytes each) to their ranks; optionally writes their unravelled index (spatial dims first,
// batch last: sparse_ops.py:132) as int64 rows of `n_dims` entries.
struct CompactDims { int64_t size[4]; int n; };

template <int kKind>
__global__ void __launch_bounds__(kCompactBlock)
compact_scatter_kernel(const void* __restrict__ src, int64_t n, int cols, double thr, const uint32_t* __restrict__ block_offsets,
                       void* __restrict__ dst, int64_t* __restrict__ locations, CompactDims dims) {
    const int64_t base = (int64_t)blockIdx.x * kCompactRows;
    bool keep[kCompactPerThread];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < kCompactPerThread; ++k) {
        const int64_t row = base + (int64_t)threadIdx.x * kCompactPerThread + k;
        keep[k] = row < n && keep_row<kKind>(src, row, cols, thr);
        mine += keep[k] ? 1u : 0u;
    }
    // exclusive scan of the per-thread counts over the block
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += t;
    }
    __shared__ uint32_t warp_sum[kCompactBlock / 32];
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    uint32_t rank = block_offsets[blockIdx.x] + incl - mine;
    for (int k = 0; k < wid; ++k) rank += warp_sum[k];
    const int elem = kKind == 0 ? 4 : 8;
#pragma unroll
    for (int k = 0; k < kCompactPerThread; ++k) {
        if (!keep[k]) continue;
        const int64_t row = base + (int64_t)threadIdx.x * kCompactPerThread + k;
        if (kKind == 0) {
            const float* a = reinterpret_cast<const float*>(src) + row * cols;
            float* b = reinterpret_cast<float*>(dst) + (int64_t)rank * cols;
            for (int c = 0; c < cols; ++c) b[c] = a[c];
        } else {
            const double* a = reinterpret_cast<const double*>(src) + row * cols;
            double* b = reinterpret_cast<double*>(dst) + (int64_t)rank * cols;
            for (int c = 0; c < cols; ++c) b[c] = a[c];
        }
        if (locations) {
            int64_t idx[4], r = row;
            for (int d = dims.n - 1; d >= 0; --d) { idx[d] = r % dims.size[d]; r /= dims.size[d]; }
            int64_t* loc = locations + (int64_t)rank * dims.n;
            for (int d = 1; d < dims.n; ++d) loc[d - 1] = idx[d];       // spatial indices ...
            loc[dims.n - 1] = idx[0];                                   // ... then the batch index
        }
        ++rank;
    }
    (void)elem;
}

// dst rows = rows of `a` then rows of `b` (the torch.cat of sparse_ops.py:41), all float64 [*, cols]
__global__ void __launch_bounds__(kBlock)
concat_rows_kernel(const double* __restrict__ a, int64_t na, const double* __restrict__ b, int64_t nb, double* __restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < na + nb; i += stride)
        dst[i] = i < na ? a[i] : b[i - na];
}

// ---- online Temporal Active Focus: one bin of a batch of recordings, state carried on the device ----------
// events f64 [N,5] (b, x, y, t, p) of one fetch step; the events with t_lo <= t < t_hi form the bin.  Per sample the
// rule of generate_taf.py:19-58 applies (a sample without events in the bin is left untouched: no ageing).
struct OnlineCell { uint32_t n; float s; };

__global__ void __launch_bounds__(kBlock)
taf_online_scatter_kernel(const double* __restrict__ ev, int64_t n, double t_lo, double t_hi, double t_span, int B, int H, int W,
                          uint32_t* __restrict__ flags, OnlineCell* __restrict__ cells) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t HW = (int64_t)H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double* r = ev + 5 * i;
        const double t = r[3];
        if (!(t >= t_lo) || !(t < t_hi)) continue;
        const long long b = (long long)r[0], x = (long long)r[1], y = (long long)r[2], p = (long long)r[4];
        if (b < 0 || b >= B || x < 0 || x >= W || y < 0 || y >= H || p < 0 || p > 1) continue;
        OnlineCell* c = cells + 2 * (b * HW + y * W + x) + p;
        atomicAdd(&c->n, 1u);
        atomicAdd(&c->s, (float)((t - t_lo) / t_span) - 1.0f);      // driver :213-215, then taf_cuda's t - 1 (:25)
        flags[b] = 1u;
    }
}

template <int K>
__global__ void __launch_bounds__(kBlock)
taf_online_update_kernel(uint32_t* __restrict__ flags, OnlineCell* __restrict__ cells, int B, int64_t HW,
                         float* __restrict__ state, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)B * HW; i += stride) {
        const int64_t b = i / HW, pix = i - b * HW;
        const bool any = flags[b] != 0u;
        uint4 c = *reinterpret_cast<uint4*>(cells + 2 * i);         // {n0, s0, n1, s1}
        if (c.x | c.z) *reinterpret_cast<uint4*>(cells + 2 * i) = make_uint4(0, 0, 0, 0);
        const uint32_t nn[2] = {c.x, c.z};
        const float ss[2] = {__uint_as_float(c.y), __uint_as_float(c.w)};
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            float v[K];
            float* cell = state + (i * 2 + p) * K;
#pragma unroll
            for (int k = 0; k < K; ++k) v[k] = cell[k];
            if (any) {
                if (nn[p]) {
                    const float mean = __fdiv_rn(ss[p], (float)nn[p] + 1e-8f);
#pragma unroll
                    for (int k = 0; k + 1 < K; ++k) v[k] = v[k + 1] - 1.0f;
                    v[K - 1] = mean;
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) v[k] -= 1.0f;
                }
#pragma unroll
                for (int k = 0; k < K; ++k) cell[k] = v[k];
            }
            if (out) {
#pragma unroll
                for (int k = 0; k < K; ++k) out[((b * K + k) * 2 + p) * HW + pix] = v[k];     // [B, 2K, H, W], channel 2k + p
            }
        }
    }
}

__global__ void clear_words_kernel(uint32_t* __restrict__ p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0u;
}

}  // namespace evrep

using namespace evrep;

extern "C" {

int evrep_sparse_splat(const double* events, int64_t n, int B, int H, int W, int C, int mode, float K, float window,
                       float iter, float infer, float* acc, evrep_stream_t stream) {
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || mode < 0 || mode > 2 || !acc || n < 0 || (n > 0 && !events) || window == 0.0f)
        return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    EVREP_CUDA(cudaMemsetAsync(acc, 0, sizeof(float) * 2 * C * (size_t)B * H * W, st));
    if (n > 0) {
        sparse_splat_kernel<<<grid_for(n), kBlock, 0, st>>>(events, n, B, H, W, C, mode, K, window, iter, infer, acc);
        EVREP_LAUNCH_CHECK();
    }
    return EVREP_OK;
}

int evrep_pixel_major_to_planar(const float* in, int B, int64_t HW, int C2, float* out, evrep_stream_t stream) {
    if (!in || !out || B <= 0 || HW <= 0 || C2 <= 0) return EVREP_ERR_ARG;
    pixel_major_to_planar_kernel<<<grid_for((int64_t)B * HW * C2), kBlock, 0, as_stream(stream)>>>(in, B, HW, C2, out);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

int evrep_sparse_agile_shift(float* past, const float* fresh, int64_t BHW, int K, float* out_state, evrep_stream_t stream) {
    if (!past || !fresh || !out_state || BHW <= 0 || K < 1) return EVREP_ERR_ARG;
    agile_shift_kernel<<<grid_for(BHW * 2), kBlock, 0, as_stream(stream)>>>(past, fresh, BHW, K, out_state);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

int evrep_sparse_taf(const double* events, int64_t n, int B, int H, int W, int C, float* out, evrep_stream_t stream) {
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || !out || n < 0 || (n > 0 && !events)) return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    const int64_t cells = (int64_t)B * C * H * W;
    EVREP_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * 2 * cells, st));
    if (n > 0) {
        sparse_taf_scatter_kernel<<<grid_for(n), kBlock, 0, st>>>(events, n, B, H, W, C, out);
        EVREP_LAUNCH_CHECK();
    }
    sparse_taf_fix_kernel<<<grid_for(cells), kBlock, 0, st>>>(out, cells);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

int evrep_sparse_event_frame(const double* events, int64_t n, int B, int H, int W, float* out, evrep_stream_t stream) {
    if (B <= 0 || H <= 0 || W <= 0 || !out || n < 0 || (n > 0 && !events)) return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    EVREP_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * 2 * (size_t)B * H * W, st));
    if (n > 0) {
        event_frame_mark_kernel<<<grid_for(n), kBlock, 0, st>>>(events, n, B, H, W, out);
        EVREP_LAUNCH_CHECK();
    }
    return EVREP_OK;
}

int evrep_sparse_to_dense(const int64_t* locations, const float* features, int64_t n, int B, int H, int W, int C,
                          float* out, evrep_stream_t stream) {
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || !out || n < 0 || (n > 0 && (!locations || !features))) return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    EVREP_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)B * H * W * C, st));
    if (n > 0) {
        sparse_to_dense_kernel<<<grid_for(n * C), kBlock, 0, st>>>(locations, features, n, B, H, W, C, out);
        EVREP_LAUNCH_CHECK();
    }
    return EVREP_OK;
}

int evrep_event_queue_tensor(const float* events, int64_t n, int Q, int B, int H, int W, const int32_t* start_times,
                             int abin, float* totals, double* out, evrep_stream_t stream) {
    if (Q <= 0 || B <= 0 || H <= 0 || W <= 0 || abin == 0 || !start_times || !totals || !out || n < 0 || (n > 0 && !events))
        return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    const int64_t cells = (int64_t)2 * B * H * W;
    zero_f32_kernel<<<grid_for(cells), kBlock, 0, st>>>(totals, cells);
    EVREP_LAUNCH_CHECK();
    if (n > 0) {
        queue_accumulate_kernel<<<grid_for(n), kBlock, 0, st>>>(events, n, B, H, W, start_times, abin, totals);
        EVREP_LAUNCH_CHECK();
    }
    queue_emit_kernel<<<grid_for(cells * Q), kBlock, 0, st>>>(totals, cells, Q, out);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}


// Stable compaction front ends.  `block_scratch`: (n / 2048 + 2) u32.  The number of kept rows ends up in
// block_scratch[n_blocks] on the device; the caller reads it back to size its views (torch.nonzero does the same).
static int compact_blocks(int64_t n) { return (int)((n + kCompactRows - 1) / kCompactRows); }

int64_t evrep_compact_scratch_bytes(int64_t n_rows) {
    if (n_rows < 0) return EVREP_ERR_ARG;
    return 4ll * (compact_blocks(n_rows) + 2);
}

int evrep_dense_to_sparse(const float* dense, const int64_t* sizes, int n_dims, int C, int64_t* locations, float* features,
                          uint32_t* block_scratch, uint32_t** count_dev, evrep_stream_t stream) {
    if (!dense || !sizes || n_dims < 1 || n_dims > 4 || C <= 0 || !locations || !features || !block_scratch) return EVREP_ERR_ARG;
    CompactDims dims;
    dims.n = n_dims;
    int64_t n = 1;
    for (int d = 0; d < 4; ++d) { dims.size[d] = d < n_dims ? sizes[d] : 1; if (d < n_dims) n *= sizes[d]; }
    const int nb = compact_blocks(n);
    cudaStream_t st = as_stream(stream);
    if (count_dev) *count_dev = block_scratch + nb;
    if (nb == 0) { clear_words_kernel<<<1, 32, 0, st>>>(block_scratch, 1); EVREP_LAUNCH_CHECK(); return EVREP_OK; }
    compact_count_kernel<0><<<nb, kCompactBlock, 0, st>>>(dense, n, C, 0.0, block_scratch);
    EVREP_LAUNCH_CHECK();
    compact_scan_kernel<<<1, 1024, 0, st>>>(block_scratch, nb);
    EVREP_LAUNCH_CHECK();
    compact_scatter_kernel<0><<<nb, kCompactBlock, 0, st>>>(dense, n, C, 0.0, block_scratch, features, locations, dims);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

int evrep_event_memory_update(const double* memory, int64_t n_memory, const double* events, int64_t n_events, double keep_from,
                              double* merged, double* memory_out, uint32_t* block_scratch, uint32_t** count_dev,
                              evrep_stream_t stream) {
    if (n_memory < 0 || n_events < 0 || !merged || !memory_out || !block_scratch) return EVREP_ERR_ARG;
    if ((n_memory > 0 && !memory) || (n_events > 0 && !events)) return EVREP_ERR_ARG;
    const int64_t n = n_memory + n_events;
    const int nb = compact_blocks(n);
    cudaStream_t st = as_stream(stream);
    if (count_dev) *count_dev = block_scratch + nb;
    if (nb == 0) { clear_words_kernel<<<1, 32, 0, st>>>(block_scratch, 1); EVREP_LAUNCH_CHECK(); return EVREP_OK; }
    concat_rows_kernel<<<grid_for(5 * n), kBlock, 0, st>>>(memory, 5 * n_memory, events, 5 * n_events, merged);
    EVREP_LAUNCH_CHECK();
    compact_count_kernel<1><<<nb, kCompactBlock, 0, st>>>(merged, n, 5, keep_from, block_scratch);
    EVREP_LAUNCH_CHECK();
    compact_scan_kernel<<<1, 1024, 0, st>>>(block_scratch, nb);
    EVREP_LAUNCH_CHECK();
    CompactDims dims; dims.n = 0; for (int d = 0; d < 4; ++d) dims.size[d] = 1;
    compact_scatter_kernel<1><<<nb, kCompactBlock, 0, st>>>(merged, n, 5, keep_from, block_scratch, memory_out, nullptr, dims);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

int64_t evrep_taf_online_scratch_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return EVREP_ERR_ARG;
    return (int64_t)(B + 3) / 4 * 16 + (int64_t)B * H * W * 2 * (int64_t)sizeof(OnlineCell);
}

int evrep_taf_online_bin(const double* events, int64_t n, double t_lo, double t_hi, double t_span, int B, int H, int W, int K,
                         float* state, float* out, void* scratch, evrep_stream_t stream) {
    if (n < 0 || (n > 0 && !events) || B <= 0 || H <= 0 || W <= 0 || !state || !scratch || !(t_span > 0.0)) return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    uint32_t* flags = reinterpret_cast<uint32_t*>(scratch);
    OnlineCell* cells = reinterpret_cast<OnlineCell*>(reinterpret_cast<char*>(scratch) + (int64_t)(B + 3) / 4 * 16);
    const int64_t HW = (int64_t)H * W;
    if (n > 0) {
        taf_online_scatter_kernel<<<grid_for(n), kBlock, 0, st>>>(events, n, t_lo, t_hi, t_span, B, H, W, flags, cells);
        EVREP_LAUNCH_CHECK();
    }
    const int grid = grid_for((int64_t)B * HW);
#define EVREP_ONLINE_K(KK) case KK: taf_online_update_kernel<KK><<<grid, kBlock, 0, st>>>(flags, cells, B, HW, state, out); break;
    switch (K) {
        EVREP_ONLINE_K(1) EVREP_ONLINE_K(2) EVREP_ONLINE_K(3) EVREP_ONLINE_K(4) EVREP_ONLINE_K(5) EVREP_ONLINE_K(6)
        EVREP_ONLINE_K(7) EVREP_ONLINE_K(8) EVREP_ONLINE_K(9) EVREP_ONLINE_K(10) EVREP_ONLINE_K(11) EVREP_ONLINE_K(12)
        default: return EVREP_ERR_ARG;
    }
#undef EVREP_ONLINE_K
    EVREP_LAUNCH_CHECK();
    clear_words_kernel<<<(B + 255) / 256, 256, 0, st>>>(flags, B);     // the cells were cleared by the update pass
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

}  // extern "C"
