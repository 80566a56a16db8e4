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

}  // extern "C"
