// Time-surface pair for optical flow: generate_opticalflow.py:72-92.
// volume2 = last timestamp per pixel, volume1 = last timestamp older than end - 50000 (end / start =
// newest / oldest timestamp of the call), both shifted to the window start, scaled by
// 255 / (end - 50000 - start) in float64 and clamped below at zero; polarity is ignored.  "Last
// writer" of the reference's sequential loop = maximum timestamp for time-ordered input.
#include "common.cuh"

namespace evrep {

constexpr int64_t kSurfaceLag = 50000;

__global__ void __launch_bounds__(kBlock)
ts_range_kernel(const uint32_t* __restrict__ t, int64_t n, uint32_t* __restrict__ range) {   // range = {~min, max}
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t v = __ldg(t + i);
        lo = min(lo, v); hi = max(hi, v);
    }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo);
    hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(range + 0, ~lo);          // the scratch is zero on entry: keep the minimum as a maximum of complements
        atomicMax(range + 1, hi);
    }
}

__global__ void __launch_bounds__(kBlock)
ts_scatter_kernel(const uint32_t* __restrict__ t, const uint16_t* __restrict__ x, const uint16_t* __restrict__ y, int64_t n,
                  int H, int W, const uint32_t* __restrict__ range, uint32_t* __restrict__ keys_old, uint32_t* __restrict__ keys_all) {
    const int64_t end = (int64_t)range[1];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t xv = __ldg(x + i), yv = __ldg(y + i), tv = __ldg(t + i);
        if (xv >= (uint32_t)W || yv >= (uint32_t)H) continue;
        const int64_t pix = (int64_t)yv * W + xv;
        atomicMax(keys_all + pix, tv + 1u);                              // :80  (0 = never written)
        if ((int64_t)tv < end - kSurfaceLag) atomicMax(keys_old + pix, tv + 1u);   // :78-79
    }
}

__global__ void __launch_bounds__(kBlock)
ts_finalize_kernel(uint32_t* __restrict__ range, uint32_t* __restrict__ keys_old, uint32_t* __restrict__ keys_all, int64_t cells,
                   double* __restrict__ out_old, double* __restrict__ out_all) {
    const double start = (double)(~range[0]), end = (double)range[1];
    const double den = end - (double)kSurfaceLag - start;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
        const uint32_t ko = keys_old[i], ka = keys_all[i];
        if (ko) keys_old[i] = 0u;
        if (ka) keys_all[i] = 0u;
        const double v1 = ((ko ? (double)(ko - 1u) : 0.0) - start) / den * 255.0;                           // :81,83
        const double v2 = ((ka ? (double)(ka - 1u) : 0.0) - start - (double)kSurfaceLag) / den * 255.0;       // :82,84
        out_old[i] = v1 < 0.0 ? 0.0 : v1;                                                                     // :89-90
        out_all[i] = v2 < 0.0 ? 0.0 : v2;
    }
}

__global__ void ts_reset_range_kernel(uint32_t* range) { range[0] = 0u; range[1] = 0u; }

}  // namespace evrep

using namespace evrep;

extern "C" {

int64_t evrep_timesurface_scratch_bytes(int H, int W) {
    if (H <= 0 || W <= 0) return EVREP_ERR_ARG;
    return (int64_t)sizeof(uint32_t) * (2 * (int64_t)H * W + 4);
}

int evrep_timesurface(const uint32_t* t, const uint16_t* x, const uint16_t* y, int64_t n, int H, int W,
                      void* scratch, double* out_old, double* out_all, evrep_stream_t stream) {
    if (n < 0 || H <= 0 || W <= 0 || !scratch || !out_old || !out_all) return EVREP_ERR_ARG;
    if (n > 0 && (!t || !x || !y)) return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    const int64_t cells = (int64_t)H * W;
    if (n == 0) {                               // :75 `if len(events) > 0`: the zero surfaces come back unchanged
        EVREP_CUDA(cudaMemsetAsync(out_old, 0, sizeof(double) * cells, st));
        EVREP_CUDA(cudaMemsetAsync(out_all, 0, sizeof(double) * cells, st));
        return EVREP_OK;
    }
    uint32_t* range = reinterpret_cast<uint32_t*>(scratch);
    uint32_t* keys_old = range + 4;
    uint32_t* keys_all = keys_old + cells;
    ts_range_kernel<<<grid_for(n, 4), kBlock, 0, st>>>(t, n, range);
    EVREP_LAUNCH_CHECK();
    ts_scatter_kernel<<<grid_for(n), kBlock, 0, st>>>(t, x, y, n, H, W, range, keys_old, keys_all);
    EVREP_LAUNCH_CHECK();
    ts_finalize_kernel<<<grid_for(cells), kBlock, 0, st>>>(range, keys_old, keys_all, cells, out_old, out_all);
    EVREP_LAUNCH_CHECK();
    ts_reset_range_kernel<<<1, 1, 0, st>>>(range);      // leave the scratch zeroed for the next call
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

}  // extern "C"
