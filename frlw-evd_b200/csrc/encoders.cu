// Per-window encoders: .dat decode, Event Count Image, Surface of Active Events, Event
// Volume and the one-bin TAF update.  All are scatter kernels into L2-resident
// accumulators (a whole 512x640 frame of counters is 2.6 MB against 126 MB of L2)
// followed by a fused dense epilogue.  HBM-bound integer/float work: no tensor cores.
#include "common.cuh"

namespace evrep {

// ------------------------------------------------------------------ D1: decode
// src/io/dat_events_tools.py:92-100.  4 records (32 B) per thread, 128-bit loads.
__global__ void __launch_bounds__(kBlock)
decode_dat_kernel(const uint2* __restrict__ rec, int64_t n, uint32_t* __restrict__ t,
                  uint16_t* __restrict__ x, uint16_t* __restrict__ y, uint8_t* __restrict__ p, bool vec_ok) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = n >> 2;
    if (vec_ok) {
        const uint4* rec4 = reinterpret_cast<const uint4*>(rec);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            uint4 a = __ldcs(rec4 + 2 * i), b = __ldcs(rec4 + 2 * i + 1);   // streamed once
            uint32_t tt[4] = {a.x, a.z, b.x, b.z};
            uint32_t ww[4] = {a.y, a.w, b.y, b.w};
            uint16_t xx[4], yy[4];
            uint8_t pp[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                xx[k] = (uint16_t)(ww[k] & 0x3FFFu);
                yy[k] = (uint16_t)((ww[k] & 0x0FFFC000u) >> 14);
                pp[k] = (uint8_t)((ww[k] & 0x10000000u) >> 28);
            }
            reinterpret_cast<uint4*>(t)[i] = make_uint4(tt[0], tt[1], tt[2], tt[3]);
            reinterpret_cast<uint2*>(x)[i] = make_uint2(xx[0] | ((uint32_t)xx[1] << 16), xx[2] | ((uint32_t)xx[3] << 16));
            reinterpret_cast<uint2*>(y)[i] = make_uint2(yy[0] | ((uint32_t)yy[1] << 16), yy[2] | ((uint32_t)yy[3] << 16));
            reinterpret_cast<uint32_t*>(p)[i] = pp[0] | (pp[1] << 8) | (pp[2] << 16) | ((uint32_t)pp[3] << 24);
        }
    }
    const int64_t done = vec_ok ? (n4 << 2) : 0;
    for (int64_t i = done + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint2 r = rec[i];
        t[i] = r.x;
        x[i] = (uint16_t)(r.y & 0x3FFFu);
        y[i] = (uint16_t)((r.y & 0x0FFFC000u) >> 14);
        p[i] = (uint8_t)((r.y & 0x10000000u) >> 28);
    }
}

__global__ void __launch_bounds__(kBlock)
soa_to_aos64_kernel(SoA ev, int64_t n, double* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double2* o = reinterpret_cast<double2*>(out + 4 * i);
        o[0] = make_double2((double)ev.x[i], (double)ev.y[i]);
        o[1] = make_double2((double)ev.t[i], (double)ev.p[i]);
    }
}

// ------------------------------------------------------------ E1: count image
// generate_eventcountimage.py:32 -- one RED.ADD.U32 per event into the planar [2,H,W]
// counter frame (L2 resident).  4 events per thread per trip for memory-level parallelism.
template <class Loader>
__global__ void __launch_bounds__(kBlock)
count_accumulate_kernel(Loader ev, int64_t n, int H, int W, uint32_t* __restrict__ counts) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t HW = (int64_t)H * W;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n; base += 4 * stride) {
        Event e[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int64_t i = base + k * stride;
            e[k].ok = false;
            if (i < n) e[k] = ev.load(i, H, W);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (e[k].ok) atomicAdd(counts + e[k].p * HW + (int64_t)e[k].y * W + e[k].x, 1u);
    }
}

// value of a cell as a function of its count: float32 running sum of 0.05, clamped, x255
__constant__ float c_count_lut[33];

__global__ void __launch_bounds__(kBlock)
count_finalize_kernel(uint32_t* __restrict__ counts, int64_t cells, float* __restrict__ out, int reset) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
        uint32_t c = counts[i];
        out[i] = c_count_lut[c < 32u ? c : 32u];
        if (reset && c) counts[i] = 0u;
    }
}

// Count LUT + nearest resize + uint8 truncation in one pass (driver epilogue,
// generate_eventcountimage.py:164,180): out[c, Y, X] = u8(LUT[count[c, ysrc[Y], xsrc[X]]]).
__global__ void __launch_bounds__(kBlock)
count_finalize_u8_kernel(const uint32_t* __restrict__ counts, int H, int W, int Ht, int Wt,
                         const int32_t* __restrict__ ysrc, const int32_t* __restrict__ xsrc, uint8_t* __restrict__ out) {
    const int64_t total = (int64_t)2 * Ht * Wt;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int X = (int)(i % Wt);
        const int64_t r = i / Wt;
        const int Y = (int)(r % Ht), c = (int)(r / Ht);
        const int ys = ysrc ? ysrc[Y] : Y, xs = xsrc ? xsrc[X] : X;
        const uint32_t n = counts[((int64_t)c * H + ys) * W + xs];
        out[i] = (uint8_t)(int)c_count_lut[n < 32u ? n : 32u];
    }
}

// The same epilogue for the count frames of evrep_count_stream (uint8 counts saturated at 255),
// all windows at once: out[w, c, Y, X] = u8(LUT[min(frames[w, c, ysrc[Y], xsrc[X]], 32)]).
template <int kVec>      // output pixels per thread: 4 (one 32-bit store) when Wt % 4 == 0, else 1
__global__ void __launch_bounds__(kBlock)
count_lut_u8_batch_kernel(const uint8_t* __restrict__ frames, int64_t frame_stride, int64_t n_windows, int H, int W,
                          int Ht, int Wt, const int32_t* __restrict__ ysrc, const int32_t* __restrict__ xsrc,
                          uint8_t* __restrict__ out) {
    const int64_t per = (int64_t)2 * Ht * Wt / kVec, total = n_windows * per;
    const int wq = Wt / kVec;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t w = i / per, j = i - w * per;
        const int X0 = (int)(j % wq) * kVec;
        const int64_t r = j / wq;
        const int Y = (int)(r % Ht), c = (int)(r / Ht);
        const int ys = ysrc ? ysrc[Y] : Y;
        const uint8_t* row = frames + w * frame_stride + ((int64_t)c * H + ys) * W;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < kVec; ++k) {
            const uint32_t n = row[xsrc ? xsrc[X0 + k] : X0 + k];
            packed |= (uint32_t)(uint8_t)(int)c_count_lut[n < 32u ? n : 32u] << (8 * k);
        }
        if (kVec == 4) reinterpret_cast<uint32_t*>(out)[i] = packed;
        else out[i] = (uint8_t)packed;
    }
}

// ------------------------------------------------------------------ A1: SAE
// Order-preserving float <-> u32 key (0 is reserved for "no event").
__device__ __forceinline__ uint32_t float_key(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

template <class Loader> struct TimeF32;
template <> struct TimeF32<SoA> {
    static __device__ __forceinline__ float get(const SoA& ev, int64_t i) { return (float)ev.time_us(i); }
};
template <> struct TimeF32<Aos64> {
    static __device__ __forceinline__ float get(const Aos64& ev, int64_t i) { return (float)ev.time_f64(i); }
};

// generate_surfaceofactiveevents.py:49 -- scatter-max of float32(t) (RED.MAX.U32 on keys).
template <class Loader>
__global__ void __launch_bounds__(kBlock)
sae_scatter_kernel(Loader ev, int64_t n, int H, int W, uint32_t* __restrict__ keys) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t HW = (int64_t)H * W;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < n; base += 4 * stride) {
        Event e[4];
        float tf[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int64_t i = base + k * stride;
            e[k].ok = false;
            if (i < n) { e[k] = ev.load(i, H, W); tf[k] = TimeF32<Loader>::get(ev, i); }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (e[k].ok) atomicMax(keys + e[k].p * HW + (int64_t)e[k].y * W + e[k].x, float_key(tf[k]));
    }
}

struct Lambdas { float v[8]; };
__device__ __forceinline__ uint8_t to_u8_trunc(float v) { return (uint8_t)(int)v; }   // numpy astype(uint8)

// :48 init, :51-54 max-merge with memory + state write, :55-63 the L decays, fused.
__global__ void __launch_bounds__(kBlock)
sae_finalize_kernel(uint32_t* __restrict__ keys, int64_t cells, float init, float now_f32,
                    Lambdas lam, int L, const float* __restrict__ mem_in, float* __restrict__ mem_out,
                    float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
        uint32_t k = keys[i];
        float latest = init;
        if (k) { latest = key_float(k); keys[i] = 0u; }
        if (mem_in) { float m = mem_in[i]; latest = latest > m ? latest : m; }
        mem_out[i] = latest;
        float rel = latest - now_f32;
        for (int l = 0; l < L; ++l) out[(int64_t)l * cells + i] = expf(lam.v[l] * rel) * 255.0f;
    }
}

// SAE epilogue for the drivers (generate_surfaceofactiveevents.py:186-204): the state is
// updated at grid resolution and the L decays are written resized + truncated to uint8,
// out[l, p, Y, X].  Output cells recompute `latest` from (keys, memory_in), so the two halves
// of the index space do not depend on each other; `keys` is cleared by the caller afterwards.
__global__ void __launch_bounds__(kBlock)
sae_finalize_u8_kernel(const uint32_t* __restrict__ keys, int H, int W, int Ht, int Wt,
                       const int32_t* __restrict__ ysrc, const int32_t* __restrict__ xsrc, float init, float now_f32,
                       Lambdas lam, int L, const float* __restrict__ mem_in, float* __restrict__ mem_out,
                       uint8_t* __restrict__ out) {
    const int64_t cells = (int64_t)2 * H * W, tcells = (int64_t)2 * Ht * Wt;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells + tcells; i += stride) {
        if (i < cells) {
            const uint32_t k = keys[i];
            float latest = k ? key_float(k) : init;
            if (mem_in) { const float m = mem_in[i]; latest = latest > m ? latest : m; }
            mem_out[i] = latest;
        } else {
            const int64_t j = i - cells;
            const int X = (int)(j % Wt);
            const int64_t r = j / Wt;
            const int Y = (int)(r % Ht), p = (int)(r / Ht);
            const int ys = ysrc ? ysrc[Y] : Y, xs = xsrc ? xsrc[X] : X;
            const int64_t src = ((int64_t)p * H + ys) * W + xs;
            const uint32_t k = keys[src];
            float latest = k ? key_float(k) : init;
            if (mem_in) { const float m = mem_in[src]; latest = latest > m ? latest : m; }
            const float rel = latest - now_f32;
            for (int l = 0; l < L; ++l) out[(int64_t)l * tcells + j] = to_u8_trunc(expf(lam.v[l] * rel) * 255.0f);
        }
    }
}

// ----------------------------------------------------------- V1: event volume
template <class Loader> struct TimeNorm;
template <> struct TimeNorm<SoA> {
    int64_t t0; double tw;
    __device__ __forceinline__ float get(const SoA& ev, int64_t i) const {
        return (float)((double)((int64_t)ev.time_us(i) - t0) / tw);
    }
};
template <> struct TimeNorm<Aos64> {
    __device__ __forceinline__ float get(const Aos64& ev, int64_t i) const { return (float)ev.time_f64(i); }
};

// generate_eventvolume.py:23-32 -- at most two of the K centres get a non-negative weight.
template <class Loader>
__global__ void __launch_bounds__(kBlock)
ev_splat_kernel(Loader ev, TimeNorm<Loader> tn, int64_t n, int H, int W, int K, float* __restrict__ acc) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t HW = (int64_t)H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Event e = ev.load(i, H, W);
        if (!e.ok) continue;
        float ts = (float)K * tn.get(ev, i);
        if (!(ts >= 0.0f) || ts > (float)(K + 1)) continue;
        int c0 = (int)floorf(ts);
        int64_t pix = (int64_t)e.y * W + e.x;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            int c = c0 + d;
            if (c < 1 || c > K) continue;
            float w = 1.0f - fabsf((float)c - ts);
            if (w > 0.0f) atomicAdd(acc + (int64_t)(2 * (c - 1) + (1 - e.p)) * HW + pix, w);
        }
    }
}

__global__ void __launch_bounds__(kBlock)
ev_scale_kernel(float4* __restrict__ acc4, int64_t n4, float* __restrict__ acc, int64_t n) {   // :37  / 5 * 255
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = acc4[i];
        v.x = div5_mul255(v.x); v.y = div5_mul255(v.y); v.z = div5_mul255(v.z); v.w = div5_mul255(v.w);
        acc4[i] = v;
    }
    for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        acc[i] = div5_mul255(acc[i]);
}

// ------------------------------------------------------------- T1: TAF, one bin
// scratch layout: [0] u32 "any valid event" flag (padded to 16 B), then per cell
// (pixel-major, polarity minor -- the reference's index p + 2x + 2Wy) {u32 n, f32 sum}.
struct TafCell { uint32_t n; float s; };

template <class Loader> struct TafTime;
template <> struct TafTime<SoA> {
    int64_t t_min; double t_span;
    __device__ __forceinline__ float get(const SoA& ev, int64_t i) const {
        return (float)((double)((int64_t)ev.time_us(i) - t_min) / t_span);
    }
};
template <> struct TafTime<Aos64> {
    __device__ __forceinline__ float get(const Aos64& ev, int64_t i) const { return (float)ev.time_f64(i); }
};

template <class Loader>
__global__ void __launch_bounds__(kBlock)
taf_scatter_kernel(Loader ev, TafTime<Loader> tt, int64_t n, int H, int W,
                   uint32_t* __restrict__ flag, TafCell* __restrict__ cells) {   // generate_taf.py:23-26
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    bool any = false;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Event e = ev.load(i, H, W);
        if (!e.ok) continue;
        TafCell* c = cells + 2 * ((int64_t)e.y * W + e.x) + e.p;
        atomicAdd(&c->n, 1u);
        atomicAdd(&c->s, tt.get(ev, i) - 1.0f);
        any = true;
    }
    if (__syncthreads_or(any) && threadIdx.x == 0) *flag = 1u;
}

// generate_taf.py:27-55.  One thread per pixel (both polarities): 2K floats of state in,
// 2K out, and -- when `out` is given -- the [2K,H,W] view, each channel store coalesced.
template <int K>
__global__ void __launch_bounds__(kBlock)
taf_update_kernel(const uint32_t* __restrict__ flag, TafCell* __restrict__ cells, int64_t HW,
                  const float* __restrict__ state_in, float* __restrict__ state_out, float* __restrict__ out) {
    const bool any = (*flag != 0u);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += stride) {
        uint4 c = *reinterpret_cast<uint4*>(cells + 2 * pix);       // {n0, s0, n1, s1}
        if (c.x | c.z) *reinterpret_cast<uint4*>(cells + 2 * pix) = make_uint4(0, 0, 0, 0);
        const uint32_t nn[2] = {c.x, c.z};
        const float ss[2] = {__uint_as_float(c.y), __uint_as_float(c.w)};
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            float v[K];
            const float* src = state_in + (pix * 2 + p) * K;
#pragma unroll
            for (int k = 0; k < K; ++k) v[k] = src[k];
            if (any) {
                if (nn[p]) {
                    float mean = __fdiv_rn(ss[p], (float)nn[p] + 1e-8f);
#pragma unroll
                    for (int k = 0; k + 1 < K; ++k) v[k] = v[k + 1] - 1.0f;
                    v[K - 1] = mean;
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) v[k] -= 1.0f;
                }
            }
            float* dst = state_out + (pix * 2 + p) * K;
            if (any || dst != src) {
#pragma unroll
                for (int k = 0; k < K; ++k) dst[k] = v[k];
            }
            if (out) {
#pragma unroll
                for (int k = 0; k < K; ++k) out[(int64_t)(2 * k + p) * HW + pix] = v[k];
            }
        }
    }
}

// ------------------------------------------------------------------ epilogues
__global__ void __launch_bounds__(kBlock)
nearest_resize_kernel(const float* __restrict__ in, int C, int H, int W, int Ht, int Wt,
                      const int32_t* __restrict__ ysrc, const int32_t* __restrict__ xsrc, float* __restrict__ out) {
    const int64_t total = (int64_t)C * Ht * Wt;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        int X = (int)(i % Wt);
        int64_t r = i / Wt;
        int Y = (int)(r % Ht);
        int c = (int)(r / Ht);
        out[i] = in[((int64_t)c * H + ysrc[Y]) * W + xsrc[X]];
    }
}

__device__ __forceinline__ uint8_t to_u8(float v, int clamp255) {
    if (clamp255 && v > 255.0f) v = 255.0f;
    return (uint8_t)(int)v;            // numpy astype(uint8): truncation toward zero
}

__global__ void __launch_bounds__(kBlock)
quantize_u8_kernel(const float* __restrict__ in, int64_t n, int clamp255, uint8_t* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = to_u8(in[i], clamp255);
}

__device__ __forceinline__ float leaky(float v) {           // generate_taf.py:69-76
    float r = 1.0f - __fdiv_rn(log1pf(-v), 8.7f);
    return (r < 0.0f ? 0.0f : r) * 255.0f;
}

__global__ void __launch_bounds__(kBlock)
leaky_kernel(const float* __restrict__ in, int64_t n, float* __restrict__ out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = leaky(in[i]);
}

// generate_taf.py:226-235: [2K,H,W] (channel 2k+p) -> u8 [K,2,Ht,Wt], slot axis flipped.
// `n` windows at once: window w reads vol + w * vol_stride and writes out + w * 2K*Ht*Wt.
// blockIdx.y walks the n * 2K output planes, blockIdx.x the pixels of a plane: no 64-bit division per element
// (8 windows of 512 x 640, 168 MB read + 42 MB written: 0.094 ms).
__global__ void __launch_bounds__(kBlock)
taf_leaky_u8_kernel(const float* __restrict__ vol, int64_t vol_stride, int n, int K, int H, int W, int Ht, int Wt,
                    const int32_t* __restrict__ ysrc, const int32_t* __restrict__ xsrc, uint8_t* __restrict__ out) {
    const uint32_t plane = (uint32_t)Ht * (uint32_t)Wt;
    const int planes = n * 2 * K;
    const bool vec = !ysrc && !xsrc && (plane % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(vol) & 15) == 0) && (vol_stride % 4 == 0);
    for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
        const int w = pl / (2 * K), ch = pl - w * 2 * K;             // destination channel 2 * slot + p
        const int k = K - 1 - (ch >> 1), p = ch & 1;
        const float* src = vol + (int64_t)w * vol_stride + (int64_t)(2 * k + p) * H * W;
        uint8_t* dst = out + (int64_t)pl * plane;
        if (vec) {      // same-size grid: 4 pixels per thread, 128-bit loads, 32-bit stores
            for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < plane / 4; q += gridDim.x * blockDim.x) {
                const float4 v = __ldcs(reinterpret_cast<const float4*>(src) + q);
                reinterpret_cast<uint32_t*>(dst)[q] = to_u8(leaky(v.x), 0) | (to_u8(leaky(v.y), 0) << 8) |
                                                      (to_u8(leaky(v.z), 0) << 16) | ((uint32_t)to_u8(leaky(v.w), 0) << 24);
            }
        } else {
            for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += gridDim.x * blockDim.x) {
                const uint32_t Y = i / (uint32_t)Wt, X = i - Y * (uint32_t)Wt;
                const int ys = ysrc ? ysrc[Y] : (int)Y, xs = xsrc ? xsrc[X] : (int)X;
                dst[i] = to_u8(leaky(src[(int64_t)ys * W + xs]), 0);
            }
        }
    }
}

static bool g_lut_ready[64] = {};

static int ensure_count_lut() {
    int dev = 0;
    EVREP_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && g_lut_ready[dev]) return EVREP_OK;
    float lut[33];
    float s = 0.0f;
    lut[0] = 0.0f;
    for (int n = 1; n <= 32; ++n) {
        volatile float next = s + 0.05f;            // float32 running sum, no contraction
        s = next;
        lut[n] = (s > 1.0f ? 1.0f : s) * 255.0f;
    }
    EVREP_CUDA(cudaMemcpyToSymbol(c_count_lut, lut, sizeof(lut)));
    if (dev < 64) g_lut_ready[dev] = true;
    return EVREP_OK;
}

template <class Loader>
static int count_accumulate(Loader ev, int64_t n, int H, int W, uint32_t* counts, cudaStream_t st) {
    if (n <= 0) return EVREP_OK;
    count_accumulate_kernel<Loader><<<grid_for(n, 4), kBlock, 0, st>>>(ev, n, H, W, counts);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

template <class Loader>
static int sae_run(Loader ev, int64_t n, int H, int W, float init, float now_f32, const float* lambdas_host,
                   int L, const float* mem_in, float* mem_out, uint32_t* keys, float* out, cudaStream_t st) {
    if (H <= 0 || W <= 0 || L < 1 || L > 8 || !lambdas_host || !mem_out || !keys || !out || n < 0) return EVREP_ERR_ARG;
    if (n > 0) {
        sae_scatter_kernel<Loader><<<grid_for(n, 4), kBlock, 0, st>>>(ev, n, H, W, keys);
        EVREP_LAUNCH_CHECK();
    }
    Lambdas lam;
    for (int l = 0; l < 8; ++l) lam.v[l] = l < L ? lambdas_host[l] : 0.0f;
    const int64_t cells = (int64_t)2 * H * W;
    sae_finalize_kernel<<<grid_for(cells), kBlock, 0, st>>>(keys, cells, init, now_f32, lam, L, mem_in, mem_out, out);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

template <class Loader>
static int ev_run(Loader ev, TimeNorm<Loader> tn, int64_t n, int H, int W, int K, float* out, cudaStream_t st) {
    if (H <= 0 || W <= 0 || K < 1 || K > 64 || !out || n < 0) return EVREP_ERR_ARG;
    const int64_t total = (int64_t)2 * K * H * W;
    EVREP_CUDA(cudaMemsetAsync(out, 0, total * sizeof(float), st));
    if (n > 0) {
        ev_splat_kernel<Loader><<<grid_for(n), kBlock, 0, st>>>(ev, tn, n, H, W, K, out);
        EVREP_LAUNCH_CHECK();
    }
    const bool vec = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const int64_t n4 = vec ? total >> 2 : 0;
    ev_scale_kernel<<<grid_for(total, 4), kBlock, 0, st>>>(reinterpret_cast<float4*>(out), n4, out, total);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

template <class Loader>
static int taf_bin_run(Loader ev, TafTime<Loader> tt, int64_t n, int H, int W, int K, const float* state_in,
                       float* state_out, float* out, void* scratch, cudaStream_t st) {
    if (H <= 0 || W <= 0 || !state_in || !state_out || !scratch || n < 0) return EVREP_ERR_ARG;
    uint32_t* flag = reinterpret_cast<uint32_t*>(scratch);
    TafCell* cells = reinterpret_cast<TafCell*>(reinterpret_cast<char*>(scratch) + 16);
    const int64_t HW = (int64_t)H * W;
    if (n > 0) {
        taf_scatter_kernel<Loader><<<grid_for(n), kBlock, 0, st>>>(ev, tt, n, H, W, flag, cells);
        EVREP_LAUNCH_CHECK();
    }
    const int grid = grid_for(HW);
#define EVREP_TAF_K(KK) case KK: taf_update_kernel<KK><<<grid, kBlock, 0, st>>>(flag, cells, HW, state_in, state_out, out); break;
    switch (K) {
        EVREP_TAF_K(1) EVREP_TAF_K(2) EVREP_TAF_K(3) EVREP_TAF_K(4) EVREP_TAF_K(5) EVREP_TAF_K(6)
        EVREP_TAF_K(7) EVREP_TAF_K(8) EVREP_TAF_K(9) EVREP_TAF_K(10) EVREP_TAF_K(11) EVREP_TAF_K(12)
        EVREP_TAF_K(13) EVREP_TAF_K(14) EVREP_TAF_K(15) EVREP_TAF_K(16)
        default: return EVREP_ERR_ARG;
    }
#undef EVREP_TAF_K
    EVREP_LAUNCH_CHECK();
    EVREP_CUDA(cudaMemsetAsync(flag, 0, 16, st));
    return EVREP_OK;
}

}  // namespace evrep

using namespace evrep;

extern "C" {

int evrep_decode_dat(const void* records, int64_t n, uint32_t* t, uint16_t* x, uint16_t* y, uint8_t* p,
                     evrep_stream_t stream) {
    if (n < 0 || (n > 0 && (!records || !t || !x || !y || !p))) return EVREP_ERR_ARG;
    if (reinterpret_cast<uintptr_t>(records) & 7) return EVREP_ERR_ARG;
    if (n == 0) return EVREP_OK;
    // 128-bit path needs 16-byte aligned records / t and 8/8/4-byte aligned x / y / p
    const bool vec = !((reinterpret_cast<uintptr_t>(records) & 15) | (reinterpret_cast<uintptr_t>(t) & 15) |
                       (reinterpret_cast<uintptr_t>(x) & 7) | (reinterpret_cast<uintptr_t>(y) & 7) |
                       (reinterpret_cast<uintptr_t>(p) & 3));
    decode_dat_kernel<<<grid_for(n, 4), kBlock, 0, as_stream(stream)>>>(
        reinterpret_cast<const uint2*>(records), n, t, x, y, p, vec);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

int evrep_soa_to_aos64(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n,
                       double* events, evrep_stream_t stream) {
    if (n < 0 || (n > 0 && (!t || !x || !y || !p || !events))) return EVREP_ERR_ARG;
    if (n == 0) return EVREP_OK;
    SoA ev{t, x, y, p, nullptr, nullptr};
    soa_to_aos64_kernel<<<grid_for(n), kBlock, 0, as_stream(stream)>>>(ev, n, events);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

int evrep_count_accumulate(const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n, int H, int W,
                           const uint16_t* xmap, const uint16_t* ymap, uint32_t* counts, evrep_stream_t stream) {
    if (H <= 0 || W <= 0 || !counts || n < 0 || (n > 0 && (!x || !y || !p))) return EVREP_ERR_ARG;
    SoA ev{nullptr, x, y, p, xmap, ymap};
    return count_accumulate(ev, n, H, W, counts, as_stream(stream));
}

int evrep_count_accumulate_aos64(const double* events, int64_t n, int ncols, int H, int W, uint32_t* counts,
                                 evrep_stream_t stream) {
    if (H <= 0 || W <= 0 || !counts || n < 0 || ncols < 4 || (n > 0 && !events)) return EVREP_ERR_ARG;
    Aos64 ev{events, ncols};
    return count_accumulate(ev, n, H, W, counts, as_stream(stream));
}

int evrep_count_finalize(uint32_t* counts, int H, int W, float* out, int reset, evrep_stream_t stream) {
    if (H <= 0 || W <= 0 || !counts || !out) return EVREP_ERR_ARG;
    int rc = ensure_count_lut();
    if (rc) return rc;
    const int64_t cells = (int64_t)2 * H * W;
    count_finalize_kernel<<<grid_for(cells), kBlock, 0, as_stream(stream)>>>(counts, cells, out, reset);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

int evrep_count_image(const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n, int H, int W,
                      const uint16_t* xmap, const uint16_t* ymap, uint32_t* counts, float* out,
                      evrep_stream_t stream) {
    int rc = evrep_count_accumulate(x, y, p, n, H, W, xmap, ymap, counts, stream);
    if (rc) return rc;
    return evrep_count_finalize(counts, H, W, out, 1, stream);
}

int evrep_count_lut_u8_batch(const uint8_t* frames, int64_t frame_stride, int64_t n_windows, int H, int W, int Ht, int Wt,
                             const int32_t* ysrc, const int32_t* xsrc, uint8_t* out, evrep_stream_t stream) {
    if (H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0 || n_windows < 0) return EVREP_ERR_ARG;
    if (n_windows == 0) return EVREP_OK;
    if (!frames || !out) return EVREP_ERR_ARG;
    if ((!ysrc || !xsrc) && (Ht != H || Wt != W)) return EVREP_ERR_ARG;
    int rc = ensure_count_lut();
    if (rc) return rc;
    if (Wt % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0)
        count_lut_u8_batch_kernel<4><<<grid_for(n_windows * 2 * Ht * Wt / 4), kBlock, 0, as_stream(stream)>>>(
            frames, frame_stride, n_windows, H, W, Ht, Wt, ysrc, xsrc, out);
    else
        count_lut_u8_batch_kernel<1><<<grid_for(n_windows * 2 * Ht * Wt), kBlock, 0, as_stream(stream)>>>(
            frames, frame_stride, n_windows, H, W, Ht, Wt, ysrc, xsrc, out);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

int evrep_count_images_u8(const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n, const int64_t* sizes_host,
                          int n_sizes, int H, int W, const uint16_t* xmap, const uint16_t* ymap, int Ht, int Wt,
                          const int32_t* ysrc, const int32_t* xsrc, uint32_t* counts, uint8_t* out, evrep_stream_t stream) {
    if (H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0 || n < 0 || n_sizes <= 0 || !sizes_host || !counts || !out) return EVREP_ERR_ARG;
    if (n > 0 && (!x || !y || !p)) return EVREP_ERR_ARG;
    if ((!ysrc || !xsrc) && (Ht != H || Wt != W)) return EVREP_ERR_ARG;
    int rc = ensure_count_lut();
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    int64_t done = 0;
    const int64_t per = (int64_t)2 * Ht * Wt;
    for (int i = 0; i < n_sizes; ++i) {
        if (i > 0 && sizes_host[i] < sizes_host[i - 1]) return EVREP_ERR_ARG;
        const int64_t take = sizes_host[i] < n ? sizes_host[i] : n;          // events[-N:]
        if (take > done) {                                                   // only the events not counted yet
            const int64_t lo = n - take, cnt = take - done;
            SoA ev{nullptr, x + lo, y + lo, p + lo, xmap, ymap};
            count_accumulate_kernel<SoA><<<grid_for(cnt, 4), kBlock, 0, st>>>(ev, cnt, H, W, counts);
            EVREP_LAUNCH_CHECK();
            done = take;
        }
        count_finalize_u8_kernel<<<grid_for(per), kBlock, 0, st>>>(counts, H, W, Ht, Wt, ysrc, xsrc, out + i * per);
        EVREP_LAUNCH_CHECK();
    }
    EVREP_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint32_t) * 2 * (size_t)H * W, st));
    return EVREP_OK;
}

int evrep_sae_u8(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n, int H, int W,
                 const uint16_t* xmap, const uint16_t* ymap, int Ht, int Wt, const int32_t* ysrc, const int32_t* xsrc,
                 float init, float now_f32, const float* lambdas_host, int L, const float* memory_in, float* memory_out,
                 uint32_t* keys, uint8_t* out, evrep_stream_t stream) {
    if (H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0 || L < 1 || L > 8 || !lambdas_host || !memory_out || !keys || !out || n < 0)
        return EVREP_ERR_ARG;
    if (n > 0 && (!t || !x || !y || !p)) return EVREP_ERR_ARG;
    if ((!ysrc || !xsrc) && (Ht != H || Wt != W)) return EVREP_ERR_ARG;
    if (memory_in == memory_out) return EVREP_ERR_ARG;         // output cells re-read memory_in
    cudaStream_t st = as_stream(stream);
    if (n > 0) {
        SoA ev{t, x, y, p, xmap, ymap};
        sae_scatter_kernel<SoA><<<grid_for(n, 4), kBlock, 0, st>>>(ev, n, H, W, keys);
        EVREP_LAUNCH_CHECK();
    }
    Lambdas lam;
    for (int l = 0; l < 8; ++l) lam.v[l] = l < L ? lambdas_host[l] : 0.0f;
    sae_finalize_u8_kernel<<<grid_for((int64_t)2 * H * W + (int64_t)2 * Ht * Wt), kBlock, 0, st>>>(
        keys, H, W, Ht, Wt, ysrc, xsrc, init, now_f32, lam, L, memory_in, memory_out, out);
    EVREP_LAUNCH_CHECK();
    EVREP_CUDA(cudaMemsetAsync(keys, 0, sizeof(uint32_t) * 2 * (size_t)H * W, st));
    return EVREP_OK;
}

int evrep_sae(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n, int H, int W,
              const uint16_t* xmap, const uint16_t* ymap, float init, float now_f32, const float* lambdas_host,
              int L, const float* memory_in, float* memory_out, uint32_t* keys, float* out, evrep_stream_t stream) {
    if (n > 0 && (!t || !x || !y || !p)) return EVREP_ERR_ARG;
    SoA ev{t, x, y, p, xmap, ymap};
    return sae_run(ev, n, H, W, init, now_f32, lambdas_host, L, memory_in, memory_out, keys, out, as_stream(stream));
}

int evrep_sae_aos64(const double* events, int64_t n, int ncols, int H, int W, float init, float now_f32,
                    const float* lambdas_host, int L, const float* memory_in, float* memory_out, uint32_t* keys,
                    float* out, evrep_stream_t stream) {
    if (ncols < 4 || (n > 0 && !events)) return EVREP_ERR_ARG;
    Aos64 ev{events, ncols};
    return sae_run(ev, n, H, W, init, now_f32, lambdas_host, L, memory_in, memory_out, keys, out, as_stream(stream));
}

int evrep_event_volume(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n,
                       int64_t t0, int64_t tw, int H, int W, int K, const uint16_t* xmap, const uint16_t* ymap,
                       float* out, evrep_stream_t stream) {
    if (tw == 0 || (n > 0 && (!t || !x || !y || !p))) return EVREP_ERR_ARG;
    SoA ev{t, x, y, p, xmap, ymap};
    TimeNorm<SoA> tn{t0, (double)tw};
    return ev_run(ev, tn, n, H, W, K, out, as_stream(stream));
}

int evrep_event_volume_aos64(const double* events, int64_t n, int ncols, int H, int W, int K, float* out,
                             evrep_stream_t stream) {
    if (ncols < 4 || (n > 0 && !events)) return EVREP_ERR_ARG;
    Aos64 ev{events, ncols};
    return ev_run(ev, TimeNorm<Aos64>{}, n, H, W, K, out, as_stream(stream));
}

int64_t evrep_taf_bin_scratch_bytes(int H, int W) {
    if (H <= 0 || W <= 0) return EVREP_ERR_ARG;
    return 16 + (int64_t)2 * H * W * (int64_t)sizeof(TafCell);
}

int evrep_taf_bin(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n,
                  int64_t t_min, double t_span, int H, int W, int K, const uint16_t* xmap, const uint16_t* ymap,
                  const float* state_in, float* state_out, float* out, void* scratch, evrep_stream_t stream) {
    if (!(t_span > 0.0) || (n > 0 && (!t || !x || !y || !p))) return EVREP_ERR_ARG;
    SoA ev{t, x, y, p, xmap, ymap};
    TafTime<SoA> tt{t_min, t_span};
    return taf_bin_run(ev, tt, n, H, W, K, state_in, state_out, out, scratch, as_stream(stream));
}

int evrep_taf_bin_aos64(const double* events, int64_t n, int ncols, int H, int W, int K, const float* state_in,
                        float* state_out, float* out, void* scratch, evrep_stream_t stream) {
    if (ncols < 4 || (n > 0 && !events)) return EVREP_ERR_ARG;
    Aos64 ev{events, ncols};
    return taf_bin_run(ev, TafTime<Aos64>{}, n, H, W, K, state_in, state_out, out, scratch, as_stream(stream));
}

int evrep_nearest_resize(const float* in, int C, int H, int W, int Ht, int Wt, const int32_t* ysrc,
                         const int32_t* xsrc, float* out, evrep_stream_t stream) {
    if (!in || !out || !ysrc || !xsrc || C <= 0 || H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0) return EVREP_ERR_ARG;
    nearest_resize_kernel<<<grid_for((int64_t)C * Ht * Wt), kBlock, 0, as_stream(stream)>>>(in, C, H, W, Ht, Wt, ysrc, xsrc, out);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

int evrep_quantize_u8(const float* in, int64_t n, int clamp255, uint8_t* out, evrep_stream_t stream) {
    if (n < 0 || (n > 0 && (!in || !out))) return EVREP_ERR_ARG;
    if (n == 0) return EVREP_OK;
    quantize_u8_kernel<<<grid_for(n), kBlock, 0, as_stream(stream)>>>(in, n, clamp255, out);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

int evrep_leaky_transform(const float* in, int64_t n, float* out, evrep_stream_t stream) {
    if (n < 0 || (n > 0 && (!in || !out))) return EVREP_ERR_ARG;
    if (n == 0) return EVREP_OK;
    leaky_kernel<<<grid_for(n), kBlock, 0, as_stream(stream)>>>(in, n, out);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

int evrep_taf_leaky_u8_batch(const float* volumes, int64_t volume_stride, int n_windows, int K, int H, int W, int Ht,
                             int Wt, const int32_t* ysrc, const int32_t* xsrc, uint8_t* out, evrep_stream_t stream);

int evrep_taf_leaky_u8(const float* volume, int K, int H, int W, int Ht, int Wt, const int32_t* ysrc,
                       const int32_t* xsrc, uint8_t* out, evrep_stream_t stream) {
    if (!volume || !out || K <= 0 || H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0) return EVREP_ERR_ARG;
    if ((!ysrc || !xsrc) && (Ht != H || Wt != W)) return EVREP_ERR_ARG;
    return evrep_taf_leaky_u8_batch(volume, 0, 1, K, H, W, Ht, Wt, ysrc, xsrc, out, stream);
}

int evrep_taf_leaky_u8_batch(const float* volumes, int64_t volume_stride, int n_windows, int K, int H, int W, int Ht,
                             int Wt, const int32_t* ysrc, const int32_t* xsrc, uint8_t* out, evrep_stream_t stream) {
    if (!volumes || !out || n_windows < 0 || K <= 0 || H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0) return EVREP_ERR_ARG;
    if ((!ysrc || !xsrc) && (Ht != H || Wt != W)) return EVREP_ERR_ARG;
    if (n_windows == 0) return EVREP_OK;
    const int64_t per_thread = (!ysrc && !xsrc) ? 4 : 1;
    const int64_t bx = ((int64_t)Ht * Wt / per_thread + kBlock - 1) / kBlock;
    const int64_t planes = (int64_t)n_windows * 2 * K;
    const dim3 grid((unsigned)(bx < 1 ? 1 : (bx > 65535 ? 65535 : bx)), (unsigned)(planes > 65535 ? 65535 : planes));
    taf_leaky_u8_kernel<<<grid, kBlock, 0, as_stream(stream)>>>(
        volumes, volume_stride, n_windows, K, H, W, Ht, Wt, ysrc, xsrc, out);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

}  // extern "C"
