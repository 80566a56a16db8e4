// Shared helpers for libevrep (sm_100a).  Host-side error plumbing and the two event
// loaders every encoder kernel is templated on.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/evrep.h"

namespace evrep {

constexpr int kBlock = 256;

// ---- error plumbing --------------------------------------------------------------
void set_cuda_error(cudaError_t e);          // api.cu: remembers the string per thread

#define EVREP_CUDA(call)                                              \
    do {                                                              \
        cudaError_t e__ = (call);                                     \
        if (e__ != cudaSuccess) {                                     \
            ::evrep::set_cuda_error(e__);                             \
            return EVREP_ERR_CUDA;                                    \
        }                                                             \
    } while (0)

#define EVREP_LAUNCH_CHECK() EVREP_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(evrep_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();                               // api.cu (cached per device)

// Grid for a grid-stride kernel: enough CTAs to cover `work` items at `per_thread`
// items per thread, capped at a few waves of the machine.
inline int grid_for(int64_t work, int per_thread = 1, int block = kBlock, int waves = 8) {
    int64_t need = (work + (int64_t)block * per_thread - 1) / ((int64_t)block * per_thread);
    int64_t cap = (int64_t)sm_count() * waves;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// v / 5 * 255 (generate_eventvolume.py:37) without the IEEE-division subroutine: one Newton
// correction of v * RN(1/5) is the correctly rounded quotient for every finite v away from the
// denormal range, so the two roundings of the reference are reproduced.
__device__ __forceinline__ float div5_mul255(float v) {
    const float q = v * 0.2f;
    const float r = fmaf(-q, 5.0f, v);
    return fmaf(r, 0.2f, q) * 255.0f;
}

// ---- event loaders ---------------------------------------------------------------
// One decoded event as the encoders see it: grid coordinates (already mapped), the
// polarity, and `ok` = inside the H x W grid with p in {0,1}.
struct Event {
    int x, y, p;
    bool ok;
};

// Structure-of-arrays stream (t u32 us, x u16, y u16, p u8) with optional coordinate LUTs.
struct SoA {
    const uint32_t* t;
    const uint16_t* x;
    const uint16_t* y;
    const uint8_t* p;
    const uint16_t* xmap;
    const uint16_t* ymap;

    __device__ __forceinline__ Event load(int64_t i, int H, int W) const {
        int xr = __ldg(x + i), yr = __ldg(y + i);
        Event e;
        e.p = __ldg(p + i);
        // LUTs hold EVREP_COORD_LUT_LEN (2^14, the .dat coordinate range) entries
        bool raw_ok = (xr < EVREP_COORD_LUT_LEN) & (yr < EVREP_COORD_LUT_LEN);
        e.x = (xmap && raw_ok) ? (int)__ldg(xmap + xr) : xr;
        e.y = (ymap && raw_ok) ? (int)__ldg(ymap + yr) : yr;
        e.ok = raw_ok & (e.x < W) & (e.y < H) & (e.p < 2);
        return e;
    }
    __device__ __forceinline__ uint32_t time_us(int64_t i) const { return __ldg(t + i); }
};

// The reference's staging matrix: float64 [N, ncols], columns (x, y, t, p[, z]).
// Coordinates / polarity are truncated toward zero like `.long()`.
struct Aos64 {
    const double* ev;
    int ncols;

    __device__ __forceinline__ Event load(int64_t i, int H, int W) const {
        const double* r = ev + i * ncols;
        long long xl = (long long)__ldg(r + 0), yl = (long long)__ldg(r + 1), pl = (long long)__ldg(r + 3);
        Event e;
        e.ok = (xl >= 0) & (xl < W) & (yl >= 0) & (yl < H) & (pl >= 0) & (pl < 2);
        e.x = (int)xl; e.y = (int)yl; e.p = (int)pl;
        return e;
    }
    __device__ __forceinline__ double time_f64(int64_t i) const { return __ldg(ev + i * ncols + 2); }
};

}  // namespace evrep
