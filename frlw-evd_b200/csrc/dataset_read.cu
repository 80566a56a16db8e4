// Training-time read path on the device: data/dataset.py:219-234 for a
// batch of samples.  A sample is the concatenation of its raw uint8 files ([C, Hs, Ws], the
// on-disk layout of W1); the kernel fuses the float conversion, the nearest resize to
// (up_h, up_w) = int(input_img_size * sr), the division by 255, the crop at (-cy, -cx) and the
// horizontal flip into one gather:
//   out[n, c, y, x] = LUT[src[n, c, ysrc(y - cy), xsrc(xf - cx)]],  xf = flip ? W_in - 1 - x : x,
//   ysrc(d) = min(floor(d * float32(Hs / up_h)), Hs - 1)   (F.interpolate 'nearest', size= form),
//   LUT[v] = float32(v) / 255 (the 256 possible results of the reference's division; one IEEE
//   division per thread at block start, kept in shared memory because the lookups diverge).
#include "common.cuh"

namespace evrep {

__global__ void __launch_bounds__(256)
load_samples_kernel(const uint8_t* __restrict__ files, int64_t file_stride, int n, int C, int Hs, int Ws,
                    const evrep_sample_aug* __restrict__ aug, int Hin, int Win, float* __restrict__ out) {
    __shared__ float s_over_255[256];
    s_over_255[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.0f);    // blockDim.x == 256
    __syncthreads();
    const int wq = Win / 4;                                  // 4 output pixels per thread when Win % 4 == 0, else 1
    const bool vec = (Win % 4) == 0;
    const int per_row = vec ? wq : Win;
    const int64_t total = (int64_t)n * C * Hin * per_row;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int xq = (int)(i % per_row);
        int64_t r = i / per_row;
        const int y = (int)(r % Hin); r /= Hin;
        const int c = (int)(r % C);
        const int s = (int)(r / C);
        const evrep_sample_aug a = aug[s];
        const float sy = (float)Hs / (float)a.up_h, sx = (float)Ws / (float)a.up_w;
        const int ys = min((int)floorf((float)(y - a.cy) * sy), Hs - 1);
        const uint8_t* row = files + s * file_stride + ((int64_t)c * Hs + ys) * Ws;
        float* o = out + (((int64_t)s * C + c) * Hin + y) * Win;
        if (vec) {
            float v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int x = xq * 4 + k;
                const int xf = a.flip ? Win - 1 - x : x;
                v[k] = s_over_255[row[min((int)floorf((float)(xf - a.cx) * sx), Ws - 1)]];
            }
            __stcs(reinterpret_cast<float4*>(o + xq * 4), make_float4(v[0], v[1], v[2], v[3]));
        } else {
            const int xf = a.flip ? Win - 1 - xq : xq;
            o[xq] = s_over_255[row[min((int)floorf((float)(xf - a.cx) * sx), Ws - 1)]];
        }
    }
}

}  // namespace evrep

using namespace evrep;

extern "C" {

int evrep_load_samples(const uint8_t* files, int64_t file_stride, int n, int C, int Hs, int Ws,
                       const evrep_sample_aug* aug, int Hin, int Win, float* out, evrep_stream_t stream) {
    if (n < 0 || C <= 0 || Hs <= 0 || Ws <= 0 || Hin <= 0 || Win <= 0) return EVREP_ERR_ARG;
    if (n == 0) return EVREP_OK;
    if (!files || !aug || !out) return EVREP_ERR_ARG;
    if (Win % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15)) return EVREP_ERR_ARG;
    const int64_t total = (int64_t)n * C * Hin * (Win % 4 == 0 ? Win / 4 : Win);
    load_samples_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(files, file_stride, n, C, Hs, Ws, aug, Hin, Win, out);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

}  // extern "C"
