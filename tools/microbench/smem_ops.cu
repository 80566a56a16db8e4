// Micro-benchmarks of the shared-memory operations the sort / tile kernels are built from (B200):
// ATOMS.ADD with and without a consumed result, MATCH.ANY, a ballot multisplit, plain scattered STS.
// Output: cycles per warp-level operation per SM at saturation (16 or 32 warps per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_ops smem_ops.cu && ./smem_ops
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kCells = 4480;
constexpr int kIters = 2048;

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void bench(uint32_t* out, long long* cycles, int keybits) {
    __shared__ uint32_t acc[kCells];
    __shared__ uint16_t warp_cnt[32][448];
    for (int i = threadIdx.x; i < kCells; i += blockDim.x) acc[i] = 0;
    for (int i = threadIdx.x; i < 32 * 448; i += blockDim.x) (&warp_cnt[0][0])[i] = 0;
    __syncthreads();
    uint32_t s = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 1u, sink = 0;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < kIters; ++it) {
        const uint32_t r = rng(s);
        if (MODE == 0) {                       // ATOMS.ADD, result unused
            atomicAdd(&acc[r % kCells], (1u << 23) + (r & 8191u));
        } else if (MODE == 1) {                // ATOMS.ADD, result consumed
            sink += atomicAdd(&acc[r % kCells], (1u << 23) + (r & 8191u));
        } else if (MODE == 2) {                // two atomics per item (count, sum)
            atomicAdd(&acc[r % (kCells / 2)], 1u);
            atomicAdd(&acc[kCells / 2 + r % (kCells / 2)], r & 8191u);
        } else if (MODE == 3) {                // MATCH.ANY on a key of `keybits` bits
            sink += __popc(__match_any_sync(0xFFFFFFFFu, r & ((1u << keybits) - 1u)));
        } else if (MODE == 4) {                // ballot multisplit + warp-private counters (no atomics)
            const uint32_t key = r & ((1u << keybits) - 1u);
            uint32_t peers = 0xFFFFFFFFu;
            for (int b = 0; b < keybits; ++b) {
                const bool bit = (key >> b) & 1u;
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, bit);
                peers &= bit ? bal : ~bal;
            }
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if (lane == leader) { base = warp_cnt[wid][key % 448u]; warp_cnt[wid][key % 448u] = (uint16_t)(base + __popc(peers)); }
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            sink += base + __popc(peers & ((1u << lane) - 1u));
        } else if (MODE == 5) {                // scattered STS
            acc[r % kCells] = r;
        } else if (MODE == 6) {                // same-address atomics (ptxas REDUX aggregation candidate)
            atomicAdd(&acc[it & 7], 1u);
        } else if (MODE == 7) {                // ATOMS.ADD consumed + ballot append (the tile kernel's accumulate)
            const uint32_t old = atomicAdd(&acc[r % kCells], (1u << 23) + (r & 8191u));
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, old == 0u);
            sink += __popc(m);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = sink + acc[threadIdx.x % kCells];
}

template <int MODE>
static void run(const char* name, int threads, int keybits = 9) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * threads);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    bench<MODE><<<sms, threads>>>(out, cyc, keybits);
    bench<MODE><<<sms, threads>>>(out, cyc, keybits);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    const double per = avg / ((double)kIters * (threads / 32));
    printf("%-46s threads=%4d keybits=%2d  %.2f cycles per warp-op per SM  (%.3f per lane)\n", name, threads, keybits, per, per / 32);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {256, 512, 1024}) {
        run<0>("ATOMS.ADD spread, result unused", threads);
        run<1>("ATOMS.ADD spread, result consumed", threads);
        run<2>("2 x ATOMS.ADD spread (count, sum)", threads);
        run<7>("ATOMS.ADD consumed + ballot", threads);
        run<3>("MATCH.ANY", threads, 9);
        run<4>("ballot multisplit + private counters", threads, 9);
        run<4>("ballot multisplit + private counters", threads, 8);
        run<5>("STS scattered", threads);
        run<6>("ATOMS.ADD same address", threads);
    }
    return 0;
}
