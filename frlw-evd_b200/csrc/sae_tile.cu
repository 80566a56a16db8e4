#include "stream_common.cuh"

namespace evrep {

// ---- Surface of Active Events over whole streams ----------------------------------------------
// generate_surfaceofactiveevents.py:44-69 for a list of consecutive windows (one per label): the
// bucketing of the TAF path sorts the events by (sensor tile, bin) -- a window is cut into bins of
// kSaeBin microseconds only so that the 18-bit time offset of a record suffices -- and one CTA per
// tile keeps the tile's state, the latest float32 timestamp per (polarity, pixel), in shared
// memory across all windows:
//   per bin     scatter-max of float32(t) into u32 keys (native shared-memory ATOMS.MAX);
//   per window  latest = max(state, float32(now) - 5e6, key)  (:48-52), state = latest (:54),
//               the [2, tile] slice of `latest` goes to the window's frame, keys are cleared.
// The L exponential decays, the nearest resize and the uint8 truncation of the driver
// (:55-63,186-204) are one batched pass over the frames (sae_decay_u8_batch_kernel).
constexpr int kSaeThreads = 512;
constexpr int kSaeTilesPerSm = 2;
constexpr int kSaeBin = 250000;          // < 2^18 us

struct SaeBin {          // one entry per global bin, built on the host
    uint32_t t0;         // time of the bin start
    int32_t window;      // window that ends with this bin, or -1
    float floor_value;   // float32(now) - 5e6 of that window
    int32_t pad;
};

struct SaeTileParams {
    StreamPlan pl;
    const SaeBin* bins;
    float* memory;           // f32 [2,H,W] state in / out
    int has_memory;
    float* latest;           // f32 [n_windows][2,H,W]
    int64_t latest_stride;
};

__device__ __forceinline__ uint32_t sae_key(float f) { return __float_as_uint(f) + 1u; }     // f >= 0; 0 = no event
__device__ __forceinline__ float sae_time(uint32_t k) { return __uint_as_float(k - 1u); }

__global__ void __launch_bounds__(kSaeThreads, kSaeTilesPerSm)
sae_tile_kernel(SaeTileParams tp) {
    const StreamPlan& pl = tp.pl;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw);                               // [kWsRing]
    uint32_t* keys = ring + kWsRing;                                                       // [2][P]
    float* state = reinterpret_cast<float*>(keys + 2 * pl.P);                              // [2][P]
    uint64_t* full = reinterpret_cast<uint64_t*>(state + 2 * pl.P);

    const int tid = threadIdx.x, tile = blockIdx.x, lane = tid & 31;
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
    // state: the caller's memory, or "older than any floor" so that the first window starts from its floor
    for (uint32_t i = tid; i < 2 * P; i += kSaeThreads) {
        const uint32_t pol = i / P, lp = i - pol * P;
        keys[i] = 0u;
        state[i] = (tp.has_memory && lp < npix) ? tp.memory[(int64_t)pol * HW + pix0 + lp] : -INFINITY;
    }
    __syncthreads();
    if (tid == 0)
        for (int c = 0; c < n_chunks && c < kWsStages; ++c) issue(c);

    // record range and descriptor of bin g: lane l of every warp holds those of bin group*32 + l,
    // the next group of 32 is loaded one group ahead
    const int TB = pl.TB;
    auto load_group = [&](int first, uint32_t& lo, uint32_t& hi, SaeBin& info) {
        const int g = min(first + lane, TB - 1);
        lo = __ldg(my_off + g); hi = __ldg(my_off + g + 1);
        info = tp.bins[g];
    };
    uint32_t b_lo, b_hi, nb_lo = 0, nb_hi = 0;
    SaeBin b_info, nb_info = {};
    load_group(0, b_lo, b_hi, b_info);

    int ready_chunk = -1, next_refill = kWsStages;
    for (int g = 0; g < TB; ++g) {
        if ((g & 31) == 0) {
            if (g) { b_lo = nb_lo; b_hi = nb_hi; b_info = nb_info; }
            if (g + 32 < TB) load_group(g + 32, nb_lo, nb_hi, nb_info);
        }
        const int src = g & 31;
        const uint32_t o0 = __shfl_sync(0xFFFFFFFFu, b_lo, src), o1 = __shfl_sync(0xFFFFFFFFu, b_hi, src);
        const uint32_t t0 = __shfl_sync(0xFFFFFFFFu, b_info.t0, src);
        const int window = __shfl_sync(0xFFFFFFFFu, b_info.window, src);
        const float floor_value = __shfl_sync(0xFFFFFFFFu, b_info.floor_value, src);
        uint32_t cur = o0;
        while (cur < o1) {
            const uint32_t avail = (uint32_t)next_refill * kWsChunkRecords;     // records requested so far
            const uint32_t limit = o1 < avail ? o1 : avail;
            const int last_c = (int)((limit - 1) / kWsChunkRecords);
            while (ready_chunk < last_c) {
                ++ready_chunk;
                mbar_wait(full + (ready_chunk % kWsStages), (uint32_t)(ready_chunk / kWsStages) & 1u);
            }
            for (uint32_t r = cur + tid; r < limit; r += kSaeThreads) {
                const uint32_t rec = ring[r & (kWsRing - 1)];
                // :76 t -> float32; record = [ d:18 | local pixel:13 | p:1 ], t = bin start + d
                const float tf = (float)(t0 + (rec >> 14));
                atomicMax(keys + (rec & 1u) * P + ((rec >> 1) & 0x1FFFu), sae_key(tf));
            }
            cur = limit;
            if (cur < o1) {                                        // the bin outgrew the ring: recycle stages
                __syncthreads();
                const int drained = (int)(cur / kWsChunkRecords);
                if (tid == 0)
                    for (int r = next_refill; r < drained + kWsStages && r < n_chunks; ++r) issue(r);
                next_refill = drained + kWsStages;
            }
        }
        if (window < 0 && o1 / kWsChunkRecords + kWsStages <= (uint32_t)next_refill) continue;   // nothing to do at this bin's end
        __syncthreads();
        {
            const int drained = (int)(o1 / kWsChunkRecords);
            if (drained + kWsStages > next_refill) {
                if (tid == 0)
                    for (int r = next_refill; r < drained + kWsStages && r < n_chunks; ++r) issue(r);
                next_refill = drained + kWsStages;
            }
        }
        if (window < 0) continue;
        // end of a window: merge keys, floor and state; emit the tile's slice of the frame
        float* o = tp.latest + (int64_t)window * tp.latest_stride + pix0;
        for (uint32_t i = tid; i < 2 * P; i += kSaeThreads) {
            const uint32_t pol = i / P, lp = i - pol * P;
            const uint32_t k = keys[i];
            float v = fmaxf(state[i], floor_value);
            if (k) { v = fmaxf(v, sae_time(k)); keys[i] = 0u; }
            state[i] = v;
            if (lp < npix) __stcs(o + (int64_t)pol * HW + lp, v);
        }
        __syncthreads();                                           // cleared keys are visible to the next scatter
    }
    // the state after the last window
    for (uint32_t i = tid; i < 2 * P; i += kSaeThreads) {
        const uint32_t pol = i / P, lp = i - pol * P;
        if (lp < npix) tp.memory[(int64_t)pol * HW + pix0 + lp] = state[i];
    }
}

struct SaeLambdas { float v[8]; };

// :55-63 + driver :186-204 for n windows at once: out[w, l, p, Y, X] =
// uint8(exp(float32(lambda_l) * (latest[w, p, ysrc[Y], xsrc[X]] - float32(now_w))) * 255).
template <int kVec>      // output pixels per thread: 4 (one 32-bit store per lambda) when Wt % 4 == 0, else 1
__global__ void __launch_bounds__(256)
sae_decay_u8_batch_kernel(const float* __restrict__ latest, int64_t latest_stride, const float* __restrict__ now_f32,
                          int64_t n_windows, int H, int W, int Ht, int Wt, const int32_t* __restrict__ ysrc,
                          const int32_t* __restrict__ xsrc, SaeLambdas lam, int L, uint8_t* __restrict__ out) {
    const int64_t tcells = (int64_t)2 * Ht * Wt, groups = tcells / kVec, total = n_windows * groups;
    const int wq = Wt / kVec;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t w = i / groups, j = i - w * groups;
        const int X0 = (int)(j % wq) * kVec;
        const int64_t r = j / wq;
        const int Y = (int)(r % Ht), p = (int)(r / Ht);
        const int ys = ysrc ? ysrc[Y] : Y;
        const float* row = latest + w * latest_stride + ((int64_t)p * H + ys) * W;
        const float now = now_f32[w];
        float rel[kVec];
#pragma unroll
        for (int k = 0; k < kVec; ++k) rel[k] = row[xsrc ? xsrc[X0 + k] : X0 + k] - now;
        uint8_t* o = out + w * L * tcells + j * kVec;
        for (int l = 0; l < L; ++l) {
            uint32_t packed = 0;
#pragma unroll
            for (int k = 0; k < kVec; ++k) {
                const uint32_t b = (uint32_t)(uint8_t)(int)(expf(lam.v[l] * rel[k]) * 255.0f);
                if (kVec == 1) o[(int64_t)l * tcells] = (uint8_t)b;
                packed |= b << (8 * k);
            }
            if (kVec == 4) *reinterpret_cast<uint32_t*>(o + (int64_t)l * tcells) = packed;
        }
    }
}

}  // namespace evrep

using namespace evrep;

static inline int64_t sae_bins_of(const evrep_sae_window& w) {
    if (w.ev_end <= w.ev_begin || w.t_last < w.t_first) return 1;          // an empty window still emits its frame
    return (w.t_last - w.t_first) / kSaeBin + 1;
}

extern "C" {

int64_t evrep_sae_stream_scratch_bytes(int64_t n_events, const evrep_sae_window* windows_host, int n_windows, int H, int W) {
    if (n_events < 0 || n_windows < 0 || H <= 0 || W <= 0 || (n_windows > 0 && !windows_host)) return EVREP_ERR_ARG;
    int64_t TB = 0;
    for (int w = 0; w < n_windows; ++w) TB += sae_bins_of(windows_host[w]);
    Layout L;
    int rc = make_layout(n_events, n_windows, TB, H, W, (int)batches_upper_bound(n_windows, TB), L, kSaeTilesPerSm);
    if (rc) return rc;
    return L.total + ((int64_t)sizeof(SaeBin) * (TB > 0 ? TB : 1) + 255) / 256 * 256;
}

int evrep_sae_stream(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                     const evrep_sae_window* windows_host, int n_windows, int H, int W,
                     const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                     float* memory_inout, int has_memory, float* latest_out, int64_t latest_stride,
                     void* scratch, int64_t scratch_bytes, evrep_stream_t stream) {
    if (n_events < 0 || n_windows < 0 || H <= 0 || W <= 0 || !scratch || !memory_inout) return EVREP_ERR_ARG;
    if (n_windows == 0) return EVREP_OK;
    if (!latest_out || !windows_host) return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    // windows -> bins of kSaeBin us starting at the window's first timestamp
    std::vector<evrep_taf_window> wins((size_t)n_windows);
    std::vector<SaeBin> bins;
    for (int w = 0; w < n_windows; ++w) {
        const evrep_sae_window& sw = windows_host[w];
        if (sw.t_first < 0 || sw.t_last > 0xFFFFFFFFll) return EVREP_ERR_RANGE;
        const int64_t nb = sae_bins_of(sw);
        wins[w].ev_begin = sw.ev_begin; wins[w].ev_end = sw.ev_end; wins[w].start_time = sw.t_first;
        wins[w].n_bins = (int32_t)nb; wins[w].fresh = 0;
        // :48  t_img = zeros(float32) + now - 5e6, evaluated in float32
        const float floor_value = (float)sw.now - 5000000.0f;
        for (int64_t z = 0; z < nb; ++z) {
            SaeBin b;
            b.t0 = (uint32_t)(sw.t_first + z * kSaeBin);
            b.window = z + 1 == nb ? w : -1;
            b.floor_value = floor_value; b.pad = 0;
            bins.push_back(b);
        }
    }
    StreamPlan pl;
    Layout L;
    int rc = prepare_stream(t, x, y, p, n_events, wins.data(), n_windows, kSaeBin, H, W, xmap, ymap, sensor_h, sensor_w,
                            scratch, scratch_bytes, st, pl, L, kSaeTilesPerSm);
    if (rc) return rc;
    const int64_t bins_bytes = (int64_t)sizeof(SaeBin) * (int64_t)bins.size();
    if (scratch_bytes < L.total + bins_bytes) return EVREP_ERR_SCRATCH;
    SaeBin* d_bins = reinterpret_cast<SaeBin*>(reinterpret_cast<char*>(scratch) + L.total);
    rc = upload_words(reinterpret_cast<const uint32_t*>(bins.data()), bins_bytes / 4, reinterpret_cast<uint32_t*>(d_bins), st);
    if (rc) return rc;
    const size_t smem = (size_t)kWsRing * 4 + (size_t)4 * L.P * 4 + 64;
    SaeTileParams tp;
    tp.pl = pl; tp.bins = d_bins; tp.memory = memory_inout; tp.has_memory = has_memory;
    tp.latest = latest_out; tp.latest_stride = latest_stride;
    EVREP_CUDA(cudaFuncSetAttribute(sae_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sae_tile_kernel<<<L.n_tiles, kSaeThreads, smem, st>>>(tp);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

int evrep_sae_decay_u8_batch(const float* latest, int64_t latest_stride, const float* now_f32, int64_t n_windows,
                             int H, int W, int Ht, int Wt, const int32_t* ysrc, const int32_t* xsrc,
                             const float* lambdas_host, int L, uint8_t* out, evrep_stream_t stream) {
    if (n_windows < 0 || H <= 0 || W <= 0 || Ht <= 0 || Wt <= 0 || L < 1 || L > 8 || !lambdas_host) return EVREP_ERR_ARG;
    if (n_windows == 0) return EVREP_OK;
    if (!latest || !now_f32 || !out) return EVREP_ERR_ARG;
    SaeLambdas lam;
    for (int l = 0; l < 8; ++l) lam.v[l] = l < L ? lambdas_host[l] : 0.0f;
    const bool vec = Wt % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
    const int64_t total = n_windows * 2 * Ht * Wt / (vec ? 4 : 1);
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
    if (vec)
        sae_decay_u8_batch_kernel<4><<<blocks, 256, 0, as_stream(stream)>>>(latest, latest_stride, now_f32, n_windows, H, W,
                                                                             Ht, Wt, ysrc, xsrc, lam, L, out);
    else
        sae_decay_u8_batch_kernel<1><<<blocks, 256, 0, as_stream(stream)>>>(latest, latest_stride, now_f32, n_windows, H, W,
                                                                             Ht, Wt, ysrc, xsrc, lam, L, out);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

}  // extern "C"
