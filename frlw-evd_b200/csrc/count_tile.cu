#include <algorithm>

#include "stream_common.cuh"

namespace evrep {

// ---- Event Count Image over whole streams -----------------------------------------------------
// The driver (generate_eventcountimage.py:130-182) encodes, for every label, the last N events
// before it for several N: windows that nest inside one label and overlap between labels.  Their
// boundaries cut the event stream into consecutive *segments*; every window is a run of segments.
// The bucketing of the TAF path sorts the events by (sensor tile, segment) -- a segment is a
// "window" of one bin for it -- and one CTA per tile then
//   per segment  counts its events per (polarity, pixel) with native shared-memory atomics and
//                stores the counts, saturated to 8 bits, in slot (segment mod D) of a ring;
//   per window   adds the slots of its segments with saturating byte adds and writes the tile's
//                [2, tile] slice of the count frame (uint8, saturated at 255).
// Saturation is harmless: the image value depends on min(count, 20) only (:32-34), and
// min(sum of min(n_g, 255), 255) >= 20 exactly when the true sum is.  The value LUT, the nearest
// resize and the uint8 output are one batched pass over the frames (evrep_count_lut_u8_batch).
constexpr int kCountThreads = 128;       // a segment brings a tile tens to hundreds of records: small CTAs, many of them
constexpr int kCountTilesPerSm = 2;      // tile size of the bucketing layout (tiles = 2 x SMs)
constexpr int kCountCtasPerSm = 8;       // resident CTAs the kernel is compiled for: groups of segments run side by side
constexpr int kCountMaxSmem = 112 * 1024;

struct CountSegment {        // per segment: the emissions that follow it
    int32_t emit_first, emit_count;
};
struct CountEmit {           // window = segments [first_segment, last_segment]; frame index = position in the list
    int32_t first_segment, last_segment;
};

struct CountTileParams {
    StreamPlan pl;
    const CountSegment* segments;
    const CountEmit* emits;
    uint8_t* frames;         // u8 [n_emits][2,H,W]
    int64_t frame_stride;
    int depth;               // ring slots
    int seg_per_group;       // blockIdx.y = group: it emits the windows that end in its segments [y * spg, (y + 1) * spg)
};

// Windows of different labels are independent, so the segment axis is cut into groups that run as separate CTAs
// (blockIdx.y): a group first counts the depth - 1 segments before its own (their slots feed its first windows,
// nothing is emitted for them), then walks its own segments.
__global__ void __launch_bounds__(kCountThreads, kCountCtasPerSm)
count_tile_kernel(CountTileParams tp) {
    const StreamPlan& pl = tp.pl;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t P = (uint32_t)pl.P;
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw);                               // [kWsRing] records
    uint32_t* acc = ring + kWsRing;                                                        // [2][P] counts of the current segment
    uint32_t* slots = acc + 2 * P;                                                         // [depth][2P / 4] bytes, 4 cells a word
    uint64_t* full = reinterpret_cast<uint64_t*>(slots + (size_t)tp.depth * (2 * P / 4));

    const int tid = threadIdx.x, tile = blockIdx.x, lane = tid & 31;
    const uint32_t HW = (uint32_t)(pl.H * pl.W);
    const uint32_t pix0 = (uint32_t)tile * P;
    const uint32_t npix = min(P, HW - pix0);
    const uint32_t words = 2 * P / 4;                                                      // P is a multiple of 32
    const uint32_t* my_off = pl.off_rel + (int64_t)tile * (pl.TB + 1);
    const uint32_t* my_records = pl.records + pl.tile_base[tile];
    const uint32_t list_len = (pl.tile_total[tile] + 3u) & ~3u;
    const int n_chunks = (int)((list_len + kWsChunkRecords - 1) / kWsChunkRecords);
    const int TB = pl.TB;
    const int g_begin = (int)blockIdx.y * tp.seg_per_group;                               // first segment this CTA emits for
    const int g_end = min(TB, g_begin + tp.seg_per_group);
    const int g_walk = max(0, g_begin - (tp.depth - 1));                                   // first segment it counts
    if (g_begin >= TB) return;
    // the ring is indexed relative to the chunk that holds the first record of segment g_walk
    const int chunk0 = (int)(__ldg(my_off + g_walk) / kWsChunkRecords);
    const uint32_t rec0 = (uint32_t)chunk0 * kWsChunkRecords;
    auto issue = [&](int c) {           // thread 0 only; c counts from chunk0
        const uint32_t first = (uint32_t)(chunk0 + c) * kWsChunkRecords;
        const uint32_t bytes = min((uint32_t)kWsChunkRecords, list_len - first) * 4u;
        uint64_t* bar = full + (c % kWsStages);
        mbar_expect_tx(bar, bytes);
        tma_load_1d(ring + (c % kWsStages) * kWsChunkRecords, my_records + first, bytes, bar);
    };
    const int n_rel = n_chunks - chunk0;                                                   // chunks from chunk0 to the end of the list
    if (tid == 0) {
        for (int s = 0; s < kWsStages; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = tid; i < 2 * P; i += kCountThreads) acc[i] = 0u;
    __syncthreads();
    if (tid == 0)
        for (int c = 0; c < n_rel && c < kWsStages; ++c) issue(c);

    // record range and emission list of segment g: lane l of every warp holds those of segment
    // group*32 + l, the next group of 32 is loaded one group ahead
    auto load_group = [&](int first, uint32_t& lo, uint32_t& hi, CountSegment& info) {
        const int g = min(first + lane, TB - 1);
        lo = __ldg(my_off + g) - rec0; hi = __ldg(my_off + g + 1) - rec0;      // record positions relative to chunk0
        info = tp.segments[g];
    };
    uint32_t b_lo, b_hi, nb_lo = 0, nb_hi = 0;
    CountSegment b_info, nb_info = {};
    load_group(g_walk, b_lo, b_hi, b_info);

    int ready_chunk = -1, next_refill = kWsStages;
    for (int g = g_walk; g < g_end; ++g) {
        const int src = (g - g_walk) & 31;
        if (src == 0) {
            if (g != g_walk) { b_lo = nb_lo; b_hi = nb_hi; b_info = nb_info; }
            if (g + 32 < g_end) load_group(g + 32, nb_lo, nb_hi, nb_info);
        }
        const uint32_t o0 = __shfl_sync(0xFFFFFFFFu, b_lo, src), o1 = __shfl_sync(0xFFFFFFFFu, b_hi, src);
        const int emit_first = __shfl_sync(0xFFFFFFFFu, b_info.emit_first, src);
        const int emit_count = g >= g_begin ? __shfl_sync(0xFFFFFFFFu, b_info.emit_count, src) : 0;
        uint32_t cur = o0;
        while (cur < o1) {
            const uint32_t avail = (uint32_t)next_refill * kWsChunkRecords;     // records requested so far
            const uint32_t limit = o1 < avail ? o1 : avail;
            const int last_c = (int)((limit - 1) / kWsChunkRecords);
            while (ready_chunk < last_c) {
                ++ready_chunk;
                mbar_wait(full + (ready_chunk % kWsStages), (uint32_t)(ready_chunk / kWsStages) & 1u);
            }
            for (uint32_t r = cur + tid; r < limit; r += kCountThreads) {
                const uint32_t rec = ring[r & (kWsRing - 1)];                   // [ d:18 | local pixel:13 | p:1 ]
                atomicAdd(acc + (rec & 1u) * P + ((rec >> 1) & 0x1FFFu), 1u);    // :32, channel = polarity (:36)
            }
            cur = limit;
            if (cur < o1) {                                        // the segment outgrew the ring: recycle stages
                __syncthreads();
                const int drained = (int)(cur / kWsChunkRecords);
                if (tid == 0)
                    for (int r = next_refill; r < drained + kWsStages && r < n_rel; ++r) issue(r);
                next_refill = drained + kWsStages;
            }
        }
        __syncthreads();                                           // the segment is counted
        {
            const int drained = (int)(o1 / kWsChunkRecords);
            if (drained + kWsStages > next_refill) {
                if (tid == 0)
                    for (int r = next_refill; r < drained + kWsStages && r < n_rel; ++r) issue(r);
                next_refill = drained + kWsStages;
            }
        }
        // counts of the segment -> its ring slot (bytes, saturated), accumulator cleared
        uint32_t* slot = slots + (size_t)(g % tp.depth) * words;
        for (uint32_t i = tid; i < words; i += kCountThreads) {
            const uint4 c = reinterpret_cast<uint4*>(acc)[i];
            reinterpret_cast<uint4*>(acc)[i] = make_uint4(0u, 0u, 0u, 0u);
            slot[i] = min(c.x, 255u) | (min(c.y, 255u) << 8) | (min(c.z, 255u) << 16) | (min(c.w, 255u) << 24);
        }
        __syncthreads();
        // the windows that end with this segment
        for (int e = 0; e < emit_count; ++e) {
            const CountEmit em = tp.emits[emit_first + e];
            uint8_t* o = tp.frames + (int64_t)(emit_first + e) * tp.frame_stride + pix0;
            for (uint32_t i = tid; i < words; i += kCountThreads) {
                uint32_t sum = 0;
                for (int s = em.first_segment; s <= em.last_segment; ++s)
                    sum = __vaddus4(sum, slots[(size_t)(s % tp.depth) * words + i]);
                const uint32_t cell = i * 4u, pol = cell / P, lp = cell - pol * P;   // 4 cells never straddle a polarity
                if (lp < npix) *reinterpret_cast<uint32_t*>(o + (int64_t)pol * HW + lp) = sum;
            }
        }
        // (no barrier: the next segment writes `acc` only, and its slot store comes after a barrier)
    }
}

}  // namespace evrep

using namespace evrep;

static int count_stream_layout(int64_t n_events, int n_segments, int H, int W, Layout& L) {
    return make_layout(n_events, n_segments, n_segments, H, W, (int)batches_upper_bound(n_segments, n_segments), L, kCountTilesPerSm);
}

extern "C" {

int64_t evrep_count_stream_scratch_bytes(int64_t n_events, int n_segments, int n_emits, int H, int W) {
    if (n_events < 0 || n_segments < 0 || n_emits < 0 || H <= 0 || W <= 0) return EVREP_ERR_ARG;
    Layout L;
    int rc = count_stream_layout(n_events, n_segments, H, W, L);
    if (rc) return rc;
    return L.total + ((int64_t)sizeof(CountSegment) * (n_segments + 1) + (int64_t)sizeof(CountEmit) * (n_emits + 1) + 511) / 256 * 256;
}

int evrep_count_stream(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                       const evrep_count_segment* segments_host, int n_segments,
                       const evrep_count_emit* emits_host, int n_emits, int H, int W,
                       const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                       uint8_t* frames_out, int64_t frame_stride, void* scratch, int64_t scratch_bytes,
                       evrep_stream_t stream) {
    if (n_events < 0 || n_segments < 0 || n_emits < 0 || H <= 0 || W <= 0 || !scratch) return EVREP_ERR_ARG;
    if (n_emits == 0) return EVREP_OK;
    if (!segments_host || !emits_host || !frames_out || n_segments == 0) return EVREP_ERR_ARG;
    if (((int64_t)H * W) % 4 || frame_stride % 4 || (reinterpret_cast<uintptr_t>(frames_out) & 3)) return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    // segments are the "windows" of the bucketing front end, one bin each (the bin index never
    // depends on the timestamps: a single bin absorbs every t)
    std::vector<evrep_taf_window> wins((size_t)n_segments);
    for (int g = 0; g < n_segments; ++g) {
        wins[g].ev_begin = segments_host[g].ev_begin; wins[g].ev_end = segments_host[g].ev_end;
        wins[g].start_time = 0; wins[g].n_bins = 1; wins[g].fresh = 0;
    }
    // emissions ordered by their last segment; per segment the range of emissions that follow it
    std::vector<CountSegment> seg((size_t)n_segments);
    std::vector<CountEmit> emits((size_t)n_emits);
    int depth = 1, e = 0;
    for (int g = 0; g < n_segments; ++g) {
        seg[g].emit_first = e; seg[g].emit_count = 0;
        while (e < n_emits && emits_host[e].last_segment == g) {
            const int first = emits_host[e].first_segment;
            if (first < 0 || first > g + 1) return EVREP_ERR_ARG;          // first == last + 1: an empty window
            emits[e].first_segment = first; emits[e].last_segment = g;
            if (g - first + 1 > depth) depth = g - first + 1;
            ++seg[g].emit_count; ++e;
        }
    }
    if (e != n_emits) return EVREP_ERR_ARG;                                // not sorted by last_segment, or out of range
    StreamPlan pl;
    Layout L;
    int rc = prepare_stream(t, x, y, p, n_events, wins.data(), n_segments, (int)kDMax, H, W, xmap, ymap, sensor_h, sensor_w,
                            scratch, scratch_bytes, st, pl, L, kCountTilesPerSm);
    if (rc) return rc;
    const size_t smem = (size_t)kWsRing * 4 + (size_t)2 * L.P * 4 + (size_t)depth * 2 * L.P + 64;
    if (smem > (size_t)kCountMaxSmem) return EVREP_ERR_RANGE;              // windows span too many segments for the ring
    const int64_t seg_bytes = (int64_t)sizeof(CountSegment) * n_segments, emit_bytes = (int64_t)sizeof(CountEmit) * n_emits;
    if (scratch_bytes < L.total + seg_bytes + emit_bytes) return EVREP_ERR_SCRATCH;
    char* extra = reinterpret_cast<char*>(scratch) + L.total;
    rc = upload_words(reinterpret_cast<const uint32_t*>(seg.data()), seg_bytes / 4, reinterpret_cast<uint32_t*>(extra), st);
    if (rc) return rc;
    rc = upload_words(reinterpret_cast<const uint32_t*>(emits.data()), emit_bytes / 4, reinterpret_cast<uint32_t*>(extra + seg_bytes), st);
    if (rc) return rc;
    CountTileParams tp;
    tp.pl = pl; tp.segments = reinterpret_cast<const CountSegment*>(extra);
    tp.emits = reinterpret_cast<const CountEmit*>(extra + seg_bytes);
    tp.frames = frames_out; tp.frame_stride = frame_stride; tp.depth = depth;
    // groups of segments: enough CTAs for two rounds of what the SMs hold at once, each group long enough that the
    // depth - 1 segments it counts again for its first windows stay a small part of its work
    const int per_sm = (int)std::min<size_t>((size_t)kCountCtasPerSm, (size_t)(224 * 1024) / (smem + 1024));
    const int want_groups = std::max(1, 2 * per_sm * sm_count() / std::max(1, L.n_tiles));
    const int min_len = std::max(8, 4 * depth);
    int groups = std::max(1, std::min(want_groups, n_segments / min_len));
    tp.seg_per_group = (n_segments + groups - 1) / groups;
    groups = (n_segments + tp.seg_per_group - 1) / tp.seg_per_group;
    EVREP_CUDA(cudaFuncSetAttribute(count_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    count_tile_kernel<<<dim3((unsigned)L.n_tiles, (unsigned)groups), kCountThreads, smem, st>>>(tp);
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

}  // extern "C"
