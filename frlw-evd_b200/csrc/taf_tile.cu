// Temporal Active Focus over whole streams (generate_taf.py:160-238): the tile kernels.
//
// The bucketing passes (bucketing.cu) have sorted the events by (sensor tile, 10 ms bin) into
// packed 4-byte records.  One CTA per sensor tile (a contiguous range of <= 2304 pixels, chosen
// so that there are <= #SM tiles when possible) keeps the tile's FIFO state -- 2K floats per
// pixel -- in REGISTERS for the whole launch, streams its own record list through a ring of TMA
// bulk copies (cp.async.bulk + mbarrier), accumulates (count, sum d) per cell with shared-memory
// atomics, applies the FIFO push / ageing rule bin by bin, and writes the [2K,H,W] tensor once
// per window with TMA bulk stores.  Tiles never talk to each other: the only cross-tile fact,
// "did any pixel see an event in this bin" (generate_taf.py:40-41), is a per-bin flag produced by
// the bucketing pass.
//
// Three kernels: taf_tile_ws_kernel (warp specialised, the default), taf_tile_kernel (single role; fallback when the
// staging tile does not fit next to two accumulator buffers) and taf_tile_pk_kernel (packed accumulators, opt-in).
//
// HBM-bound byte/float work: no tensor cores.  Sums of d are exact integers, so the result does
// not depend on the order in which records are accumulated.
#include "stream_common.cuh"

#include <type_traits>

namespace evrep {

struct TileParams {
    StreamPlan pl;
    float* state;          // [H,W,2,K]
    float* out;            // window w at out + w * out_stride
    int64_t out_stride;
    int emit_state;        // write the state after every window (always after the last)
    int n_emits;           // number of windows (batches that end a window)
    int bulk_out;          // out rows are 16-byte aligned: emit through smem + TMA bulk stores
    float span;            // f32(abin + 1e-8)
    const uint32_t* src;   // bin-major records (kRuns kernels): first record of the (tile, bin) run, [n_tiles][TB]
    uint32_t pk_hotmask[33];   // packed accumulators with c count bits: count bits that mark a cell whose sum of d may not be exact
};

// Shared-memory carve-up of the tile kernel (all offsets multiples of 128 bytes).
struct TileSmem {
    int ring, acc, stage, bars, off, any, meta, total;
    __host__ __device__ TileSmem(int P, int K) {
        int o = 0;
        ring = o;  o += kStages * kChunkRecords * 4;
        acc = o;   o += 2 * P * (int)sizeof(uint2);                 // {n, sum d} per (pixel, polarity)
        stage = o; o += 2 * K * P * 4;                              // [2K][P] output staging
        bars = o;  o += 128;
        off = o;   o += 2 * (kBatchBins + 1) * 4; o = (o + 127) / 128 * 128;
        any = o;   o += 2 * kBatchBins * 4;
        meta = o;  o += 2 * (int)sizeof(Batch); o = (o + 127) / 128 * 128;
        total = o;
    }
};

template <int K, int SLOTS>
__global__ void __launch_bounds__(kTafThreads, 1)
taf_tile_kernel(TileParams tp) {
    const StreamPlan& pl = tp.pl;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmem lay(pl.P, K);
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw + lay.ring);     // [kStages][kChunkRecords]
    uint2* acc = reinterpret_cast<uint2*>(smem_raw + lay.acc);             // [2P]
    float* stage = reinterpret_cast<float*>(smem_raw + lay.stage);         // [2K][P]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);
    uint32_t* s_off = reinterpret_cast<uint32_t*>(smem_raw + lay.off);     // [2][kBatchBins+1]
    uint32_t* s_any = reinterpret_cast<uint32_t*>(smem_raw + lay.any);     // [2][kBatchBins]
    Batch* s_meta = reinterpret_cast<Batch*>(smem_raw + lay.meta);         // [2]

    const int tid = threadIdx.x, tile = blockIdx.x;
    const int64_t HW = (int64_t)pl.H * pl.W;
    const int64_t pix0 = (int64_t)tile * pl.P;
    const int npix = (int)min((int64_t)pl.P, HW - pix0);
    const uint32_t* my_records = pl.records + pl.tile_base[tile];
    const uint32_t list_len = (pl.tile_total[tile] + 3u) & ~3u;
    const int n_chunks = (int)((list_len + kChunkRecords - 1) / kChunkRecords);
    const uint32_t* my_off = pl.off_rel + (int64_t)tile * (pl.TB + 1);

    auto issue = [&](int c) {           // thread 0 only
        const uint32_t first = (uint32_t)c * kChunkRecords;
        const uint32_t bytes = min((uint32_t)kChunkRecords, list_len - first) * 4u;
        uint64_t* bar = full + (c % kStages);
        mbar_expect_tx(bar, bytes);
        tma_load_1d(ring + (c % kStages) * kChunkRecords, my_records + first, bytes, bar);
    };

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 2 * pl.P; i += kTafThreads) acc[i] = make_uint2(0u, 0u);
    if (tid == 0) s_meta[0] = pl.batches[0];
    __syncthreads();
    if (tid == 0)
        for (int c = 0; c < n_chunks && c < kStages; ++c) issue(c);
    {
        const Batch m0 = s_meta[0];
        if (tid <= m0.nb) s_off[tid] = my_off[m0.gbin0 + tid];
        if (tid < m0.nb) s_any[tid] = pl.bin_any[m0.gbin0 + tid];
    }

    // FIFO state of this thread's pixels, as float2 pairs for the packed f32x2 adds of
    // sm_100: element k of (slot, polarity) is v[s][p][k / 2].{x,y}; k = K-1 is the newest.
    static_assert(K % 4 == 0, "K must be a multiple of 4");
    float2 v[SLOTS][2][K / 2];
    const bool first_fresh = (pl.batches[0].flags & 1) != 0;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int lp = s * kTafThreads + tid;
        if (lp < npix && !first_fresh) {
            const float4* src = reinterpret_cast<const float4*>(tp.state + (pix0 + lp) * 2 * K);
#pragma unroll
            for (int q = 0; q < 2 * K / 4; ++q) {
                const float4 f = src[q];
                v[s][(q * 4) / K][((q * 4) % K) / 2 + 0] = make_float2(f.x, f.y);
                v[s][(q * 4) / K][((q * 4) % K) / 2 + 1] = make_float2(f.z, f.w);
            }
        } else {
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int k = 0; k < K / 2; ++k) v[s][p][k] = make_float2(kTafInit, kTafInit);
        }
    }
    __syncthreads();

    int ready_chunk = -1;      // highest chunk this thread has observed complete
    int next_refill = kStages; // next chunk to load; its stage is free once chunk (next_refill - kStages) is drained
    bool staged_once = false;
    // batch descriptors are fetched two batches ahead, their offsets / flags one batch ahead
    Batch nmeta = pl.batches[pl.n_batches > 1 ? 1 : 0];
    for (int j = 0; j < pl.n_batches; ++j) {
        const int buf = j & 1;
        const Batch meta = s_meta[buf];
        uint32_t pre_off = 0, pre_any = 0;
        Batch nnmeta = nmeta;
        if (j + 1 < pl.n_batches) {
            if (tid <= nmeta.nb) pre_off = my_off[nmeta.gbin0 + tid];
            if (tid < nmeta.nb) pre_any = pl.bin_any[nmeta.gbin0 + tid];
            if (j + 2 < pl.n_batches) nnmeta = pl.batches[j + 2];
        }
        if (meta.flags & 1) {
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int k = 0; k < K / 2; ++k) v[s][p][k] = make_float2(kTafInit, kTafInit);
        }
        for (int b = 0; b < meta.nb; ++b) {
            const uint32_t o0 = s_off[buf * (kBatchBins + 1) + b], o1 = s_off[buf * (kBatchBins + 1) + b + 1];
            if (!s_any[buf * kBatchBins + b]) continue;          // nobody saw an event: no ageing
            const bool have = o1 > o0;
            if (have) {
                uint32_t cur = o0;
                while (cur < o1) {
                    const int c = (int)(cur / kChunkRecords);
                    const uint32_t chunk_end = (uint32_t)(c + 1) * kChunkRecords;
                    const uint32_t seg_end = o1 < chunk_end ? o1 : chunk_end;
                    if (c >= next_refill) {
                        // a single bin longer than the whole ring: recycle drained stages now
                        __syncthreads();
                        if (tid == 0)
                            for (int r = next_refill; r <= c && r < n_chunks; ++r) issue(r);
                        next_refill = c + 1;
                    }
                    if (c > ready_chunk) { mbar_wait(full + (c % kStages), (uint32_t)(c / kStages) & 1u); ready_chunk = c; }
                    const uint32_t* chunk = ring + (c % kStages) * kChunkRecords;
                    for (uint32_t r = cur + tid; r < seg_end; r += kTafThreads) {
                        const uint32_t rec = chunk[r & (kChunkRecords - 1)];
                        uint2* cell = acc + (rec & 0x3FFFu);         // 2 * local pixel + p
                        atomicAdd(&cell->x, 1u);
                        atomicAdd(&cell->y, rec >> 14);
                    }
                    cur = seg_end;
                }
                __syncthreads();                                  // all records of the bin are in `acc`
                // every chunk that ends at or before o1 is drained: refill those ring stages
                const int drained = (int)(o1 / kChunkRecords);    // chunks [0, drained) fully consumed
                if (tid == 0)
                    for (int r = next_refill; r < drained + kStages && r < n_chunks; ++r) issue(r);
                if (drained + kStages > next_refill) next_refill = drained + kStages;
            }
            const float2 minus1 = make_float2(-1.0f, -1.0f);
            if (!have) {
                // the tile saw nothing in this bin, but some other tile did: everything ages
#pragma unroll
                for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                    for (int p = 0; p < 2; ++p)
#pragma unroll
                        for (int k = 0; k < K / 2; ++k) v[s][p][k] = __fadd2_rn(v[s][p][k], minus1);
            } else {
                // read and clear this thread's accumulators, then release `acc` for the next bin
                // BEFORE the arithmetic: the long update phase runs without a barrier behind it
                uint4 a[SLOTS];
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kTafThreads + tid;
                    a[s] = make_uint4(0u, 0u, 0u, 0u);
                    // every slot but the last lies inside the tile's accumulator array
                    if (s < SLOTS - 1 || lp < pl.P) {
                        a[s] = *reinterpret_cast<uint4*>(acc + 2 * lp);           // {n0, S0, n1, S1}
                        if (a[s].x | a[s].z) *reinterpret_cast<uint4*>(acc + 2 * lp) = make_uint4(0u, 0u, 0u, 0u);
                    }
                }
                __syncthreads();
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const uint32_t nn[2] = {a[s].x, a[s].z}, ss[2] = {a[s].y, a[s].w};
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        // mean(t_norm) - 1 = S / (n span) - 1 (generate_taf.py:23-27); for n == 0
                        // the value is NaN and is never selected
                        const bool active = nn[p] != 0u;
                        float r;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)nn[p] * tp.span));
                        const float mean = fmaf((float)ss[p], r, -1.0f);
                        float2 aged[K / 2];
#pragma unroll
                        for (int k = 0; k < K / 2; ++k) aged[k] = __fadd2_rn(v[s][p][k], minus1);
#pragma unroll
                        for (int k = 0; k < K / 2; ++k) {
                            const float next = (k + 1 < K / 2) ? aged[k + 1].x : mean;
                            v[s][p][k].x = active ? aged[k].y : aged[k].x;
                            v[s][p][k].y = active ? next : aged[k].y;
                        }
                    }
                }
            }
        }
        if (meta.flags & 2) {
            const bool write_state = tp.emit_state || (j == pl.n_batches - 1);
            float* o = tp.out + (int64_t)meta.win * tp.out_stride + pix0;
            if (tp.bulk_out) {
                // [2K][npix] staging tile -> one TMA bulk store per channel row
                if (staged_once) {
                    if (tid < 2 * K) bulk_wait_read();             // previous window's rows have left smem
                    __syncthreads();
                }
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kTafThreads + tid;
                    if (s == SLOTS - 1 && lp >= pl.P) continue;     // columns >= npix are staged but never stored
#pragma unroll
                    for (int k = 0; k < K / 2; ++k)
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            stage[(4 * k + p) * pl.P + lp] = v[s][p][k].x;
                            stage[(4 * k + 2 + p) * pl.P + lp] = v[s][p][k].y;
                        }
                }
                fence_async_smem();
                __syncthreads();
                if (tid < 2 * K) {
                    bulk_store_1d(o + (int64_t)tid * HW, stage + tid * pl.P, (uint32_t)npix * 4u);
                    bulk_commit();
                }
                staged_once = true;
            } else {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kTafThreads + tid;
                    if (lp >= npix) continue;
#pragma unroll
                    for (int k = 0; k < K / 2; ++k)
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            __stcs(o + (int64_t)(4 * k + p) * HW + lp, v[s][p][k].x);
                            __stcs(o + (int64_t)(4 * k + 2 + p) * HW + lp, v[s][p][k].y);
                        }
                }
            }
            if (write_state) {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kTafThreads + tid;
                    if (lp >= npix) continue;
                    float4* dst = reinterpret_cast<float4*>(tp.state + (pix0 + lp) * 2 * K);
#pragma unroll
                    for (int q = 0; q < 2 * K / 4; ++q) {
                        const float2 lo = v[s][(q * 4) / K][((q * 4) % K) / 2], hi = v[s][(q * 4) / K][((q * 4) % K) / 2 + 1];
                        dst[q] = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
                }
            }
        }
        if (j + 1 < pl.n_batches) {
            const int nb = buf ^ 1;
            if (tid == 0) s_meta[nb] = nmeta;
            if (tid <= nmeta.nb) s_off[nb * (kBatchBins + 1) + tid] = pre_off;
            if (tid < nmeta.nb) s_any[nb * kBatchBins + tid] = pre_any;
            nmeta = nnmeta;
        }
        __syncthreads();
    }
    if (tp.bulk_out && tid < 2 * K) bulk_wait_all();               // smem must outlive the bulk reads
}

// ---- warp-specialised tile kernel -------------------------------------------------------------
// Same algorithm as taf_tile_kernel, split into two roles so that the latency-bound record
// bookkeeping runs ahead of, and concurrently with, the arithmetic:
//   * producer warpgroup (warps 0-3, 56 registers after setmaxnreg.dec).  Three accumulate warps
//     walk the bins, wait for the TMA ring, accumulate (n, sum d) of the next bins into one of two
//     accumulator buffers and refill the ring; the fourth, the store warp, sends every staged
//     window tensor with TMA bulk stores;
//   * consumer warpgroups (warps 4-15, 152 registers after setmaxnreg.inc): own the FIFO state of
//     the tile (6 pixels per thread), read + clear the accumulator of a bin, apply the update, and
//     copy the window tensor into the staging tile.  They never synchronise among themselves.
// Hand-over uses named barriers (bar.arrive / bar.sync): FULL[buf] accumulate -> consumers,
// EMPTY[buf] consumers -> accumulate, STAGED consumers -> store warp, STAGE_FREE store warp ->
// consumers.  4 warps per SM sub-partition: 1 producer-group warp + 3 consumers.
// kRuns: the records are laid out bin-major (slices.cu, bin-major sort): a tile's list is the concatenation of its
// per-bin runs, each padded to 16 bytes with null records, and lives at src[tile][bin] instead of one contiguous
// range.  The ring, the offsets (off_rel counts the padded records) and the roles are the same; a 2 KB chunk of the
// list is fetched with one bulk copy per run it touches (one or two).
template <int K, int SLOTS, bool kRuns>
__global__ void __launch_bounds__(kWsThreads, 1)
taf_tile_ws_kernel(TileParams tp) {
    const StreamPlan& pl = tp.pl;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmemWS lay(pl.P, K);
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw + lay.ring);     // [kWsStages][kWsChunkRecords]
    uint2* acc = reinterpret_cast<uint2*>(smem_raw + lay.acc);             // [2][2P]
    float* stage = reinterpret_cast<float*>(smem_raw + lay.stage);         // [2K][P]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);     // [kWsStages] record chunks
    uint64_t* staged = full + kWsStages;                                   // consumers -> store warp
    uint64_t* stage_free = staged + 1;                                     // store warp -> consumers
    volatile uint32_t* drained_w = reinterpret_cast<volatile uint32_t*>(stage_free + 1);   // ring chunks the accumulate warps are done with

    const int tid = threadIdx.x, tile = blockIdx.x;
    const int64_t HW = (int64_t)pl.H * pl.W;
    const int64_t pix0 = (int64_t)tile * pl.P;
    const int npix = (int)min((int64_t)pl.P, HW - pix0);
    const uint32_t* my_off = pl.off_rel + (int64_t)tile * (pl.TB + 1);

    if (tid == 0) {
        for (int s = 0; s < kWsStages; ++s) mbar_init(full + s, 1);
        mbar_init(staged, kConsumerThreads / 32);
        mbar_init(stage_free, 1);
        *drained_w = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 4 * pl.P; i += kWsThreads) acc[i] = make_uint2(0u, 0u);
    __syncthreads();

    if (tid < kProducerThreads) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (tid >= kAccumThreads) {
            // ================================== store warp ==================================
            // Keeps the record ring full -- TMA loads into the stages the accumulate warps have drained (they publish
            // `drained_w`; issuing the copies from accumulate thread 0 cost that warp ~350 cycles per bin) -- and sends
            // every window tensor the consumers stage with TMA bulk stores.  kRuns: the accumulate warps issue the loads.
            const int lane = tid - kAccumThreads;
            const uint32_t* my_records = pl.records + (kRuns ? 0u : pl.tile_base[tile]);
            const uint32_t list_len = (pl.tile_total[tile] + 3u) & ~3u;
            const int n_chunks = kRuns ? 0 : (int)((list_len + kWsChunkRecords - 1) / kWsChunkRecords);
            const int n_emits = tp.bulk_out ? tp.n_emits : 0;
            int issued = 0, emitted = 0, j = 0;
            while (issued < n_chunks || emitted < n_emits) {
                bool worked = false;
                if (issued < n_chunks) {
                    const int limit = min(n_chunks, (int)*drained_w + kWsStages);
                    if (limit > issued) {
                        if (lane == 0)
                            for (int c = issued; c < limit; ++c) {
                                const uint32_t first = (uint32_t)c * kWsChunkRecords;
                                const uint32_t bytes = min((uint32_t)kWsChunkRecords, list_len - first) * 4u;
                                uint64_t* bar = full + (c % kWsStages);
                                mbar_expect_tx(bar, bytes);
                                tma_load_1d(ring + (c % kWsStages) * kWsChunkRecords, my_records + first, bytes, bar);
                                // DRAM reads queue behind the window stores for microseconds and the ring holds 16 KB:
                                // pull the chunk kWsPrefetch ahead into L2 now
                                const uint32_t ahead = first + kWsPrefetch * kWsChunkRecords;
                                if (ahead < list_len)
                                    l2_prefetch(my_records + ahead, min((uint32_t)kWsChunkRecords, list_len - ahead) * 4u);
                            }
                        __syncwarp();
                        issued = limit;
                        worked = true;
                    }
                }
                if (emitted < n_emits) {
                    const bool ready = __shfl_sync(0xFFFFFFFFu, (int)mbar_test(staged, (uint32_t)emitted & 1u), 0) != 0;
                    if (ready) {
                        while (!(pl.batches[j].flags & 2)) ++j;
                        const Batch m = pl.batches[j];
                        ++j;
                        if (lane < 2 * K) {
                            float* o = tp.out + (int64_t)m.win * tp.out_stride + pix0;
                            bulk_store_1d(o + (int64_t)lane * HW, stage + lane * pl.P, (uint32_t)npix * 4u);
                            bulk_commit();
                            bulk_wait_read();                        // the rows have left shared memory
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(stage_free);
                        ++emitted;
                        worked = true;
                    }
                }
                if (!worked) __nanosleep(32);
            }
            if (lane < 2 * K) bulk_wait_all();
            return;
        }
        // ================================ accumulate warps ================================
        const uint32_t* my_records = kRuns ? pl.records : pl.records + pl.tile_base[tile];
        const uint32_t list_len = (pl.tile_total[tile] + 3u) & ~3u;
        const int n_chunks = (int)((list_len + kWsChunkRecords - 1) / kWsChunkRecords);
        // kRuns, thread 0: cursor over the tile's runs -- bin `ib` holds list positions [v_lo, v_hi) at record src_ib;
        // the next bin's end and source are fetched one bin ahead
        const uint32_t* my_src = kRuns ? tp.src + (int64_t)tile * pl.TB : nullptr;
        int ib = 0;
        uint32_t v_lo = 0, v_hi = 0, src_ib = 0, v_hi_next = 0, src_next = 0;
        if (kRuns && tid == 0 && pl.TB > 0) {
            v_hi = my_off[1]; src_ib = my_src[0];
            if (pl.TB > 1) { v_hi_next = my_off[2]; src_next = my_src[1]; }
        }
        auto issue = [&](int c) {           // thread 0 only; chunks are requested in increasing order
            const uint32_t first = (uint32_t)c * kWsChunkRecords;
            const uint32_t bytes = min((uint32_t)kWsChunkRecords, list_len - first) * 4u;
            uint64_t* bar = full + (c % kWsStages);
            mbar_expect_tx(bar, bytes);
            if (!kRuns) {
                tma_load_1d(ring + (c % kWsStages) * kWsChunkRecords, my_records + first, bytes, bar);
                return;
            }
            uint32_t cur = first;
            const uint32_t end = first + bytes / 4u;
            while (cur < end) {
                while (v_hi <= cur) {               // the run of bin ib is used up (or empty): on to the next bin
                    ++ib;
                    v_lo = v_hi; v_hi = v_hi_next; src_ib = src_next;
                    if (ib + 1 < pl.TB) { v_hi_next = __ldg(my_off + ib + 2); src_next = __ldg(my_src + ib + 1); }
                }
                const uint32_t piece_end = v_hi < end ? v_hi : end;
                tma_load_1d(ring + (c % kWsStages) * kWsChunkRecords + (cur - first), my_records + src_ib + (cur - v_lo),
                            (piece_end - cur) * 4u, bar);
                cur = piece_end;
            }
        };
        if (kRuns && tid == 0)
            for (int c = 0; c < n_chunks && c < kWsStages; ++c) issue(c);
        BatchFeed feed;
        feed.init(smem_raw + lay.feed_p, &pl, my_off, tid, kBarProducers, kAccumThreads);
        int ready_chunk = -1, next_refill = kWsStages, buf = 0;
        int uses0 = 0, uses1 = 0;
        constexpr int kHandOver = kAccumThreads + kConsumerThreads;
        for (int j = 0; j < pl.n_batches; ++j) {
            const Batch meta = feed.begin(j);
            const int jb = j & 1;
            for (int b = 0; b < meta.nb; ++b) {
                const uint32_t o0 = feed.s_off[jb * (kBatchBins + 1) + b], o1 = feed.s_off[jb * (kBatchBins + 1) + b + 1];
                if (!feed.s_any[jb * kBatchBins + b] || o1 <= o0) continue;
                uint2* my_acc = acc + buf * 2 * pl.P;
                const int last_c = (int)((o1 - 1) / kWsChunkRecords);
                const bool had_use = (buf ? uses1 : uses0) > 0;
                if (last_c < next_refill) {
                    // common case: every chunk of the bin is already in flight
                    while (ready_chunk < last_c) {
                        ++ready_chunk;
                        mbar_wait(full + (ready_chunk % kWsStages), (uint32_t)(ready_chunk / kWsStages) & 1u);
                    }
                    // records are pulled into registers BEFORE waiting for the accumulator buffer
                    constexpr int kPre = 8;
                    constexpr uint32_t kNoRec = 0xFFFFFFFFu;        // d = 2^18-1, pixel 8191: never produced for P <= 2560
                    uint32_t pre[kPre];
                    // rounds of kAccumThreads records the bin needs (uniform): a small bin (GEN1: ~70
                    // records) takes a two-round path instead of walking all eight rounds of the unrolled code
                    const int n_pre = (int)min((uint32_t)kPre, (o1 - o0 + kAccumThreads - 1) / kAccumThreads);
                    // only compiled into the small-tile instantiations (GEN1-size grids): on large tiles the
                    // extra branch alone cost 4 % of the kernel
                    const bool small_bin = SLOTS <= 2 && n_pre <= 2;
                    if (small_bin) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const uint32_t r = o0 + tid + i * kAccumThreads;
                            pre[i] = (i < n_pre && r < o1) ? ring[r & (kWsRing - 1)] : kNoRec;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < kPre; ++i) {
                            const uint32_t r = o0 + tid + i * kAccumThreads;
                            pre[i] = r < o1 ? ring[r & (kWsRing - 1)] : kNoRec;
                        }
                    }
                    const bool all_pre = (o1 - o0) <= (uint32_t)(kPre * kAccumThreads);
                    // the consumers must have drained this buffer (its first use needs no wait); the
                    // barrier also tells that every accumulate thread is done with all earlier bins
                    if (had_use) named_sync(kBarEmpty0 + buf, kHandOver);
                    else named_sync(kBarProducers, kAccumThreads);
                    {
                        // ring stages whose chunk ends before the first record still to be read are free
                        const int drained = (int)((all_pre ? o1 : o0) / kWsChunkRecords);
                        if (kRuns) {
                            if (tid == 0)
                                for (int r = next_refill; r < drained + kWsStages && r < n_chunks; ++r) issue(r);
                        } else if (tid == 0) {
                            *drained_w = (uint32_t)drained;           // the store warp refills them
                        }
                        if (drained + kWsStages > next_refill) next_refill = drained + kWsStages;
                    }
                    if (small_bin) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            if (pre[i] != kNoRec) {
                                uint2* cell = my_acc + (pre[i] & 0x3FFFu);   // 2 * local pixel + p
                                atomicAdd(&cell->x, 1u);
                                atomicAdd(&cell->y, pre[i] >> 14);
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < kPre; ++i) {
                            if (pre[i] != kNoRec) {
                                uint2* cell = my_acc + (pre[i] & 0x3FFFu);
                                atomicAdd(&cell->x, 1u);
                                atomicAdd(&cell->y, pre[i] >> 14);
                            }
                        }
                    }
                    for (uint32_t r = o0 + tid + kPre * kAccumThreads; r < o1; r += kAccumThreads) {
                        const uint32_t rec = ring[r & (kWsRing - 1)];
                        if (kRuns && rec == kNoRec) continue;           // run padding
                        uint2* cell = my_acc + (rec & 0x3FFFu);
                        atomicAdd(&cell->x, 1u);
                        atomicAdd(&cell->y, rec >> 14);
                    }
                } else {
                    // a single bin longer than the ring: go chunk by chunk, recycling drained stages
                    if (had_use) named_sync(kBarEmpty0 + buf, kHandOver);
                    uint32_t cur = o0;
                    while (cur < o1) {
                        const int c = (int)(cur / kWsChunkRecords);
                        const uint32_t chunk_end = (uint32_t)(c + 1) * kWsChunkRecords;
                        const uint32_t seg_end = o1 < chunk_end ? o1 : chunk_end;
                        if (c >= next_refill) {
                            named_sync(kBarProducers, kAccumThreads);
                            if (kRuns) {
                                if (tid == 0)
                                    for (int r = next_refill; r <= c && r < n_chunks; ++r) issue(r);
                                next_refill = c + 1;
                            } else {
                                // everything before `cur` has been read by every accumulate thread
                                if (tid == 0) *drained_w = cur / kWsChunkRecords;
                                next_refill = (int)(cur / kWsChunkRecords) + kWsStages;
                            }
                        }
                        while (ready_chunk < c) {
                            ++ready_chunk;
                            mbar_wait(full + (ready_chunk % kWsStages), (uint32_t)(ready_chunk / kWsStages) & 1u);
                        }
                        for (uint32_t r = cur + tid; r < seg_end; r += kAccumThreads) {
                            const uint32_t rec = ring[r & (kWsRing - 1)];
                            if (kRuns && rec == 0xFFFFFFFFu) continue;      // run padding
                            uint2* cell = my_acc + (rec & 0x3FFFu);
                            atomicAdd(&cell->x, 1u);
                            atomicAdd(&cell->y, rec >> 14);
                        }
                        cur = seg_end;
                    }
                }
                named_arrive(kBarFull0 + buf, kHandOver);            // hand the accumulator to the consumers
                if (buf) ++uses1; else ++uses0;
                buf ^= 1;
            }
            feed.end(j);
        }
        // match the consumers' last EMPTY arrivals so that no barrier phase is left open
        if (uses0 > 0) named_sync(kBarEmpty0, kHandOver);
        if (uses1 > 0) named_sync(kBarEmpty1, kHandOver);
        return;
    }

    // =================================== consumer warpgroups ===================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int ctid = tid - kProducerThreads;
    constexpr int kHandOver = kAccumThreads + kConsumerThreads;
    static_assert(K % 4 == 0, "K must be a multiple of 4");
    float2 v[SLOTS][2][K / 2];
    const bool first_fresh = (pl.batches[0].flags & 1) != 0;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int lp = s * kConsumerThreads + ctid;
        if (lp < npix && !first_fresh) {
            const float4* src = reinterpret_cast<const float4*>(tp.state + (pix0 + lp) * 2 * K);
#pragma unroll
            for (int q = 0; q < 2 * K / 4; ++q) {
                const float4 f = src[q];
                v[s][(q * 4) / K][((q * 4) % K) / 2 + 0] = make_float2(f.x, f.y);
                v[s][(q * 4) / K][((q * 4) % K) / 2 + 1] = make_float2(f.z, f.w);
            }
        } else {
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int k = 0; k < K / 2; ++k) v[s][p][k] = make_float2(kTafInit, kTafInit);
        }
    }
    // batch descriptors and the tile's per-batch bit masks come straight from global memory,
    // one batch ahead: the consumers never synchronise among themselves
    const uint32_t* my_bits = pl.tile_bits + (int64_t)tile * pl.n_batches;
    Batch meta = pl.batches[0];
    uint32_t bits = my_bits[0];
    int buf = 0, emitted = 0;
    const float2 minus1 = make_float2(-1.0f, -1.0f);
    for (int j = 0; j < pl.n_batches; ++j) {
        Batch nmeta = meta;
        uint32_t nbits = 0;
        if (j + 1 < pl.n_batches) { nmeta = pl.batches[j + 1]; nbits = my_bits[j + 1]; }
        if (meta.flags & 1) {
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int k = 0; k < K / 2; ++k) v[s][p][k] = make_float2(kTafInit, kTafInit);
        }
        for (int b = 0; b < meta.nb; ++b) {
            if (!((bits >> (16 + b)) & 1u)) continue;               // nobody saw an event: no ageing
            if (!((bits >> b) & 1u)) {
                // the tile saw nothing in this bin, but some other tile did: everything ages
#pragma unroll
                for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                    for (int p = 0; p < 2; ++p)
#pragma unroll
                        for (int k = 0; k < K / 2; ++k) v[s][p][k] = __fadd2_rn(v[s][p][k], minus1);
                continue;
            }
            named_sync(kBarFull0 + buf, kHandOver);                  // the producers filled this accumulator
            uint2* my_acc = acc + buf * 2 * pl.P;
            uint4 a[SLOTS];
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int lp = s * kConsumerThreads + ctid;
                a[s] = make_uint4(0u, 0u, 0u, 0u);
                if (s < SLOTS - 1 || lp < pl.P) {                    // every slot but the last lies inside the array
                    a[s] = *reinterpret_cast<uint4*>(my_acc + 2 * lp);               // {n0, S0, n1, S1}
                    if (a[s].x | a[s].z) *reinterpret_cast<uint4*>(my_acc + 2 * lp) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            named_arrive(kBarEmpty0 + buf, kHandOver);               // clean again: give it back
            buf ^= 1;
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const uint32_t nn[2] = {a[s].x, a[s].z}, ss[2] = {a[s].y, a[s].w};
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    // mean(t_norm) - 1 = S / (n span) - 1 (generate_taf.py:23-27); NaN for n == 0, never selected
                    const bool active = nn[p] != 0u;
                    float r;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)nn[p] * tp.span));
                    const float mean = fmaf((float)ss[p], r, -1.0f);
                    float2 aged[K / 2];
#pragma unroll
                    for (int k = 0; k < K / 2; ++k) aged[k] = __fadd2_rn(v[s][p][k], minus1);
#pragma unroll
                    for (int k = 0; k < K / 2; ++k) {
                        const float next = (k + 1 < K / 2) ? aged[k + 1].x : mean;
                        v[s][p][k].x = active ? aged[k].y : aged[k].x;
                        v[s][p][k].y = active ? next : aged[k].y;
                    }
                }
            }
        }
        if (meta.flags & 2) {
            const bool write_state = tp.emit_state || (j == pl.n_batches - 1);
            if (tp.bulk_out) {
                // stage the [2K][P] tile; the store warp sends it.  No consumer waits for another:
                // each warp streams its columns, fences, signals and moves on to the next bin.
                if (emitted > 0) mbar_wait(stage_free, (uint32_t)(emitted - 1) & 1u);
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kConsumerThreads + ctid;
                    if (s == SLOTS - 1 && lp >= pl.P) continue;     // columns >= npix are staged but never stored
#pragma unroll
                    for (int k = 0; k < K / 2; ++k)
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            stage[(4 * k + p) * pl.P + lp] = v[s][p][k].x;
                            stage[(4 * k + 2 + p) * pl.P + lp] = v[s][p][k].y;
                        }
                }
                fence_async_smem();
                __syncwarp();
                if ((ctid & 31) == 0) mbar_arrive(staged);
                ++emitted;
            } else {
                float* o = tp.out + (int64_t)meta.win * tp.out_stride + pix0;
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kConsumerThreads + ctid;
                    if (lp >= npix) continue;
#pragma unroll
                    for (int k = 0; k < K / 2; ++k)
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            __stcs(o + (int64_t)(4 * k + p) * HW + lp, v[s][p][k].x);
                            __stcs(o + (int64_t)(4 * k + 2 + p) * HW + lp, v[s][p][k].y);
                        }
                }
            }
            if (write_state) {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kConsumerThreads + ctid;
                    if (lp >= npix) continue;
                    float4* dst = reinterpret_cast<float4*>(tp.state + (pix0 + lp) * 2 * K);
#pragma unroll
                    for (int q = 0; q < 2 * K / 4; ++q) {
                        const float2 lo = v[s][(q * 4) / K][((q * 4) % K) / 2], hi = v[s][(q * 4) / K][((q * 4) % K) / 2 + 1];
                        dst[q] = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
                }
            }
        }
        meta = nmeta;
        bits = nbits;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// taf_tile_pk_kernel: the warp-specialised kernel with PACKED accumulators (opt-in: EVREP_TAF_TILE_KERNEL=pk).
//
// Measured on the ws kernel with in-kernel cycle counters (tools/diag_tile_timing.py): the accumulate warps, not the
// consumers, were the critical path -- two shared-memory atomics per record on 8-byte cells (16 usable banks: ~5-way
// conflicts), the TMA refills issued by accumulate thread 0 (~350 cycles per bin) and record chunks that arrived late
// behind the bursts of window stores (16 KB ring).  Here:
//   * one accumulator word per (pixel, polarity): {sum d : 32 - c | n : c}, ONE atomic per record on 4-byte cells.  c is
//     chosen per (tile, bin) hand-over so that the count field cannot overflow (c = 10 up to 1023 records, one more bit
//     per doubling); the sum field is exact whenever n <= nmax(c) = (2^(32-c) - 1) / (abin - 1) (419 for 10 ms bins and
//     c = 10), and a consumer warp that finds a larger count (a hot pixel) recomputes that cell's sum from the bin's
//     records in global memory, so results stay exact for every input;
//   * the accumulators are half as large, which pays for THREE buffers (the accumulate warps run two bins ahead) and a
//     32 KB record ring;
//   * the store warp issues the record TMA loads as well (it polls a `drained` word the accumulate warps publish), and
//     all hand-overs are mbarriers: nobody but the waiting side blocks.
// The arithmetic (FIFO push / ageing, window tensor, state) is the ws kernel's, bit for bit.
// Measured (tools/taf_tile_variants.py, profiles/r2_tile_variants.txt): 0.78 vs 1.06 ms at 30 Mev/s (2000 records per
// (tile, bin): the ws kernel's accumulate warps are the limit there), but 1.12 vs 1.04 ms on the 10 Mev/s headline stream
// (the consumers are bound by the ALU pipe either way and unpacking costs them two more instructions per cell), and several
// times slower when moving edges put >= 32 events per bin on many cells (every such cell is rescanned).  Hence opt-in.
constexpr int kPkStages = 16;
constexpr int kPkRing = kWsChunkRecords * kPkStages;
static_assert((kPkRing & (kPkRing - 1)) == 0, "ring size must be a power of two");
constexpr int kPkBufs = 3;
constexpr int kPkMinCountBits = 10;

struct TileSmemPK {
    int ring, acc, stage, bars, info, feed_p, total;
    __host__ __device__ TileSmemPK(int P, int K) {
        int o = 0;
        ring = o;   o += kPkStages * kWsChunkRecords * 4;
        acc = o;    o += kPkBufs * 2 * P * 4;                       // three buffers of one word per (pixel, polarity)
        stage = o;  o += 2 * K * P * 4;                             // [2K][P] output staging
        bars = o;   o += (kPkStages + 2 * kPkBufs + 2) * 8;
        info = o;   o += kPkBufs * 16 + 16;                         // per buffer {first record, records, count bits, nmax}; drained chunks
        feed_p = o; o += (TileSmemWS::kFeedBytes + 15) / 16 * 16;
        total = o;
    }
};

#ifdef EVREP_TILE_TIMING
// Diagnostic build only, cycles per tile CTA: [0] accumulate thread 0 total, [1] waiting for record chunks, [2] preload +
// producers' barrier, [3] waiting for the accumulator buffer, [4] atomics + hand-over, [5] batch feed; [8] consumer thread 0
// total, [9] waiting for FULL, [10] read + clear, [11] update, [12] waiting for the staging tile, [13] staging.
__device__ unsigned long long g_tile_timing[1024][16];
#define TT_DECL long long tt[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const long long tt_start = clock64(); long long tq = tt_start
#define TT_MARK() do { tq = clock64(); } while (0)
#define TT_ADD(slot) do { const long long now_ = clock64(); tt[slot] += now_ - tq; tq = now_; } while (0)
#else
#define TT_DECL
#define TT_MARK()
#define TT_ADD(slot)
#endif


// Warp-cooperative exact sum of d over the records of one hand-over that fall on `cell` (hot pixels only).
__device__ __noinline__ uint32_t pk_rescan(const uint32_t* recs, uint32_t count, uint32_t cell) {
    uint32_t tot = 0;
    for (uint32_t i = threadIdx.x & 31u; i < count; i += 32u) {
        const uint32_t rec = __ldg(recs + i);
        if ((rec & 0x3FFFu) == cell) tot += rec >> 14;
    }
    return __reduce_add_sync(0xFFFFFFFFu, tot);
}

template <int K, int SLOTS>
__global__ void __launch_bounds__(kWsThreads, 1)
taf_tile_pk_kernel(TileParams tp) {
    const StreamPlan& pl = tp.pl;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmemPK lay(pl.P, K);
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw + lay.ring);     // [kPkStages][kWsChunkRecords]
    uint32_t* acc = reinterpret_cast<uint32_t*>(smem_raw + lay.acc);       // [kPkBufs][2P]
    float* stage = reinterpret_cast<float*>(smem_raw + lay.stage);         // [2K][P]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);     // [kPkStages] record chunks
    uint64_t* acc_full = full + kPkStages;                                 // [kPkBufs] accumulate warps -> consumers
    uint64_t* acc_empty = acc_full + kPkBufs;                              // [kPkBufs] consumers -> accumulate warps
    uint64_t* staged = acc_empty + kPkBufs;                                // consumers -> store warp
    uint64_t* stage_free = staged + 1;                                     // store warp -> consumers
    volatile uint32_t* info = reinterpret_cast<volatile uint32_t*>(smem_raw + lay.info);   // [kPkBufs][4]
    volatile uint32_t* drained_w = info + kPkBufs * 4;

    const int tid = threadIdx.x, tile = blockIdx.x;
    const int64_t HW = (int64_t)pl.H * pl.W;
    const int64_t pix0 = (int64_t)tile * pl.P;
    const int npix = (int)min((int64_t)pl.P, HW - pix0);
    const uint32_t* my_off = pl.off_rel + (int64_t)tile * (pl.TB + 1);

    if (tid == 0) {
        for (int s = 0; s < kPkStages; ++s) mbar_init(full + s, 1);
        for (int b = 0; b < kPkBufs; ++b) { mbar_init(acc_full + b, kAccumThreads / 32); mbar_init(acc_empty + b, kConsumerThreads / 32); }
        mbar_init(staged, kConsumerThreads / 32);
        mbar_init(stage_free, 1);
        *drained_w = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < kPkBufs * 2 * pl.P; i += kWsThreads) acc[i] = 0u;
    __syncthreads();

    if (tid < kProducerThreads) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (tid >= kAccumThreads) {
            // ================================== store warp ==================================
            // Keeps the record ring full (TMA loads into the stages the accumulate warps have drained) and sends the
            // window tensors the consumers stage (TMA bulk stores).
            const int lane = tid - kAccumThreads;
            const uint32_t* my_records = pl.records + pl.tile_base[tile];
            const uint32_t list_len = (pl.tile_total[tile] + 3u) & ~3u;
            const int n_chunks = (int)((list_len + kWsChunkRecords - 1) / kWsChunkRecords);
            const int n_emits = tp.bulk_out ? tp.n_emits : 0;
            int issued = 0, emitted = 0, j = 0;
            while (issued < n_chunks || emitted < n_emits) {
                bool worked = false;
                if (issued < n_chunks) {
                    const int limit = min(n_chunks, (int)*drained_w + kPkStages);
                    if (limit > issued) {
                        if (lane == 0)
                            for (int c = issued; c < limit; ++c) {
                                const uint32_t first = (uint32_t)c * kWsChunkRecords;
                                const uint32_t bytes = min((uint32_t)kWsChunkRecords, list_len - first) * 4u;
                                uint64_t* bar = full + (c % kPkStages);
                                mbar_expect_tx(bar, bytes);
                                tma_load_1d(ring + (c % kPkStages) * kWsChunkRecords, my_records + first, bytes, bar);
                            }
                        __syncwarp();
                        issued = limit;
                        worked = true;
                    }
                }
                if (emitted < n_emits) {
                    const bool ready = __shfl_sync(0xFFFFFFFFu, (int)mbar_test(staged, (uint32_t)emitted & 1u), 0) != 0;
                    if (ready) {
                        while (!(pl.batches[j].flags & 2)) ++j;
                        const Batch m = pl.batches[j];
                        ++j;
                        if (lane < 2 * K) {
                            float* o = tp.out + (int64_t)m.win * tp.out_stride + pix0;
                            bulk_store_1d(o + (int64_t)lane * HW, stage + lane * pl.P, (uint32_t)npix * 4u);
                            bulk_commit();
                            bulk_wait_read();                        // the rows have left shared memory
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(stage_free);
                        ++emitted;
                        worked = true;
                    }
                }
                if (!worked) __nanosleep(32);
            }
            if (lane < 2 * K) bulk_wait_all();
            return;
        }
        // ================================ accumulate warps ================================
        BatchFeed feed;
        feed.init(smem_raw + lay.feed_p, &pl, my_off, tid, kBarProducers, kAccumThreads);
        const uint32_t base_idx = pl.tile_base[tile];
        int ready_chunk = -1, buf = 0;
        uint32_t round = 0;                                          // times the buffers have gone round
        TT_DECL;
        for (int j = 0; j < pl.n_batches; ++j) {
            TT_MARK();
            const Batch meta = feed.begin(j);
            TT_ADD(5);
            const int jb = j & 1;
            for (int b = 0; b < meta.nb; ++b) {
                const uint32_t o0 = feed.s_off[jb * (kBatchBins + 1) + b], o1 = feed.s_off[jb * (kBatchBins + 1) + b + 1];
                if (!feed.s_any[jb * kBatchBins + b] || o1 <= o0) continue;
                // count bits of this hand-over: the count field holds every record of the bin
                int cbits = kPkMinCountBits;
                while (cbits < 32 && (o1 - o0) >> cbits) ++cbits;
                uint32_t* my_acc = acc + buf * 2 * pl.P;
                constexpr int kPre = 8;
                constexpr uint32_t kNoRec = 0xFFFFFFFFu;            // d = 2^18-1, pixel 8191: never produced for P <= 2560
                // the bin goes through the registers in pieces of kPre x 96 records (one piece for all but dense bins)
                for (uint32_t q0 = o0; q0 < o1; q0 += kPre * kAccumThreads) {
                    const uint32_t q1 = min(o1, q0 + (uint32_t)(kPre * kAccumThreads));
                    const int last_c = (int)((q1 - 1) / kWsChunkRecords);
                    TT_MARK();
                    while (ready_chunk < last_c) {
                        ++ready_chunk;
                        mbar_wait(full + (ready_chunk % kPkStages), (uint32_t)(ready_chunk / kPkStages) & 1u);
                    }
                    TT_ADD(1);
                    // records are pulled into registers BEFORE waiting for the accumulator buffer
                    uint32_t pre[kPre];
                    const int n_pre = (int)((q1 - q0 + kAccumThreads - 1) / kAccumThreads);
                    const bool small_bin = SLOTS <= 2 && n_pre <= 2;        // GEN1-size grids: ~70 records per bin
                    if (small_bin) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const uint32_t r = q0 + tid + i * kAccumThreads;
                            pre[i] = (i < n_pre && r < q1) ? ring[r & (kPkRing - 1)] : kNoRec;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < kPre; ++i) {
                            const uint32_t r = q0 + tid + i * kAccumThreads;
                            pre[i] = r < q1 ? ring[r & (kPkRing - 1)] : kNoRec;
                        }
                    }
                    // every accumulate thread is done with the ring up to q1: tell the store warp
                    named_sync(kBarProducers, kAccumThreads);
                    if (tid == 0) *drained_w = q1 / kWsChunkRecords;
                    TT_ADD(2);
                    if (q0 == o0) {
                        if (round > 0) mbar_wait(acc_empty + buf, (round - 1u) & 1u);   // the consumers have drained this buffer
                        if (tid == 0) {
                            info[buf * 4 + 0] = base_idx + o0;
                            info[buf * 4 + 1] = o1 - o0;
                            info[buf * 4 + 2] = (uint32_t)cbits;
                            info[buf * 4 + 3] = tp.pk_hotmask[cbits];
                        }
                    }
                    TT_ADD(3);
                    if (small_bin) {
#pragma unroll
                        for (int i = 0; i < 2; ++i)
                            if (pre[i] != kNoRec) atomicAdd(my_acc + (pre[i] & 0x3FFFu), ((pre[i] >> 14) << cbits) + 1u);
                    } else {
#pragma unroll
                        for (int i = 0; i < kPre; ++i)
                            if (pre[i] != kNoRec) atomicAdd(my_acc + (pre[i] & 0x3FFFu), ((pre[i] >> 14) << cbits) + 1u);
                    }
                    TT_ADD(4);
                }
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(acc_full + buf);              // hand the accumulator to the consumers
                if (++buf == kPkBufs) { buf = 0; ++round; }
            }
            TT_MARK();
            feed.end(j);
            TT_ADD(5);
        }
#ifdef EVREP_TILE_TIMING
        if (tid == 0) {
            g_tile_timing[tile][0] = clock64() - tt_start;
            for (int i = 1; i < 8; ++i) g_tile_timing[tile][i] = tt[i];
        }
#endif
        return;
    }

    // =================================== consumer warpgroups ===================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int ctid = tid - kProducerThreads;
    static_assert(K % 4 == 0, "K must be a multiple of 4");
    float2 v[SLOTS][2][K / 2];
    const bool first_fresh = (pl.batches[0].flags & 1) != 0;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        const int lp = s * kConsumerThreads + ctid;
        if (lp < npix && !first_fresh) {
            const float4* src = reinterpret_cast<const float4*>(tp.state + (pix0 + lp) * 2 * K);
#pragma unroll
            for (int q = 0; q < 2 * K / 4; ++q) {
                const float4 f = src[q];
                v[s][(q * 4) / K][((q * 4) % K) / 2 + 0] = make_float2(f.x, f.y);
                v[s][(q * 4) / K][((q * 4) % K) / 2 + 1] = make_float2(f.z, f.w);
            }
        } else {
#pragma unroll
            for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int k = 0; k < K / 2; ++k) v[s][p][k] = make_float2(kTafInit, kTafInit);
        }
    }
    const uint32_t* my_bits = pl.tile_bits + (int64_t)tile * pl.n_batches;
    Batch meta = pl.batches[0];
    uint32_t bits = my_bits[0];
    int buf = 0, emitted = 0;
    uint32_t round = 0;
    const float2 minus1 = make_float2(-1.0f, -1.0f);
    TT_DECL;
    for (int j = 0; j < pl.n_batches; ++j) {
        Batch nmeta = meta;
        uint32_t nbits = 0;
        if (j + 1 < pl.n_batches) { nmeta = pl.batches[j + 1]; nbits = my_bits[j + 1]; }
        if (meta.flags & 1) {
#pragma unroll
            for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                for (int p = 0; p < 2; ++p)
#pragma unroll
                    for (int k = 0; k < K / 2; ++k) v[s][p][k] = make_float2(kTafInit, kTafInit);
        }
        for (int b = 0; b < meta.nb; ++b) {
            if (!((bits >> (16 + b)) & 1u)) continue;               // nobody saw an event: no ageing
            if (!((bits >> b) & 1u)) {
                // the tile saw nothing in this bin, but some other tile did: everything ages
#pragma unroll
                for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                    for (int p = 0; p < 2; ++p)
#pragma unroll
                        for (int k = 0; k < K / 2; ++k) v[s][p][k] = __fadd2_rn(v[s][p][k], minus1);
                continue;
            }
            TT_MARK();
            mbar_wait(acc_full + buf, round & 1u);                   // the accumulate warps filled this buffer
            TT_ADD(9);
            uint32_t* my_acc = acc + buf * 2 * pl.P;
            const uint32_t cbits = info[buf * 4 + 2], hotmask = info[buf * 4 + 3];
            uint2 w[SLOTS];
            uint32_t seen = 0u;
#pragma unroll
            for (int s = 0; s < SLOTS; ++s) {
                const int lp = s * kConsumerThreads + ctid;
                w[s] = make_uint2(0u, 0u);
                if (s < SLOTS - 1 || lp < pl.P) {                    // every slot but the last lies inside the array
                    w[s] = *reinterpret_cast<uint2*>(my_acc + 2 * lp);
                    if (w[s].x | w[s].y) *reinterpret_cast<uint2*>(my_acc + 2 * lp) = make_uint2(0u, 0u);
                }
                seen |= w[s].x | w[s].y;
            }
            const uint32_t cmask = cbits >= 32u ? 0xFFFFFFFFu : (1u << cbits) - 1u;
            // one test per thread: some count has a bit at or above the power of two below nmax(c) + 1
            const bool any_hot = __any_sync(0xFFFFFFFFu, (seen & hotmask) != 0u);
            // FIFO push / ageing of the tile's cells; kFix: the sums of hot cells come from `fix`
            uint32_t fix[SLOTS][2];
            auto update = [&](auto use_fix) {
                constexpr bool kFix = decltype(use_fix)::value;
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        // mean(t_norm) - 1 = S / (n span) - 1 (generate_taf.py:23-27); NaN for n == 0, never selected
                        const uint32_t word = p ? w[s].y : w[s].x;
                        const uint32_t n = word & cmask;
                        uint32_t sd = cbits >= 32u ? 0u : word >> cbits;
                        if (kFix && (n & hotmask)) sd = fix[s][p];
                        const bool active = n != 0u;
                        float r;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)n * tp.span));
                        const float mean = fmaf((float)sd, r, -1.0f);
                        float2 aged[K / 2];
#pragma unroll
                        for (int k = 0; k < K / 2; ++k) aged[k] = __fadd2_rn(v[s][p][k], minus1);
#pragma unroll
                        for (int k = 0; k < K / 2; ++k) {
                            const float next = (k + 1 < K / 2) ? aged[k + 1].x : mean;
                            v[s][p][k].x = active ? aged[k].y : aged[k].x;
                            v[s][p][k].y = active ? next : aged[k].y;
                        }
                    }
                }
            };
            if (any_hot) {
                // a count beyond the exact range of the packed sum: recompute those sums from the bin's records
                const uint32_t* recs = pl.records + info[buf * 4 + 0];
                const uint32_t i_count = info[buf * 4 + 1];
                const int warp_ctid = ctid & ~31;
#pragma unroll
                for (int s = 0; s < SLOTS; ++s)
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        fix[s][p] = 0u;
                        uint32_t mask = __ballot_sync(0xFFFFFFFFu, ((p ? w[s].y : w[s].x) & cmask & hotmask) != 0u);
                        while (mask) {
                            const int src = __ffs(mask) - 1;
                            mask &= mask - 1u;
                            const uint32_t cell = 2u * (uint32_t)(s * kConsumerThreads + warp_ctid + src) + (uint32_t)p;
                            const uint32_t tot = pk_rescan(recs, i_count, cell);
                            if ((ctid & 31) == src) fix[s][p] = tot;
                        }
                    }
            }
            __syncwarp();
            if ((ctid & 31) == 0) mbar_arrive(acc_empty + buf);      // clean again (and its records no longer needed): give it back
            TT_ADD(10);
            if (++buf == kPkBufs) { buf = 0; ++round; }
            if (any_hot) update(std::true_type{});
            else update(std::false_type{});
            TT_ADD(11);
        }
        if (meta.flags & 2) {
            const bool write_state = tp.emit_state || (j == pl.n_batches - 1);
            if (tp.bulk_out) {
                // stage the [2K][P] tile; the store warp sends it
                TT_MARK();
                if (emitted > 0) mbar_wait(stage_free, (uint32_t)(emitted - 1) & 1u);
                TT_ADD(12);
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kConsumerThreads + ctid;
                    if (s == SLOTS - 1 && lp >= pl.P) continue;     // columns >= npix are staged but never stored
#pragma unroll
                    for (int k = 0; k < K / 2; ++k)
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            stage[(4 * k + p) * pl.P + lp] = v[s][p][k].x;
                            stage[(4 * k + 2 + p) * pl.P + lp] = v[s][p][k].y;
                        }
                }
                fence_async_smem();
                __syncwarp();
                if ((ctid & 31) == 0) mbar_arrive(staged);
                TT_ADD(13);
                ++emitted;
            } else {
                float* o = tp.out + (int64_t)meta.win * tp.out_stride + pix0;
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kConsumerThreads + ctid;
                    if (lp >= npix) continue;
#pragma unroll
                    for (int k = 0; k < K / 2; ++k)
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            __stcs(o + (int64_t)(4 * k + p) * HW + lp, v[s][p][k].x);
                            __stcs(o + (int64_t)(4 * k + 2 + p) * HW + lp, v[s][p][k].y);
                        }
                }
            }
            if (write_state) {
#pragma unroll
                for (int s = 0; s < SLOTS; ++s) {
                    const int lp = s * kConsumerThreads + ctid;
                    if (lp >= npix) continue;
                    float4* dst = reinterpret_cast<float4*>(tp.state + (pix0 + lp) * 2 * K);
#pragma unroll
                    for (int q = 0; q < 2 * K / 4; ++q) {
                        const float2 lo = v[s][(q * 4) / K][((q * 4) % K) / 2], hi = v[s][(q * 4) / K][((q * 4) % K) / 2 + 1];
                        dst[q] = make_float4(lo.x, lo.y, hi.x, hi.y);
                    }
                }
            }
        }
        meta = nmeta;
        bits = nbits;
    }
#ifdef EVREP_TILE_TIMING
    if (ctid == 0) {
        g_tile_timing[tile][8] = clock64() - tt_start;
        for (int i = 9; i < 16; ++i) g_tile_timing[tile][i] = tt[i];
    }
#endif
}

template <int K>
static int launch_tiles_pk(const TileParams& tp, cudaStream_t st) {
    const int slots = (tp.pl.P + kConsumerThreads - 1) / kConsumerThreads;
    const size_t smem = (size_t)TileSmemPK(tp.pl.P, K).total;
#define EVREP_TILE_PK(S)                                                                                  \
    case S:                                                                                               \
        EVREP_CUDA(cudaFuncSetAttribute(taf_tile_pk_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        taf_tile_pk_kernel<K, S><<<tp.pl.n_tiles, kWsThreads, smem, st>>>(tp);                            \
        break;
    switch (slots) {
        EVREP_TILE_PK(1) EVREP_TILE_PK(2) EVREP_TILE_PK(3) EVREP_TILE_PK(4) EVREP_TILE_PK(5) EVREP_TILE_PK(6)
        default: return EVREP_ERR_RANGE;
    }
#undef EVREP_TILE_PK
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

template <int K, bool kRuns>
static int launch_tiles_ws(const TileParams& tp, cudaStream_t st) {
    const int slots = (tp.pl.P + kConsumerThreads - 1) / kConsumerThreads;
    const size_t smem = (size_t)TileSmemWS(tp.pl.P, K).total;
#define EVREP_TILE_WS(S)                                                                                  \
    case S:                                                                                               \
        EVREP_CUDA(cudaFuncSetAttribute(taf_tile_ws_kernel<K, S, kRuns>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        taf_tile_ws_kernel<K, S, kRuns><<<tp.pl.n_tiles, kWsThreads, smem, st>>>(tp);                     \
        break;
    switch (slots) {
        EVREP_TILE_WS(1) EVREP_TILE_WS(2) EVREP_TILE_WS(3) EVREP_TILE_WS(4) EVREP_TILE_WS(5) EVREP_TILE_WS(6)
        default: return EVREP_ERR_RANGE;
    }
#undef EVREP_TILE_WS
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

template <int K>
static int launch_tiles(const TileParams& tp, int slots, size_t smem, cudaStream_t st) {
#define EVREP_TILE(S)                                                                                     \
    case S:                                                                                               \
        EVREP_CUDA(cudaFuncSetAttribute(taf_tile_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        taf_tile_kernel<K, S><<<tp.pl.n_tiles, kTafThreads, smem, st>>>(tp);                              \
        break;
    switch (slots) {
        EVREP_TILE(1) EVREP_TILE(2) EVREP_TILE(3) EVREP_TILE(4) EVREP_TILE(5) EVREP_TILE(6)
        default: return EVREP_ERR_RANGE;
    }
#undef EVREP_TILE
    EVREP_LAUNCH_CHECK();
    return EVREP_OK;
}

}  // namespace evrep

using namespace evrep;

extern "C" {

int64_t evrep_taf_stream_scratch_bytes(int64_t n_events, int n_windows, int64_t total_bins, int H, int W) {
    if (n_events < 0 || n_windows < 0 || total_bins < 0 || H <= 0 || W <= 0) return EVREP_ERR_ARG;
    Layout L;
    int rc = make_layout(n_events, n_windows, total_bins, H, W, (int)batches_upper_bound(n_windows, total_bins), L);
    if (rc) return rc;
    return L.total;
}

int evrep_taf_stream(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                     const evrep_taf_window* windows_host, int n_windows, int abin, int H, int W, int K,
                     const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                     float* state_inout, int emit_state_every_window,
                     float* out, int64_t out_stride, void* scratch, int64_t scratch_bytes,
                     void* ev_tiles_begin, void* ev_tiles_end, evrep_stream_t stream) {
    if (n_events < 0 || n_windows < 0 || H <= 0 || W <= 0 || abin <= 0 || !state_inout || !scratch) return EVREP_ERR_ARG;
    if (K != 4 && K != 8) return EVREP_ERR_ARG;
    if (n_windows == 0) return EVREP_OK;
    if (!out || (reinterpret_cast<uintptr_t>(state_inout) & 15)) return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    StreamPlan pl;
    Layout L;
    int rc = prepare_stream(t, x, y, p, n_events, windows_host, n_windows, abin, H, W, xmap, ymap, sensor_h, sensor_w,
                            scratch, scratch_bytes, st, pl, L);
    if (rc) return rc;

    TileParams tp;
    tp.pl = pl; tp.state = state_inout; tp.out = out; tp.out_stride = out_stride;
    tp.emit_state = emit_state_every_window;
    tp.n_emits = n_windows;
    tp.span = (float)((double)abin + 1e-8);
    tp.bulk_out = (((int64_t)H * W) % 4 == 0 && out_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
    tp.src = nullptr;
    for (int c = 0; c <= 32; ++c) {
        // the sum field (32 - c bits) is exact for counts up to nmax; cells are treated as hot from the power of two below nmax + 1
        const uint64_t dmax = abin > 1 ? (uint64_t)abin - 1u : 1u;
        const uint64_t nmax = c >= 32 ? 0u : (((uint64_t)1 << (32 - c)) - 1u) / dmax;
        const uint64_t cmask = c >= 32 ? 0xFFFFFFFFull : ((uint64_t)1 << c) - 1u;
        uint64_t T = 1;
        while (T * 2 <= nmax + 1) T *= 2;
        tp.pk_hotmask[c] = (uint32_t)(cmask & ~(T - 1u));
    }

    const size_t smem = (size_t)TileSmem(L.P, K).total;
    if (ev_tiles_begin) EVREP_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev_tiles_begin), st));
    const char* legacy = getenv("EVREP_TAF_TILE_KERNEL");            // A/B runs: "single" = the non-specialised kernel, "pk"
    const bool ws = !(legacy && strcmp(legacy, "single") == 0) && (size_t)TileSmemWS(L.P, K).total <= 232448;
    // "pk" = the packed-accumulator variant: faster for 1500-3000 records per (tile, bin), slower below and with hot pixels
    const bool pk = ws && legacy && strcmp(legacy, "pk") == 0 && (size_t)TileSmemPK(L.P, K).total <= 232448;
    if (pk) rc = K == 8 ? launch_tiles_pk<8>(tp, st) : launch_tiles_pk<4>(tp, st);
    else if (ws) rc = K == 8 ? launch_tiles_ws<8, false>(tp, st) : launch_tiles_ws<4, false>(tp, st);
    else rc = K == 8 ? launch_tiles<8>(tp, L.slots, smem, st) : launch_tiles<4>(tp, L.slots, smem, st);
    if (rc) return rc;
    if (ev_tiles_end) EVREP_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev_tiles_end), st));
    return EVREP_OK;
}


int64_t evrep_taf_stream_ordered_scratch_bytes(int64_t n_events, int n_windows, int64_t total_bins, int H, int W) {
    if (n_events < 0 || n_windows < 0 || total_bins < 0 || H <= 0 || W <= 0) return EVREP_ERR_ARG;
    return binmajor_scratch_bytes(n_events, n_windows, total_bins, H, W);
}

int64_t evrep_taf_stream_ordered_status_offset(int64_t n_events, int n_windows, int64_t total_bins, int H, int W) {
    if (n_events < 0 || n_windows < 0 || total_bins < 0 || H <= 0 || W <= 0) return EVREP_ERR_ARG;
    return binmajor_status_offset(n_events, n_windows, total_bins, H, W);
}

int evrep_taf_stream_ordered(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                             const evrep_taf_window* windows_host, int n_windows, int abin, int H, int W, int K,
                             const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                             float* state_inout, int emit_state_every_window,
                             float* out, int64_t out_stride, void* scratch, int64_t scratch_bytes,
                             void* ev_tiles_begin, void* ev_tiles_end, evrep_stream_t stream) {
    if (n_events < 0 || n_windows < 0 || H <= 0 || W <= 0 || abin <= 0 || !state_inout || !scratch) return EVREP_ERR_ARG;
    if (K != 4 && K != 8) return EVREP_ERR_ARG;
    if (n_windows == 0) return EVREP_OK;
    if (!out || (reinterpret_cast<uintptr_t>(state_inout) & 15)) return EVREP_ERR_ARG;
    cudaStream_t st = as_stream(stream);
    StreamPlan pl;
    Layout L;
    const uint32_t* src = nullptr;
    uint32_t* status = nullptr;
    int rc = prepare_stream_binmajor(t, x, y, p, n_events, windows_host, n_windows, abin, H, W, xmap, ymap, sensor_h, sensor_w,
                                     scratch, scratch_bytes, st, pl, L, src, status);
    if (rc) return rc;
    if ((size_t)TileSmemWS(L.P, K).total > 232448) return EVREP_ERR_RANGE;
    TileParams tp;
    tp.pl = pl; tp.state = state_inout; tp.out = out; tp.out_stride = out_stride;
    tp.emit_state = emit_state_every_window;
    tp.n_emits = n_windows;
    tp.span = (float)((double)abin + 1e-8);
    tp.bulk_out = (((int64_t)H * W) % 4 == 0 && out_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
    tp.src = src;
    for (int c = 0; c <= 32; ++c) tp.pk_hotmask[c] = 0u;
    if (ev_tiles_begin) EVREP_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev_tiles_begin), st));
    rc = K == 8 ? launch_tiles_ws<8, true>(tp, st) : launch_tiles_ws<4, true>(tp, st);
    if (rc) return rc;
    if (ev_tiles_end) EVREP_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(ev_tiles_end), st));
    return EVREP_OK;
}

#ifdef EVREP_TILE_TIMING
int evrep_debug_tile_timing(unsigned long long* host, int n_tiles) {
    EVREP_CUDA(cudaDeviceSynchronize());
    EVREP_CUDA(cudaMemcpyFromSymbol(host, g_tile_timing, sizeof(unsigned long long) * 16 * (size_t)n_tiles));
    return EVREP_OK;
}
#endif

}  // extern "C"
