/*
 * evrep.h -- C ABI of libevrep.so, the B200 (sm_100a) event-representation encoders.
 *
 * This is the drop-in boundary for the preprocessing hot path of HarmoniaLeo/FRLW-EvD.
 * The reference has no FFI of its own for this path: its encoders are Python functions
 * built from ATen ops.  Each entry point below therefore cites the reference function
 * (file:line, relative to the reference tree) whose GPU work it replaces; the Python
 * modules in frlw-evd_b200/ re-export the reference's names on top of these symbols and
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - plain C symbols, `int` return: 0 = ok, < 0 = error (evrep_strerror()).
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`.
 *   - the library never allocates or frees device memory: inputs, state, outputs and
 *     scratch are caller-owned (sizes from the evrep_*_scratch_bytes() queries).
 *   - every call is stream-ordered on `stream` (a cudaStream_t) and never synchronises;
 *     the Python shims synchronise where the reference's functions block.
 *   - event streams are structure-of-arrays: t u32 microseconds, x u16, y u16, p u8
 *     (the decoded record of src/io/psee_loader.py:39-44).  `*_aos64` variants read the
 *     reference's staging layout instead: a float64 [N, ncols] matrix with columns
 *     (x, y, t, p[, z]) (generate_taf.py:195), coordinates truncated like `.long()`.
 *   - `xmap` / `ymap` (nullable) are coordinate look-up tables applied to raw sensor
 *     coordinates before encoding: the gen4 policy `coord * ratio` in float64, truncated
 *     (generate_taf.py:103-104,216-218).  NULL = identity.  A LUT always holds
 *     EVREP_COORD_LUT_LEN u16 entries (the 14-bit coordinate range of a .dat record);
 *     entries >= W (or H) mark coordinates to drop.
 *   - events whose (mapped) coordinates fall outside H x W, or whose polarity is not 0/1,
 *     are dropped (the reference only filters in the SAE encoder,
 *     generate_surfaceofactiveevents.py:72, and device-asserts elsewhere).
 *   - no CPU fallback exists: without a CUDA device every compute entry point fails
 *     with EVREP_ERR_CUDA.
 */
#ifndef EVREP_H
#define EVREP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EVREP_VERSION 100
#define EVREP_COORD_LUT_LEN 16384

#define EVREP_OK            0
#define EVREP_ERR_ARG      -1   /* null pointer / non-positive size / unsupported K */
#define EVREP_ERR_CUDA     -2   /* a CUDA runtime call or launch failed */
#define EVREP_ERR_SCRATCH  -3   /* scratch buffer too small */
#define EVREP_ERR_RANGE    -4   /* a size exceeds what the packed formats can hold */

typedef void* evrep_stream_t;   /* cudaStream_t */

int         evrep_version(void);
const char* evrep_strerror(int code);
/* Last CUDA error string seen by the calling thread ("" if none). */
const char* evrep_last_cuda_error(void);
/* SM count / compute capability of the current device (any pointer may be NULL). */
int         evrep_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------- D1: .dat decode ----
 * src/io/dat_events_tools.py:82-100 (stream_td_data) with EV_TYPE :16 -- 8-byte records
 * (u32 t, i32 w): x = w & 0x3FFF, y = (w & 0x0FFFC000) >> 14, p = (w & 0x10000000) >> 28.
 * `records` must be 8-byte aligned. */
int evrep_decode_dat(const void* records, int64_t n,
                     uint32_t* t, uint16_t* x, uint16_t* y, uint8_t* p, evrep_stream_t stream);
/* D4 inverse (tests / compatibility): SoA -> the float64 [N,4] (x, y, t, p) staging matrix
 * of generate_taf.py:195. */
int evrep_soa_to_aos64(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p,
                       int64_t n, double* events, evrep_stream_t stream);

/* ------------------------------------------------------ E1: Event Count Image --------
 * generate_eventcountimage.py:19-41 (generate_eventframe).  Counts events per (p, y, x)
 * into `counts` (u32 [2,H,W]) and writes out[p,y,x] = f(count), f = the float32 running
 * sum of 0.05 clamped to 1, times 255 (bit-exact with the reference's index_add_).
 * `counts` must be zero before the first accumulate of a window.  evrep_count_image()
 * = accumulate + finalize(reset=1).  Nested last-N windows (driver :156) call
 * accumulate on the extra events and finalize(reset=0) per N. */
int evrep_count_accumulate(const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n,
                           int H, int W, const uint16_t* xmap, const uint16_t* ymap,
                           uint32_t* counts, evrep_stream_t stream);
int evrep_count_accumulate_aos64(const double* events, int64_t n, int ncols, int H, int W,
                                 uint32_t* counts, evrep_stream_t stream);
int evrep_count_finalize(uint32_t* counts, int H, int W, float* out, int reset, evrep_stream_t stream);
int evrep_count_image(const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n,
                      int H, int W, const uint16_t* xmap, const uint16_t* ymap,
                      uint32_t* counts, float* out, evrep_stream_t stream);

/* Driver epilogue fused (generate_eventcountimage.py:156-180): the nested last-N windows of one
 * label -- sizes_host ascending, window i = the last min(sizes[i], n) of the n events given -- each
 * event counted once, every window written as uint8 [2,Ht,Wt] (count LUT, nearest resize with the
 * int32 maps ysrc/xsrc or NULL when Ht x Wt == H x W, truncation) into out + i * 2*Ht*Wt.
 * `counts` zero on entry, zero on return. */
int evrep_count_images_u8(const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n,
                          const int64_t* sizes_host, int n_sizes, int H, int W,
                          const uint16_t* xmap, const uint16_t* ymap, int Ht, int Wt,
                          const int32_t* ysrc, const int32_t* xsrc, uint32_t* counts, uint8_t* out,
                          evrep_stream_t stream);

/* ------------------------------------------------- A1: Surface of Active Events ------
 * generate_surfaceofactiveevents.py:71-80 (generate_leaky_cuda) + :44-69 (taf_cuda).
 * latest[p,y,x] = max float32(t) over the events (the sequential last-writer result for
 * time-sorted input), `init` where no event; merged by max with `memory_in` when not
 * NULL; written to `memory_out` (absolute float32 timestamps, may alias memory_in);
 * out[l,p,y,x] = expf(lambdas[l] * (latest - now_f32)) * 255, out is [2L,H,W].
 * `init` = f32(f32(now) - 5e6) and `now_f32` = f32(now) are computed by the caller.
 * `keys` is a u32 [2,H,W] scratch, zero on entry, zero on return.  L <= 8. */
int evrep_sae(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n,
              int H, int W, const uint16_t* xmap, const uint16_t* ymap,
              float init, float now_f32, const float* lambdas_host, int L,
              const float* memory_in, float* memory_out, uint32_t* keys, float* out,
              evrep_stream_t stream);
/* evrep_sae with the driver epilogue fused (generate_surfaceofactiveevents.py:186-204): the state
 * (memory_out, f32 [2,H,W], must not alias memory_in) is updated at grid resolution and the decays
 * are written resized + truncated as uint8 [L,2,Ht,Wt]. */
int evrep_sae_u8(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n,
                 int H, int W, const uint16_t* xmap, const uint16_t* ymap, int Ht, int Wt,
                 const int32_t* ysrc, const int32_t* xsrc, float init, float now_f32,
                 const float* lambdas_host, int L, const float* memory_in, float* memory_out,
                 uint32_t* keys, uint8_t* out, evrep_stream_t stream);
int evrep_sae_aos64(const double* events, int64_t n, int ncols, int H, int W,
                    float init, float now_f32, const float* lambdas_host, int L,
                    const float* memory_in, float* memory_out, uint32_t* keys, float* out,
                    evrep_stream_t stream);

/* ------------------------------------------------------------ V1: Event Volume -------
 * generate_eventvolume.py:15-42 (generate_agile_event_volume_cuda).  Temporal bilinear
 * splat: t* = K * f32(t_norm), centres 1..K, weight 1 - |c - t*| into channel
 * 2(c-1) + (1-p); out [2K,H,W] = sum / 5 * 255.  SoA form: t_norm = (t - t0) / tw in
 * float64 (driver :141).  `out` doubles as the accumulator (it is zeroed by the call). */
int evrep_event_volume(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p,
                       int64_t n, int64_t t0, int64_t tw, int H, int W, int K,
                       const uint16_t* xmap, const uint16_t* ymap, float* out, evrep_stream_t stream);
int evrep_event_volume_aos64(const double* events, int64_t n, int ncols, int H, int W, int K,
                             float* out, evrep_stream_t stream);

/* ------------------------------------------- T1: Temporal Active Focus, one bin ------
 * generate_taf.py:60-67 (generate_taf_cuda) + :19-58 (taf_cuda).  One 10 ms step of the
 * per-(y,x,p) K-deep FIFO: active cells shift and push mean(f32(t_norm) - 1), inactive
 * cells age by 1; a bin without any (valid) event leaves the state untouched.
 * state layout f32 [H,W,2,K]; out (nullable) f32 [2K,H,W], channel 2k + p.
 * SoA form: t_norm = (t - t_min) / t_span in float64 (driver :213-215, t_span = abin+1e-8).
 * `scratch`: evrep_taf_bin_scratch_bytes(H, W) bytes, zero on entry, zero on return.
 * state_out may alias state_in.  K in {1..16}. */
int64_t evrep_taf_bin_scratch_bytes(int H, int W);
int evrep_taf_bin(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p,
                  int64_t n, int64_t t_min, double t_span, int H, int W, int K,
                  const uint16_t* xmap, const uint16_t* ymap,
                  const float* state_in, float* state_out, float* out, void* scratch,
                  evrep_stream_t stream);
int evrep_taf_bin_aos64(const double* events, int64_t n, int ncols, int H, int W, int K,
                        const float* state_in, float* state_out, float* out, void* scratch,
                        evrep_stream_t stream);

/* ------------------------------------- T2: Temporal Active Focus, whole streams ------
 * The driver loop of generate_taf.py:160-238 for a list of windows, in two launches
 * groups: (1) events are bucketed by (sensor tile, 10 ms bin) into packed 4-byte records,
 * (2) one persistent kernel walks all windows: each CTA owns a tile of the sensor, keeps
 * that tile's FIFO state in registers, consumes its records bin by bin (TMA bulk copies
 * into shared memory), and emits the [2K,H,W] tensor (+ the state) at every window end.
 *
 * A window w covers events [ev_begin, ev_end) (indices into the SoA arrays) and bins
 * i = 0..n_bins-1 of width `abin` starting at start_time; an event belongs to bin
 * clamp(floor((t - start_time) / abin), 0, n_bins - 1) -- the reference's inclusive
 * edges with the later bin winning (:201-202).  `fresh` resets the state to -6000
 * (:205-209) before the window.  Windows must be ordered and non-overlapping.
 * Per-cell mean: exact integer sum of (t - t_min) over the bin, divided in float64 by
 * n * (abin + 1e-8), minus 1 -- order independent, hence deterministic; equals the
 * reference's sequential float32 sum to rounding (bit-exact for n == 1).
 *
 * out: f32 [n_windows][2K,H,W] (window w at out + w * out_stride floats).
 * state_inout: f32 [H,W,2,K]; read unless windows[0].fresh, written after the last window.
 * The state after window w is, by definition, the [H,W,2,K] permutation of out[w]
 * (generate_taf.py:55), so nothing is lost by keeping it on chip in between; pass
 * emit_state_every_window != 0 to have the tensor rewritten after every window anyway.
 * sensor_h / sensor_w: raw coordinate range covered by ymap / xmap (ignored when the maps
 * are NULL); raw coordinates beyond it are dropped.
 * K must be 4 or 8; abin <= 262143; at most 2048 tiles of 2304 pixels.
 * ev_tiles_begin / ev_tiles_end: optional cudaEvent_t handles recorded on `stream` right
 * before / after the tile kernel (the dominant launch), for live roofline timing. */
typedef struct {
    int64_t ev_begin;
    int64_t ev_end;
    int64_t start_time;
    int32_t n_bins;
    int32_t fresh;
} evrep_taf_window;

int64_t evrep_taf_stream_scratch_bytes(int64_t n_events, int n_windows, int64_t total_bins, int H, int W);
int evrep_taf_stream(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p,
                     int64_t n_events, const evrep_taf_window* windows_host, int n_windows,
                     int abin, int H, int W, int K,
                     const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                     float* state_inout, int emit_state_every_window,
                     float* out, int64_t out_stride,
                     void* scratch, int64_t scratch_bytes,
                     void* ev_tiles_begin, void* ev_tiles_end, evrep_stream_t stream);

/* ------------- T2 (time-ordered input, bin-major variant): one-pass sort + the tile kernel of evrep_taf_stream ------
 * Measured slower than evrep_taf_stream; kept as an alternative.  Same contract and results as evrep_taf_stream for streams whose timestamps are non-decreasing over
 * [windows[0].ev_begin, windows[n-1].ev_end).  Every bin of a window is then one contiguous index range: the bins'
 * ranges come from bisection, every bin is cut into slices of <= 8188 events, one CTA sorts a slice by sensor tile in
 * shared memory, the CTAs of a bin's slices exchange their tile counts and write every (bin, tile) run contiguously
 * (9 bytes read + 4 written per event, one pass instead of two); the register-resident tile kernel then fetches a
 * 2 KB chunk of a tile's list with one bulk copy per run it touches.  Windows of more than 128 x 8188 events return
 * EVREP_ERR_RANGE (use evrep_taf_stream).  Order violations are counted, not repaired: after the call (stream order)
 * the status block inside `scratch` says how many events lie outside the bin their position implies; pass
 * scratch + evrep_taf_stream_ordered_status_offset(...) to evrep_stream_order_violations. */
int64_t evrep_taf_stream_ordered_scratch_bytes(int64_t n_events, int n_windows, int64_t total_bins, int H, int W);
/* byte offset of the status block inside `scratch`: u32 [0] order violations, [2] sort CTAs that gave up waiting */
int64_t evrep_taf_stream_ordered_status_offset(int64_t n_events, int n_windows, int64_t total_bins, int H, int W);
int evrep_taf_stream_ordered(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p,
                             int64_t n_events, const evrep_taf_window* windows_host, int n_windows,
                             int abin, int H, int W, int K,
                             const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                             float* state_inout, int emit_state_every_window,
                             float* out, int64_t out_stride, void* scratch, int64_t scratch_bytes,
                             void* ev_tiles_begin, void* ev_tiles_end, evrep_stream_t stream);

/* ------------- T2 (time-ordered input, shared-memory variant): slice sort + tile kernel with lazy ageing ------
 * Same contract and results as evrep_taf_stream for streams whose timestamps are non-decreasing
 * over [windows[0].ev_begin, windows[n-1].ev_end) -- what src/io/psee_loader.py hands out, and
 * what generate_taf.py:188-193 assumes when it seeks by time.  Every bin of a window is then one
 * contiguous index range, so the two bucketing passes collapse into one: bin ranges by bisection,
 * one CTA sorts a slice of <= 8188 events by sensor tile in shared memory and writes it back with
 * one TMA bulk store (9 bytes read + 4 written per event).  The tile kernel keeps the FIFO state
 * in shared memory as circular buffers with lazy ageing and touches only the cells that received
 * events (one packed shared-memory atomic per record, one push per active cell per bin).
 *
 * out (nullable): f32 [n_windows][2K,H,W].  out_u8 (nullable): u8 [n_windows][K,2,H,W], the bytes of
 * the bins{K/2} / bins{K} files when no resize follows (leaky transform, slot flip, truncation:
 * generate_taf.py:226-235), written straight from the tile kernel.  At least one of the two.
 * Order violations are counted, not repaired: after the call (stream order) the first 4 bytes of
 * `scratch` hold the number of events whose timestamp lies outside the bin their position implies;
 * evrep_stream_order_violations copies that word to the host (it synchronises the stream).  With a
 * non-zero count the result is not the reference's: use evrep_taf_stream for unordered input.
 * K must be 4 or 8; abin <= 262143; tiles of at most 4096 pixels. */
int64_t evrep_taf_stream_sliced_scratch_bytes(int64_t n_events, int n_windows, int64_t total_bins,
                                               int H, int W, int K);
int evrep_taf_stream_sliced(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p,
                             int64_t n_events, const evrep_taf_window* windows_host, int n_windows,
                             int abin, int H, int W, int K,
                             const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                             float* state_inout, int emit_state_every_window,
                             float* out, int64_t out_stride, uint8_t* out_u8, int64_t out_u8_stride,
                             void* scratch, int64_t scratch_bytes,
                             void* ev_tiles_begin, void* ev_tiles_end, evrep_stream_t stream);
int evrep_stream_order_violations(const void* scratch, uint32_t* host_out, evrep_stream_t stream);
/* Counts the adjacent pairs t[i] > t[i+1] into *violations_dev (device memory, 4 bytes); 0 = the
 * stream may use the *_ordered entry points. */
int evrep_events_order_check(const uint32_t* t, int64_t n_events, uint32_t* violations_dev, evrep_stream_t stream);

/* ----------------------------------------- V2: Event Volume, whole streams ----------------
 * generate_eventvolume.py:15-42 for a list of ordered, non-overlapping windows in one call:
 * window w holds events [ev_begin, ev_end) and t_norm = (t - t0) / tw (float64, driver :141).
 * Same two steps as evrep_taf_stream: bucketing by (sensor tile, window) into 4-byte records,
 * then one CTA per tile that accumulates its [2K, tile] slice in shared memory and writes it
 * with TMA bulk stores.  out: f32 [n_windows][2K,H,W] (window w at out + w * out_stride).
 * tw <= 262143 us (use evrep_event_volume per window beyond that); 2K * tile floats must fit
 * in shared memory (K <= 12 at 512x640). */
typedef struct {
    int64_t ev_begin;
    int64_t ev_end;
    int64_t t0;
} evrep_ev_window;

int64_t evrep_event_volume_stream_scratch_bytes(int64_t n_events, int n_windows, int H, int W);
int evrep_event_volume_stream(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p,
                              int64_t n_events, const evrep_ev_window* windows_host, int n_windows,
                              int64_t tw, int H, int W, int K,
                              const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                              float* out, int64_t out_stride, void* scratch, int64_t scratch_bytes,
                              evrep_stream_t stream);

/* -------------------- V2 (time-ordered input): Event Volume for overlapping, nested windows ------
 * generate_eventvolume.py:118-169 encodes, for every label, the events of the last 250 / 500 /
 * 1000 ms: windows that nest inside a label and overlap between labels.  The caller cuts the
 * stream at every window boundary into consecutive segments -- events [ev_begin, ev_end), all of
 * them with start_time <= t < start_time + 262144 -- and gives every window ("span") as a run of
 * segments first_segment .. last_segment (inclusive; first > last = no events) with its own origin
 * t0 and length tw: t_norm = (t - t0) / tw in float64 (:141).  The events are sorted once by
 * (segment, sensor tile) with the one-pass slice sort of evrep_taf_stream_sliced; every span
 * then re-reads the 4-byte records of its segments and splats them with its own normalisation
 * (shared-memory fixed-point accumulators as in evrep_event_volume_stream, one CTA per tile).
 * No limit on tw below 2^31 us.  Timestamps must be non-decreasing over the segments.
 * out (nullable): f32 [n_spans][2K,H,W], values / 5 * 255 (:37).  out_u8 (nullable): u8
 * [n_spans][2K,H,W], the file bytes when no resize follows (clamped at 255, truncated, :158-160).
 * evrep_event_volume_u8_batch turns a batch of float volumes into those bytes with the optional
 * nearest resize of the gen1 policy (:150). */
typedef struct {
    int64_t ev_begin;
    int64_t ev_end;
    int64_t start_time;
} evrep_ev_segment;

typedef struct {
    int32_t first_segment;
    int32_t last_segment;
    int64_t t0;
    int64_t tw;
} evrep_ev_span;

int64_t evrep_event_volume_spans_scratch_bytes(int64_t n_events, int n_segments, int n_spans, int H, int W, int K);
int evrep_event_volume_spans(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p,
                             int64_t n_events, const evrep_ev_segment* segments_host, int n_segments,
                             const evrep_ev_span* spans_host, int n_spans, int H, int W, int K,
                             const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                             float* out, int64_t out_stride, uint8_t* out_u8, int64_t out_u8_stride,
                             void* scratch, int64_t scratch_bytes,
                             void* ev_tiles_begin, void* ev_tiles_end, evrep_stream_t stream);
int evrep_event_volume_u8_batch(const float* volumes, int64_t volume_stride, int n, int C, int H, int W,
                                int Ht, int Wt, const int32_t* ysrc, const int32_t* xsrc, uint8_t* out,
                                evrep_stream_t stream);

/* ------------------------------------------------- E1 + E2 over a whole stream -------
 * generate_eventcountimage.py:130-182 for many labels in one call.  The driver's windows (the
 * last N events before a label, several N per label) nest and overlap; the caller cuts the
 * stream at every window boundary into consecutive segments [ev_begin, ev_end) and lists the
 * windows as runs of segments [first_segment, last_segment], ordered by last_segment
 * (first_segment == last_segment + 1 denotes an empty window).
 * frames_out: u8 [n_emits][2,H,W] event counts per (polarity, pixel), saturated at 255 --
 * frame i belongs to emits_host[i]; H*W and frame_stride must be multiples of 4.
 * A window may span as many segments as fit the shared-memory ring (EVREP_ERR_RANGE otherwise;
 * e.g. 39 segments at 512x640, 187 at 240x304).
 * evrep_count_lut_u8_batch: value LUT (:32-41), nearest resize and uint8 truncation (:164,180)
 * for all frames: out u8 [n_windows][2,Ht,Wt]. */
typedef struct {
    int64_t ev_begin;
    int64_t ev_end;
} evrep_count_segment;

typedef struct {
    int32_t first_segment;
    int32_t last_segment;
} evrep_count_emit;

int64_t evrep_count_stream_scratch_bytes(int64_t n_events, int n_segments, int n_emits, int H, int W);
int evrep_count_stream(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                       const evrep_count_segment* segments_host, int n_segments,
                       const evrep_count_emit* emits_host, int n_emits, int H, int W,
                       const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                       uint8_t* frames_out, int64_t frame_stride, void* scratch, int64_t scratch_bytes,
                       evrep_stream_t stream);
int evrep_count_lut_u8_batch(const uint8_t* frames, int64_t frame_stride, int64_t n_windows, int H, int W,
                             int Ht, int Wt, const int32_t* ysrc, const int32_t* xsrc, uint8_t* out,
                             evrep_stream_t stream);

/* ------------------------------------------------- A1 + A2 over a whole stream -------
 * generate_surfaceofactiveevents.py:44-69 for every label of a recording in one call: window w
 * holds the events [ev_begin, ev_end) the driver (:147-175) hands to the encoder for label
 * `now`; t_first / t_last are the timestamps of its first and last event (ignored when the
 * window is empty).  Windows are consecutive and do not overlap.  The per-pixel state (latest
 * float32 timestamp, :54) is carried from window to window in shared memory and written back to
 * memory_inout (f32 [2,H,W]; read only when has_memory != 0) at the end.
 * latest_out: f32 [n_windows][2,H,W], window w at latest_out + w * latest_stride -- the value the
 * reference calls t_img after the merge with `memory` (:52).
 * evrep_sae_decay_u8_batch: the L decays exp(lambda (t_img - now)) * 255 (:55-63), the nearest
 * resize and the uint8 truncation of the driver (:186-204) for all windows:
 * out u8 [n_windows][L,2,Ht,Wt]; now_f32: device array of float32(now) per window. */
typedef struct {
    int64_t ev_begin;
    int64_t ev_end;
    int64_t now;
    int64_t t_first;
    int64_t t_last;
} evrep_sae_window;

int64_t evrep_sae_stream_scratch_bytes(int64_t n_events, const evrep_sae_window* windows_host, int n_windows,
                                       int H, int W);
int evrep_sae_stream(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n_events,
                     const evrep_sae_window* windows_host, int n_windows, int H, int W,
                     const uint16_t* xmap, const uint16_t* ymap, int sensor_h, int sensor_w,
                     float* memory_inout, int has_memory, float* latest_out, int64_t latest_stride,
                     void* scratch, int64_t scratch_bytes, evrep_stream_t stream);
int evrep_sae_decay_u8_batch(const float* latest, int64_t latest_stride, const float* now_f32, int64_t n_windows,
                             int H, int W, int Ht, int Wt, const int32_t* ysrc, const int32_t* xsrc,
                             const float* lambdas_host, int L, uint8_t* out, evrep_stream_t stream);

/* ------------------------------------------------- time-surface pair (8f) ------------
 * generate_opticalflow.py:72-92 (generate_timesurface with zero-initialised volumes): out_all =
 * last timestamp per pixel, out_old = last timestamp older than (newest - 50000), both shifted to
 * the oldest timestamp of the call, scaled by 255 / (newest - 50000 - oldest) in float64 and
 * clamped below at 0.  out_*: f64 [H,W].  Events outside the grid are dropped (the driver filters
 * them, :176); polarity is not used.  scratch: evrep_timesurface_scratch_bytes bytes, zero on
 * entry, left zeroed. */
int64_t evrep_timesurface_scratch_bytes(int H, int W);
int evrep_timesurface(const uint32_t* t, const uint16_t* x, const uint16_t* y, int64_t n, int H, int W,
                      void* scratch, double* out_old, double* out_all, evrep_stream_t stream);

/* ------------------------------------------------- training-time read path (8f) -----
 * data/dataset.py:219-234 for a batch of samples already in device memory.  Sample s is the
 * raw uint8 content of its file(s), [C, Hs, Ws] at files + s * file_stride (for TAF K = 8 the
 * bins4 file followed by the bins8 file, :294-308).  One pass does: float conversion, nearest
 * resize to (up_h, up_w) = int(input_img_size * sr) (F.interpolate legacy index rule), / 255,
 * crop img[:, -cy : Hin - cy, -cx : Win - cx] (cy, cx <= 0 as drawn at :153-161) and the
 * horizontal flip.  out: f32 [n, C, Hin, Win] (the reference returns [C, Hin, Win, 1, 1] per
 * sample).  aug: DEVICE array of n descriptors.  The crop must lie inside the resized image:
 * Hin - cy <= up_h and Win - cx <= up_w. */
typedef struct {
    int32_t up_h;
    int32_t up_w;
    int32_t cy;
    int32_t cx;
    int32_t flip;
} evrep_sample_aug;

int evrep_load_samples(const uint8_t* files, int64_t file_stride, int n, int C, int Hs, int Ws,
                       const evrep_sample_aug* aug, int Hin, int Win, float* out, evrep_stream_t stream);

/* ------------------------------------------------- T3 / R1 / W1: output epilogues ----
 * evrep_nearest_resize: F.interpolate(mode='nearest') as used at generate_taf.py:222 --
 * out[c, Y, X] = in[c, ysrc[Y], xsrc[X]] with the legacy index maps (int32, device).
 * evrep_quantize_u8: the `.astype(np.uint8)` of the drivers (float -> uint8 truncation),
 * with the Event Volume clamp `np.where(v > 255, 255, v)` (generate_eventvolume.py:155-157)
 * when clamp255 != 0.
 * evrep_taf_leaky_u8: generate_taf.py:226-235 fused -- leaky_transform (:69-76),
 * view [K,2,Ht,Wt], flip of the slot axis, nearest resize, uint8 truncation.  `out` is
 * u8 [K,2,Ht,Wt] with slot 0 = newest: its first half is the bins{K/2} file and its
 * second half the bins{K} file (data/dataset.py:294-308). */
int evrep_nearest_resize(const float* in, int C, int H, int W, int Ht, int Wt,
                         const int32_t* ysrc, const int32_t* xsrc, float* out, evrep_stream_t stream);
int evrep_quantize_u8(const float* in, int64_t n, int clamp255, uint8_t* out, evrep_stream_t stream);
int evrep_taf_leaky_u8(const float* volume, int K, int H, int W, int Ht, int Wt,
                       const int32_t* ysrc, const int32_t* xsrc, uint8_t* out, evrep_stream_t stream);
/* n_windows tensors at once: window w at volumes + w * volume_stride floats -> out + w * 2K*Ht*Wt. */
int evrep_taf_leaky_u8_batch(const float* volumes, int64_t volume_stride, int n_windows, int K, int H, int W,
                             int Ht, int Wt, const int32_t* ysrc, const int32_t* xsrc, uint8_t* out,
                             evrep_stream_t stream);
int evrep_leaky_transform(const float* in, int64_t n, float* out, evrep_stream_t stream);

/* ----------------------------------- S1-S6: data/sparse_ops.py (batched plugin surface) ----
 * The reference's online encoders, called as `to_volume(events, B, shape, iter, memory,
 * events_window, volume_bins, infer_time)` at data/fetcher.py:53.  `events` is the float64
 * [N,5] matrix (b, x, y, t, p) ([N,7] (b, x, y, t, c, p, feature) for evrep_sparse_taf).
 *
 * evrep_sparse_splat: temporal bilinear splat with centres 0..C-1 into the pixel-major
 *   accumulator acc f32 [B*H*W, C, 2] (zeroed by the call), slot 1-p.
 *   mode 0: t* = (K t) / window             sparse_ops.py:12  (generate_agile_event_volume_cuda, full)
 *   mode 1: t* = ((t - iter) + infer) / window * K      :15  (incremental, C = 2)
 *   mode 2: t* = ((K - 1) t) / window                   :56  (generate_event_volume_cuda)
 * evrep_pixel_major_to_planar: [B*HW, C2] -> [B, C2, HW]  (the permute(0,3,1,2,4) of :34,:68).
 * evrep_sparse_agile_shift: :25-32 -- past[:, K-1] += fresh[:, 0] IN PLACE (the reference mutates
 *   the caller's tensor), out_state = cat(past[:, 1:], fresh[:, 1:]).  Layouts [BHW, K, 2] / [BHW, 2, 2].
 * evrep_sparse_taf: :72-85, out f32 [B, C, H, W, 2]; plane 1 becomes -1e8 where 0, +1 elsewhere.
 * evrep_sparse_event_frame: :88-107, out f32 [B, 2, H, W], 255 where any event.
 * evrep_sparse_to_dense: :109-121, locations int64 [N,3] (b, y, x), features f32 [N,C] -> [B,H,W,C]. */
int evrep_sparse_splat(const double* events, int64_t n, int B, int H, int W, int C, int mode, float K,
                       float window, float iter, float infer, float* acc, evrep_stream_t stream);
int evrep_pixel_major_to_planar(const float* in, int B, int64_t HW, int C2, float* out, evrep_stream_t stream);
int evrep_sparse_agile_shift(float* past, const float* fresh, int64_t BHW, int K, float* out_state,
                             evrep_stream_t stream);
int evrep_sparse_taf(const double* events, int64_t n, int B, int H, int W, int C, float* out, evrep_stream_t stream);
int evrep_sparse_event_frame(const double* events, int64_t n, int B, int H, int W, float* out, evrep_stream_t stream);
int evrep_sparse_to_dense(const int64_t* locations, const float* features, int64_t n, int B, int H, int W, int C,
                          float* out, evrep_stream_t stream);

/* S2 / S6 as device code, and an online TAF bin for the streaming driver (data/fetcher.py:35-62).
 * evrep_dense_to_sparse: data/sparse_ops.py:123-135 -- rows of the dense tensor [sizes..., C] (1 to 4
 *   leading dims, batch first) with a non-zero |.|-sum, in row-major order: locations int64 [N, n_dims]
 *   (spatial indices, then the batch index) and features f32 [N, C].  Both outputs are sized for all
 *   rows by the caller; N is left on the device at *count_dev (u32) -- read it back like
 *   torch.nonzero does.  block_scratch: evrep_compact_scratch_bytes(rows) bytes.
 * evrep_event_memory_update: :40-42 -- merged = cat(memory, events) (float64 [*,5]) and
 *   memory_out = the rows of merged with column 3 >= keep_from, order kept; count at *count_dev.
 * evrep_taf_online_bin: one 10 ms bin of a BATCH of recordings with the FIFO state carried on the
 *   device: events float64 [N,5] (b, x, y, t, p) of a fetch step, of which t_lo <= t < t_hi form the
 *   bin, t_norm = (t - t_lo) / t_span; per sample the rule of generate_taf.py:19-58 (a sample
 *   without events in the bin is not aged).  state f32 [B,H,W,2,K] updated in place; out (nullable)
 *   f32 [B,2K,H,W], channel 2k + p.  scratch: evrep_taf_online_scratch_bytes, zero on entry and return. */
int64_t evrep_compact_scratch_bytes(int64_t n_rows);
int evrep_dense_to_sparse(const float* dense, const int64_t* sizes, int n_dims, int C, int64_t* locations,
                          float* features, uint32_t* block_scratch, uint32_t** count_dev, evrep_stream_t stream);
int evrep_event_memory_update(const double* memory, int64_t n_memory, const double* events, int64_t n_events,
                              double keep_from, double* merged, double* memory_out, uint32_t* block_scratch,
                              uint32_t** count_dev, evrep_stream_t stream);
int64_t evrep_taf_online_scratch_bytes(int B, int H, int W);
int evrep_taf_online_bin(const double* events, int64_t n, double t_lo, double t_hi, double t_span, int B, int H,
                         int W, int K, float* state, float* out, void* scratch, evrep_stream_t stream);

/* -------------------- N1: data/event_representation_tool/src/event_queue_tensor.cpp:10-118 ----
 * As the extension behaves (its deques are never fed during the event loop): every event
 * (b, x, y, t, p, z) f32 [N,6] adds 1 - (start[b] + abin (z+1) - t) / abin to cell (p, b, y, x);
 * positive totals land in queue slot Q-1 of plane 0, plane 1 is -1.  out f64 [2, Q, 2, B, H, W];
 * totals: f32 [2*B*H*W] scratch. */
int evrep_event_queue_tensor(const float* events, int64_t n, int Q, int B, int H, int W,
                             const int32_t* start_times, int abin, float* totals, double* out,
                             evrep_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* EVREP_H */
