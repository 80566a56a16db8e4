"""Parity of the CUDA encoders (through the C ABI) with the CPU oracle and the golden
vectors recorded from the reference.  Integer results bit-exact; float results within
the north-star tolerance 1e-5 relative / 1e-6 absolute."""
import numpy as np
import pytest
import torch

from frlw_evd_b200 import ops, synth
from oracle import encoders as oe
from oracle import psee_io

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6
LAMBDAS = [0.00001, 0.0000025, 0.000001]
DEV = "cuda"


def close(a, b):
    a = a.cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return np.allclose(a, b, rtol=RTOL, atol=ATOL)


def exact(a, b):
    a = a.cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    return np.array_equal(a, b)


def stream(H, W, dur, rate, seed):
    t, x, y, p = synth.make_stream(H, W, dur, rate, seed)
    aos = torch.from_numpy(np.stack([x, y, t, p], 1).astype(np.float64))
    return (t, x, y, p), aos


def test_decode_dat_bit_exact():
    (t, x, y, p), _ = stream(720, 1280, 20000, 5e6, 5)
    rec = synth.pack_dat_records(t, x, y, p)
    for n in (len(t), len(t) - 3, 1, 0):
        raw = torch.from_numpy(rec[:n].view(np.uint8).copy()).to(DEV)
        ev = ops.decode_dat(raw)
        want = psee_io.decode_records(rec[:n].view(psee_io.RECORD))
        assert exact(ev.t, want["t"]) and exact(ev.x, want["x"]) and exact(ev.y, want["y"]) and exact(ev.p, want["p"])
    ev = ops.EventStream.from_numpy(t, x, y, p)
    assert exact(ev.to_aos64(), np.stack([x, y, t, p], 1).astype(np.float64))


def test_count_image_golden_and_stream(golden):
    H, W = [int(v) for v in golden["shape"]]
    got = ops.count_image_aos64(torch.from_numpy(golden["events"]).to(DEV), (H, W))
    assert exact(got, golden["eci"])
    (t, x, y, p), aos = stream(240, 304, 50000, 4e6, 1000)
    want = oe.count_image(aos, (240, 304))
    assert exact(ops.count_image_aos64(aos.to(DEV), (240, 304)), want)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    assert exact(ops.count_image(ev, (240, 304)), want)
    # nested last-N windows == independent encodes of events[-N:]
    outs = ops.count_images_nested(ev, [50000, 100000, 300000], (240, 304))
    for n, o in zip([50000, 100000, 300000], outs):
        assert exact(o, oe.count_image(aos[-n:], (240, 304)))
    # scratch is left clean
    assert exact(ops.count_image(ev.slice(0, 0), (240, 304)), np.zeros((2, 240, 304), np.float32))


def test_count_image_gen4_policy():
    (t, x, y, p), aos = stream(720, 1280, 40000, 10e6, 1002)
    scaled = aos.clone()
    scaled[:, 0] *= 640 / 1280
    scaled[:, 1] *= 512 / 720
    want = oe.count_image(scaled, (512, 640))
    maps = ops.make_coord_maps((720, 1280), (512, 640), DEV)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    assert exact(ops.count_image(ev, (512, 640), maps), want)
    assert exact(ops.count_image_aos64(scaled.to(DEV), (512, 640)), want)


def test_sae_golden(golden):
    H, W = [int(v) for v in golden["shape"]]
    ev = torch.from_numpy(golden["sae_events"]).to(DEV)
    o0, m0 = ops.sae_aos64(ev[:4000], (H, W), LAMBDAS, None, np.int64(40000))
    assert exact(m0, golden["sae_mem0"]) and close(o0, golden["sae_out0"])
    o1, m1 = ops.sae_aos64(ev[4000:], (H, W), LAMBDAS, m0, np.int64(50000))
    assert exact(m1, golden["sae_mem1"]) and close(o1, golden["sae_out1"])


def test_sae_stream_state_bit_exact():
    (t, x, y, p), aos = stream(240, 304, 6_000_000, 2e5, 1001)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    cut = len(t) // 2
    now0, now1 = int(t[cut - 1]) + 1, int(t[-1]) + 1
    w0, wm0 = oe.sae_surfaces(aos[:cut], (240, 304), LAMBDAS, None, np.int64(now0))
    w1, wm1 = oe.sae_surfaces(aos[cut:], (240, 304), LAMBDAS, wm0, np.int64(now1))
    g0, gm0 = ops.sae(ev.slice(0, cut), (240, 304), LAMBDAS, None, now0)
    g1, gm1 = ops.sae(ev.slice(cut, len(t)), (240, 304), LAMBDAS, gm0, now1)
    assert exact(gm0, wm0) and exact(gm1, wm1)          # float32-rounded timestamps: bit-exact
    assert close(g0, w0) and close(g1, w1)


def test_count_whole_stream_nested_overlapping_windows_bit_exact():
    """evrep_count_stream + evrep_count_lut_u8_batch: the driver's last-N windows of consecutive
    labels (nested inside a label, overlapping between labels, one empty, one reaching back to
    event 0) equal the oracle's count image byte for byte, with and without the gen1 resize."""
    (t, x, y, p), aos = stream(240, 304, 400_000, 1.5e6, 41)
    # a hot pixel so that counts pass the saturation point of the LUT and of the 8-bit ring slots
    x[::7], y[::7], p[::7] = 100, 50, 1
    aos = torch.from_numpy(np.stack([x, y, t, p], 1).astype(np.float64))
    ev = ops.EventStream.from_numpy(t, x, y, p)
    ends = [0, 30_000, 90_000, 140_000, 200_000, 290_000, len(t)]
    sizes = (20_000, 50_000, 120_000)
    windows = [(max(e - n, 0), e) for e in ends for n in sizes]
    frames = ops.count_stream(ev, windows, (240, 304))
    native = ops.count_lut_u8_batch(frames)
    resized = ops.count_lut_u8_batch(frames, (256, 320))
    for i, (lo, hi) in enumerate(windows):
        want = oe.count_image(aos[lo:hi], (240, 304))
        assert exact(native[i], want.numpy().astype(np.uint8)), i
        assert exact(resized[i], oe.nearest_resize(want, (256, 320)).numpy().astype(np.uint8)), i


def test_count_whole_stream_with_coordinate_maps_equals_per_label_path():
    (t, x, y, p), _ = stream(720, 1280, 300_000, 8e6, 42)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    maps = ops.make_coord_maps((720, 1280), (512, 640), DEV)
    sizes = (400_000, 800_000, 1_200_000)
    ends = [500_000, 900_000, 1_700_000, len(t)]
    windows = [(max(e - n, 0), e) for e in ends for n in sizes]
    got = ops.count_lut_u8_batch(ops.count_stream(ev, windows, (512, 640), maps))
    for j, e in enumerate(ends):
        lo = max(e - max(sizes), 0)
        want = ops.count_images_u8(ev.slice(lo, e), sizes, (512, 640), (512, 640), maps)
        assert exact(got[3 * j:3 * j + 3], want), j


def _sae_window_list(t, bounds, nows):
    return [(lo, hi, now, int(t[lo]) if hi > lo else 0, int(t[hi - 1]) if hi > lo else 0)
            for (lo, hi), now in zip(bounds, nows)]


def test_sae_whole_stream_matches_oracle_window_by_window():
    """evrep_sae_stream: a first window of several seconds (cut into 250 ms bins inside), then
    short ones, one of them empty; the frames and the carried state are bit-exact (float32-rounded
    timestamps), the uint8 decays equal up to truncation flips."""
    torch.set_num_threads(1)
    (t, x, y, p), aos = stream(240, 304, 3_000_000, 6e5, 77)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    edges = [0, int(np.searchsorted(t, 2_400_000)), int(np.searchsorted(t, 2_450_000)), int(np.searchsorted(t, 2_450_000)),
             int(np.searchsorted(t, 2_700_000)), len(t)]
    bounds = list(zip(edges[:-1], edges[1:]))
    nows = [2_400_000, 2_450_000, 2_500_000, 2_700_000, 3_000_001]
    latest, mem = ops.sae_stream(ev, _sae_window_list(t, bounds, nows), (240, 304))
    u8 = ops.sae_decay_u8_batch(latest, nows, LAMBDAS, (256, 320))
    memory = None
    for i, ((lo, hi), now) in enumerate(zip(bounds, nows)):
        want, memory = oe.sae_surfaces(aos[lo:hi], (240, 304), LAMBDAS, memory, np.int64(now))
        assert exact(latest[i], memory), i
        want_u8 = oe.nearest_resize(want, (256, 320)).numpy().astype(np.uint8).reshape(3, 2, 256, 320)
        diff = np.abs(u8[i].cpu().numpy().astype(np.int16) - want_u8.astype(np.int16))
        assert diff.max() <= 1 and (diff != 0).mean() < 1e-3, i
    assert exact(mem, memory)
    # a second call continues from the returned state
    (t2, x2, y2, p2), aos2 = stream(240, 304, 200_000, 6e5, 78)
    t2 = (t2 + 3_000_001).astype(np.uint32)
    aos2[:, 2] += 3_000_001
    ev2 = ops.EventStream.from_numpy(t2, x2, y2, p2)
    latest2, mem2 = ops.sae_stream(ev2, _sae_window_list(t2, [(0, len(t2))], [3_300_000]), (240, 304), mem)
    _, want_mem2 = oe.sae_surfaces(aos2, (240, 304), LAMBDAS, memory, np.int64(3_300_000))
    assert exact(latest2[0], want_mem2) and exact(mem2, want_mem2) and exact(mem, memory)      # input state untouched


def test_sae_whole_stream_with_coordinate_maps_equals_per_window_path():
    """1MP events on the 512x640 grid: the stream kernel and the per-window kernel agree bit for bit."""
    (t, x, y, p), _ = stream(720, 1280, 400_000, 8e6, 79)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    maps = ops.make_coord_maps((720, 1280), (512, 640), DEV)
    cuts = [0] + [int(np.searchsorted(t, v)) for v in (150_000, 200_000, 250_000, 400_000)]
    nows = [150_000, 200_000, 250_000, 400_000]
    bounds = list(zip(cuts[:-1], cuts[1:]))
    latest, mem = ops.sae_stream(ev, _sae_window_list(t, bounds, nows), (512, 640), None, maps)
    u8 = ops.sae_decay_u8_batch(latest, nows, LAMBDAS)
    memory = None
    for i, ((lo, hi), now) in enumerate(zip(bounds, nows)):
        want_u8, memory = ops.sae_u8(ev.slice(lo, hi), (512, 640), (512, 640), LAMBDAS, memory, now, maps)
        assert exact(latest[i], memory), i
        assert exact(u8[i], want_u8), i
    assert exact(mem, memory)


@pytest.mark.parametrize("K", [5, 8])
def test_event_volume_golden(golden, K):
    H, W = [int(v) for v in golden["shape"]]
    got = ops.event_volume_aos64(torch.from_numpy(golden["events_norm"]).to(DEV), (H, W), K)
    assert close(got, golden["ev_k%d" % K])


def test_event_volume_config1_gen1_k8():
    """BASELINE config 1: Event Volume K=8, GEN1, one 50 ms window (~200k events)."""
    (t, x, y, p), aos = stream(240, 304, 50000, 4e6, 1000)
    norm = aos.clone()
    norm[:, 2] = (norm[:, 2] - 0) / 50000
    want = oe.event_volume(norm, (240, 304), 8)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    assert close(ops.event_volume(ev, 0, 50000, (240, 304), 8), want)
    assert close(ops.event_volume_aos64(norm.to(DEV), (240, 304), 8), want)


def test_taf_bin_golden_sequence(golden):
    H, W = [int(v) for v in golden["shape"]]
    ev = golden["events"]
    K = 8
    state = ops.taf_fresh_state((H, W), K, DEV)
    for it in range(6):
        sel = (ev[:, 2] >= it * 10000) & (ev[:, 2] < (it + 1) * 10000)
        e5 = np.concatenate([ev[sel], np.full((int(sel.sum()), 1), float(it))], axis=1)
        if it in (2, 5):
            e5 = e5[:0]
        e5[:, 2] = (e5[:, 2] - it * 10000) / (10000 + 1e-8)
        out, state = ops.taf_bin_aos64(torch.from_numpy(e5).to(DEV), (H, W), K, state)
        assert close(out, golden["taf_out"][it]), it
        assert close(state, golden["taf_state"][it]), it
    assert close(ops.leaky_transform(torch.from_numpy(golden["taf_out"][4]).to(DEV)), golden["taf_leaky"].reshape(16, H, W))


def test_taf_bin_soa_gen4_policy():
    (t, x, y, p), aos = stream(720, 1280, 30000, 10e6, 1002)
    maps = ops.make_coord_maps((720, 1280), (512, 640), DEV)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    K = 8
    want_state = oe.taf_fresh_state((512, 640), K)
    got_state = ops.taf_fresh_state((512, 640), K, DEV)
    for it in range(3):
        lo, hi = np.searchsorted(t, [it * 10000, (it + 1) * 10000])
        e5 = torch.cat([aos[lo:hi], torch.zeros(hi - lo, 1, dtype=torch.float64)], 1).clone()
        e5[:, 2] = (e5[:, 2] - it * 10000) / (10000 + 1e-8)
        e5[:, 0] *= 0.5
        e5[:, 1] *= 512 / 720
        want_out, want_state = oe.taf_bin_update(e5, (512, 640), want_state, K)
        got_out, got_state = ops.taf_bin(ev.slice(lo, hi), it * 10000, 10000 + 1e-8, (512, 640), K, got_state, maps)
        assert close(got_out, want_out) and close(got_state, want_state)


def test_epilogues():
    g = torch.Generator().manual_seed(0)
    vol = torch.rand((16, 240, 304), generator=g) * 300 - 20
    vol_d = vol.to(DEV)
    want = torch.nn.functional.interpolate(vol[None], size=(256, 320), mode="nearest")[0]
    assert exact(ops.nearest_resize(vol_d, (256, 320)), want)
    pos = vol.clamp(min=0)
    assert exact(ops.quantize_u8(pos.to(DEV), clamp255=True), np.where(pos.numpy() > 255, 255, pos.numpy()).astype(np.uint8))
    state = -torch.rand((16, 240, 304), generator=g) * 7
    state[::3] = -6000.0
    ref = oe.leaky_transform(oe.nearest_resize(state, (256, 320)).view(8, 2, 256, 320)).numpy()
    ref = np.flip(ref, axis=0).astype(np.uint8)
    got = ops.taf_leaky_u8(state.to(DEV), 8, (256, 320)).cpu().numpy()
    # uint8 truncation of a float: allow the 1-LSB flips of log1pf vs the CPU kernel
    diff = np.abs(got.astype(int) - ref.astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3


def test_stream_encoders_edge_cases():
    """No windows, windows without events, events that all fall off the grid."""
    (t, x, y, p), aos = stream(240, 304, 100_000, 5e5, 91)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    n = len(t)
    # no windows at all
    assert ops.count_stream(ev, [], (240, 304)).shape == (0, 2, 240, 304)
    latest, mem = ops.sae_stream(ev, [], (240, 304))
    assert latest.shape == (0, 2, 240, 304)
    # only empty windows
    frames = ops.count_stream(ev, [(5, 5), (n, n)], (240, 304))
    assert int(frames.sum()) == 0
    latest, mem = ops.sae_stream(ev, [(7, 7, 60_000, 0, 0)], (240, 304))
    want, want_mem = oe.sae_surfaces(aos[7:7], (240, 304), LAMBDAS, None, np.int64(60_000))
    assert exact(latest[0], want_mem) and exact(mem, want_mem)
    # every event outside the grid: dropped, like the per-window kernels
    far = ops.EventStream.from_numpy(t, (x + 400).astype(np.uint16), y, p)
    assert int(ops.count_stream(far, [(0, n)], (240, 304)).sum()) == 0
    latest, _ = ops.sae_stream(far, [(0, n, 100_000, int(t[0]), int(t[-1]))], (240, 304))
    floor = np.float32(np.float32(100_000.0) - np.float32(5_000_000.0))
    assert np.all(latest.cpu().numpy() == floor)
    vol = ops.event_volume_stream(far, [(0, n, 0)], 100_000, (240, 304), 5)
    assert float(vol.abs().sum()) == 0.0
