"""Time-surface pair against the golden vectors recorded from the
reference's ``generate_timesurface`` and against the oracle on a larger stream.  float64
arithmetic in the reference's operation order: bit-exact."""
import numpy as np
import pytest
import torch

from frlw_evd_b200 import generate_opticalflow as gof
from frlw_evd_b200 import ops, synth
from oracle import encoders as oe

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ts_golden():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "timesurface.npz"))


@pytest.mark.parametrize("tag", ["a", "b"])
def test_golden(ts_golden, tag):
    H, W = [int(v) for v in ts_golden["ts_%s_shape" % tag]]
    ev = ts_golden["ts_%s_events" % tag]
    v1, v2 = gof.generate_timesurface(ev, np.zeros((H, W)), np.zeros((H, W)), 0.0)
    assert v1.dtype == np.float64 and np.array_equal(v1, ts_golden["ts_%s_v1" % tag])
    assert np.array_equal(v2, ts_golden["ts_%s_v2" % tag])
    t1, t2 = gof.generate_timesurface(torch.from_numpy(ev).cuda(), torch.zeros(H, W), torch.zeros(H, W))
    assert np.array_equal(t1.cpu().numpy(), v1) and np.array_equal(t2.cpu().numpy(), v2)


def test_gen1_half_second_window_matches_oracle_and_leaves_scratch_clean():
    t, x, y, p = synth.make_stream(240, 304, 500_000, 1e6, 17)
    t = (t + 1_234_567).astype(np.uint32)
    ev = ops.EventStream.from_numpy(t, x, y, p)
    aos = np.stack([x, y, t, p], 1).astype(np.float64)
    want1, want2 = oe.timesurface_pair(aos, (240, 304))
    for _ in range(2):                                   # second call: the scratch was left zeroed
        got1, got2 = ops.timesurface(ev, (240, 304))
        assert np.array_equal(got1.cpu().numpy(), want1) and np.array_equal(got2.cpu().numpy(), want2)
    # events off the grid are dropped; no events: the zero surfaces come back
    v1, v2 = gof.generate_timesurface(np.zeros((0, 4)), np.zeros((4, 5)), np.zeros((4, 5)))
    assert not v1.any() and not v2.any()
