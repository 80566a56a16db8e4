"""Training-time read path: the CUDA kernel behind ``evrep_load_samples``
against the outputs of the reference's own dataset class (tests/golden/dataset_read.npz) and
against the oracle on larger, seeded inputs.  Every value is one of 256 floats: bit-exact."""
import numpy as np
import pytest
import torch

from frlw_evd_b200.data import dataset_gpu as dg
from oracle import dataset_read as dr
from oracle import make_golden as mg

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_golden_samples_bit_exact(dataset_golden, tmp_path):
    g = dataset_golden
    for i in range(len(mg.DATASET_CASES)):
        h, w, hin, win, cx, cy, flip = [int(v) for v in g["ds%d_params" % i]]
        sr = float(g["ds%d_sr" % i][0])
        # through the files, like the reference
        for sub, key in (("bins4", "ds%d_bins4"), ("bins8", "ds%d_bins8")):
            (tmp_path / "train" / sub).mkdir(parents=True, exist_ok=True)
            g[key % i].tofile(str(tmp_path / "train" / sub / ("rec_%d.npy" % i)))
        paths = dg.taf_file_names(str(tmp_path), "train", "rec", i, 8)
        got = dg.load_sample(paths, 16, (h, w), (hin, win), sr, cx, cy, bool(flip), DEV)
        want = g["ds%d_out" % i]
        assert tuple(got.shape) == want.shape
        assert np.array_equal(got.cpu().numpy(), want), i


@pytest.mark.parametrize("img_size,in_size", [((256, 320), (256, 320)), ((512, 640), (512, 640)), ((30, 37), (41, 53))])
def test_batch_matches_oracle(img_size, in_size):
    rng = np.random.default_rng(3)
    n, C = 5, 16 if img_size[0] < 500 else 4
    vols = rng.integers(0, 256, (n, C) + tuple(img_size), dtype=np.uint8)
    params = []
    for i in range(n):
        sr = 1.0 if i == 0 else float(rng.uniform(1.0, 1.5))
        lo_x, lo_y = int(in_size[1] - sr * in_size[1]), int(in_size[0] - sr * in_size[0])
        cx = int(rng.uniform(lo_x, 0)) if sr > 1.0 else 0
        cy = int(rng.uniform(lo_y, 0)) if sr > 1.0 else 0
        params.append((sr, cx, cy, bool(i % 2)))
    got = dg.augment_batch(torch.from_numpy(vols).to(DEV), in_size, params).cpu().numpy()
    for i, (sr, cx, cy, flip) in enumerate(params):
        want = dr.augment_sample(vols[i].astype(np.float32), in_size, sr, cx, cy, flip)[:, :, :, 0, 0]
        assert np.array_equal(got[i], want), (i, sr, cx, cy, flip)


def test_rejects_a_crop_outside_the_resized_image():
    vols = torch.zeros((1, 2, 8, 8), dtype=torch.uint8, device=DEV)
    with pytest.raises(ValueError):
        dg.augment_batch(vols, (8, 8), [(1.0, -1, 0, False)])
