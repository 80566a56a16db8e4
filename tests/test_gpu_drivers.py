"""The four command-line drivers end to end on the GPU against the oracle drivers (which
reproduce the reference scripts byte for byte, see test_oracle_golden.py): same files,
same names; Event Count Image files byte-identical (and equal to the recorded reference
digests), float-derived uint8 files equal up to rare 1-LSB truncation flips."""
import json
import os

import numpy as np
import pytest

from frlw_evd_b200 import generate_eventcountimage, generate_eventvolume, generate_surfaceofactiveevents, generate_taf
from oracle import drivers as od
from oracle import make_golden as mg

pytestmark = pytest.mark.gpu

PRODUCT = {"count_image": generate_eventcountimage.main, "sae": generate_surfaceofactiveevents.main,
           "event_volume": generate_eventvolume.main, "taf": generate_taf.main}
ORACLE = {"count_image": od.run_count_image, "sae": od.run_sae, "event_volume": od.run_event_volume, "taf": od.run_taf}


def tree(root):
    out = {}
    for d, _, files in os.walk(root):
        for f in files:
            out[os.path.relpath(os.path.join(d, f), root)] = os.path.join(d, f)
    return out


@pytest.mark.parametrize("case", mg.DRIVER_CASES, ids=[c[0] for c in mg.DRIVER_CASES])
@pytest.mark.parametrize("rep", ["count_image", "sae", "event_volume", "taf"])
def test_driver_files(case, rep, tmp_path):
    import torch
    torch.set_num_threads(1)                 # the oracle's SAE scatter is sequential by definition
    raw = str(tmp_path / "raw")
    mg.write_case(raw, case)
    got_dir, want_dir = str(tmp_path / "got"), str(tmp_path / "want")
    PRODUCT[rep](["-raw_dir", raw, "-label_dir", raw, "-target_dir", got_dir, "-dataset", case[2]])
    ORACLE[rep](raw, raw, want_dir, case[2])
    got, want = tree(got_dir), tree(want_dir)
    assert sorted(got) == sorted(want) and len(want) > 0
    flips = total = 0
    for name in want:
        a = np.fromfile(got[name], dtype=np.uint8).astype(np.int16)
        b = np.fromfile(want[name], dtype=np.uint8).astype(np.int16)
        assert a.shape == b.shape, name
        if rep == "count_image":
            assert np.array_equal(a, b), name
        else:
            diff = np.abs(a - b)
            assert diff.max() <= 1, (name, int(diff.max()))
            flips += int((diff != 0).sum())
            total += diff.size
    if rep == "count_image":
        with open(os.path.join(os.path.dirname(__file__), "golden", "drivers_digest.json")) as fh:
            digests = json.load(fh)["%s/%s" % (case[0], rep)]
        assert mg._digest_tree(got_dir) == digests          # byte-identical to the reference's own files
    else:
        assert flips / total < 1e-3, (flips, total)


@pytest.mark.parametrize("rep", ["count_image", "sae", "event_volume", "taf"])
def test_multi_gpu_worker_writes_the_same_files(rep, tmp_path):
    """`multi_gpu.encode_one` (what every rank runs for its recordings) == the command line, file by file."""
    from frlw_evd_b200 import multi_gpu
    case = mg.DRIVER_CASES[0]
    raw = str(tmp_path / "raw")
    mg.write_case(raw, case)
    cli_dir, rank_dir = str(tmp_path / "cli"), str(tmp_path / "rank")
    PRODUCT[rep](["-raw_dir", raw, "-label_dir", raw, "-target_dir", cli_dir, "-dataset", case[2]])
    stats = [multi_gpu.encode_one(rep, case[2], mode, name, event_file, label_file, rank_dir)
             for mode, name, event_file, label_file, _size in multi_gpu.list_recordings(raw, raw)]
    assert mg._digest_tree(cli_dir) == mg._digest_tree(rank_dir)
    written = sum(os.path.getsize(p) for p in tree(rank_dir).values())
    assert sum(s["bytes_written"] for s in stats) == written and all(s["windows"] > 0 for s in stats)
