"""The N > 1 path on CPU: two gloo ranks shard a list of recordings and reduce the
statistics vector (the path's only collective).  The encoders themselves are stubbed --
they need a GPU -- so this covers the host-side logic of frlw_evd_b200.multi_gpu."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from frlw_evd_b200 import multi_gpu, synth


def test_lpt_assignment_is_a_balanced_partition():
    sizes = [160, 160, 160, 160, 800, 90, 90, 10]
    plan = multi_gpu.assign_recordings(sizes, 3)
    flat = sorted(i for p in plan for i in p)
    assert flat == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in p) for p in plan]
    assert max(loads) == 800 and min(loads) >= 410
    # equal sizes degrade to round robin: 64 recordings over 8 ranks -> 8 each
    assert [len(p) for p in multi_gpu.assign_recordings([20] * 64, 8)] == [8] * 8
    assert multi_gpu.assign_recordings([5, 5, 5], 1) == [[0, 1, 2]]


def test_list_recordings_is_sorted_and_sized(tmp_path):
    for mode, name, seed in (("train", "b", 1), ("train", "a", 2), ("val", "c", 3)):
        synth.write_recording(str(tmp_path), str(tmp_path), mode, name, "gen1", 20000, 1e6, seed)
    found = multi_gpu.list_recordings(str(tmp_path), str(tmp_path))
    assert [(m, n) for m, n, *_ in found] == [("train", "a"), ("train", "b"), ("val", "c")]
    assert all(size > 8 * 19000 for *_, size in found)


def _worker(rank, world, port, sizes, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    recordings = [("train", "r%d" % i, "", "", s) for i, s in enumerate(sizes)]
    seen = []

    def encode(rec):
        seen.append(rec[1])
        return {"events": rec[-1] // 8, "windows": 3, "bytes_written": 100}

    totals = multi_gpu.run(recordings, encode, torch.device("cpu"), rank, world)
    torch.save({"totals": totals, "seen": seen}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


def test_two_gloo_ranks_shard_and_reduce(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    sizes = [800, 160, 160, 160, 480, 80, 80]
    mp.spawn(_worker, args=(2, port, sizes, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(tmp_path / ("rank%d.pt" % r)) for r in (0, 1))
    assert sorted(r0["seen"] + r1["seen"]) == sorted("r%d" % i for i in range(len(sizes)))
    assert not set(r0["seen"]) & set(r1["seen"])
    for totals in (r0["totals"], r1["totals"]):          # identical on every rank after the all-reduce
        assert totals["recordings"] == len(sizes)
        assert totals["events"] == sum(s // 8 for s in sizes)
        assert totals["windows"] == 3 * len(sizes) and totals["bytes_written"] == 100 * len(sizes)
        assert totals["seconds"] >= 0
    assert r0["totals"] == r1["totals"]
