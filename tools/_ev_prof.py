import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
from frlw_evd_b200 import ops
t, x, y, p = bench.get_stream(1002, 2.0, 1e7)
dev = torch.device('cuda', 0)
ev = ops.EventStream.from_numpy(t, x, y, p, dev)
maps = ops.make_coord_maps(bench.SENSOR, bench.GRID, dev)
edges = np.searchsorted(t, np.arange(0, 2000001, 50000))
windows = [(int(edges[i]), int(edges[i + 1]), i * 50000) for i in range(len(edges) - 1)]
out = torch.empty((len(windows), 16, 512, 640), dtype=torch.float32, device=dev)
for _ in range(3):
    ops.event_volume_stream(ev, windows, 50000, (512, 640), 8, maps, out)
torch.cuda.synchronize()
