"""Test helpers: oracle-side execution of a TAF window list (the reference driver's
inner loops, generate_taf.py:195-222, on in-memory events)."""
import math

import numpy as np
import torch

from oracle import encoders as oe


def staged(t, x, y, p, lo, hi):
    return torch.from_numpy(np.stack([x[lo:hi], y[lo:hi], t[lo:hi], p[lo:hi]], 1).astype(np.float64))


def oracle_taf_windows(t, x, y, p, windows, abin, grid, K, scale=None, state=None):
    """windows: (ev_begin, ev_end, start_time, n_bins, fresh).  Returns (list of [2K,H,W]
    outputs, final state).  ``scale`` = (rw, rh) for the gen4 coordinate policy."""
    outs = []
    if state is None:
        state = oe.taf_fresh_state(grid, K)
    vol = None
    for (lo, hi, start, n_bins, fresh) in windows:
        ev = staged(t, x, y, p, lo, hi)
        z = torch.zeros_like(ev[:, 0])
        for i in range(n_bins):
            a, b = start + i * abin, start + (i + 1) * abin
            z = torch.where((ev[:, 2] >= a) & (ev[:, 2] <= b), torch.zeros_like(z) + i, z)
        if fresh:
            state = oe.taf_fresh_state(grid, K)
        for i in range(n_bins):
            e = ev[z == i].clone()
            t_min = start + i * abin
            e[:, 2] = (e[:, 2] - t_min) / (abin + 1e-8)
            if scale is not None:
                e[:, 0] *= scale[0]
                e[:, 1] *= scale[1]
            e5 = torch.cat([e, torch.full((e.shape[0], 1), float(i), dtype=torch.float64)], 1)
            vol, state = oe.taf_bin_update(e5, grid, state, K)
        if n_bins == 0:
            vol = state.permute(3, 2, 0, 1).contiguous().view(2 * K, grid[0], grid[1])
        outs.append(vol.clone())
    return outs, state
