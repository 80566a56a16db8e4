"""Test helpers."""
import numpy as np
import torch

from oracle import drivers as od


def staged(t, x, y, p, lo=0, hi=None):
    hi = len(t) if hi is None else hi
    return torch.from_numpy(np.stack([x[lo:hi], y[lo:hi], t[lo:hi], p[lo:hi]], 1).astype(np.float64))


def oracle_taf_windows(t, x, y, p, windows, abin, grid, K, scale=None, state=None):
    """Oracle execution of a TAF window list on in-memory events."""
    outs, state = od.taf_windows_in_memory(staged(t, x, y, p), windows, abin, grid, K, scale, state)
    return [o.clone() for o in outs], state
