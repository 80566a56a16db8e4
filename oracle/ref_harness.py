"""Loader for the REAL reference functions (build container only).

TEST INFRASTRUCTURE.  ``/root/reference`` is mounted read-only in the build
container and is absent on the GPU box, so this module is used only by
``oracle/make_golden.py`` (to produce ``tests/golden/``) and by the
``requires_reference`` tests that validate the in-repo restatement against the
unmodified reference.  Nothing here is copied from the reference: the functions
are extracted from its sources at run time with ``ast`` and executed as-is.

Shims:
  * the encoder scripts import modules that are absent here (sqlalchemy, h5py,
    tkinter, sklearn...) at module level, so only the ``FunctionDef`` nodes are
    compiled, in a namespace holding ``torch, np, time, math``;
  * ``torch.cuda.synchronize`` is a no-op and ``.cuda()`` is identity on CPU;
  * numpy-2: ``parse_header`` returns ``np.uint8`` sizes that overflow in
    ``psee_loader.py:49`` and ``np.lib.format._read_array_header`` is gone.
"""
from __future__ import annotations

import ast
import contextlib
import importlib
import math
import os
import runpy
import sys
import time
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("EVREP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "generate_taf.py"))


@contextlib.contextmanager
def _cpu_cuda_shims():
    """Make the reference's ``.cuda()`` / ``synchronize`` calls CPU no-ops."""
    saved = (torch.Tensor.cuda, torch.cuda.synchronize, torch.cuda.empty_cache)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.empty_cache = lambda *a, **k: None
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.cuda.synchronize, torch.cuda.empty_cache = saved


def load_functions(script: str) -> dict:
    """Compile every top-level ``def`` of ``<reference>/<script>`` unmodified."""
    path = os.path.join(REFERENCE_ROOT, script)
    with open(path, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    tree.body = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
    ns = {"torch": torch, "np": np, "time": time, "math": math}
    exec(compile(tree, path, "exec"), ns)  # noqa: S102 - executing the reference is the point
    fns = {k: v for k, v in ns.items() if isinstance(v, types.FunctionType)}

    def wrap(fn):
        def call(*a, **k):
            with _cpu_cuda_shims():
                return fn(*a, **k)
        call.__name__ = fn.__name__
        return call

    return {k: wrap(v) for k, v in fns.items()}


def load_sparse_ops():
    """``data/sparse_ops.py`` imports cleanly (torch + numpy only)."""
    spec = importlib.util.spec_from_file_location(
        "_ref_sparse_ops", os.path.join(REFERENCE_ROOT, "data", "sparse_ops.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _numpy2_shims():
    fmt = np.lib.format
    if not hasattr(fmt, "_read_array_header"):
        def _read_array_header(fp, version, max_header_size=10000):
            if tuple(version) == (1, 0):
                return fmt.read_array_header_1_0(fp)
            return fmt.read_array_header_2_0(fp)
        fmt._read_array_header = _read_array_header


def load_io():
    """Import the reference's ``src.io`` package (dat/npy tools + PSEELoader)."""
    _numpy2_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    dat = importlib.import_module("src.io.dat_events_tools")
    npy = importlib.import_module("src.io.npy_events_tools")
    psee = importlib.import_module("src.io.psee_loader")
    if not getattr(dat, "_evrep_patched", False):
        orig = dat.parse_header

        def parse_header(f):
            bod, ev_type, ev_size, size = orig(f)
            return bod, int(ev_type), int(ev_size), size
        dat.parse_header = parse_header
        dat._evrep_patched = True
    return dat, npy, psee


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return None


def run_script(script: str, argv: list) -> None:
    """Run a reference ``generate_*.py`` end to end on CPU (its ``__main__``)."""
    load_io()
    stubs = {}
    for name in ("sqlalchemy", "h5py", "tkinter", "sklearn", "sklearn.datasets"):
        try:
            importlib.import_module(name)
        except Exception:
            stubs[name] = _Stub(name)
    if "sklearn" in stubs:
        stubs["sklearn"].datasets = stubs.get("sklearn.datasets")

    class _Quiet:
        def __init__(self, *a, **k): pass
        def update(self, *a, **k): pass
        def close(self): pass

    tqdm_mod = importlib.import_module("tqdm")
    saved_tqdm = tqdm_mod.tqdm
    saved_argv = sys.argv
    saved_cwd = os.getcwd()
    sys.modules.update(stubs)
    tqdm_mod.tqdm = _Quiet
    sys.argv = [script] + list(argv)
    try:
        with _cpu_cuda_shims(), open(os.devnull, "w") as devnull, contextlib.redirect_stdout(devnull):
            try:
                runpy.run_path(os.path.join(REFERENCE_ROOT, script), run_name="__main__")
            except (ZeroDivisionError, NameError):
                # The scripts' trailing "Average Representation time" print divides by a
                # count that is zero / reads names that are unset when no `test` split
                # exists; the files have already been written at that point.
                pass
    finally:
        sys.argv = saved_argv
        tqdm_mod.tqdm = saved_tqdm
        for name in stubs:
            sys.modules.pop(name, None)
        os.chdir(saved_cwd)
