"""``.npy`` label / event files -- same API as ``src/io/npy_events_tools.py``."""
from __future__ import annotations

import numpy as np


def stream_td_data(file_handle, buffer, dtype, ev_count=-1):
    dat = np.fromfile(file_handle, dtype=dtype, count=ev_count)
    for name, _ in dtype:
        buffer[name][:len(dat)] = dat[name]


def parse_header(fhandle):
    """Return ``(data offset, field list, record size, (None, None), numpy dtype)``.

    As in the reference (:30-61) the field list is PACKED -- rebuilt from (name, format)
    pairs -- with ``ts`` renamed ``t`` and ``confidence`` renamed ``class_confidence``;
    callers read the payload with ``np.fromfile(f, dtype=field_list)``."""
    fmt = np.lib.format
    version = fmt.read_magic(fhandle)
    if tuple(version) == (1, 0):
        shape, fortran, dtype = fmt.read_array_header_1_0(fhandle)
    else:
        shape, fortran, dtype = fmt.read_array_header_2_0(fhandle)
    assert not fortran, "Fortran order arrays not supported"
    ev_size = dtype.itemsize
    assert ev_size != 0
    start = fhandle.tell()
    rename = {"ts": "t", "confidence": "class_confidence"}
    ev_type = [(rename.get(name, name), str(dtype.fields[name][0])) for name in dtype.names]
    return start, ev_type, ev_size, (None, None), dtype


def read_label_times(path) -> np.ndarray:
    """Sorted unique label timestamps of a ``*_bbox.npy`` file
    (``generate_taf.py:146-151``)."""
    with open(path, "rb") as fh:
        _start, v_type, _size, _hw, _dtype = parse_header(fh)
        boxes = np.fromfile(fh, dtype=v_type, count=-1)
    return np.unique(boxes["t"])
