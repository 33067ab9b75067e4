"""TEST INFRASTRUCTURE ONLY -- reader for the binary dump written by oracle/refdump.cpp (format in its header)."""
import struct
from typing import List

import numpy as np


def read_refdump(path: str) -> List[dict]:
    b = open(path, "rb").read()
    assert b[:4] == b"RDMP"
    o = 4
    (nsp,) = struct.unpack_from("<I", b, o); o += 4
    out = []
    for _ in range(nsp):
        (nl,) = struct.unpack_from("<I", b, o); o += 4
        name = b[o:o + nl].decode(); o += nl
        ncols, ldp = struct.unpack_from("<II", b, o); o += 8
        cols = []
        for _ in range(ncols):
            ln, rc, thr, ml = struct.unpack_from("<IIfI", b, o); o += 16
            cols.append(dict(len=ln, rc=bool(rc), thr=np.float32(thr), name=b[o:o + ml].decode())); o += ml
        P = np.frombuffer(b, dtype="<f4", count=ldp * ncols, offset=o).reshape(ncols, ldp).copy(); o += 4 * ldp * ncols
        (nh,) = struct.unpack_from("<Q", b, o); o += 8
        hits = np.frombuffer(b, dtype=np.dtype([("seq", "<u4"), ("pos", "<u8"), ("col", "<u4"), ("score", "<f4"), ("naive", "<f4")]),
                             count=nh, offset=o).copy(); o += 24 * nh
        out.append(dict(name=name, cols=cols, P=P, hits=hits))
    assert o == len(b)
    return out
