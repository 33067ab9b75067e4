"""ctypes bindings of the two C ABIs (include/b200scan.h, include/blamm_host.h).

Nothing here computes scores: every call goes into libb200scan.so (CUDA) or libblammhost.so (C++ host
model).  Loading fails loudly when the libraries are missing -- there is no Python or CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .build import lib_dir

ENGINE_AUTO, ENGINE_GATHER, ENGINE_TENSOR = 0, 1, 2
LOWER_ZERO, LOWER_FOLD = 0, 1
NUM_SLOTS = 3

HIT_DTYPE = np.dtype([("pos", "<u8"), ("col", "<u4"), ("score", "<f4")])
HIT12_DTYPE = np.dtype([("pos", "<u4"), ("col", "<u4"), ("score", "<f4")])      # b200scan_hit12
HIT8_DTYPE = np.dtype([("key", "<u4"), ("score", "<f4")])                      # b200scan_hit8: key = (pos & 255) << 24 | col
HITS_16, HITS_12, HITS_8 = 16, 12, 8
BUCKET_SHIFT = 8


class Timing(ctypes.Structure):
    _fields_ = [("h2d_ms", ctypes.c_float), ("pack_ms", ctypes.c_float), ("score_ms", ctypes.c_float),
                ("rescore_ms", ctypes.c_float), ("d2h_ms", ctypes.c_float), ("n_candidates", ctypes.c_uint64),
                ("n_hits", ctypes.c_uint64), ("engine_used", ctypes.c_int32), ("kernel_launches", ctypes.c_int32),
                ("order_ms", ctypes.c_float), ("reserved", ctypes.c_float)]

    def as_dict(self) -> dict:
        return {k: getattr(self, k) for k, _ in self._fields_}


_scan = None
_host = None
_u64p = ctypes.POINTER(ctypes.c_uint64)


def scan_lib() -> ctypes.CDLL:
    global _scan
    if _scan is None:
        path = os.environ.get("B200SCAN_LIB") or os.path.join(lib_dir(), "libb200scan.so")     # override: diagnostic builds only
        if not os.path.exists(path):
            raise RuntimeError("libb200scan.so is missing (run `make` / __graft_entry__.build()); there is no fallback path")
        L = ctypes.CDLL(path)
        vp, i32, u64, f32p = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint64, ctypes.POINTER(ctypes.c_float)
        L.b200scan_abi_version.restype = ctypes.c_int
        L.b200scan_device_count.restype = ctypes.c_int
        L.b200scan_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int, u64, u64]
        L.b200scan_destroy.argtypes = [vp]; L.b200scan_destroy.restype = None
        L.b200scan_last_error.argtypes = [vp]; L.b200scan_last_error.restype = ctypes.c_char_p
        L.b200scan_set_engine.argtypes = [vp, ctypes.c_int]
        L.b200scan_set_tensor_accumulator.argtypes = [vp, ctypes.c_int]
        L.b200scan_tensor_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(ctypes.c_double)]
        L.b200scan_set_motifs.argtypes = [vp, vp, i32, i32, vp, vp]
        L.b200scan_host_alloc.argtypes = [u64]; L.b200scan_host_alloc.restype = vp
        L.b200scan_host_free.argtypes = [vp]; L.b200scan_host_free.restype = None
        L.b200scan_submit_ascii.argtypes = [vp, ctypes.c_int, vp, u64, u64, vp, u64, ctypes.c_int]
        L.b200scan_submit_packed.argtypes = [vp, ctypes.c_int, vp, vp, u64, u64, vp, u64]
        L.b200scan_collect.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp), _u64p, ctypes.POINTER(Timing)]
        L.b200scan_collect12.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp), _u64p, ctypes.POINTER(Timing)]
        L.b200scan_collect8.argtypes = [vp, ctypes.c_int, ctypes.POINTER(vp), _u64p, ctypes.POINTER(vp), _u64p, ctypes.POINTER(Timing)]
        L.b200scan_set_hit_format.argtypes = [vp, ctypes.c_int]
        L.b200scan_rerun_resident.argtypes = [vp, ctypes.c_int, ctypes.c_int, f32p, f32p, _u64p]
        L.b200scan_hist_begin.argtypes = [vp, vp, vp, ctypes.c_uint32]
        L.b200scan_hist_block_ascii.argtypes = [vp, vp, u64, u64, vp, u64, ctypes.c_int]
        L.b200scan_hist_read.argtypes = [vp, vp, u64]
        L.b200scan_flush_l2.argtypes = [vp]
        L.b200scan_tensor_work.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        L.b200scan_describe.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32), _u64p]
        _scan = L
    return _scan


def host_lib() -> ctypes.CDLL:
    global _host
    if _host is None:
        path = os.path.join(lib_dir(), "libblammhost.so")
        if not os.path.exists(path):
            raise RuntimeError("libblammhost.so is missing (run `make` / __graft_entry__.build())")
        L = ctypes.CDLL(path)
        vp, u64 = ctypes.c_void_p, ctypes.c_uint64
        L.blamm_host_last_error.restype = ctypes.c_char_p
        L.blamm_motifs_load.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp)]
        L.blamm_motifs_free.argtypes = [vp]; L.blamm_motifs_free.restype = None
        L.blamm_motifs_count.argtypes = [vp]; L.blamm_motifs_max_len.argtypes = [vp]
        L.blamm_motifs_generate_matrix.argtypes = [vp, _u64p, ctypes.c_float]
        L.blamm_motifs_get_matrix.argtypes = [vp, vp, u64]
        L.blamm_motifs_get_columns.argtypes = [vp, vp, vp, vp, vp]
        L.blamm_motifs_name.argtypes = [vp, ctypes.c_int]; L.blamm_motifs_name.restype = ctypes.c_char_p
        L.blamm_motifs_set_thresholds.argtypes = [vp, ctypes.c_int, ctypes.c_float, ctypes.c_char_p, ctypes.c_char_p, vp]
        L.blamm_motifs_write_histograms.argtypes = [vp, _u64p, ctypes.c_float, u64, u64, ctypes.c_char_p, ctypes.c_char_p]
        L.blamm_fasta_open.argtypes = [ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, u64, ctypes.POINTER(vp)]
        L.blamm_fasta_close.argtypes = [vp]; L.blamm_fasta_close.restype = None
        L.blamm_fasta_set_parallel.argtypes = [vp, ctypes.c_uint, u64]
        L.blamm_fasta_next.argtypes = [vp, u64, u64, ctypes.POINTER(vp), _u64p, _u64p, _u64p, ctypes.POINTER(vp),
                                       ctypes.POINTER(vp), ctypes.POINTER(vp), _u64p]
        L.blamm_fasta_num_sequences.argtypes = [vp]
        L.blamm_fasta_sequence_name.argtypes = [vp, ctypes.c_int]; L.blamm_fasta_sequence_name.restype = ctypes.c_char_p
        L.blamm_fasta_counts.argtypes = [vp, _u64p]
        L.blamm_format_score.argtypes = [ctypes.c_float, ctypes.c_char_p]
        L.blamm_pack_ascii.argtypes = [vp, u64, ctypes.c_int, vp, vp]
        L.blamm_fasta_pack.argtypes = [vp, ctypes.c_int, vp, vp]
        _host = L
    return _host


class ScanError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("b200scan error %d: %s" % (code, msg))
        self.code = code


class Scanner:
    """One b200scan context (one CUDA device)."""

    def __init__(self, device: int = 0, max_block_nt: int = 1 << 26, max_hits: int = 1 << 22):
        self._L = scan_lib()
        self._ctx = ctypes.c_void_p()
        rc = self._L.b200scan_create(ctypes.byref(self._ctx), device, max_block_nt, max_hits)
        if rc != 0:
            raise ScanError(rc, (self._L.b200scan_last_error(None) or b"").decode())
        self._keep = {}
        self._hit_format = HITS_16
        self._slot_format = {}

    def close(self) -> None:
        if self._ctx:
            self._L.b200scan_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc: int) -> None:
        if rc != 0:
            raise ScanError(rc, (self._L.b200scan_last_error(self._ctx) or b"").decode())

    def set_engine(self, engine: int) -> None:
        self._chk(self._L.b200scan_set_engine(self._ctx, engine))

    def set_tensor_accumulator(self, bits: int) -> None:
        self._chk(self._L.b200scan_set_tensor_accumulator(self._ctx, bits))

    def set_hit_format(self, fmt: int) -> None:
        """HITS_16 (b200scan_hit, default), HITS_12 (b200scan_hit12) or HITS_8 (b200scan_hit8, ordered on the device) for
        blocks submitted from now on."""
        self._chk(self._L.b200scan_set_hit_format(self._ctx, fmt))
        self._hit_format = fmt

    def tensor_info(self) -> dict:
        b, m = ctypes.c_int32(), ctypes.c_double()
        self._chk(self._L.b200scan_tensor_info(self._ctx, ctypes.byref(b), ctypes.byref(m)))
        return dict(accumulator_bits=b.value, mean_margin=m.value)

    def set_motifs(self, P: np.ndarray, col_len: np.ndarray, thr: np.ndarray) -> None:
        """P: (n_cols, ldp) C-contiguous float32 == column-major ldp x n_cols."""
        P = np.ascontiguousarray(P, dtype=np.float32)
        col_len = np.ascontiguousarray(col_len, dtype=np.int32)
        thr = np.ascontiguousarray(thr, dtype=np.float32)
        assert P.ndim == 2 and P.shape[0] == len(col_len) == len(thr)
        self._chk(self._L.b200scan_set_motifs(self._ctx, P.ctypes.data, P.shape[1], P.shape[0], col_len.ctypes.data, thr.ctypes.data))

    def submit_ascii(self, slot: int, block, n_total: Optional[int] = None, n_payload: Optional[int] = None,
                     frag_starts: Optional[np.ndarray] = None, lower: int = LOWER_ZERO) -> None:
        """block: bytes, numpy uint8 array, or an int address of (pinned) host memory (then n_total is required)."""
        if isinstance(block, int):
            ptr = block
        elif isinstance(block, (bytes, bytearray)):
            buf = np.frombuffer(block, dtype=np.uint8)
            ptr, n_total = buf.ctypes.data, (len(buf) if n_total is None else n_total)
            self._keep[slot] = (buf, block)
        else:
            buf = np.ascontiguousarray(block, dtype=np.uint8)
            ptr, n_total = buf.ctypes.data, (len(buf) if n_total is None else n_total)
            self._keep[slot] = buf
        n_payload = n_total if n_payload is None else n_payload
        fs = np.ascontiguousarray(frag_starts if frag_starts is not None else [], dtype=np.uint64)
        self._chk(self._L.b200scan_submit_ascii(self._ctx, slot, ptr, n_total, n_payload, fs.ctypes.data if len(fs) else None,
                                                len(fs), lower))
        self._slot_format[slot] = self._hit_format

    def submit_packed(self, slot: int, codes2: np.ndarray, zero_mask: Optional[np.ndarray], n_total: int,
                      n_payload: Optional[int] = None, frag_starts: Optional[np.ndarray] = None) -> None:
        codes2 = np.ascontiguousarray(codes2, dtype=np.uint32)
        zm = None if zero_mask is None else np.ascontiguousarray(zero_mask, dtype=np.uint32)
        n_payload = n_total if n_payload is None else n_payload
        fs = np.ascontiguousarray(frag_starts if frag_starts is not None else [], dtype=np.uint64)
        self._chk(self._L.b200scan_submit_packed(self._ctx, slot, codes2.ctypes.data, None if zm is None else zm.ctypes.data,
                                                 n_total, n_payload, fs.ctypes.data if len(fs) else None, len(fs)))
        self._slot_format[slot] = self._hit_format

    def submit_packed_ptr(self, slot: int, codes_ptr: int, zmask_ptr: Optional[int], n_total: int, n_payload: int,
                          frag_starts: Optional[np.ndarray] = None) -> None:
        """b200scan_submit_packed on raw addresses (page-locked buffers of b200scan_host_alloc: they must stay unchanged until
        the block has been collected)."""
        fs = np.ascontiguousarray(frag_starts if frag_starts is not None else [], dtype=np.uint64)
        self._chk(self._L.b200scan_submit_packed(self._ctx, slot, codes_ptr, zmask_ptr, n_total, n_payload,
                                                 fs.ctypes.data if len(fs) else None, len(fs)))
        self._slot_format[slot] = self._hit_format

    def collect(self, slot: int, copy: bool = True, fmt: Optional[int] = None) -> Tuple[np.ndarray, dict]:
        """Hits of the block on `slot`, as HIT_DTYPE or HIT12_DTYPE records -- the format the block was submitted under
        (`fmt` overrides the choice of the collecting function: only the state tests do that)."""
        hp, n, t = ctypes.c_void_p(), ctypes.c_uint64(), Timing()
        fmt = self._slot_format.get(slot, self._hit_format) if fmt is None else fmt
        if fmt == HITS_8:
            h8, bstart, tm = self.collect8(slot, copy=copy)
            return expand_hits8(h8, bstart), tm
        fn, dt = (self._L.b200scan_collect12, HIT12_DTYPE) if fmt == HITS_12 else (self._L.b200scan_collect, HIT_DTYPE)
        self._chk(fn(self._ctx, slot, ctypes.byref(hp), ctypes.byref(n), ctypes.byref(t)))
        self._keep.pop(slot, None)
        if n.value == 0:
            return np.zeros(0, dtype=dt), t.as_dict()
        raw = (ctypes.c_uint8 * (n.value * dt.itemsize)).from_address(hp.value)
        hits = np.frombuffer(raw, dtype=dt)
        return (hits.copy() if copy else hits), t.as_dict()

    def collect8(self, slot: int, copy: bool = True) -> Tuple[np.ndarray, np.ndarray, dict]:
        """b200scan_collect8: the block's hits as HIT8_DTYPE records in (position, column) order + bucket_start (one entry
        per 256 window positions, plus the total)."""
        hp, bp, n, nb, t = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint64(), ctypes.c_uint64(), Timing()
        self._chk(self._L.b200scan_collect8(self._ctx, slot, ctypes.byref(hp), ctypes.byref(n), ctypes.byref(bp), ctypes.byref(nb),
                                            ctypes.byref(t)))
        self._keep.pop(slot, None)
        hits = (np.frombuffer((ctypes.c_uint8 * (n.value * 8)).from_address(hp.value), dtype=HIT8_DTYPE) if n.value
                else np.zeros(0, dtype=HIT8_DTYPE))
        bstart = (np.frombuffer((ctypes.c_uint32 * (nb.value + 1)).from_address(bp.value), dtype=np.uint32) if bp.value
                  else np.zeros(1, dtype=np.uint32))
        return (hits.copy() if copy else hits), (bstart.copy() if copy else bstart), t.as_dict()

    def scan(self, block, frag_starts=None, n_payload=None, lower: int = LOWER_ZERO, slot: int = 0) -> Tuple[np.ndarray, dict]:
        self.submit_ascii(slot, block, n_payload=n_payload, frag_starts=frag_starts, lower=lower)
        return self.collect(slot)

    def rerun_resident(self, slot: int, iters: int) -> Tuple[float, float, int]:
        tot, sc, nh = ctypes.c_float(), ctypes.c_float(), ctypes.c_uint64()
        self._chk(self._L.b200scan_rerun_resident(self._ctx, slot, iters, ctypes.byref(tot), ctypes.byref(sc), ctypes.byref(nh)))
        return tot.value, sc.value, nh.value

    def hist_begin(self, col_min: np.ndarray, col_max: np.ndarray, num_bins: int) -> None:
        mn = np.ascontiguousarray(col_min, dtype=np.float32); mx = np.ascontiguousarray(col_max, dtype=np.float32)
        self._hist_shape = (len(mn), num_bins)
        self._chk(self._L.b200scan_hist_begin(self._ctx, mn.ctypes.data, mx.ctypes.data, num_bins))

    def hist_block(self, block, frag_starts=None, n_payload: Optional[int] = None, lower: int = LOWER_ZERO) -> None:
        buf = np.frombuffer(block, dtype=np.uint8) if isinstance(block, (bytes, bytearray)) else np.ascontiguousarray(block, dtype=np.uint8)
        n_payload = len(buf) if n_payload is None else n_payload
        fs = np.ascontiguousarray(frag_starts if frag_starts is not None else [], dtype=np.uint64)
        self._chk(self._L.b200scan_hist_block_ascii(self._ctx, buf.ctypes.data if len(buf) else None, len(buf), n_payload,
                                                    fs.ctypes.data if len(fs) else None, len(fs), lower))

    def hist_read(self) -> np.ndarray:
        out = np.zeros(self._hist_shape, dtype=np.uint64)
        self._chk(self._L.b200scan_hist_read(self._ctx, out.ctypes.data, out.size))
        return out

    def tensor_work(self) -> dict:
        """Tensor-core operations per window: as issued (padded tiles) and algorithmic (8 x sum L)."""
        a, b = ctypes.c_double(), ctypes.c_double()
        self._chk(self._L.b200scan_tensor_work(self._ctx, ctypes.byref(a), ctypes.byref(b)))
        return dict(mma_ops_per_window=a.value, algorithmic_ops_per_window=b.value)

    def flush_l2(self) -> None:
        self._chk(self._L.b200scan_flush_l2(self._ctx))

    def describe(self) -> dict:
        a, b, c, d, e = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_uint64()
        self._chk(self._L.b200scan_describe(self._ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d), ctypes.byref(e)))
        return dict(n_cols=a.value, max_len=b.value, n_tiles=c.value, sm_count=d.value, sum_len=e.value)


def expand_hits8(hits8: np.ndarray, bucket_start: np.ndarray) -> np.ndarray:
    """Ordered b200scan_hit8 records + bucket index -> HIT12_DTYPE records with full block positions (same order)."""
    out = np.zeros(len(hits8), dtype=HIT12_DTYPE)
    if len(hits8):
        per_bucket = np.diff(bucket_start.astype(np.int64))
        bucket = np.repeat(np.arange(len(per_bucket), dtype=np.uint32), per_bucket)
        out["pos"] = (bucket << np.uint32(BUCKET_SHIFT)) + (hits8["key"] >> np.uint32(24))
        out["col"] = hits8["key"] & np.uint32(0xFFFFFF)
        out["score"] = hits8["score"]
    return out


class HostError(RuntimeError):
    pass


class MotifSet:
    """MotifContainer mirror: load (+ reverse complements), matrix P for a background, thresholds."""

    def __init__(self, path: str, revcompl: bool, load_permutations: bool = True):
        self._L = host_lib()
        self._h = ctypes.c_void_p()
        if self._L.blamm_motifs_load(path.encode(), int(load_permutations), int(revcompl), ctypes.byref(self._h)) != 0:
            raise HostError(self._L.blamm_host_last_error().decode())
        self.n_cols = self._L.blamm_motifs_count(self._h)
        self.max_len = self._L.blamm_motifs_max_len(self._h)
        self.names = [self._L.blamm_motifs_name(self._h, c).decode() for c in range(self.n_cols)]

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.blamm_motifs_free(self._h)
            self._h = None

    def _chk(self, rc: int) -> None:
        if rc != 0:
            raise HostError(self._L.blamm_host_last_error().decode())

    def generate_matrix(self, bg: Sequence[int], pseudo: float = 0.25) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        bga = np.array(bg, dtype=np.uint64)
        self._chk(self._L.blamm_motifs_generate_matrix(self._h, bga.ctypes.data_as(_u64p), pseudo))
        P = np.zeros((self.n_cols, 4 * self.max_len), dtype=np.float32)
        self._chk(self._L.blamm_motifs_get_matrix(self._h, P.ctypes.data, P.size))
        col_len = np.zeros(self.n_cols, dtype=np.int32); rc = np.zeros(self.n_cols, dtype=np.uint8)
        self._chk(self._L.blamm_motifs_get_columns(self._h, col_len.ctypes.data, rc.ctypes.data, None, None))
        return P, col_len, rc.astype(bool)

    def min_max(self) -> Tuple[np.ndarray, np.ndarray]:
        mn = np.zeros(self.n_cols, dtype=np.float32); mx = np.zeros(self.n_cols, dtype=np.float32)
        self._chk(self._L.blamm_motifs_get_columns(self._h, None, None, mn.ctypes.data, mx.ctypes.data))
        return mn, mx

    def thresholds(self, mode: str, value: float, species: str = "", histdir: str = "") -> np.ndarray:
        thr = np.zeros(self.n_cols, dtype=np.float32)
        self._chk(self._L.blamm_motifs_set_thresholds(self._h, {"at": 0, "rt": 1, "pt": 2}[mode], value, species.encode(),
                                                      histdir.encode(), thr.ctypes.data))
        return thr

    def write_histograms(self, bg: Sequence[int], species: str, histdir: str, pseudo: float = 0.25, num_bins: int = 250,
                         max_length: int = 10_000_000) -> None:
        bga = np.array(bg, dtype=np.uint64)
        self._chk(self._L.blamm_motifs_write_histograms(self._h, bga.ctypes.data_as(_u64p), pseudo, num_bins, max_length,
                                                        species.encode(), histdir.encode()))


def pack_ascii(block, lower: int = LOWER_ZERO) -> Tuple[np.ndarray, np.ndarray, bool]:
    """Host 2-bit packer (blamm_pack_ascii): (codes2, zero_mask, has_zero) of a block, the inputs of Scanner.submit_packed."""
    L = host_lib()
    buf = np.frombuffer(block, dtype=np.uint8) if isinstance(block, (bytes, bytearray)) else np.ascontiguousarray(block, dtype=np.uint8)
    n = len(buf)
    codes = np.zeros((n + 15) // 16, dtype=np.uint32); zm = np.zeros((n + 31) // 32, dtype=np.uint32)
    rc = L.blamm_pack_ascii(buf.ctypes.data if n else None, n, int(lower == LOWER_FOLD), codes.ctypes.data, zm.ctypes.data)
    if rc < 0:
        raise HostError(L.blamm_host_last_error().decode())
    return codes, zm, bool(rc)


class FastaStream:
    """FastaBatch mirror: filtered stream of a group's FASTA files, chunk by chunk."""

    def __init__(self, files: Sequence[str], max_filtered: int = 2 ** 63, threads: int = 1, segment_bytes: int = 0):
        self._L = host_lib()
        self._h = ctypes.c_void_p()
        arr = (ctypes.c_char_p * len(files))(*[f.encode() for f in files])
        if self._L.blamm_fasta_open(arr, len(files), max_filtered, ctypes.byref(self._h)) != 0:
            raise HostError(self._L.blamm_host_last_error().decode())
        if (threads != 1 or segment_bytes) and self._L.blamm_fasta_set_parallel(self._h, threads, segment_bytes) != 0:
            raise HostError(self._L.blamm_host_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.blamm_fasta_close(self._h)
            self._h = None

    def next(self, payload: int, halo: int) -> Optional[dict]:
        chars, fs, fq, fp = (ctypes.c_void_p() for _ in range(4))
        nt, npay, st, nf = (ctypes.c_uint64() for _ in range(4))
        rc = self._L.blamm_fasta_next(self._h, payload, halo, ctypes.byref(chars), ctypes.byref(nt), ctypes.byref(npay),
                                      ctypes.byref(st), ctypes.byref(fs), ctypes.byref(fq), ctypes.byref(fp), ctypes.byref(nf))
        if rc < 0:
            raise HostError(self._L.blamm_host_last_error().decode())
        if rc == 0:
            return None

        def arr(p, n):
            return np.frombuffer((ctypes.c_uint64 * n).from_address(p.value), dtype=np.uint64).copy() if n else np.zeros(0, np.uint64)

        return dict(chars=ctypes.string_at(chars.value, nt.value), n_total=nt.value, n_payload=npay.value, stream_start=st.value,
                    frag_start=arr(fs, nf.value), frag_seq=arr(fq, nf.value), frag_pos=arr(fp, nf.value))

    def pack(self, n_total: int, lower: int = LOWER_ZERO) -> Tuple[np.ndarray, np.ndarray, bool]:
        """2-bit codes, zero mask and has_zero of the chunk the last next() returned (packed on the parser threads)."""
        codes = np.zeros((n_total + 15) // 16, dtype=np.uint32); zm = np.zeros((n_total + 31) // 32, dtype=np.uint32)
        rc = self._L.blamm_fasta_pack(self._h, int(lower == LOWER_FOLD), codes.ctypes.data, zm.ctypes.data)
        if rc < 0:
            raise HostError(self._L.blamm_host_last_error().decode())
        return codes, zm, bool(rc)

    def seq_names(self) -> List[str]:
        return [self._L.blamm_fasta_sequence_name(self._h, i).decode() for i in range(self._L.blamm_fasta_num_sequences(self._h))]

    def counts(self) -> List[int]:
        c = (ctypes.c_uint64 * 4)()
        self._L.blamm_fasta_counts(self._h, c)
        return list(c)
