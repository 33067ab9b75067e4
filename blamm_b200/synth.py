"""Seeded synthetic inputs for tests and bench.py.

Real JASPAR CORE is not available offline, so "full JASPAR" is a JASPAR-LIKE set: `n` motifs whose lengths
follow clip(round(Gamma(9, 1.4)), 5, 35) (mean ~12.5, like CORE vertebrates), integer count matrices drawn
from Dirichlet-multinomials.  Every report that uses it says so.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def jaspar_like_lengths(n: int, rng: np.random.Generator, lo: int = 5, hi: int = 35) -> np.ndarray:
    L = np.clip(np.rint(rng.gamma(9.0, 1.4, size=n)), lo, hi).astype(int)
    if n >= 3:
        L[0], L[1], L[2] = hi, min(30, hi), lo
    return L


def random_pfms(lengths: Sequence[int], rng: np.random.Generator) -> List[np.ndarray]:
    out = []
    for L in lengths:
        N = int(rng.integers(20, 2000))
        alphas = rng.choice([0.1, 0.3, 1.0, 3.0], size=L)
        pfm = np.zeros((L, 4), dtype=np.int64)
        for j in range(L):
            p = rng.dirichlet(np.full(4, alphas[j]))
            pfm[j] = rng.multinomial(N, p)
        out.append(pfm)
    return out


def write_jaspar(path: str, pfms: Sequence[np.ndarray], prefix: str = "SY") -> List[str]:
    names = []
    with open(path, "w") as f:
        for i, pfm in enumerate(pfms):
            name = "%s%04d.1" % (prefix, i)
            names.append(name)
            f.write(">%s\tsynthetic%d\n" % (name, i))
            for b, letter in enumerate("ACGT"):
                f.write("%s  [%s ]\n" % (letter, "".join(" %6d" % int(v) for v in pfm[:, b])))
    return names


def make_jaspar_like(path: str, n: int, seed: int, uniform_len: Tuple[int, int] = None) -> List[np.ndarray]:
    rng = np.random.default_rng(seed)
    lengths = rng.integers(uniform_len[0], uniform_len[1] + 1, size=n) if uniform_len else jaspar_like_lengths(n, rng)
    pfms = random_pfms(lengths, rng)
    write_jaspar(path, pfms)
    return pfms


def random_acgt(n: int, seed: int, probs: Sequence[float] = (0.25, 0.25, 0.25, 0.25)) -> np.ndarray:
    """n upper-case ACGT characters as uint8; probabilities are quantised to 1/256 (one random byte per base)."""
    rng = np.random.default_rng(seed)
    edges = np.rint(np.cumsum(probs) * 256).astype(int)
    lut = np.empty(256, dtype=np.uint8)
    lo = 0
    for k in range(4):
        lut[lo:edges[k] if k < 3 else 256] = ACGT[k]
        lo = edges[k]
    out = np.empty(n, dtype=np.uint8)
    step = 1 << 24
    for s in range(0, n, step):
        e = min(n, s + step)
        out[s:e] = lut[rng.integers(0, 256, size=e - s, dtype=np.uint8)]
    return out


def write_fasta(path: str, records: Sequence[Tuple[str, np.ndarray]], width: int = 60) -> None:
    with open(path, "wb") as f:
        for name, seq in records:
            f.write(b">" + name.encode() + b"\n")
            n = len(seq)
            full = (n // width) * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = np.asarray(seq[:full], dtype=np.uint8).reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(np.asarray(seq[full:], dtype=np.uint8).tobytes() + b"\n")


def counts_of(seq: np.ndarray) -> List[int]:
    """Nucleotide counts the way `blamm dict` takes them (case-insensitive ACGT)."""
    h = np.bincount(seq, minlength=256)
    return [int(h[ord(c)] + h[ord(c.lower())]) for c in "ACGT"]
