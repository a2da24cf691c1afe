"""Frame fingerprints and comparison helpers (test / tooling code).

A *fingerprint* is everything the parity contract covers for one frame, in a
form that does not depend on arena addresses or on the order in which tiles
were compacted (both differ run to run, even inside the reference):

  image     final stages[3].filled                       (exact)
  normals   final normals image                          (exact; +-1 LSB allowed vs CPU)
  filled[s] per-level filled images                      (exact)
  active[s] sorted positions of tiles still ambiguous after level s
  tapes[s]  for those tiles, hash + length of the *logical* tape (non-JUMP cells)

Works on any object with the accessor set of oracle._Base / capi.Context
(image(), normals(), filled(s), tiles(s), arena()).
"""
from __future__ import annotations

import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402  (test infrastructure)

CHAIN = {2: [0, 2], 3: [0, 1, 2]}   # interval levels live in these stages


def _arena(obj):
    if hasattr(obj, "arena"):
        return obj.arena()
    return obj.tape_data()


def fingerprint(obj, dim: int, with_tapes: bool = True) -> dict:
    fp = {"dim": dim, "size": int(obj.image().shape[0])}
    fp["image"] = np.array(obj.image(), dtype=np.int32, copy=True)
    if dim == 3:
        fp["normals"] = np.array(obj.normals(), dtype=np.uint32, copy=True)
    arena = np.ascontiguousarray(_arena(obj)) if with_tapes else None
    for s in CHAIN[dim]:
        f = np.asarray(obj.filled(s))
        if dim == 2 and s == 2:   # 2D keeps its 8-px level in the first (S/8)^2 entries of stage 2
            side = fp["size"] // 8
            f = f.reshape(-1)[: side * side].reshape(side, side)
        fp[f"filled{s}"] = np.array(f, dtype=np.int32, copy=True)
        t = obj.tiles(s)
        act = t[t["position"] != -1]
        order = np.argsort(act["position"], kind="stable")
        act = act[order]
        fp[f"active{s}"] = np.array(act["position"], dtype=np.int32, copy=True)
        if with_tapes:
            h, ln = oracle.tape_hashes(arena, np.array(act["tape"], dtype=np.int32))
            fp[f"tape_hash{s}"] = h
            fp[f"tape_len{s}"] = ln
    return fp


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


def summarize(fp: dict) -> dict:
    """Small JSON-able summary (hashes + counts) for committing as a golden record."""
    out = {"dim": fp["dim"], "size": fp["size"]}
    for k, v in fp.items():
        if isinstance(v, np.ndarray):
            out[k] = {"sha": digest(v), "n": int(v.size)}
            if k == "image":
                out[k]["nonzero"] = int((v != 0).sum())
                out[k]["sum"] = int(v.astype(np.int64).sum())
            if k.startswith("tape_len"):
                out[k]["total"] = int(v.astype(np.int64).sum())
    return out


def compare(a: dict, b: dict, normals_lsb: int = 0) -> dict:
    """Field-by-field comparison; returns {field: mismatch description} (empty = equal)."""
    bad = {}
    for k in a:
        if k in ("dim", "size"):
            if a[k] != b.get(k):
                bad[k] = f"{a[k]} vs {b.get(k)}"
            continue
        if k not in b:
            continue
        x, y = a[k], b[k]
        if x.shape != y.shape:
            bad[k] = f"shape {x.shape} vs {y.shape}"
            continue
        if k == "normals" and normals_lsb > 0:
            xb = x.view(np.uint8).astype(np.int16)
            yb = y.view(np.uint8).astype(np.int16)
            d = np.abs(xb - yb)
            n = int((d > normals_lsb).sum())
            if n:
                bad[k] = f"{n} bytes differ by more than {normals_lsb} LSB (max {int(d.max())})"
            continue
        n = int((x != y).sum())
        if n:
            bad[k] = f"{n} of {x.size} entries differ"
    return bad


def compare_summary(a: dict, b: dict) -> dict:
    bad = {}
    for k, v in a.items():
        if k not in b:
            continue
        if isinstance(v, dict):
            if v.get("sha") != b[k].get("sha"):
                bad[k] = f"{v} vs {b[k]}"
        elif v != b[k]:
            bad[k] = f"{v} vs {b[k]}"
    return bad


def load_tape(model: str) -> np.ndarray:
    return np.fromfile(ROOT / "tests" / "golden" / "tapes" / f"{model}.u64", dtype="<u8")
