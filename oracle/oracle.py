"""ctypes face of oracle/flat_oracle.c plus a small numpy twin.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under c99_vectordb_b200/ does.  PARITY UNPINNED for the search
arithmetic (no faiss in this image, no golden vectors in the reference) — see flat_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "flat_oracle.c"
LIB = HERE / "_build" / "liboracle.so"

METRIC_IP, METRIC_L2 = 0, 1
ORDER_SIMD, ORDER_DEVICE, ORDER_DEVICE16, ORDER_DEVICE8 = 0, 1, 2, 3


def device_lanes(d: int, chunk: int) -> int:
    """Lanes sharing a row in the kernels' arithmetic for rows of d elements stored in chunks of `chunk`
    (4 = fp32 rows, 8 = bf16 rows) — the rule of csrc/cabi.cu: pick_lpr restated."""
    nvec = -(-d // chunk)
    if nvec <= 16 and nvec % 8 == 0:
        return 8
    if nvec <= 48 and nvec % 16 == 0:
        return 16
    if nvec % 32 == 0 or nvec > 64:
        return 32
    best, best_slots = 32, -(-nvec // 32) * 32
    for lanes in (16, 8):
        slots = -(-nvec // lanes) * lanes
        if slots <= best_slots:
            best, best_slots = lanes, slots
    return best


def device_order(d: int, chunk: int) -> int:
    """The order code of that arithmetic."""
    return {32: ORDER_DEVICE, 16: ORDER_DEVICE16, 8: ORDER_DEVICE8}[device_lanes(d, chunk)]

_lib = None


def build(force: bool = False) -> Path:
    if not force and LIB.exists() and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    subprocess.run(
        ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared", "-std=c11",
         "-ffp-contract=off", "-o", str(LIB), str(SRC), "-lm"],
        check=True,
    )
    return LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        fp, ip, dp = C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_double)
        L.oracle_synth_rows.argtypes = [fp, C.c_int64, C.c_int, C.c_uint64, C.c_int64]
        L.oracle_synth_rows.restype = None
        L.oracle_round_bf16.argtypes = [fp, C.c_int64]
        L.oracle_round_bf16.restype = None
        L.oracle_scores.argtypes = [C.c_int, C.c_int, C.c_int, fp, C.c_int64, C.c_int, fp, fp]
        L.oracle_scores.restype = None
        L.oracle_scores_f64.argtypes = [C.c_int, fp, C.c_int64, C.c_int, fp, dp]
        L.oracle_scores_f64.restype = None
        for name in ("oracle_search", "oracle_search_rowpar"):
            f = getattr(L, name)
            f.argtypes = [C.c_int, C.c_int, C.c_int, fp, C.c_int64, C.c_int, ip, fp, C.c_int64, C.c_int64, fp, ip]
            f.restype = C.c_int
        L.oracle_normalize_rows.argtypes = [fp, C.c_int64, C.c_int, C.c_int]
        L.oracle_normalize_rows.restype = None
        L.oracle_merge_topk.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, fp, ip, fp, ip]
        L.oracle_merge_topk.restype = None
        L.oracle_max_threads.restype = C.c_int
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_set_threads.restype = None
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _pf(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _pi(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def set_threads(n: int) -> None:
    """Thread count of the oracle's OpenMP loops (overrides an inherited OMP_NUM_THREADS)."""
    lib().oracle_set_threads(int(n))


def synth_rows(n: int, d: int, seed: int, first_row: int = 0) -> np.ndarray:
    out = np.empty((n, d), dtype=np.float32)
    lib().oracle_synth_rows(_pf(out), n, d, seed, first_row)
    return out


def round_bf16(x: np.ndarray) -> np.ndarray:
    y = _f32(x).copy()
    lib().oracle_round_bf16(_pf(y), y.size)
    return y


def normalize_rows(x: np.ndarray, order: int = ORDER_SIMD) -> np.ndarray:
    """memo_cli.py:131-135 applied to every row."""
    y = np.atleast_2d(_f32(x)).copy()
    lib().oracle_normalize_rows(_pf(y), y.shape[0], y.shape[1], order)
    return y.reshape(np.shape(x))


def scores(metric: int, db: np.ndarray, q: np.ndarray, order: int = ORDER_SIMD, chunk: int = 4) -> np.ndarray:
    db, q = _f32(db), _f32(q).reshape(-1)
    out = np.empty(db.shape[0], dtype=np.float32)
    if order == ORDER_DEVICE:
        order = device_order(db.shape[1], chunk)
    lib().oracle_scores(metric, order, chunk, _pf(db), db.shape[0], db.shape[1], _pf(q), _pf(out))
    return out


def scores_f64(metric: int, db: np.ndarray, q: np.ndarray) -> np.ndarray:
    db, q = _f32(db), _f32(q).reshape(-1)
    out = np.empty(db.shape[0], dtype=np.float64)
    lib().oracle_scores_f64(metric, _pf(db), db.shape[0], db.shape[1], _pf(q), out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def search(metric: int, db: np.ndarray, q: np.ndarray, k: int, ids: np.ndarray | None = None,
           order: int = ORDER_SIMD, chunk: int = 4, rowpar: bool = False):
    """index.search(q, k) over a flat index holding `db` (memo_cli.py:292)."""
    db = _f32(db)
    q = np.atleast_2d(_f32(q))
    n, d = db.shape if db.ndim == 2 else (0, q.shape[1])
    nq = q.shape[0]
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    idp = None
    if ids is not None:
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        idp = _pi(ids)
    if order == ORDER_DEVICE:
        order = device_order(d, chunk)
    fn = lib().oracle_search_rowpar if rowpar else lib().oracle_search
    rc = fn(metric, order, chunk, _pf(db), n, d, idp, _pf(q), nq, k, _pf(D), _pi(I))
    if rc:
        raise MemoryError("oracle_search")
    return D, I


def merge_topk(metric: int, D_parts: np.ndarray, I_parts: np.ndarray):
    """[G,nq,k] shard-major best-first lists -> [nq,k] (K4 restatement)."""
    Dp, Ip = _f32(D_parts), np.ascontiguousarray(I_parts, dtype=np.int64)
    G, nq, k = Dp.shape
    Do = np.empty((nq, k), dtype=np.float32)
    Io = np.empty((nq, k), dtype=np.int64)
    lib().oracle_merge_topk(metric, G, nq, k, _pf(Dp), _pi(Ip), _pf(Do), _pi(Io))
    return Do, Io


# ---- numpy twin (cross-checks the C oracle; fp64 ranking) ----------------------------------------

def np_search_f64(metric: int, db: np.ndarray, q: np.ndarray, k: int, ids: np.ndarray | None = None):
    """Brute-force ranking in float64 with the stated tie rule (stable sort => smaller row first)."""
    db64 = np.asarray(db, dtype=np.float64)
    q64 = np.atleast_2d(np.asarray(q, dtype=np.float64))
    nq, n = q64.shape[0], db64.shape[0]
    D = np.full((nq, k), -np.finfo(np.float32).max if metric == METRIC_IP else np.finfo(np.float32).max, dtype=np.float64)
    I = np.full((nq, k), -1, dtype=np.int64)
    for i in range(nq):
        if metric == METRIC_IP:
            s = db64 @ q64[i]
            order = np.argsort(-s, kind="stable")
        else:
            s = ((db64 - q64[i]) ** 2).sum(axis=1)
            order = np.argsort(s, kind="stable")
        m = min(k, n)
        rows = order[:m]
        D[i, :m] = s[rows]
        I[i, :m] = rows if ids is None else np.asarray(ids)[rows]
    return D, I


def check_topk_against_truth(metric: int, db: np.ndarray, q: np.ndarray, D: np.ndarray, I: np.ndarray,
                             rows_of_ids=None, rel_eps: float = 1e-5, abs_floor: float | None = None):
    """Adjudicated parity check of one result list against the fp64 truth (SURVEY.md §7 hard part 1).

    Tolerance of a score with fp64 value t:  tol(t) = rel_eps * |t| + abs_floor.
      rel_eps   1e-5, the relative tolerance north_star states for fp32 distances — applied PER ELEMENT to |t|;
      abs_floor the explicit absolute term for scores near zero, where a relative bound is meaningless: by default
                2^-24 * sum_i |q_i * y_i| bounded by 2^-24 * ||q|| * max||y|| (half an fp32 ulp of the largest
                possible magnitude of the sum) — 6e-8 for unit vectors.
    Passes iff (a) every returned distance is within tol of the fp64 score of its row, (b) the returned order is
    best-first up to tol of the pair, and (c) the result is complete: no row outside it has an fp64 score better
    than the worst returned one by more than tol.  Returns human-readable violations (empty == pass).
    """
    q = np.asarray(q, dtype=np.float32).reshape(-1)
    truth = scores_f64(metric, db, q)
    sign = -1.0 if metric == METRIC_IP else 1.0  # smaller sign*score is better
    bad: list[str] = []
    rows = np.asarray(I if rows_of_ids is None else [rows_of_ids[int(i)] if i >= 0 else -1 for i in I], dtype=np.int64)
    valid = rows >= 0
    got = rows[valid]
    if len(set(got.tolist())) != len(got):
        bad.append("duplicate rows in result")
    if len(got) != min(len(rows), db.shape[0]):
        bad.append(f"expected {min(len(rows), db.shape[0])} valid results, got {len(got)}")
    if abs_floor is None:
        qn = float(np.linalg.norm(q.astype(np.float64)))
        yn = float(np.sqrt(np.max(np.einsum("ij,ij->i", db.astype(np.float64), db.astype(np.float64))))) if db.size else 0.0
        mag = qn * yn if metric == METRIC_IP else (qn + yn) ** 2
        abs_floor = 2.0 ** -24 * mag

    def tol(t):
        return rel_eps * abs(float(t)) + abs_floor

    for pos, (r, dist) in enumerate(zip(rows, np.asarray(D, dtype=np.float64))):
        if r < 0:
            continue
        if abs(dist - truth[r]) > tol(truth[r]):
            bad.append(f"pos {pos}: distance {dist} vs fp64 {truth[r]} (tolerance {tol(truth[r])})")
    t = sign * truth[got]
    for j in range(1, len(t)):
        if t[j] < t[j - 1] - tol(t[j - 1]):
            bad.append(f"result not best-first within tolerance at position {j}")
            break
    if len(got):
        worst = np.max(t)
        mask = np.ones(db.shape[0], dtype=bool)
        mask[got] = False
        if mask.any():
            best_out = np.min(sign * truth[mask])
            if best_out < worst - tol(worst):
                bad.append(f"a better row was left out: outside {best_out} vs worst kept {worst}")
    return bad
