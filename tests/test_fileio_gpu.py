"""`.memo` payload I/O through b200_index_add_file / b200_index_write_file (faiss
layout + fast load): file -> pinned ring -> device and back, no intermediate host copies.  The
file bytes are checked against numpy's own serialisation of the same arrays (faiss layout,
SURVEY.md App. A.5) and searches after a reload must be bit-identical."""
import struct

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200(gpu):
    import c99_vectordb_b200 as m

    return m


def _header(d, n, metric):
    return struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, metric)


def _expected_file(d, metric, db, ids):
    n = db.shape[0]
    flat = (b"IxFI" if metric == 0 else b"IxF2") + _header(d, n, metric) + struct.pack("<Q", n * d) + db.tobytes()
    if ids is None:
        return flat
    return b"IxM2" + _header(d, n, metric) + flat + struct.pack("<Q", n) + ids.tobytes()


@pytest.mark.parametrize("n,d,metric,with_ids", [
    (1, 384, 1, True),          # memo's first save (memo_cli.py:437, :448)
    (777, 10, 0, True),         # padded pitch (d % 4 != 0): file rows go through K1
    (5000, 768, 0, False),      # flat without id map
    (300_000, 96, 1, True),     # 115 MB payload: several 64 MB ring chunks, threaded pread / pwrite
])
def test_file_bytes_and_reload(b200, tmp_path, n, d, metric, with_ids):
    db = oracle.synth_rows(n, d, 21)
    ids = (np.arange(n, dtype=np.int64) * 5 + 3) if with_ids else None
    base = b200.IndexFlat(d, metric)
    idx = b200.IndexIDMap2(base) if with_ids else base
    if with_ids:
        idx.add_with_ids(db, ids)
    else:
        idx.add(db)
    p = tmp_path / "db.memo"
    b200.write_index(idx, str(p))
    assert p.read_bytes() == _expected_file(d, metric, db, ids)
    back = b200.read_index(str(p))
    assert type(back) is type(idx) and back.ntotal == n and back.d == d and back.metric_type == metric
    if with_ids:
        np.testing.assert_array_equal(b200.vector_to_array(back.id_map), ids)
        np.testing.assert_array_equal(back.index.reconstruct_n(0, n), db)
    else:
        np.testing.assert_array_equal(back.reconstruct_n(0, n), db)
    q = oracle.synth_rows(2, d, 22)
    k = min(n, 10)
    Da, Ia = idx.search(q, k)
    Db, Ib = back.search(q, k)
    np.testing.assert_array_equal(Ia, Ib)
    np.testing.assert_array_equal(Da, Db)
    # appending after a reload keeps working (memo save: load_index -> add_with_ids -> write_index)
    if with_ids:
        back.add_with_ids(db[:1], np.array([10**9], dtype=np.int64))
        b200.write_index(back, str(p))
        again = b200.read_index(str(p))
        assert again.ntotal == n + 1 and b200.vector_to_array(again.id_map)[-1] == 10**9


def test_bf16_rows_are_written_widened(b200, tmp_path):
    d, n = 64, 1000
    db = oracle.synth_rows(n, d, 23)
    idx = b200.IndexIDMap2(b200.IndexFlat(d, 0, store="bf16"))
    idx.add_with_ids(db, np.arange(n, dtype=np.int64))
    p = tmp_path / "bf16.memo"
    b200.write_index(idx, str(p))
    assert p.read_bytes() == _expected_file(d, 0, oracle.round_bf16(db), np.arange(n, dtype=np.int64))


def test_empty_index_roundtrip(b200, tmp_path):
    p = tmp_path / "empty.memo"
    b200.write_index(b200.IndexIDMap2(b200.IndexFlatL2(384)), str(p))
    assert p.read_bytes() == _expected_file(384, 1, np.zeros((0, 384), np.float32), np.zeros((0,), np.int64))
    back = b200.read_index(str(p))
    assert isinstance(back, b200.IndexIDMap2) and back.ntotal == 0


@pytest.mark.parametrize("cut", ["rows", "ids", "count"])
def test_truncated_payload_raises(b200, tmp_path, cut):
    """memo catches Exception around read_index and starts a fresh index (memo_cli.py:254-257)."""
    d, n = 32, 100
    db = oracle.synth_rows(n, d, 24)
    ids = np.arange(n, dtype=np.int64)
    blob = _expected_file(d, 1, db, ids)
    keep = {"rows": len(blob) - n * 8 - 8 - 100, "ids": len(blob) - 8, "count": 37 + 4 + 33 + 4}[cut]
    p = tmp_path / "cut.memo"
    p.write_bytes(blob[:keep])
    with pytest.raises(RuntimeError):
        b200.read_index(str(p))
    flat = _expected_file(d, 1, db, None)
    p.write_bytes(flat[: len(flat) - 4])
    with pytest.raises(RuntimeError, match="shorter|read error"):
        b200.read_index(str(p))


def test_write_to_unwritable_path_raises(b200, tmp_path):
    idx = b200.IndexFlatL2(8)
    idx.add(np.ones((3, 8), np.float32))
    with pytest.raises((OSError, RuntimeError), match="No such file"):
        b200.write_index(idx, str(tmp_path / "no_such_dir" / "x.memo"))


def test_c_level_save_and_load(b200, tmp_path):
    """b200_index_save / b200_index_load: what a cgo / JNI host binds instead of faiss_write_index_fname /
    faiss_read_index_fname [upstream C API] — whole files, no Python header code involved."""
    import ctypes as C

    from c99_vectordb_b200 import _cabi

    L = _cabi.load()
    d, n = 48, 3000
    db = oracle.synth_rows(n, d, 41)
    ids = np.arange(n, dtype=np.int64) * 11 + 5
    idx = b200.IndexIDMap2(b200.IndexFlat(d, 0))
    idx.add_with_ids(db, ids)
    p = str(tmp_path / "c.memo").encode()
    _cabi.check(L.b200_index_save(idx.index._h, p, 2))
    assert open(p, "rb").read() == _expected_file(d, 0, db, ids)
    h, info = C.c_void_p(), _cabi.MemoInfo()
    _cabi.check(L.b200_index_load(C.byref(h), p, 0, 0, C.byref(info)))
    try:
        assert (info.kind, info.d, info.metric, info.ntotal) == (2, d, 0, n)
        assert L.b200_index_ntotal(h) == n and L.b200_index_has_ids(h) == 1
        q = oracle.synth_rows(2, d, 42)
        D = np.empty((2, 7), np.float32)
        I = np.empty((2, 7), np.int64)
        _cabi.check(L.b200_index_search(h, q.ctypes.data, 2, 7, D.ctypes.data, I.ctypes.data))
        Dw, Iw = idx.search(q, 7)
        np.testing.assert_array_equal(I, Iw)
        np.testing.assert_array_equal(D, Dw)
    finally:
        L.b200_index_destroy(h)
    # a bare flat index, bf16-resident after load
    _cabi.check(L.b200_index_save(idx.index._h, p, 0))
    _cabi.check(L.b200_index_load(C.byref(h), p, 1, 0, None))
    try:
        assert L.b200_index_ntotal(h) == n and L.b200_index_has_ids(h) == 0 and L.b200_index_store(h) == 1
    finally:
        L.b200_index_destroy(h)
    open(p, "wb").write(b"IxF2 nonsense")
    assert L.b200_index_load(C.byref(h), p, 0, 0, None) != 0 and not h.value


def test_c99_example_program(b200, tmp_path):
    """examples/memo_recall.c, compiled with gcc -std=c99 -Wpedantic -Werror, loads a .memo file and searches."""
    import subprocess
    import sys
    from pathlib import Path

    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from test_c_example_cpu import build_example, write_memo

    exe = build_example(tmp_path)
    d, n = 384, 500
    db = oracle.normalize_rows(oracle.synth_rows(n, d, 51))
    ids = np.arange(n, dtype=np.int64) + 1000
    write_memo(tmp_path / "e.memo", db, ids)
    r = subprocess.run([str(exe), str(tmp_path / "e.memo"), "3"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert "id-mapped index, d=384, L2, 500 rows" in lines[0]
    assert lines[1].split()[:3] == ["1", "id", "1000"] and float(lines[1].split()[-1]) == 0.0
    Dw, Iw = oracle.search(1, db, db[:1], 3, ids=ids, order=oracle.ORDER_DEVICE)
    assert [int(l.split()[2]) for l in lines[1:4]] == Iw[0].tolist()
