"""Resident index service on the CUDA index: a separate service process holds the
rows in HBM; client processes attach through the faiss-shaped proxy classes and must see exactly
what the in-process index returns."""
import os
import subprocess
import sys
import time
from pathlib import Path

import numpy as np
import pytest

from c99_vectordb_b200 import index as ix
from c99_vectordb_b200 import resident

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture()
def service(tmp_path, gpu):
    sock = str(tmp_path / "svc.sock")
    proc = subprocess.Popen([sys.executable, "-m", "c99_vectordb_b200.resident", "serve", "--socket", sock],
                            cwd=str(ROOT))
    deadline = time.monotonic() + 180
    c = None
    while time.monotonic() < deadline:
        try:
            c = resident.ResidentClient(sock, autostart=False)
            break
        except ConnectionError:
            assert proc.poll() is None, "service exited early"
            time.sleep(0.1)
    assert c is not None, "service did not come up"
    resident.set_client(c)
    yield sock
    resident.set_client(None)
    try:
        c.call("shutdown")
    except Exception:
        pass
    c.close()
    try:
        proc.wait(timeout=30)
    except subprocess.TimeoutExpired:
        proc.kill()


def test_service_matches_in_process_index(service, tmp_path):
    d, n = 384, 20000  # memo's DIM (memo_cli.py:17)
    rng = np.random.default_rng(11)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    ids = (np.arange(n, dtype=np.int64) * 2 + 1)
    q = x[[5, 77, 19999]] + 0.01 * rng.standard_normal((3, d)).astype(np.float32)

    local = ix.IndexIDMap2(ix.IndexHNSWFlat(d, 32))
    local.add_with_ids(x, ids)
    Dl, Il = local.search(q, 10)
    Dfull, Ifull = local.search(q[:1], n)  # memo's k = ntotal (memo_cli.py:291)

    remote = resident.IndexIDMap2(resident.IndexHNSWFlat(d, 32))
    remote.add_with_ids(x[: n // 2], ids[: n // 2])
    remote.add_with_ids(x[n // 2:], ids[n // 2:])
    path = tmp_path / "db.memo"
    resident.write_index(remote, str(path))
    del remote

    # a later "CLI process": new connection, attach by path
    c2 = resident.ResidentClient(service, autostart=False)
    resident.set_client(c2)
    again = resident.read_index(str(path))
    assert isinstance(again, resident.IndexIDMap2) and again.ntotal == n
    np.testing.assert_array_equal(resident.vector_to_array(again.id_map), ids)
    D, I = again.search(q, 10)
    np.testing.assert_array_equal(I, Il)
    np.testing.assert_array_equal(D, Dl)
    D2, I2 = again.search(q[:1], n)
    np.testing.assert_array_equal(I2, Ifull)
    np.testing.assert_array_equal(D2, Dfull)
    allowed = ids[100:200]
    Df, If = again.search(q, 10, ids_allowed=allowed)  # filter push-down through the service
    Dlf, Ilf = local.search(q, 10, ids_allowed=allowed)
    np.testing.assert_array_equal(If, Ilf)
    np.testing.assert_array_equal(Df, Dlf)
    st = c2.call("stats")[0]
    assert st["loads"] == 0 and st["hits"] == 1 and st["resident"] == [str(path)]

    # the file the service wrote is a plain faiss-layout flat index the in-process reader accepts
    disk = ix.read_index(str(path))
    assert disk.ntotal == n
    Dd, Id = disk.search(q, 10)
    np.testing.assert_array_equal(Id, Il)
    np.testing.assert_array_equal(Dd, Dl)


def test_service_errors_are_loud(service, tmp_path):
    with pytest.raises(RuntimeError):
        resident.read_index(str(tmp_path / "missing.memo"))
    idx = resident.IndexIDMap2(resident.IndexFlatIP(16))
    with pytest.raises(AssertionError):
        idx.add_with_ids(np.zeros((2, 16), np.float32), np.arange(3))  # ids/rows mismatch: wrapper assertion as in faiss
    idx.add_with_ids(np.ones((2, 16), np.float32), np.arange(2))
    with pytest.raises(RuntimeError):
        resident.write_index(idx, str(tmp_path / "no_such_dir" / "x.memo"))
    assert idx.ntotal == 2  # the connection and the index survive
