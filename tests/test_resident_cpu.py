"""Resident index service — host logic on CPU.

The service is started in-process with the CPU stub as its backend (tests/stub_faiss_oracle.py, test
infrastructure) so that the protocol, the path cache (mtime/size validation, dirty entries, LRU
eviction), error propagation and the faiss-shaped client surface are covered without a GPU.  The
same scenarios run against the CUDA index in tests/test_resident_gpu.py."""
import importlib
import os
import sys
import threading
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
import stub_faiss_oracle as stub  # noqa: E402

from c99_vectordb_b200 import resident  # noqa: E402


@pytest.fixture()
def service(tmp_path):
    sock = str(tmp_path / "svc.sock")
    srv = resident.make_server(sock, backend=stub, max_resident=2)
    th = threading.Thread(target=srv.serve_forever, kwargs={"poll_interval": 0.02}, daemon=True)
    th.start()
    c = resident.ResidentClient(sock, autostart=False)
    resident.set_client(c)
    yield SimpleService(srv, sock, c)
    resident.set_client(None)
    c.close()
    srv.shutdown()
    srv.server_close()


class SimpleService:
    def __init__(self, srv, sock, client):
        self.srv, self.sock, self.client = srv, sock, client

    def new_process(self):
        """What a fresh CLI process looks like to the service: a new connection, no handles."""
        c = resident.ResidentClient(self.sock, autostart=False)
        resident.set_client(c)
        return c

    def stats(self):
        return self.client.call("stats")[0]


def _rows(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def test_socket_is_private(service):
    assert (os.stat(service.sock).st_mode & 0o777) == 0o600


def test_create_add_search_matches_backend(service):
    d = 32
    x, q = _rows(200, d, 1), _rows(3, d, 2)
    ids = np.arange(200, dtype=np.int64) * 3 + 7
    base = resident.IndexHNSWFlat(d, 32)
    base.hnsw.efConstruction = 200  # memo_cli.py:246-247
    base.hnsw.efSearch = 64
    idx = resident.IndexIDMap2(base)
    assert isinstance(idx, resident.IndexIDMap2) and idx.ntotal == 0
    assert resident.vector_to_array(idx.id_map).shape == (0,)
    idx.add_with_ids(x[:120], ids[:120])
    idx.add_with_ids(x[120:], ids[120:])
    assert idx.ntotal == 200
    np.testing.assert_array_equal(resident.vector_to_array(idx.id_map), ids)
    D, I = idx.search(q, 10)
    ref = stub.IndexIDMap2(stub.IndexFlat(d, stub.METRIC_L2))
    ref.add_with_ids(x, ids)
    Dr, Ir = ref.search(q, 10)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D, Dr)
    assert D.dtype == np.float32 and I.dtype == np.int64 and D.flags.writeable


def test_flat_ip_without_idmap(service):
    d = 16
    x, q = _rows(50, d, 3), _rows(2, d, 4)
    idx = resident.IndexFlatIP(d)
    idx.add(x)
    D, I = idx.search(q, 5)
    ref = stub.IndexFlatIP(d)
    ref.add(x)
    Dr, Ir = ref.search(q, 5)
    np.testing.assert_array_equal(I, Ir)
    np.testing.assert_array_equal(D, Dr)


def test_rows_stay_resident_across_processes(service, tmp_path):
    d = 24
    x, q = _rows(300, d, 5), _rows(1, d, 6)
    path = tmp_path / "db.memo"
    idx = resident.IndexIDMap2(resident.IndexFlatL2(d))
    idx.add_with_ids(x, np.arange(300, dtype=np.int64))
    resident.write_index(idx, str(path))
    assert path.exists()
    D0, I0 = idx.search(q, 7)
    del idx
    for _ in range(3):  # three later CLI invocations
        c = service.new_process()
        again = resident.read_index(str(path))
        assert isinstance(again, resident.IndexIDMap2) and again.ntotal == 300 and again.d == d
        D, I = again.search(q, 7)
        np.testing.assert_array_equal(I, I0)
        np.testing.assert_array_equal(D, D0)
        del again
        c.close()
    st = service.stats()
    assert st["loads"] == 0 and st["hits"] == 3 and st["writes"] == 1  # never re-read from disk
    assert st["resident"] == [str(path)]


def test_changed_file_is_reloaded(service, tmp_path):
    d = 8
    path = tmp_path / "db.memo"
    a = stub.IndexIDMap2(stub.IndexFlat(d, stub.METRIC_L2))
    a.add_with_ids(_rows(10, d, 7), np.arange(10, dtype=np.int64))
    stub.write_index(a, str(path))
    assert resident.read_index(str(path)).ntotal == 10
    assert resident.read_index(str(path)).ntotal == 10
    assert service.stats()["loads"] == 1 and service.stats()["hits"] == 1
    a.add_with_ids(_rows(5, d, 8), np.arange(10, 15, dtype=np.int64))
    stub.write_index(a, str(path))  # another writer replaced the file behind the service's back
    os.utime(path, ns=(1, 1))       # even with an older timestamp the stamp differs
    assert resident.read_index(str(path)).ntotal == 15
    assert service.stats()["loads"] == 2


def test_unwritten_mutation_is_dropped(service, tmp_path):
    d = 8
    path = tmp_path / "db.memo"
    idx = resident.IndexIDMap2(resident.IndexFlatL2(d))
    idx.add_with_ids(_rows(10, d, 9), np.arange(10, dtype=np.int64))
    resident.write_index(idx, str(path))
    same = resident.read_index(str(path))
    same.add_with_ids(_rows(1, d, 10), np.array([99], dtype=np.int64))  # ...and the process dies before write_index
    assert same.ntotal == 11
    fresh = resident.read_index(str(path))  # the file is the truth (memo_cli.py:251-261)
    assert fresh.ntotal == 10
    assert 99 not in resident.vector_to_array(fresh.id_map).tolist()


def test_lru_eviction(service, tmp_path):
    d = 8
    paths = []
    for i in range(3):
        idx = resident.IndexIDMap2(resident.IndexFlatL2(d))
        idx.add_with_ids(_rows(4 + i, d, 20 + i), np.arange(4 + i, dtype=np.int64))
        p = tmp_path / f"db{i}.memo"
        resident.write_index(idx, str(p))
        paths.append(str(p))
    st = service.stats()
    assert st["evictions"] == 1 and st["resident"] == sorted(paths[1:])  # max_resident = 2
    assert resident.read_index(paths[0]).ntotal == 4  # evicted -> loaded again
    assert service.stats()["loads"] == 1


def test_errors_travel_as_runtime_error(service, tmp_path):
    with pytest.raises(RuntimeError, match="FileNotFoundError"):
        resident.read_index(str(tmp_path / "missing.memo"))
    bad = tmp_path / "bad.memo"
    bad.write_bytes(b"not an index")
    with pytest.raises(RuntimeError):
        resident.read_index(str(bad))
    idx = resident.IndexIDMap2(resident.IndexFlatL2(8))
    with pytest.raises(AssertionError):
        idx.add_with_ids(np.zeros((2, 9), np.float32), np.arange(2))  # faiss wrapper's shape assertion
    with pytest.raises(RuntimeError, match="add_with_ids"):
        idx.add(np.zeros((1, 8), np.float32))
    with pytest.raises(RuntimeError, match="unknown op"):
        service.client.call("no_such_op", {"h": 1})
    # the connection survives every one of those
    assert service.client.call("ping")[0]["pid"] == os.getpid()


def test_bad_frames_do_not_kill_the_service(service):
    import socket as so

    s = so.socket(so.AF_UNIX, so.SOCK_STREAM)
    s.connect(service.sock)
    s.sendall(b"GET / HTTP/1.0\r\n\r\n")
    s.close()
    s = so.socket(so.AF_UNIX, so.SOCK_STREAM)
    s.connect(service.sock)
    s.sendall(resident.MAGIC + (1 << 30).to_bytes(4, "little"))  # absurd header length
    s.close()
    assert service.new_process().call("ping")[0]["backend"] == stub.__name__


def test_no_service_raises_connection_error(tmp_path):
    with pytest.raises(ConnectionError):
        resident.ResidentClient(str(tmp_path / "nobody.sock"), autostart=False)


def test_second_service_on_same_socket_is_refused(service):
    with pytest.raises(RuntimeError, match="already listens"):
        resident.make_server(service.sock, backend=stub)


def test_shim_resident_exports_what_memo_touches():
    shim = Path(resident.__file__).parent / "shim_resident"
    sys.path.insert(0, str(shim))
    saved = sys.modules.pop("faiss", None)
    try:
        f = importlib.import_module("faiss")
        for name in ("IndexHNSWFlat", "IndexIDMap2", "read_index", "write_index", "vector_to_array",
                     "IndexFlatIP", "IndexFlatL2"):  # memo_cli.py:245 :248 :255 :361 :268
            assert hasattr(f, name), name
        assert f.IndexIDMap2 is resident.IndexIDMap2
    finally:
        sys.path.remove(str(shim))
        sys.modules.pop("faiss", None)
        if saved is not None:
            sys.modules["faiss"] = saved


@pytest.mark.skipif(not Path("/root/reference/memo_cli.py").exists(), reason="reference tree not present")
def test_reference_adapter_runs_unmodified_over_the_service(service, tmp_path, monkeypatch):
    """The reference's own create_index / rebuild_index_from_texts / load_index / get_existing_ids /
    search_all (memo_cli.py:244-298) running against the resident client classes give what they give
    against the in-process stub."""
    monkeypatch.setenv("PYTHONHASHSEED", "0")
    texts = ["alpha beta gamma", None, "peanut allergy note", "   ", "gamma delta", "beta beta alpha"]

    def load_memo(faiss_module):
        sys.modules.pop("memo_cli", None)
        monkeypatch.setitem(sys.modules, "faiss", faiss_module)
        monkeypatch.syspath_prepend("/root/reference")
        return importlib.import_module("memo_cli")

    m = load_memo(resident)
    idx = m.rebuild_index_from_texts(texts, False)
    assert isinstance(idx, resident.IndexIDMap2)
    path = tmp_path / "ref.memo"
    resident.write_index(idx, str(path))
    service.new_process()
    loaded = m.load_index(path, False)
    assert isinstance(loaded, resident.IndexIDMap2)
    assert m.get_existing_ids(loaded) == {0, 2, 4, 5}
    qv = m.embed_text_hash("alpha gamma")
    got = m.search_all(loaded, qv)
    assert service.stats()["loads"] == 0  # attached to the resident rows

    m2 = load_memo(stub)
    ref = m2.search_all(m2.rebuild_index_from_texts(texts, False), qv)
    assert [r.doc_id for r in got] == [r.doc_id for r in ref]
    assert [r.score for r in got] == [r.score for r in ref]
    sys.modules.pop("memo_cli", None)


def test_concurrent_clients_are_serialised_correctly(service, tmp_path):
    """Several CLI processes at once: each connection has its own handle table, requests are
    serialised by the service lock, nobody sees another client's rows."""
    d = 16
    path = tmp_path / "shared.memo"
    shared = resident.IndexIDMap2(resident.IndexFlatL2(d))
    xs = _rows(400, d, 40)
    shared.add_with_ids(xs, np.arange(400, dtype=np.int64))
    resident.write_index(shared, str(path))
    ref = stub.IndexIDMap2(stub.IndexFlat(d, stub.METRIC_L2))
    ref.add_with_ids(xs, np.arange(400, dtype=np.int64))
    errors = []

    def worker(seed):
        try:
            c = resident.ResidentClient(service.sock, autostart=False)
            info, _ = c.call("open", {"path": str(path)})
            mine, _ = c.call("create", {"d": d, "metric": 1, "idmap": True})
            own = _rows(20 + seed, d, 100 + seed)
            c.call("add", {"h": mine["h"], "with_ids": True}, (own, np.arange(20 + seed, dtype=np.int64) + 1000 * seed))
            for it in range(25):
                q = _rows(2, d, 1000 * seed + it)
                _, (D, I) = c.call("search", {"h": info["h"], "k": 5}, (q,))
                Dr, Ir = ref.search(q, 5)
                assert np.array_equal(I, Ir) and np.array_equal(D, Dr)
                _, (D2, I2) = c.call("search", {"h": mine["h"], "k": 3}, (q,))
                assert ((I2 >= 1000 * seed) & (I2 < 1000 * seed + 20 + seed)).all()
            assert c.call("ntotal", {"h": mine["h"]})[0]["ntotal"] == 20 + seed
            c.close()
        except Exception as ex:  # surfaced in the main thread
            errors.append(repr(ex))

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(1, 7)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    st = service.stats()
    assert st["loads"] == 0 and st["hits"] == 6


def test_unsaved_mutations_are_private_to_their_connection(service, tmp_path):
    """Every process of the reference holds a private copy of the index it read.  Two connections that opened the
    same cached file: an add on one is invisible to the other until it is written; a sole holder mutates in place
    (no reload) and the entry returns to the cache on write."""
    d = 16
    x = _rows(40, d, 3)
    ids = np.arange(40, dtype=np.int64)
    path = str(tmp_path / "shared.memo")
    first = resident.IndexIDMap2(resident.IndexHNSWFlat(d, 32))
    first.add_with_ids(x[:30], ids[:30])
    resident.write_index(first, path)

    a_conn = service.new_process()
    a = resident.read_index(path)
    b_conn = service.new_process()
    b = resident.read_index(path)
    assert a.ntotal == 30 and b.ntotal == 30
    loads_before = service.stats()["loads"]
    resident.set_client(a_conn)
    a.add_with_ids(x[30:35], ids[30:35])          # shared entry: a gets a private copy
    assert a.ntotal == 35
    resident.set_client(b_conn)
    assert b.ntotal == 30                          # b still sees what it read
    Db, Ib = b.search(x[31:32], 1)
    assert Ib[0, 0] != 31
    c_conn = service.new_process()
    c = resident.read_index(path)                  # a fresh process reads the file: 30 rows
    assert c.ntotal == 30
    assert service.stats()["loads"] == loads_before + 1   # the private copy; c and b share the cached entry
    resident.set_client(a_conn)
    resident.write_index(a, path)
    d_conn = service.new_process()
    dd = resident.read_index(path)
    assert dd.ntotal == 35
    for cn in (a_conn, b_conn, c_conn, d_conn):
        cn.close()

    # sole holder: in place, no extra load
    e_conn = service.new_process()
    e = resident.read_index(path)
    loads = service.stats()["loads"]
    e.add_with_ids(x[35:], ids[35:])
    assert e.ntotal == 40 and service.stats()["loads"] == loads
    f_conn = service.new_process()
    f = resident.read_index(path)                  # the dirty entry left the cache: the file is the truth
    assert f.ntotal == 35
    e_conn.close()
    f_conn.close()


def test_default_socket_lives_in_a_private_directory(monkeypatch, tmp_path):
    monkeypatch.delenv("B200_RESIDENT_SOCKET", raising=False)
    monkeypatch.delenv("XDG_RUNTIME_DIR", raising=False)
    p = resident.default_socket_path()
    st = os.stat(os.path.dirname(p))
    assert st.st_uid == os.getuid() and (st.st_mode & 0o077) == 0
    open_dir = tmp_path / "open"
    open_dir.mkdir(mode=0o755)
    os.chmod(open_dir, 0o755)
    with pytest.raises(PermissionError):
        resident._private_dir(str(open_dir))
