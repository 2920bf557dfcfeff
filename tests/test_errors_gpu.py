"""Error behaviour at the boundary mirrors faiss's: RuntimeError for engine failures (status != 0 from the
C ABI with the message of b200_last_error), AssertionError for shape checks in the Python surface."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def b200(gpu):
    import c99_vectordb_b200 as m

    return m


def test_shape_and_argument_checks(b200):
    idx = b200.IndexFlatL2(16)
    x = oracle.synth_rows(10, 16, 1)
    idx.add(x)
    with pytest.raises(AssertionError):
        idx.add(np.zeros((3, 15), np.float32))          # wrong dimension
    with pytest.raises(AssertionError):
        idx.search(np.zeros((1, 17), np.float32), 3)
    with pytest.raises(AssertionError):
        idx.search(x[:1], 0)                             # k must be positive (faiss asserts k > 0)
    with pytest.raises(RuntimeError):
        idx.add_with_ids(x, np.arange(10))               # not implemented for a bare flat index, as faiss
    with pytest.raises(AssertionError):
        idx.search(x[:1], 3, row_mask=np.ones(9, bool))  # mask must cover every row
    with pytest.raises(RuntimeError):
        idx.set_option("no_such_option", 1)
    with pytest.raises(ValueError):
        b200.IndexFlat(16, 1, store="fp8")
    with pytest.raises(RuntimeError):
        b200.IndexFlat(0, 1)


def test_idmap_rules(b200):
    base = b200.IndexFlatIP(8)
    w = b200.IndexIDMap2(base)
    x = oracle.synth_rows(4, 8, 2)
    with pytest.raises(RuntimeError):
        w.add(x)                                         # "add does not make sense with IndexIDMap"
    with pytest.raises(AssertionError):
        w.add_with_ids(x, np.arange(3))                  # ids/vectors count mismatch
    w.add_with_ids(x, np.array([7, 7, 8, 9]))            # duplicate ids are allowed, as in faiss
    D, I = w.search(x[:1], 4)
    assert sorted(I[0].tolist()) == [7, 7, 8, 9]
    with pytest.raises(RuntimeError):
        b200.IndexIDMap2(base)                           # "index must be empty on input"
    with pytest.raises(RuntimeError):
        w.reconstruct(12345)
    np.testing.assert_array_equal(w.reconstruct(9), x[3])


def test_c_abi_reports_errors_with_messages(b200):
    from c99_vectordb_b200 import _cabi

    L = _cabi.load()
    h = C.c_void_p()
    assert L.b200_index_create(C.byref(h), 8, 5, 0, 0) != 0 and b"metric" in L.b200_last_error()
    assert L.b200_index_create(C.byref(h), 8, 0, 0, 99) != 0 and b"device" in L.b200_last_error()
    assert L.b200_index_create(C.byref(h), 8, 0, 0, 0) == 0
    D = np.empty((1, 3), np.float32); I = np.empty((1, 3), np.int64); q = np.zeros((1, 8), np.float32)
    assert L.b200_index_search(h, q.ctypes.data, 1, -1, D.ctypes.data, I.ctypes.data) != 0 and b"k must be positive" in L.b200_last_error()
    assert L.b200_index_search(h, None, 1, 3, D.ctypes.data, I.ctypes.data) != 0
    assert L.b200_index_search(h, q.ctypes.data, 1, 3, D.ctypes.data, I.ctypes.data) == 0 and (I == -1).all()  # empty index pads
    x = np.ones((2, 8), np.float32); ids = np.array([1, 2], np.int64)
    assert L.b200_index_add(h, x.ctypes.data, 2, None, 0) == 0
    assert L.b200_index_add(h, x.ctypes.data, 2, ids.ctypes.data, 0) != 0 and b"mix" in L.b200_last_error()
    assert L.b200_index_get_rows(h, 1, 5, x.ctypes.data) != 0 and b"out of bounds" in L.b200_last_error()
    assert L.b200_index_search_exchange_dev(h, None, 1, 3, None, None, None) != 0   # exchange not configured
    assert L.b200_index_reset(h) == 0 and L.b200_index_ntotal(h) == 0
    assert L.b200_index_destroy(h) == 0
