"""Drop-in `faiss` module for memo: put c99_vectordb_b200/shim on PYTHONPATH (ahead of any real
faiss) and `import faiss` in memo_cli.py:13 resolves here.  See INTEGRATION.md."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from c99_vectordb_b200.index import *  # noqa: F401,F403,E402
from c99_vectordb_b200.index import (  # noqa: F401,E402
    METRIC_INNER_PRODUCT,
    METRIC_L2,
    Index,
    IndexFlat,
    IndexFlatIP,
    IndexFlatL2,
    IndexHNSWFlat,
    IndexIDMap,
    IndexIDMap2,
    Int64Vector,
    normalize_L2,
    read_index,
    vector_to_array,
    write_index,
)
