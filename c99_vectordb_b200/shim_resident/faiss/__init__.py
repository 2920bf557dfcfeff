"""Drop-in `faiss` module for memo that talks to the resident index service: put
c99_vectordb_b200/shim_resident on PYTHONPATH and `import faiss` in memo_cli.py:13 resolves here.
Every CLI process then attaches to rows already resident in HBM instead of re-reading and
re-uploading the .memo file.  Socket: $B200_RESIDENT_SOCKET; set B200_RESIDENT_AUTOSTART=1 to spawn
the service on first use.  See INTEGRATION.md §8."""
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
if _root not in _sys.path:
    _sys.path.insert(0, _root)

from c99_vectordb_b200.resident import (  # noqa: F401,E402
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
    read_index,
    vector_to_array,
    write_index,
)
