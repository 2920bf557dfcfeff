"""c99_vectordb_b200 — B200-native flat vector recall path of memo (mikesmullin/c99-vectordb v2).

The importable spelling of the package named `c99-vectordb_b200` (a hyphen cannot be imported).
Holds only what the hot path needs: csrc/ (sm_100a kernels + the C ABI of include/b200_flat.h),
the faiss-shaped host surface memo_cli.py drives (index.py), the mirror of memo's index adapter
(memo_adapter.py) and the row-sharded multi-GPU index (sharded.py).
"""
from .index import (  # noqa: F401
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

__version__ = "0.1.0"
