"""Where a one-shot process spends its read_index time (memo_cli.py:251-261): device start-up, row
allocation, pinned ring, file -> device.  Run in a fresh process per measurement."""
import json, os, subprocess, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CHILD = r"""
import sys, time, json, ctypes as C
t0 = time.perf_counter()
import numpy as np
from c99_vectordb_b200 import index as ix, _cabi
t1 = time.perf_counter()
L = _cabi.load()
t2 = time.perf_counter()
probe = ix.IndexFlat(8, 1)          # CUDA start-up lands here
t3 = time.perf_counter()
idx = ix.read_index(sys.argv[1])
t4 = time.perf_counter()
q = np.zeros((1, idx.d), np.float32); q[0, 0] = 1
D, I = idx.search(q, 10)
t5 = time.perf_counter()
D, I = idx.search(q, 10)
t6 = time.perf_counter()
print(json.dumps(dict(import_s=t1-t0, dlopen_s=t2-t1, cuda_start_s=t3-t2, read_index_s=t4-t3, first_search_s=t5-t4, second_search_s=t6-t5)))
"""

def main():
    import numpy as np
    from c99_vectordb_b200 import index as ix
    n, d = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 384
    tmp = tempfile.mkdtemp(prefix="b200cold_")
    path = os.path.join(tmp, "db.memo")
    idx = ix.IndexIDMap2(ix.IndexHNSWFlat(d, 32))
    idx.index.add_synthetic(n, 1234, with_ids=True)
    t0 = time.perf_counter(); ix.write_index(idx, path); wt = time.perf_counter() - t0
    del idx
    print(json.dumps(dict(n=n, d=d, file_gb=os.path.getsize(path) / 1e9, write_index_s=wt)), flush=True)
    for env in ({}, {"B200_UPLOAD_THREADS": "4"}, {"B200_UPLOAD_THREADS": "32"}, {"CUDA_MODULE_LOADING": "EAGER"}, {}):
        out = subprocess.run([sys.executable, "-c", CHILD, path], env=dict(os.environ, PYTHONPATH=str(ROOT), **env),
                             check=True, capture_output=True, text=True)
        r = json.loads(out.stdout.strip().splitlines()[-1]); r["env"] = env
        print(json.dumps(r), flush=True)

main()
