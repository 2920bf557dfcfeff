"""One-shot CLI latency with and without the resident service.

Models what memo does per invocation (memo_cli.py:883-949 -> load_index :251-261 -> search_all
:288-298): a FRESH python process imports the faiss module, read_index()es the .memo file and runs
one search with k = 10 (and one with k = ntotal, the reference's shape).  Arm A uses the in-process
shim (CUDA start-up + file read + upload every time); arm B uses the resident shim against a warm
service.  Wall clock of the whole child process and of its read_index+search part are reported.

usage: python tools/bench_resident.py [--n 1000000] [--d 384] [--runs 5]
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

CHILD = r"""
import sys, time, json
t0 = time.perf_counter()
import numpy as np
import faiss
t1 = time.perf_counter()
idx = faiss.read_index(sys.argv[1])
t2 = time.perf_counter()
q = np.load(sys.argv[2])
D, I = idx.search(q, 10)
t3 = time.perf_counter()
D2, I2 = idx.search(q, int(idx.ntotal))
t4 = time.perf_counter()
print(json.dumps({"import_s": t1 - t0, "read_index_s": t2 - t1, "search_k10_s": t3 - t2, "search_kall_s": t4 - t3,
                  "top": int(I[0, 0]), "ntotal": int(idx.ntotal)}))
"""


def run_child(shim: str, path: str, qpath: str, env_extra: dict) -> dict:
    env = dict(os.environ, PYTHONPATH=str(ROOT / "c99_vectordb_b200" / shim) + os.pathsep + str(ROOT), **env_extra)
    t0 = time.perf_counter()
    out = subprocess.run([sys.executable, "-c", CHILD, path, qpath], env=env, check=True, capture_output=True, text=True)
    wall = time.perf_counter() - t0
    r = json.loads(out.stdout.strip().splitlines()[-1])
    r["process_wall_s"] = wall
    return r


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=384)
    ap.add_argument("--runs", type=int, default=5)
    a = ap.parse_args()
    sys.path.insert(0, str(ROOT))
    import numpy as np

    from c99_vectordb_b200 import index as ix
    from c99_vectordb_b200 import resident

    tmp = tempfile.mkdtemp(prefix="b200res_")
    path, qpath, sock = os.path.join(tmp, "db.memo"), os.path.join(tmp, "q.npy"), os.path.join(tmp, "svc.sock")
    idx = ix.IndexIDMap2(ix.IndexHNSWFlat(a.d, 32))
    idx.index.add_synthetic(a.n, 1234, with_ids=True)
    ix.write_index(idx, path)
    q = idx.index.reconstruct_n(12345, 1) * 0.5
    np.save(qpath, q)
    del idx
    file_gb = os.path.getsize(path) / 1e9

    svc = subprocess.Popen([sys.executable, "-m", "c99_vectordb_b200.resident", "serve", "--socket", sock], cwd=str(ROOT))
    c = None
    for _ in range(1800):
        try:
            c = resident.ResidentClient(sock, autostart=False)
            break
        except ConnectionError:
            time.sleep(0.1)
    assert c is not None
    try:
        cold = [run_child("shim", path, qpath, {}) for _ in range(a.runs)]
        first = run_child("shim_resident", path, qpath, {"B200_RESIDENT_SOCKET": sock})  # loads into the service
        warm = [run_child("shim_resident", path, qpath, {"B200_RESIDENT_SOCKET": sock}) for _ in range(a.runs)]
        stats = c.call("stats")[0]
    finally:
        try:
            c.call("shutdown")
        except Exception:
            pass
        svc.wait(timeout=60)
    assert all(r["top"] == cold[0]["top"] for r in cold + warm + [first])

    def med(rs, key):
        return statistics.median(r[key] for r in rs)

    keys = ("process_wall_s", "import_s", "read_index_s", "search_k10_s", "search_kall_s")
    print(json.dumps({
        "workload": f"{a.n}x{a.d} fp32 L2 IDMap2 .memo file ({file_gb:.2f} GB), one-shot process: read_index + search k=10 + search k=ntotal",
        "runs": a.runs,
        "in_process_shim": {k: med(cold, k) for k in keys},
        "resident_first_call": {k: first[k] for k in keys},
        "resident_warm": {k: med(warm, k) for k in keys},
        "speedup_process_wall": med(cold, "process_wall_s") / med(warm, "process_wall_s"),
        "service_stats": {k: stats[k] for k in ("loads", "hits", "searches")},
    }))
    return 0


if __name__ == "__main__":
    sys.exit(main())
