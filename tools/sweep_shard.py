"""Fixed-cost sensitivities of the single-query scan at the 8-GPU shard size of the headline database
(1.25M x 768 fp32 = 3.84 GB, ideal 0.519 ms at 7.4 TB/s): claim run length, warps, stages, fused tail."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import c99_vectordb_b200 as m
from c99_vectordb_b200 import _cabi
import ctypes as C

n, d, k = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000, 768, 10
idx = m.IndexFlat(d, 0)
idx.add_synthetic(n, 1234)
q = torch.empty((1, d), dtype=torch.float32, device="cuda")
_cabi.check(_cabi.load().b200_synth_rows_dev(q.data_ptr(), 1, d, 5678, 0, 0, C.c_void_p(1)))
D = torch.empty((1, k), dtype=torch.float32, device="cuda"); I = torch.empty((1, k), dtype=torch.int64, device="cuda")
defaults = dict(scan_variant=0, scan_warps=16, scan_stages=0, scan_tile_rows=0, scan_claim_chunk=0, scan_fused_tail=-1, scan_dynamic_tiles=-1)
configs = [dict(), dict(scan_claim_chunk=1), dict(scan_claim_chunk=2), dict(scan_claim_chunk=8), dict(scan_claim_chunk=16),
           dict(scan_warps=6), dict(scan_warps=8), dict(scan_warps=12), dict(scan_stages=2), dict(scan_stages=3),
           dict(scan_tile_rows=4), dict(scan_tile_rows=16), dict(scan_fused_tail=0), dict(scan_dynamic_tiles=0)]
res = {i: [] for i in range(len(configs))}
for rnd in range(3):
    for ci, cfg in enumerate(configs):
        for kname, v in {**defaults, **cfg}.items():
            idx.set_option(kname, v)
        for _ in range(3): idx.search_device(q, k, D=D, I=I)
        torch.cuda.synchronize()
        ts = []
        for _ in range(30):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); idx.search_device(q, k, D=D, I=I); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
        res[ci].append(sorted(ts)[15])
ideal = n * d * 4 / 7.4e12 * 1e6
for ci, cfg in enumerate(configs):
    r = sorted(res[ci])
    print(json.dumps(dict(n=n, cfg=cfg, p50_us=round(r[1], 1), rounds=[round(x, 1) for x in r], over_ideal_us=round(r[1] - ideal, 1))), flush=True)
