"""Real multi-GPU run of the row-sharded index (NCCL all-gather + K4 merge kernel).  Needs >= 2
GPUs (gpurun --gpus 2); on a 1-GPU box the test is skipped — the host logic is covered on CPU by
tests/test_sharded_gloo_cpu.py."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    from oracle import oracle
    from c99_vectordb_b200.sharded import ShardedIndexFlat

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        for metric, n, d, k, nq in ((0, 200_003, 384, 10, 1), (1, 50_000, 768, 100, 5), (1, 7, 16, 5, 2)):
            base = oracle.synth_rows(n // 2, d, 9)
            db = np.concatenate([base, base, oracle.synth_rows(n - 2 * (n // 2), d, 10)])  # cross-shard exact ties
            ids = np.arange(n, dtype=np.int64) * 3 + 5
            q = oracle.synth_rows(nq, d, 8)
            idx = ShardedIndexFlat(d, metric)
            idx.add_with_ids(db, ids)
            D, I = idx.search(q, k)
            Dw, Iw = oracle.search(metric, db, q, k, ids=ids, order=oracle.ORDER_DEVICE)
            np.testing.assert_array_equal(I, Iw)
            np.testing.assert_array_equal(D, Dw)
        # device-generated shards: global ids are global row positions
        idx = ShardedIndexFlat(384, 1)
        idx.add_synthetic(300_000, 1234)
        q = oracle.synth_rows(3, 384, 5678)
        D, I = idx.search(q, 10)
        Dw, Iw = oracle.search(1, oracle.synth_rows(300_000, 384, 1234), q, 10, order=oracle.ORDER_DEVICE)
        np.testing.assert_array_equal(I, Iw)
        np.testing.assert_array_equal(D, Dw)
        # fused exchange: the scan kernel stores into the peers' buffers over NVLink and merges itself
        assert idx.enable_fused_exchange()
        for nq, k in ((1, 10), (1, 100), (3, 10), (8, 100), (11, 7)):
            q = oracle.synth_rows(nq, 384, 999 + nq)
            Df, If = idx.search(q, k)
            if nq > 1:  # the fused kernel itself also handles query blocks (forced through the C ABI)
                import ctypes as C
                import torch
                from c99_vectordb_b200 import _cabi
                qt = torch.from_numpy(q).cuda()
                Dt = torch.empty((nq, k), dtype=torch.float32, device="cuda")
                It = torch.empty((nq, k), dtype=torch.int64, device="cuda")
                _cabi.check(_cabi.load().b200_index_search_exchange_dev(idx.local.index._h, qt.data_ptr(), nq, k, Dt.data_ptr(),
                                                                        It.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream or 1)))
                torch.cuda.synchronize()
                np.testing.assert_array_equal(It.cpu().numpy(), If)
                np.testing.assert_array_equal(Dt.cpu().numpy(), Df)
            Dw, Iw = oracle.search(1, oracle.synth_rows(300_000, 384, 1234), q, k, order=oracle.ORDER_DEVICE)
            np.testing.assert_array_equal(If, Iw)
            np.testing.assert_array_equal(Df, Dw)
        # a large batch: every shard runs the tensor-core path (K3), results merged through NCCL + K4
        idx.local.index.set_option("gemm_min_nq", 32)
        qb = oracle.synth_rows(70, 384, 4242)
        Db, Ib = idx.search(qb, 10)
        assert idx.local.index.get_option("stat_gemm_used") == 1
        Dw2, Iw2 = oracle.search(1, oracle.synth_rows(300_000, 384, 1234), qb, 10, order=oracle.ORDER_DEVICE)
        np.testing.assert_array_equal(Ib, Iw2)
        np.testing.assert_array_equal(Db, Dw2)
        # starved thresholds: the certificate taken after the merge rejects queries, the widened retry and the exact
        # scan finish them — the answer stays exact and every rank takes the same decisions
        idx.local.index.set_option("gemm_min_nq", 2)
        idx.local.index.set_option("gemm_min_rows", 4096)
        idx.local.index.set_option("gemm_emit_factor", 2)
        qs = oracle.synth_rows(48, 384, 777)
        Ds, Is = idx.search(qs, 40)
        Dw3, Iw3 = oracle.search(1, oracle.synth_rows(300_000, 384, 1234), qs, 40, order=oracle.ORDER_DEVICE)
        np.testing.assert_array_equal(Is, Iw3)
        np.testing.assert_array_equal(Ds, Dw3)
        if idx.local.index.get_option("stat_gemm_used") == 1:
            assert idx.last_batch_uncertified is not None and idx.last_batch_uncertified > 0
        idx.local.index.set_option("gemm_emit_factor", 8)
        q1 = oracle.synth_rows(1, 384, 31337)
        Dw1, Iw1 = oracle.search(1, oracle.synth_rows(300_000, 384, 1234), q1, 10, order=oracle.ORDER_DEVICE)
        for rep in range(20):  # back-to-back fused searches exercise the double-buffered slots
            Df2, If2 = idx.search(q1, 10)
            np.testing.assert_array_equal(If2, Iw1)
        np.save(os.path.join(out_dir, f"ok_{rank}.npy"), I)
    finally:
        dist.destroy_process_group()


def test_sharded_nccl_matches_unsharded_oracle(tmp_path, gpu):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    outs = [np.load(tmp_path / f"ok_{r}.npy") for r in range(world)]
    for o in outs[1:]:
        np.testing.assert_array_equal(o, outs[0])
