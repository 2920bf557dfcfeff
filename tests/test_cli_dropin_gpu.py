"""CLI drop-in on the GPU (SURVEY.md §4 tier 3, memo_cli.py:883-949): the UNMODIFIED reference CLI runs a whole
session — save (append, overwrite-by-id -> rebuild), reindex, recall -k / --filter / --yaml, a corrupt .memo, clean —
with c99_vectordb_b200/shim (and the resident-service shim) on PYTHONPATH as its `faiss` module, and its stdout,
return codes and error lines are compared with the same session over the oracle-backed stub:
 * stub in the kernels' summation order: transcripts identical byte for byte (scores included);
 * stub in the faiss-like SIMD order: identical up to near-tie order and 1e-5 relative scores.
The reference CLI comes from /root/reference, or from baseline/_ref where __graft_entry__.build() installed it
(git-ignored, shipped with the repo snapshot to the GPU box)."""
import os
import subprocess
import sys
import time

import pytest

from cli_dropin_script import ROOT, find_reference_cli, run_session
from test_cli_dropin_cpu import assert_transcripts_equivalent, make_stub_dir

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cli(gpu):
    c = find_reference_cli()
    if c is None:
        pytest.skip("the reference CLI is neither under /root/reference nor installed in baseline/_ref")
    return c


def test_reference_cli_on_gpu_shim(cli, tmp_path):
    stub = make_stub_dir(tmp_path)
    extra = [ROOT / "tests", ROOT]
    gpu_run = run_session(cli, ROOT / "c99_vectordb_b200" / "shim", tmp_path / "w_gpu")
    assert all(x[4] == "" for x in gpu_run), [x for x in gpu_run if x[4]]
    dev = run_session(cli, stub, tmp_path / "w_dev", {"STUB_FAISS_ORDER": "device"}, extra_path=extra)
    simd = run_session(cli, stub, tmp_path / "w_simd", extra_path=extra)
    for g, o in zip(gpu_run, dev):
        assert g[:4] == o[:4], (g[0], g, o)  # label, rc, stdout, error lines: byte for byte
    assert_transcripts_equivalent(gpu_run, simd)
    # the index files the GPU shim wrote are what the session ended with (clean removed db; sub/dir/db2 remains)
    assert (tmp_path / "w_gpu" / "sub" / "dir" / "db2.memo").stat().st_size > 5 * 384 * 4


def test_reference_cli_on_resident_shim(cli, tmp_path):
    """The same session against the resident service (shim_resident): one long-lived process owns the device
    indexes, every CLI invocation attaches over a socket."""
    sock = tmp_path / "svc.sock"
    env = dict(os.environ, B200_RESIDENT_SOCKET=str(sock))
    svc = subprocess.Popen([sys.executable, "-m", "c99_vectordb_b200.resident", "serve", "--socket", str(sock)], cwd=ROOT, env=env,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    try:
        for _ in range(600):
            if sock.exists():
                break
            if svc.poll() is not None:
                pytest.fail("resident service exited: " + svc.stdout.read().decode()[-2000:])
            time.sleep(0.1)
        assert sock.exists(), "resident service did not come up"
        res = run_session(cli, ROOT / "c99_vectordb_b200" / "shim_resident", tmp_path / "w_res", {"B200_RESIDENT_SOCKET": str(sock)})
        assert all(x[4] == "" for x in res), [x for x in res if x[4]]
        stub = make_stub_dir(tmp_path)
        dev = run_session(cli, stub, tmp_path / "w_dev", {"STUB_FAISS_ORDER": "device"}, extra_path=[ROOT / "tests", ROOT])
        for g, o in zip(res, dev):
            assert g[:4] == o[:4], (g[0], g, o)
    finally:
        svc.terminate()
        try:
            svc.wait(timeout=10)
        except Exception:
            svc.kill()
