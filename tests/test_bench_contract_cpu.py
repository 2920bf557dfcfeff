"""bench.py contract pieces that need no GPU: the reference arm prints one JSON line with the keys
the driver reads, and the scan-pass arithmetic mirrors the native query blocking."""
import json
import subprocess
import sys

from conftest import ROOT

sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def test_reference_arm_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--workload", "10kx384_ip_f32_k10_nq100"], capture_output=True, text=True, check=True, timeout=300).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and line["gpu_launches"] == 0 and line["vs_baseline"] is None


def test_reference_arm_non_zero_ranks_exit_quietly():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, check=True, timeout=60, env={"RANK": "1", "PATH": "/usr/bin:/bin"}).stdout
    assert out.strip() == ""


def test_scan_passes_mirror_query_blocking():
    assert [bench.scan_passes(n) for n in (1, 2, 3, 4, 5, 8, 9, 13, 100)] == [1, 1, 1, 1, 2, 1, 2, 3, 13]
    assert bench.DEFAULT_WORKLOAD in bench.WORKLOADS and bench.WORKLOADS[bench.DEFAULT_WORKLOAD][:2] == (10_000_000, 768)


def test_l2_policy_says_what_fits():
    """The `config.l2_policy` string has to say which case a workload is in (timing rules of the bench contract)."""
    assert "no flush needed" in bench.l2_policy(bench.DEFAULT_WORKLOAD, 1) and ">> 126 MB" in bench.l2_policy(bench.DEFAULT_WORKLOAD, 8)
    small = bench.l2_policy("10kx384_ip_f32_k10_nq100", 1)
    assert "fits the 126 MB L2" in small and "no bandwidth claim" in small
    mid = bench.l2_policy("1Mx768_cos_f32_k10_nq1", 8)  # 384 MB shard: larger than L2, streamed front to back
    assert "> 126 MB L2" in mid and "fits" not in mid
    assert bench.workload_config(bench.DEFAULT_WORKLOAD, 2)["l2_policy"] == bench.l2_policy(bench.DEFAULT_WORKLOAD, 2)
    # config 0 is the CPU-runnable case: part of the one-GPU line only
    assert bench.OTHER_CONFIGS[0] == "10kx384_ip_f32_k10_nq100" and bench.WORKLOADS[bench.OTHER_CONFIGS[0]][0] < 1_000_000
