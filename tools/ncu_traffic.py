"""Record dram__bytes_read.sum + dram__bytes_write.sum per launch of the scan kernel from a transposed `ncu --set full`
summary (tools/ncu_summary.py) into profiles/ncu_traffic.json, keyed by the hash of the kernel sources the capture was
taken from (bench.py drops the figure when csrc/scan_topk.cuh or common.cuh change afterwards).
usage: python tools/ncu_traffic.py profiles/r2_ncu_scan.csv 10Mx768_ip_f32_k10_nq1 [launch column, default last]"""
import csv, json, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    from bench import kernel_source_sha16

    src, workload = sys.argv[1], sys.argv[2]
    rows = {r[0]: r for r in csv.reader(open(src)) if r}
    col = int(sys.argv[3]) + 2 if len(sys.argv) > 3 else len(rows["Kernel Name"]) - 1
    total = sum(float(rows[m][col]) * UNIT[rows[m][1]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    out = ROOT / "profiles" / "ncu_traffic.json"
    try:
        j = json.loads(out.read_text())
    except Exception:
        j = {}
    sha = kernel_source_sha16()
    if j.get("kernel_source_sha16") != sha:
        j = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch of scan_topk_kernel, from `ncu --set full` captures",
             "kernel_source_sha16": sha, "workloads": {}, "source": ""}
    j["workloads"][workload] = int(round(total))
    j["source"] = f"ncu --set full, {Path(src).name}, kernel {rows['Kernel Name'][col][:60]}"
    out.write_text(json.dumps(j, indent=1) + "\n")
    print(workload, int(round(total)), "bytes;", rows["gpu__time_duration.sum"][col], rows["gpu__time_duration.sum"][1])


main()
