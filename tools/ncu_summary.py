"""Turn an `ncu --set full` report into the transposed CSV kept under profiles/ (one row per metric, one
column per captured launch), keeping the metrics the roofline discussion uses.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.csv"""
import csv, io, re, subprocess, sys

KEEP = re.compile(r"^(Kernel Name|gpu__time_duration|dram__bytes|dram__throughput|gpu__dram_throughput|sm__cycles_active|"
                  r"sm__throughput|sm__warps_active|sm__inst_executed\.(sum|avg\.per_cycle)|smsp__cycles_active|launch__|"
                  r"lts__t_sector_hit_rate|lts__t_bytes|l1tex__t_bytes|sm__pipe_tensor|sm__inst_executed_pipe_(tensor|uniform|lsu)|"
                  r"smsp__warp_issue_stalled.*_per_warp_active|FBSP|TPC)")

def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, launches = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(launches))])
        for c, name in enumerate(names):
            if KEEP.match(name):
                w.writerow([name, units[c]] + [l[c] if c < len(l) else "" for l in launches])
    print(f"{out}: {len(launches)} launches")

main()
