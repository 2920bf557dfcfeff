"""Write the Blackwell-instruction evidence of the built library under profiles/: for every kernel in
c99_vectordb_b200/_b200flat.so whose SASS holds tcgen05 / TMA / mbarrier / PDL instructions, the mangled name, its
register count and the count of each mnemonic family, followed by the first few matching SASS lines.
usage: python tools/sass_evidence.py  (needs cuobjdump; no GPU)"""
import collections, re, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "c99_vectordb_b200" / "_b200flat.so"
FAMILIES = [("UTCHMMA", "tcgen05.mma kind::f16"), ("UTCBAR", "tcgen05.commit"), ("LDTM", "tcgen05.ld"), ("UTCATOMSWS", "tcgen05.alloc"),
            ("UTMALDG", "cp.async.bulk.tensor (TMA tile load)"), ("UBLKCP", "cp.async.bulk (TMA bulk copy)"),
            ("UTMAPF", "prefetch.tensormap"), ("SYNCS", "mbarrier"), ("ACQBULK", "griddepcontrol.wait"), ("PREEXIT", "griddepcontrol.launch_dependents"),
            ("FMNMX3", "3-input max"), ("UCGABAR", "barrier.cluster")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], check=True, capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", str(LIB)], check=True, capture_output=True, text=True).stdout
    regs = {}
    for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+)", res):
        regs[m.group(1)] = int(m.group(2))
    out = {"scan": [], "gemm": [], "other": []}
    cur, lines = None, collections.defaultdict(list)
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        if cur and re.search(r"/\*[0-9a-f]{4}\*/", ln):
            lines[cur].append(ln.strip())
    demangle = subprocess.run(["cu++filt"] + list(lines), capture_output=True, text=True).stdout.splitlines()
    pretty = dict(zip(lines, demangle)) if len(demangle) == len(lines) else {}
    for fn, body in lines.items():
        counts = {fam: sum(1 for l in body if re.search(r"\b" + fam, l)) for fam, _ in FAMILIES}
        if not any(counts[f] for f in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP")):
            continue
        key = "scan" if "scan_topk" in fn else "gemm" if ("gemm_topk" in fn or "gemm_rows_topk" in fn) else "other"
        txt = [f"== {pretty.get(fn, fn)[:200]}", f"   mangled {fn[:160]}", f"   registers {regs.get(fn, '?')}, SASS instructions {len(body)}",
               "   " + ", ".join(f"{fam} x{c}" for fam, c in counts.items() if c)]
        shown = set()
        for l in body:
            for fam, _ in FAMILIES[:7]:
                if re.search(r"\b" + fam, l) and fam not in shown:
                    shown.add(fam)
                    txt.append("     " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", l))
        out[key].append("\n".join(txt))
    legend = "\n".join(f"  {fam:10s} = {what}" for fam, what in FAMILIES)
    for key, fname in (("scan", "sass_scan_topk.txt"), ("gemm", "sass_gemm_topk.txt"), ("other", "sass_other_kernels.txt")):
        if not out[key]:
            continue
        head = (f"cuobjdump -sass {LIB.relative_to(ROOT)} (sm_100a), kernels holding tcgen05 / TMA instructions; "
                f"written by tools/sass_evidence.py\nSASS mnemonic -> PTX:\n{legend}\n\n")
        (ROOT / "profiles" / fname).write_text(head + "\n\n".join(out[key]) + "\n")
        print(fname, len(out[key]), "kernels")


main()
