"""The CLI session both drop-in tests replay through the UNMODIFIED reference CLI (memo_cli.py:883-949):
save (append, overwrite-by-id -> rebuild, unknown id), recall (-k, --yaml, --filter), reindex, a corrupt .memo,
clean, and the argument edge cases of SURVEY.md App. B.  Test infrastructure: `run_session` runs it in a scratch
directory with a given `faiss` module directory on PYTHONPATH and returns the transcript."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

NOTES_1 = """\
---
body: My daughter's name is Sarah and she is allergic to peanuts.
metadata: {kind: family, priority: 3}
---
body: The wifi password at the office is hunter2
metadata: {kind: work, priority: 1}
---
body: Remember to rotate the API keys every 90 days
metadata: {kind: work, priority: 2}
---
body: |-
  Shopping list:
  peanuts are NOT allowed, buy almonds
---
body: flat index search on B200 with fused top-k
metadata: {kind: work, priority: 5, tags: [gpu, search]}
"""
NOTES_2 = """\
---
id: 1
body: The wifi password at the office changed to correct-horse-battery
metadata: {kind: work, priority: 4}
---
body: Sarah's school starts at 8:15
metadata: {kind: family, priority: 2}
"""
NOTES_BAD_ID = """\
---
id: 99
body: this id does not exist
"""
NOTES_MANY = "".join(
    f"---\nbody: note number {i} about topic {i % 7} and subject {i % 11} with words alpha{i % 5} beta{i % 3}\n"
    f"metadata: {{priority: {i % 6}, kind: k{i % 4}}}\n" for i in range(120))

# (label, argv after `memo_cli.py`, optional hook run before the step)
STEPS = [
    ("save-1", ["-f", "db", "save", "notes1.yaml"]),
    ("recall-k3", ["-f", "db", "recall", "-k", "3", "peanuts", "allergies"]),
    ("recall-yaml", ["-f", "db", "recall", "-k", "5", "--yaml", "wifi", "password"]),
    ("recall-filter", ["-f", "db", "recall", "-k", "5", "--filter", "{priority: {$gte: 2}}", "password", "keys"]),
    ("save-overwrite", ["-f", "db", "-v", "save", "notes2.yaml"]),
    ("recall-after-overwrite", ["-f", "db", "recall", "-k", "10", "wifi", "password", "office"]),
    ("save-bad-id", ["-f", "db", "save", "bad_id.yaml"]),
    ("save-many", ["-f", "db", "save", "many.yaml"]),
    ("recall-many", ["-f", "db", "recall", "-k", "25", "topic", "3", "subject", "alpha2"]),
    ("recall-many-yaml-filter", ["-f", "db", "recall", "-k", "40", "--yaml", "--filter", "{kind: k2, priority: {$lte: 3}}", "note", "beta1"]),
    ("reindex", ["-f", "db", "reindex"]),
    ("recall-after-reindex", ["-f", "db", "recall", "-k", "7", "rotate", "keys"]),
    ("recall-corrupt-memo", ["-f", "db", "recall", "-k", "2", "peanuts"], "corrupt"),
    ("reindex-repairs", ["-f", "db", "-v", "reindex"]),
    ("recall-repaired", ["-f", "db", "recall", "-k", "2", "peanuts"]),
    ("recall-k0", ["-f", "db", "recall", "-k", "0", "peanuts"]),
    ("recall-k1000", ["-f", "db", "recall", "-k", "1000", "note"]),
    ("recall-kabc", ["-f", "db", "recall", "-k", "abc", "peanuts"]),
    ("recall-noquery", ["-f", "db", "recall"]),
    ("subdir-save", ["-f", "sub/dir/db2", "save", "notes1.yaml"]),
    ("subdir-recall", ["-f", "sub/dir/db2", "recall", "almonds"]),
    ("clean", ["-f", "db", "clean"]),
    ("recall-empty", ["-f", "db", "recall", "anything"]),
    ("clean-again", ["-f", "db", "clean"]),
    ("no-f", ["recall", "x"]),
]


def find_reference_cli() -> Path | None:
    """The unmodified reference CLI: /root/reference in the build container, baseline/_ref (installed from it by
    __graft_entry__.build(), git-ignored, shipped to the GPU box) elsewhere."""
    for cand in (Path("/root/reference/memo_cli.py"), ROOT / "baseline" / "_ref" / "memo_cli.py"):
        if cand.exists():
            return cand
    return None


def run_session(cli: Path, faiss_dir: Path, work: Path, extra_env: dict | None = None, extra_path: list | None = None):
    """Replay STEPS; returns [(label, rc, stdout, stderr_tail)].  Absolute scratch paths are replaced by <WORK>."""
    work.mkdir(parents=True, exist_ok=True)
    (work / "notes1.yaml").write_text(NOTES_1)
    (work / "notes2.yaml").write_text(NOTES_2)
    (work / "bad_id.yaml").write_text(NOTES_BAD_ID)
    (work / "many.yaml").write_text(NOTES_MANY)
    env = dict(os.environ)
    env["PYTHONHASHSEED"] = "0"
    env["PYTHONPATH"] = os.pathsep.join([str(faiss_dir)] + [str(p) for p in (extra_path or [])])
    env.update(extra_env or {})
    out = []
    for step in STEPS:
        label, argv = step[0], step[1]
        if len(step) > 2 and step[2] == "corrupt":
            (work / "db.memo").write_bytes(b"this is not an index file\x00\x01\x02" * 10)
        r = subprocess.run([sys.executable, str(cli), *argv], cwd=work, env=env, capture_output=True, text=True, timeout=300)
        clean = lambda t: t.replace(str(work.resolve()), "<WORK>").replace(str(work), "<WORK>")
        # the CLI's own diagnostics (errors and -v lines, memo_cli.py vlog); other stderr noise is not part of the contract
        err_lines = [ln for ln in clean(r.stderr).splitlines() if ln.startswith(("Error", "Rebuilt", "Loaded", "[memo]"))]
        out.append((label, r.returncode, clean(r.stdout), "\n".join(err_lines), r.stderr[-2000:] if r.returncode not in (0, 1) else ""))
    return out
