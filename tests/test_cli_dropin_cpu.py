"""The CLI drop-in harness itself, on CPU: the unmodified reference CLI over the oracle-backed `faiss` stub in both
summation orders.  The two transcripts must agree up to near-tie order and 1e-5 relative score differences — the
canonicalisation the GPU test (tests/test_cli_dropin_gpu.py) relies on — and the session must exercise what
SURVEY.md App. B observed (rebuild on overwrite, unknown id, corrupt .memo, k clamping, clean)."""
import re

import pytest

from cli_dropin_script import ROOT, find_reference_cli, run_session

SCORE = re.compile(r"^  \[(\d+)\] Score: (-?\d+\.\d+) \|$")


def make_stub_dir(tmp_path):
    d = tmp_path / "stubfaiss"
    d.mkdir()
    (d / "faiss.py").write_text("from stub_faiss_oracle import *  # noqa\nfrom stub_faiss_oracle import IndexIDMap2, IndexHNSWFlat, read_index, write_index, vector_to_array  # noqa\n")
    return d


def parse_results(stdout: str):
    """Non-YAML recall output -> [(id, score, body lines)]."""
    res, cur = [], None
    for ln in stdout.splitlines():
        m = SCORE.match(ln)
        if m:
            cur = [int(m.group(1)), float(m.group(2)), []]
            res.append(cur)
        elif cur is not None and ln.startswith("      "):
            cur[2].append(ln)
    return res


def assert_transcripts_equivalent(a, b, eps=1.2e-4):
    """Same labels / return codes / error lines; recall listings equal up to order inside groups of scores closer
    than eps (4 printed decimals) and score differences below eps; YAML results likewise (1e-5 relative)."""
    import yaml

    assert [x[0] for x in a] == [x[0] for x in b]
    for (label, rc_a, out_a, err_a, tb_a), (_, rc_b, out_b, err_b, tb_b) in zip(a, b):
        assert rc_a == rc_b, (label, rc_a, rc_b, tb_a, tb_b)
        assert err_a == err_b, (label, err_a, err_b)
        if out_a == out_b:
            continue
        if out_a.startswith("results:") and out_b.startswith("results:"):
            ra, rb = yaml.safe_load(out_a)["results"], yaml.safe_load(out_b)["results"]
            assert len(ra) == len(rb), label
            pa = [(r["id"], r["score"], r["body"]) for r in ra]
            pb = [(r["id"], r["score"], r["body"]) for r in rb]
            tol = lambda s: 1e-5 * max(1.0, abs(s))
        else:
            pa, pb = parse_results(out_a), parse_results(out_b)
            assert out_a.splitlines()[0] == out_b.splitlines()[0], label  # "Top k results:"
            assert len(pa) == len(pb), label
            tol = lambda s: eps
        start = 0
        for i in range(1, len(pa) + 1):
            if i == len(pa) or abs(pa[i][1] - pa[i - 1][1]) > 2 * tol(pa[i][1]):
                if i == len(pa):
                    # the last group may be cut by k in the middle of a near tie: which of the tied records made it
                    # is summation-order dependent, their scores are not
                    for x, y in zip(sorted(r[1] for r in pa[start:i]), sorted(r[1] for r in pb[start:i])):
                        assert abs(x - y) <= 2 * tol(x), (label, x, y)
                else:
                    assert sorted((r[0], str(r[2])) for r in pa[start:i]) == sorted((r[0], str(r[2])) for r in pb[start:i]), (label, start, i)
                start = i
        sb = {r[0]: r for r in pb}
        for x in pa:
            if x[0] in sb:
                assert abs(x[1] - sb[x[0]][1]) <= tol(x[1]) and str(x[2]) == str(sb[x[0]][2]), (label, x, sb[x[0]])


def test_reference_cli_over_oracle_stub(tmp_path):
    cli = find_reference_cli()
    if cli is None:
        pytest.skip("the reference CLI is neither under /root/reference nor installed in baseline/_ref")
    stub = make_stub_dir(tmp_path)
    extra = [ROOT / "tests", ROOT]
    simd = run_session(cli, stub, tmp_path / "w_simd", extra_path=extra)
    dev = run_session(cli, stub, tmp_path / "w_dev", {"STUB_FAISS_ORDER": "device"}, extra_path=extra)
    t = {x[0]: x for x in simd}
    assert all(x[4] == "" for x in simd), [x for x in simd if x[4]]
    assert t["save-1"][1] == 0 and t["save-1"][2].count("Memorized:") == 5
    assert "Top 3 results:" in t["recall-k3"][2] and "peanuts" in t["recall-k3"][2]
    assert t["recall-yaml"][2].startswith("results:")
    assert "Rebuilt index with 6 vectors" in t["save-overwrite"][3]  # -v diagnostics go to stderr
    assert "correct-horse-battery" in t["recall-after-overwrite"][2] and "hunter2" not in t["recall-after-overwrite"][2]
    assert t["save-bad-id"][1] == 1 and "override id 99 does not exist" in t["save-bad-id"][3]
    assert t["recall-many"][2].count("Score:") == 25
    assert "Wrote index" in t["reindex"][2]
    assert t["recall-corrupt-memo"][2].strip() == "Top 2 results:"  # corrupt index = empty index (memo_cli.py:254-257)
    assert t["recall-repaired"][2].count("Score:") == 2
    assert t["recall-k0"][2].startswith("Top 1 results:") and t["recall-k1000"][2].startswith("Top 100 results:")
    assert t["recall-kabc"][1] == 1 and t["recall-noquery"][1] == 1 and t["no-f"][1] == 1
    assert "almonds" in t["subdir-recall"][2]
    assert "Cleared memory database" in t["clean"][2] and "already empty" in t["clean-again"][2]
    assert t["recall-empty"][2].strip() == "Top 2 results:"
    assert_transcripts_equivalent(simd, dev)
