"""include/b200_flat.h is plain C99 and examples/memo_recall.c builds against the shared library with
-Wpedantic -Werror.  Without a device the example still parses the .memo headers (host code) and then
fails loudly at load — there is no CPU fallback."""
import struct
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def build_example(tmp_path) -> Path:
    from c99_vectordb_b200 import _cabi

    _cabi.load()  # makes sure the library exists
    exe = tmp_path / "memo_recall"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Wpedantic", "-Werror", "-O2", "-I", str(ROOT / "include"),
                    str(ROOT / "examples" / "memo_recall.c"), "-o", str(exe), str(_cabi.LIB_PATH),
                    f"-Wl,-rpath,{_cabi.LIB_PATH.parent}", "-lm"], check=True)
    return exe


def write_memo(path, x, ids, metric=1):
    n, d = x.shape
    hdr = struct.pack("<iqqqBi", d, n, 1 << 20, 1 << 20, 1, metric)
    path.write_bytes(b"IxM2" + hdr + (b"IxFI" if metric == 0 else b"IxF2") + hdr + struct.pack("<Q", n * d) + x.tobytes()
                     + struct.pack("<Q", n) + ids.tobytes())


def test_header_is_c99_and_example_fails_loudly_without_a_device(tmp_path):
    exe = build_example(tmp_path)
    x = np.arange(24, dtype=np.float32).reshape(3, 8)
    write_memo(tmp_path / "t.memo", x, np.array([5, 6, 7], dtype=np.int64))
    r = subprocess.run([str(exe), str(tmp_path / "t.memo"), "2"], capture_output=True, text=True)
    assert "id-mapped index, d=8, L2, 3 rows" in r.stdout
    from c99_vectordb_b200 import _cabi

    if _cabi.device_count_or_zero() == 0:
        assert r.returncode == 1 and r.stderr.startswith("load:")
    r = subprocess.run([str(exe), str(tmp_path / "missing.memo")], capture_output=True, text=True)
    assert r.returncode == 1 and "could not open" in r.stderr
