"""The built library really carries the Blackwell instructions the design claims, and the tensor-core kernels issue
them warp-uniformly: no ELECT / R2UR / BRA.U.ANY waterfall around UTCHMMA or UTMALDG (DESIGN.md 7.6 — a divergent
`if (lane == 0)` issue region cost the batched path a quarter of the tensor pipe).  cuobjdump only, no GPU."""
import shutil
import subprocess
from pathlib import Path

import pytest

from conftest import ROOT

LIB = ROOT / "c99_vectordb_b200" / "_b200flat.so"
KERNELS = {
    # mangled name: instructions that must be there
    "_Z16gemm_topk_kernelILi2ELb0ELi1EEv14CUtensorMap_stS0_10GemmParams": ("UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "LDTM.x32", "UTCBAR.2CTA.MULTICAST"),
    "_Z16gemm_topk_kernelILi2ELb0ELi0EEv14CUtensorMap_stS0_10GemmParams": ("UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "LDTM.x32"),
    "_Z21gemm_rows_topk_kernelILi1ELb0EEv14CUtensorMap_stS0_14GemmRowsParams": ("UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "LDTM.x32"),
}


def _cuobjdump():
    fallback = Path("/usr/local/cuda/bin/cuobjdump")
    return shutil.which("cuobjdump") or (str(fallback) if fallback.exists() else None)


@pytest.mark.parametrize("fn", sorted(KERNELS))
def test_tensor_core_kernels_issue_warp_uniformly(fn):
    exe = _cuobjdump()
    if exe is None or not LIB.exists():
        pytest.skip("cuobjdump or the built library is missing")
    sass = subprocess.run([exe, "-sass", "-fun", fn, str(LIB)], capture_output=True, text=True, timeout=300).stdout
    assert "Function : " + fn in sass, "kernel not found in the built library"
    for ins in KERNELS[fn]:
        assert ins in sass, f"{ins} missing from {fn}"
    assert "BRA.U.ANY" not in sass, "a tcgen05 / TMA instruction is issued from a divergent region again (R2UR waterfall)"
