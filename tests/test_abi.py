"""The C-ABI library loads, exports every symbol include/ibx.h declares, and fails loudly without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(ib):
    lib = ctypes.CDLL(os.path.join(ROOT, "immersedboundary.jl_b200", "libibx.so"))
    protos = ib._lib.parse_header()
    assert len(protos) > 100
    missing = [n for n in protos if not hasattr(lib, n)]
    assert not missing, missing
    assert ib._lib.lib.ibx_version().decode().startswith("ibx-b200")


def test_built_for_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", os.path.join(ROOT, "immersedboundary.jl_b200", "libibx.so")],
                         capture_output=True, text=True).stdout
    archs = {tok for line in out.splitlines() for tok in line.replace(".", " ").split() if tok.startswith("sm_")}
    assert archs == {"sm_100a"}, archs


def test_error_reporting_and_argument_checks(ib):
    with pytest.raises(ib.IbxError, match="nd must be 2 or 3"):
        ib.Stereolitography(np.zeros((3, 4)))
    with pytest.raises(ib.IbxError, match="cannot open"):
        ib.Stereolitography("/nonexistent/file.stl")
    s = ib.Stereolitography(np.array([[0.0, 0.0], [1.0, 0.0]]))
    msh = ib.Mesh([0.0, 0.0], [1.0, 1.0], ("w", s, np.float32(0.1)))
    with pytest.raises(ib.IbxError, match="dimension out of range"):
        ib.Domain(msh, hypercube_families=[("bad", [(5, True)])], upload=False)


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="this check is for machines without a GPU")
def test_no_cpu_fallback(ib):
    import immersedboundary_jl_b200.domain as d
    d._ctx = None
    with pytest.raises(ib.IbxError, match="no CPU fallback"):
        ib.context()
    s = ib.Stereolitography(np.array([[0.0, 0.0], [1.0, 0.0]]))
    msh = ib.Mesh([0.0, 0.0], [1.0, 1.0], ("w", s, np.float32(0.1)))
    with pytest.raises(ib.IbxError):
        ib.Domain(msh)  # upload=True needs the device


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "immersedboundary.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh", ".jl")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
