"""CPU-side checks of the drop-in boundary: the C-ABI library builds for
sm_100a, loads, and exports every symbol include/casm_monte_gpu.h declares;
without a GPU the product fails loudly instead of falling back to the CPU."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT, have_cuda


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "casm_monte_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmg_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from casmcode_monte_b200 import _capi

    decl = declared_symbols()
    assert len(decl) >= 45
    assert sorted(_capi.SIGNATURES) == decl


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g

    g.build()
    from casmcode_monte_b200 import _capi

    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert _capi.load().cmg_abi_version() == 1


def test_library_holds_sm100a_code_only():
    from casmcode_monte_b200 import _capi

    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "casm_monte_gpu.h"\nint main(void){ return CMG_ABI_VERSION - 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


@pytest.mark.skipif(have_cuda(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_a_device():
    import casmcode_monte_b200 as m

    with pytest.raises(m.CmgError) as ei:
        m.IsingLatticeGPU([8, 8])
    assert ei.value.code == -2  # CMG_ENODEVICE


def test_product_does_not_reference_the_oracle():
    # the oracle is test infrastructure: nothing under the package or include/ may use it
    pat = re.compile(r"oracle", re.I)
    for base in ("casmcode_monte_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cpp")):
                    text = open(os.path.join(dirpath, f)).read()
                    hits = [ln for ln in text.splitlines() if pat.search(ln) and "build_oracle" not in ln and "CPU oracle is test" not in ln and "make" not in ln]
                    assert not hits, (f, hits[:3])
