"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports exactly the
entry points include/pandrs_b200.h declares (no compute calls: there is no GPU here)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pandrs_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pdrs_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def native():
    import pandrs_b200._native as n
    if not os.path.exists(n.LIB_PATH):
        n.build()
    return n


def test_header_and_binding_agree(native):
    assert _declared() == sorted(native.SIGNATURES)


def test_library_exports_every_declared_symbol(native):
    L = native.lib()
    for name in _declared():
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", native.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (pdrs_\w+)", out))
    assert exported == set(_declared())
    assert L.pdrs_abi_version() == 1


def test_header_compiles_as_plain_c(tmp_path):
    # the boundary is C: plain pointers and sizes, no C++ / torch types
    src = tmp_path / "t.c"
    src.write_text('#include "pandrs_b200.h"\nint main(void) { pdrs_col c; (void)c; return sizeof(pdrs_options) == 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.dirname(HEADER), "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_struct_layouts_match_ctypes(native, tmp_path):
    import ctypes as C
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include "pandrs_b200.h"\nint main(void) { printf("%zu %zu %zu %zu\\n", sizeof(pdrs_col), sizeof(pdrs_agg), sizeof(pdrs_options), sizeof(pdrs_stats)); return 0; }\n')
    exe = tmp_path / "s"
    subprocess.run(["gcc", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(native.PdrsCol), C.sizeof(native.PdrsAgg), C.sizeof(native.PdrsOptions), C.sizeof(native.PdrsStats)]


def test_no_device_means_a_loud_error_not_a_fallback(native):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import pandrs_b200 as pb
    with pytest.raises(pb.PandrsError) as e:
        pb.Context(device=0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    # oracle/ is test infrastructure: nothing under pandrs_b200/ may reference it
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pandrs_b200")):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)
                assert "libpandrs_oracle" not in text and "orc_" not in text, os.path.join(dirpath, f)
