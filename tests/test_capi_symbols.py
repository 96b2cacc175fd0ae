"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly what include/dvm_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dvm_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dvm_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from dv_matcher_b200 import build
    return build.build()


def test_header_symbols_exported(lib_path):
    names = _declared()
    assert len(names) >= 20
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in dvm_b200.h but not exported: {missing}"
    extra = sorted(n for n in exported if n.startswith("dvm_") and n not in names)
    assert not extra, f"exported but not declared in dvm_b200.h: {extra}"


def test_ctypes_binding_covers_header(lib_path):
    from dv_matcher_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.dvm_version() == 100
    # pure host-side calls only (no GPU here): workspace sizing and argument validation
    assert lib.dvm_softmap_workspace_bytes(1, 5000, 5000, 128, 0) > 0
    assert lib.dvm_softmap_workspace_bytes(0, 5000, 5000, 128, 0) == 0
    assert lib.dvm_fps_workspace_bytes(2, 50000) >= 2 * 50000 * 4
    rc = lib.dvm_softmap_fwd(None, None, None, 1, 8, 8, 8, 0, 1.0, 10, 1, 0, None, None, None, None, None, None, None, None, None, 0, None)
    assert rc == -1 and b"non-null" in lib.dvm_last_error_string()
    rc = lib.dvm_knn3(None, None, 1, 8, 8, 3, 0, None, None, None, None, None, 0, None)
    assert rc == -1
    assert lib.dvm_knn3_workspace_bytes(2, 5000, 5000) > 0 and lib.dvm_knn3_workspace_bytes(2, 5000, 500) == 0
    assert lib.dvm_chamfer_workspace_bytes(1, 4995, 2200) == lib.dvm_knn3_workspace_bytes(1, 2200, 4995)


def test_only_sm100a_code_is_embedded(lib_path):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "dv_matcher_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
