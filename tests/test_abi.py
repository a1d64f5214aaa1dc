"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly what include/pgd_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pgd_b200.h")


def _declared():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pgd_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib_path():
    from pgdrome_b200 import _build

    return _build.build_library()


def test_header_symbols_exported(lib_path):
    names = _declared()
    assert len(names) >= 24
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (pgd_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    extra = sorted(n for n in exported if n not in names and n != "pgd_free_pattern")
    assert not extra, extra


def test_ctypes_table_matches_header(lib_path):
    from pgdrome_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load_library()
    assert lib.pgd_abi_version() == 1
    # argument counts agree with the header prototypes
    txt = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, args in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, txt, flags=re.S)
        assert m, name
        params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        assert len(params) == len(args), (name, len(params), len(args))


def test_library_targets_sm100a(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback():
    import torch

    from pgdrome_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.PGDB200Error):
        _lib.handle()
    with pytest.raises(_lib.PGDB200Error):
        _lib.spmv(torch.zeros(2, dtype=torch.int32), torch.zeros(1, dtype=torch.int32), torch.zeros(1, dtype=torch.float64),
                  torch.zeros(1, dtype=torch.float64))
    h = ctypes.c_void_p()
    assert _lib.load_library().pgd_create(0, ctypes.byref(h)) != 0
