"""The C-ABI library builds, loads, and exports every symbol include/holocron_b200.h declares, with the argument
lists the ctypes binding assumes (no compute calls: runs without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

from holocron_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "holocron_b200.h"


def _header_decls():
    hdr = (ROOT / "include" / "holocron_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    out = {}
    for name, args in re.findall(r"(?:int|size_t) (hb_\w+)\((.*?)\);", hdr, flags=re.S):
        args = [a.strip() for a in args.replace("\n", " ").split(",")]
        if args == ["void"]:
            args = []
        sig = ""
        for a in args:
            if "*" in a:
                sig += "p"
            elif a.startswith("size_t"):
                sig += "z"
            elif a.startswith("long long"):
                sig += "q"
            elif a.startswith("float"):
                sig += "f"
            elif a.startswith("int"):
                sig += "i"
            else:
                raise AssertionError(f"unexpected argument type in header: {a!r}")
        out[name] = sig
    return out


def test_header_matches_binding_table():
    decls = _header_decls()
    assert decls, "no declarations parsed from the header"
    assert set(decls) == set(_lib.SIGNATURES), set(decls) ^ set(_lib.SIGNATURES)
    for name, sig in decls.items():
        assert _lib.SIGNATURES[name] == sig, name


def test_library_exports_every_declared_symbol():
    assert _lib.lib_path().exists(), "libholocron_b200.so has not been built (python -m holocron_b200.csrc.build)"
    handle = ctypes.CDLL(str(_lib.lib_path()))
    for name in list(_header_decls()) + ["hb_launch_count", "hb_launch_count_reset", "hb_version"]:
        assert hasattr(handle, name), f"{name} is declared in the header but not exported by the library"
    # the binding installs argtypes for all of them without error
    assert _lib.lib() is not None
    assert _lib.lib().hb_optim_chunk_elems() == 4096


def test_every_reference_citation_in_header():
    hdr = (ROOT / "include" / "holocron_b200.h").read_text()
    assert hdr.count("holocron/") >= 6  # each group cites the reference file:line it replaces


def test_header_is_valid_c_and_cxx(tmp_path):
    """include/holocron_b200.h must compile on its own as C and as C++ (it is the contract other hosts bind)."""
    import shutil
    import subprocess
    for compiler, lang in (("gcc", "c"), ("g++", "c++")):
        if shutil.which(compiler) is None:
            pytest.skip(f"{compiler} not available")
        src = tmp_path / f"use_header.{'c' if lang == 'c' else 'cpp'}"
        src.write_text('#include "holocron_b200.h"\nint main(void) { return 0; }\n')
        r = subprocess.run([compiler, "-fsyntax-only", "-Wall", f"-I{HEADER.parent}", str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
