"""The C-ABI shared library loads without a GPU and exports every symbol include/magic_b200.h declares."""
import ctypes
import os
import re
import subprocess

import magic_b200  # noqa: F401
from magic_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    import __graft_entry__ as g
    return g.build()


def test_header_symbols_are_exported():
    lib_path = _ensure_built()
    decls = _lib.parse_header()
    assert len(decls) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (magic_\w+)", out))
    missing = sorted(set(decls) - exported)
    assert not missing, missing
    lib = _lib.load()
    assert lib.magic_version() >= 100
    assert isinstance(lib.magic_last_error(), bytes)


def test_header_is_plain_c_abi():
    src = open(_lib.HEADER).read()
    assert 'extern "C"' in src
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)   # no torch / C++ types in the signatures
    assert "torch" not in code and "at::" not in code and "std::" not in code and "#include" not in code
    for name, (ret, args) in _lib.parse_header().items():
        assert ret in (ctypes.c_int, ctypes.c_char_p), name


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    from magic_b200 import synth
    from magic_b200.config import make_config
    m = magic_b200.GlocalTextPathCMTPreTraining(make_config(128))
    with pytest.raises(Exception):
        m(synth.make_batch("sap", 2), "sap", True)   # no CPU fallback exists


def test_header_compiles_as_c_and_links_from_a_c_program(tmp_path):
    """The boundary is usable from plain C (what a cgo / JNI / FFI binding sees): include/magic_b200.h compiles with
    `gcc -std=c99 -pedantic -Wall -Werror`, and a C program links against the shared library and calls the entry
    points that need no device (version, last error, the GEMM capability query) without a GPU."""
    lib_path = _ensure_built()
    c = tmp_path / "abi.c"
    c.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "magic_b200.h"\n'
        "int main(void) {\n"
        "  int v = magic_version();\n"
        "  const char* e = magic_last_error();\n"
        '  printf("%d %d %d\\n", v, e != NULL, magic_gemm_tc_supported(5120, 768, 768));\n'
        "  return v >= 100 ? 0 : 1;\n"
        "}\n")
    exe = tmp_path / "abi"
    libdir = os.path.dirname(lib_path)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o",
                    str(exe), "-L", libdir, "-lmagic_b200", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) >= 100 and out[1] == "1" and out[2] in ("0", "1")
