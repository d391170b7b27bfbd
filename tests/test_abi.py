"""The C-ABI shared library: loads, exports every symbol the header declares, and has no CPU path."""
import ctypes as C
import pathlib
import re

import pytest

from conftest import has_gpu
from odeintr_b200 import _abi, codegen
from systems import LORENZ, LORENZ_PARS

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "odeintr_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(odeintr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "libodeintr_b200.so does not export " + n
    assert sorted(_abi.SIGNATURES) == names, "ctypes table and header disagree"


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_a_device():
    lib = _abi.load()
    g = codegen.generate_sys("lorenz", LORENZ, LORENZ_PARS, True, "rk4")
    h = C.c_void_p()
    rc = lib.odeintr_compile(g.code.encode(), b"lorenz", 3, 0, 1e-6, 1e-6, 0, C.byref(h))
    assert rc == 3 and not h.value
    assert "no CPU fallback" in _abi.last_error()
    import odeintr_b200

    with pytest.raises(odeintr_b200.OdeintrError, match="no CPU fallback"):
        odeintr_b200.compile_sys("lorenz", LORENZ, LORENZ_PARS, True, method="rk4")


def test_product_never_touches_the_oracle():
    for p in (ROOT / "odeintr_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h", ".cpp") and p.is_file():
            text = p.read_text()
            assert "build_oracle" not in text and "odeint_restated" not in text, p
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), p


def test_python_flag_constants_match_the_header():
    import re

    header = (ROOT / "include" / "odeintr_b200.h").read_text()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+ODEINTR_FLAG_(\w+)\s+(\d+)u", header)}
    assert len(flags) >= 9
    for name, value in flags.items():
        assert getattr(_abi, "FLAG_" + name) == value, name
    assert len(set(flags.values())) == len(flags)  # one bit each
    assert all(v & (v - 1) == 0 for v in flags.values())


def test_rcpp_glue_compiles_against_the_header():
    # src/gpu_glue.cpp is the reference-side binding (the Rcpp translation unit that replaces the module sourceCpp used
    # to build).  R and Rcpp do not exist in this image: type-check it against include/odeintr_b200.h with a mock of the
    # Rcpp interface it uses, so that the glue cannot drift away from the C-ABI unnoticed.
    import subprocess

    root = pathlib.Path(__file__).resolve().parent.parent
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", str(root / "include"),
                        "-I", str(root / "tests" / "mock_rcpp"), str(root / "src" / "gpu_glue.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # every generated function of compile_sys_template.cpp:58-171 has its export in the glue and its closure in R/
    glue = (root / "src" / "gpu_glue.cpp").read_text()
    rsrc = (root / "R" / "odeintr_gpu.R").read_text()
    for fn in ("run", "adap", "at", "continue_at", "no_record", "reset_observer", "get_state", "set_state", "get_output",
               "set_params", "get_params"):
        assert "odeintr_gpu_%s(" % fn in glue and "odeintr_gpu_%s(" % fn in rsrc, fn
