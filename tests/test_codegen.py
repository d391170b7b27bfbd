"""Host generator (mirror of R/odeintr.R + R/utils.R of the reference): text-level behaviour."""
import ctypes as C
import re

import pytest

import odeintr_b200
from odeintr_b200 import _abi, codegen
from odeintr_b200.codegen import OdeintrError
from systems import LOGISTIC, LORENZ, LORENZ_PARS, ROBERTSON, ROBERTSON_PARS, STIFF, VDP


def test_r_number_formatting():
    # R/utils.R:181-187 pastes doubles with 15 significant digits; README.md:259 shows the 1e-06 style
    cases = {8 / 3: "2.66666666666667", 10: "10", 28.0: "28", 1e-6: "1e-06", 1e5: "1e+05", 1e4: "10000",
             0.0001: "1e-04", 0.001: "0.001", 3e7: "3e+07", 0.04: "0.04", 1 / 3: "0.333333333333333", -0.5: "-0.5",
             0.1 + 0.2: "0.3", 1234567.0: "1234567", 0.5: "0.5", 100.0: "100", 1e-15: "1e-15"}
    for v, txt in cases.items():
        assert codegen.r_num(v) == txt


def test_parameter_declarations():
    p = codegen.handle_pars(LORENZ_PARS, const=True)
    assert p.globals == "const double sigma = 10, R = 28, b = 2.66666666666667;"
    assert p.n_runtime == 0
    p = codegen.handle_pars(LORENZ_PARS)
    assert p.globals == "double sigma = 10, R = 28, b = 2.66666666666667;" and p.n_runtime == 3
    p = codegen.handle_pars("mu")
    assert p.globals == "double mu;" and p.names == ["mu"]
    p = codegen.handle_pars(["a", "b"])
    assert p.globals == "double a, b;"
    p = codegen.handle_pars(3)
    assert p.kind == "vector" and p.n_runtime == 3
    with pytest.raises(OdeintrError, match="All parameters must be named"):
        codegen.handle_pars({"a": 1, "": 2})
    with pytest.raises(OdeintrError, match="Invalid parameter specification"):
        codegen.handle_pars(object())


def test_sys_dim_and_vectorisation():
    assert codegen.get_sys_dim(LORENZ) == 3
    assert codegen.get_sys_dim("dxdt[ 7 ] = 1") == 8
    assert codegen.get_sys_dim(LOGISTIC) is None
    assert codegen.vectorize_1d_sys(LOGISTIC) == "dxdt[0] = x[0] * (1 - x[0])"
    g = codegen.generate_sys("logistic", LOGISTIC)
    assert g.sys_dim == 1 and "dxdt[0] = x[0] * (1 - x[0])" in g.code
    assert g.stepper.method == "rk5_i"  # the reference's default (R/odeintr.R:296)


def test_method_table_and_invalid_methods():
    assert codegen.make_stepper_type("rk4") == "runge_kutta4<state_type>"
    assert codegen.make_stepper_type("rk5_i") == "runge_kutta_dopri5<state_type>"
    assert codegen.make_stepper_constr("rk5_i", 1e-6, 1e-6) == "odeint::make_dense_output(1e-06, 1e-06, stepper_type())"
    assert codegen.make_stepper_constr("rk54_a", 1e-8, 1e-6) == "odeint::make_controlled(1e-08, 1e-06, stepper_type())"
    for bad in ("euler_a", "rk4_i", "rk4_a", "rk54_i", "rk78_i", "bs_a", "bsd_i"):
        with pytest.raises(OdeintrError, match="Invalid integration method"):
            codegen.stepper_spec(bad)
    with pytest.raises(OdeintrError):
        codegen.stepper_spec("nonsense")


def test_symbolic_jacobian_matches_readme_golden_text():
    # /root/reference/README.md:180-187
    assert codegen.JacobianCpp(STIFF) == ("J(0, 0) = -101;\nJ(0, 1) = -100;\nJ(1, 0) = 1;\nJ(1, 1) = 0;\n"
                                          "dfdt[0] = 0.0;\ndfdt[1] = 0.0;")
    rob = codegen.JacobianCpp(ROBERTSON)
    assert "J(1, 1) = -beta*x[2] - 2*gamma*x[1];" in rob and rob.count("J(") == 9 and rob.count("dfdt[") == 3


def _cubin(code: str, flags: int = 0) -> int:
    lib = _abi.load()
    size = C.c_size_t()
    rc = lib.odeintr_compile_cubin(code.encode(), flags, None, 0, C.byref(size))
    assert rc == 0, _abi.last_error()
    return size.value


@pytest.mark.parametrize("method", ["euler", "rk4", "rk54", "rk5", "rk78", "rk5_a", "rk5_i", "rk54_a", "rk78_a", "bs", "bsd"])
def test_generated_code_compiles_for_sm100a_without_a_device(method):
    g = codegen.generate_sys("lorenz", LORENZ, LORENZ_PARS, False, method)
    assert "__SYS__" not in g.code and "__GLOBALS__" not in g.code
    assert _cubin(g.code) > 1000
    assert _cubin(g.code, _abi.FLAG_FMAD | _abi.FLAG_KEEP_ZERO_TERMS) > 1000


def test_snippets_may_use_std_math_globals_and_vector_pars():
    g = codegen.generate_sys("osc", "dxdt[0] = x[1]; dxdt[1] = -pars[0] * std::sin(x[0]) - damp(x[1]) + std::pow(t, 2.0)",
                             pars=2, method="rk4", globals="double damp(double v) { return pars[1] * std::abs(v) * v; }")
    assert g.pars.n_runtime == 2
    assert _cubin(g.code) > 1000


def test_bad_snippet_reports_nvrtc_log():
    g = codegen.generate_sys("bad", "dxdt[0] = undefined_symbol * x[0]", method="rk4")
    lib = _abi.load()
    size = C.c_size_t()
    rc = lib.odeintr_compile_cubin(g.code.encode(), 0, None, 0, C.byref(size))
    assert rc == 2 and "undefined_symbol" in _abi.last_error()


def test_compile_false_returns_code_without_touching_the_gpu():
    code = odeintr_b200.compile_sys("logitest", LOGISTIC, compile=False, method="rk4")
    assert isinstance(code, str) and "logitest" in code and code.system is None


def test_implicit_and_large_system_code_compiles_without_a_device():
    from systems import BISTABLE_1D, BISTABLE_2D, BISTABLE_PARS, CHAIN, CHAIN_GLOBALS, CHAIN_PARS
    assert _cubin(codegen.generate_implicit("stiff", STIFF).code) > 1000
    assert _cubin(codegen.generate_implicit("rob", ROBERTSON, ROBERTSON_PARS).code) > 1000
    for method in ("rk4", "euler"):
        assert _cubin(codegen.generate_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, True, method, sys_dim=1 << 20).code) > 1000
    assert _cubin(codegen.generate_sys_openmp("b2", BISTABLE_2D, BISTABLE_PARS, False, "rk4", sys_dim=64 * 64).code) > 1000
    g = codegen.generate_sys_openmp("c", CHAIN, CHAIN_PARS, True, "rk4", sys_dim=1 << 20, globals=CHAIN_GLOBALS)
    assert "odeintr_k[1]" in g.code and _cubin(g.code) > 1000


def _kernels(code: str):
    """Names of the __global__ functions in the sm_100a cubin NVRTC builds from `code` (no device needed)."""
    import subprocess
    import tempfile

    lib = _abi.load()
    size = C.c_size_t(0)
    assert lib.odeintr_compile_cubin(code.encode(), 0, None, 0, C.byref(size)) == 0, _abi.last_error()
    buf = C.create_string_buffer(size.value)
    assert lib.odeintr_compile_cubin(code.encode(), 0, buf, size.value, C.byref(size)) == 0
    with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
        f.write(buf.raw[:size.value]); f.flush()
        out = subprocess.run(["cuobjdump", "-elf", f.name], capture_output=True, text=True).stdout
    return {w[len(".text."):] for line in out.splitlines() for w in line.split() if w.startswith(".text.odeintr_")}


def test_large_system_translation_units_carry_the_kernels_their_method_can_use():
    # which hand-written kernels a generated large-system TU instantiates (the host picks among them at run time)
    from systems import BISTABLE_1D, BISTABLE_2D, BISTABLE_DOC, BISTABLE_PARS, CHAIN, CHAIN_GLOBALS, CHAIN_PARS
    staged = {"odeintr_lattice_stage", "odeintr_lattice_stage_sharded", "odeintr_lattice_combo", "odeintr_lattice_info",
              "odeintr_lattice_probe"}
    tiled = {"odeintr_lattice_rk4_tiled", "odeintr_lattice_rk4_tiled_interior", "odeintr_lattice_rk4_tiled_tma",
             "odeintr_lattice_rk4_warp_h1", "odeintr_lattice_rk4_warp_h2", "odeintr_lattice_warp_info"}
    k = _kernels(codegen.generate_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, True, "rk4", sys_dim=1 << 20).code)
    assert k == staged | tiled  # a 2^20-cell lattice is too long for one block: no ensemble kernel
    k = _kernels(codegen.generate_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, False, "rk4", sys_dim=1024).code)
    assert k == staged | tiled | {"odeintr_lattice_ensemble"}
    k = _kernels(codegen.generate_sys_openmp("c", CHAIN, CHAIN_PARS, True, "rk4", sys_dim=2 * 300, globals=CHAIN_GLOBALS).code)
    assert k == staged | tiled | {"odeintr_lattice_ensemble"}
    dopri5 = {"odeintr_lattice_dopri5_tiled", "odeintr_lattice_dopri5_tiled_h1", "odeintr_lattice_dopri5_tiled_h2",
              "odeintr_lattice_dopri5_warp_h1", "odeintr_lattice_dopri5_warp_h2", "odeintr_lattice_dopri5_warp_info"}
    for method in ("rk5_i", "rk5_a", "rk5"):
        k = _kernels(codegen.generate_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, True, method, sys_dim=1 << 20).code)
        assert k == staged | dopri5
    # lattices one block can hold also get the adaptive one-warp / one-block-per-trajectory kernels (K5a)
    k = _kernels(codegen.generate_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, False, "rk5_i", sys_dim=1024).code)
    assert k == staged | dopri5 | {"odeintr_lattice_ensemble_dopri5_dense", "odeintr_lattice_ensemble_dopri5_controlled",
                                   "odeintr_lattice_ensemble_plain"}
    for method in ("euler", "rk54", "rk54_a", "rk78", "rk78_a", "bs", "bsd"):
        k = _kernels(codegen.generate_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, True, method, sys_dim=1 << 20).code)
        assert k == staged
    k = _kernels(codegen.generate_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, False, "rk78_a", sys_dim=300).code)
    assert k == staged | {"odeintr_lattice_ensemble_rk78_controlled", "odeintr_lattice_ensemble_plain"}
    k = _kernels(codegen.generate_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, False, "bs", sys_dim=300).code)
    assert k == staged | {"odeintr_lattice_ensemble_bs_controlled", "odeintr_lattice_ensemble_bsd_dense"}
    for src in (BISTABLE_2D, BISTABLE_DOC):
        k = _kernels(codegen.generate_sys_openmp("b2", src, BISTABLE_PARS, True, "rk4", sys_dim=96 * 96).code)
        assert k == staged | {"odeintr_lattice_rk4_tiled2d_h1", "odeintr_lattice_rk4_tiled2d_h2",
                              "odeintr_lattice_rk4_stream2d", "odeintr_lattice_stream2d_info"}
    k = _kernels(codegen.generate_sys_openmp("b2", BISTABLE_2D, BISTABLE_PARS, True, "rk5_i", sys_dim=96 * 96).code)
    assert k == staged | {"odeintr_lattice_dopri5_stream2d", "odeintr_lattice_dopri5_stream2d_info"}


def test_snippets_that_treat_the_loop_variable_as_an_int_still_build():
    # the tiled kernels hand the snippet a symbolic loop index (lattice_index.cuh).  Clamped boundaries written with
    # std::min / std::max keep compiling (mixed-type overloads); a snippet the symbolic index cannot serve at all (a
    # user template that needs both arguments of one type) is rebuilt with the stage-per-launch kernels alone
    clamp = "for (int i = 0; i < N; ++i) dxdt[i] = x[std::min(i + 1, (int)N - 1)] + x[std::max(i - 1, 0)] - 2.0 * x[i];"
    k = _kernels(codegen.generate_sys_openmp("c", clamp, None, False, "rk4", sys_dim=5000).code)
    assert "odeintr_lattice_rk4_tiled_tma" in k and "odeintr_lattice_probe" in k
    k = _kernels(codegen.generate_sys_openmp("c", clamp, None, False, "rk5_i", sys_dim=5000).code)
    assert "odeintr_lattice_dopri5_tiled_h1" in k
    pick = "template <class T> __device__ T pick(T a, T b) { return a < b ? a : b; }"
    odd = "for (int i = 0; i < N; ++i) dxdt[i] = x[pick(i + 1, (int)N - 1)] - x[i];"
    for method in ("rk4", "rk5_i"):
        k = _kernels(codegen.generate_sys_openmp("o", odd, None, False, method, sys_dim=5000, globals=pick).code)
        assert k == {"odeintr_lattice_stage", "odeintr_lattice_stage_sharded", "odeintr_lattice_combo", "odeintr_lattice_info"}
    # a snippet that is wrong either way reports the first diagnosis
    lib = _abi.load()
    size = C.c_size_t(0)
    bad = codegen.generate_sys_openmp("b", "for (int i = 0; i < N; ++i) dxdt[i] = undefined_name * x[i];", sys_dim=64)
    assert lib.odeintr_compile_cubin(bad.code.encode(), 0, None, 0, C.byref(size)) != 0
    assert "undefined_name" in _abi.last_error()


def test_cell_loop_parser():
    from systems import BISTABLE_2D, CHAIN
    loops = codegen.parse_cell_loops(CHAIN)
    assert loops.bound == "L" and loops.start == "0" and loops.var_type == "int"
    assert [o.replace(" ", "") for o in loops.offsets] == ["odeintr_c", "L+odeintr_c"]
    loops = codegen.parse_cell_loops(BISTABLE_2D)
    assert len(loops.bodies) == 2 and len(loops.offsets) == 1 and "odeintr_k[0] +=" in loops.bodies[1]
    from systems import BISTABLE_DOC
    loops = codegen.parse_cell_loops(BISTABLE_DOC)
    assert len(loops.bodies) == 2 and "laplace2D(x, odeintr_i / M, odeintr_i % M)" in loops.bodies[0]
    with pytest.raises(OdeintrError, match="same range"):
        codegen.parse_cell_loops("for (int i = 0; i < N; ++i) dxdt[i] = 1; for (int i = 1; i < N; ++i) dxdt[i] += 1;")
    g = codegen.generate_sys_openmp("b", "for (int i = 0; i < N; ++i) dxdt[i] = -x[i];", sys_dim=8)  # default rk5_i
    assert g.stepper.method == "rk5_i" and g.stepper.mode == "dense" and _cubin(g.code) > 1000
    for method, mode in (("bs", "controlled"), ("bsd", "dense")):  # steppers by type, not by suffix (R/utils.R:57-58)
        g = codegen.generate_sys_openmp("b", "for (int i = 0; i < N; ++i) dxdt[i] = -x[i];", sys_dim=8, method=method)
        assert g.stepper.mode == mode and _cubin(g.code) > 1000
    with pytest.raises(OdeintrError, match="multistep"):
        codegen.generate_sys_openmp("b", "for (int i = 0; i < N; ++i) dxdt[i] = -x[i];", sys_dim=8, method="abm4")
