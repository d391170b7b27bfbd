"""compile_sys / compile_sys_openmp / compile_implicit -- the reference's user-facing generator calls
(``/root/reference/R/odeintr.R:293-357,418-470,559-612``) with the same arguments and defaults.

Like the reference, each call returns the generated code (a ``str``; the reference returns it
invisibly) and, when ``compile=True``, installs the generated functions ``NAME``, ``NAME_at``, ... into
``env`` (a dict, the analogue of the R environment that gets ``attach()``-ed under ``name``).  The
returned string also carries ``.system`` (the CompiledSystem) and ``.env``.
"""
from __future__ import annotations

from typing import Dict, Optional

from . import codegen
from .codegen import OdeintrError

#: name -> env, the analogue of R's search path entries created by attach(env, name = name)
attached: Dict[str, Dict[str, object]] = {}


class GeneratedCode(str):
    """The code string returned by compile_*; ``.system`` / ``.env`` are set when it was compiled."""
    system = None
    env = None


def _install(gen: codegen.Generated, compile: bool, env: Optional[dict], flags: int, observer=None) -> GeneratedCode:
    code = GeneratedCode(gen.code)
    if compile:
        from .system import CompiledSystem

        system = CompiledSystem(gen, flags=flags, observer=observer)
        if env is None:
            env = {}
        env.update(system.exports())
        attached[gen.name] = env  # re-attaching a name replaces the previous entry (R/odeintr.R:350-352)
        code.system = system
        code.env = env
    return code


def compile_sys(name, sys, pars=None, const=False, method="rk5_i", sys_dim=-1, atol=1e-6, rtol=1e-6, globals="",
                headers="", footers="", compile=True, observer=None, env=None, flags=0):
    """R/odeintr.R:293-357.  ``observer`` is a callable ``f(x, t)`` (an R closure in the reference): it is replayed on
    the host over the samples the GPU recorded, and NAME_get/set_observer and NAME_get/set_output_processor are
    installed as the reference's observer template does (R/odeintr.R:343-347)."""
    if observer is not None and not callable(observer):
        raise OdeintrError("observer must be a function of (x, t)")
    gen = codegen.generate_sys(name, sys, pars, const, method, sys_dim, atol, rtol, globals, headers, footers)
    return _install(gen, compile, env, flags, observer)


def compile_sys_openmp(name, sys, pars=None, const=False, method="rk5_i", sys_dim=-1, atol=1e-6, rtol=1e-6,
                       globals="", headers="", footers="", compile=True, env=None, flags=0, halo=None):
    """R/odeintr.R:418-470: the large-system variant (Boost openmp_range_algebra in the reference).

    ``halo`` (extension, optional): the stencil radius of the loop body on a 1-D lattice -- how many cells to the
    left / right of cell i the body reads.  When it is small, a fixed-step rk4 step runs as one temporally blocked
    kernel (8x less HBM traffic, same bits).  ``None`` (default): measured from the snippet's index expressions at the
    first integrate call (and verified on every access; a snippet that outgrows it is re-integrated by the
    stage-per-launch kernels); a positive number: declared, accesses beyond it are reported as an error; ``0``:
    always the stage-per-launch kernels."""
    gen = codegen.generate_sys_openmp(name, sys, pars, const, method, sys_dim, atol, rtol, globals, headers, footers)
    code = _install(gen, compile, env, flags)
    if compile and method == "rk4":
        code.system.set_halo(-1 if halo is None else int(halo), gen.n_fields)
    return code


def compile_implicit(name, sys, pars=None, const=False, sys_dim=-1, jacobian=None, atol=1e-6, rtol=1e-6, globals="",
                     headers="", footers="", compile=True, env=None, flags=0):
    """R/odeintr.R:559-612: rosenbrock4 with controller and dense output, symbolic Jacobian by default."""
    gen = codegen.generate_implicit(name, sys, pars, const, sys_dim, jacobian, atol, rtol, globals, headers, footers)
    return _install(gen, compile, env, flags)
