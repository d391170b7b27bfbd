"""odeintr_b200 -- B200-native integration engine behind odeintr's compile_sys / compile_implicit API.

Public surface = the reference's exported generator functions (``/root/reference/NAMESPACE:3-10`` minus
the R-closure ``integrate_sys`` path): ``compile_sys``, ``compile_sys_openmp``,
``compile_implicit``, ``JacobianCpp``.  See ``api.py``.
"""
from .codegen import JacobianCpp, OdeintrError, r_num  # noqa: F401
from .api import compile_sys, compile_sys_openmp, compile_implicit, attached  # noqa: F401
from .system import CompiledSystem, EnsembleRecord, proc_output  # noqa: F401

__all__ = ["compile_sys", "compile_sys_openmp", "compile_implicit", "JacobianCpp", "OdeintrError",
           "CompiledSystem", "EnsembleRecord", "attached"]
