"""ctypes binding of include/odeintr_b200.h (libodeintr_b200.so).

This is the stub a maintainer of the reference would write as Rcpp glue (INTEGRATION.md shows the R
version); every product computation goes through it.  There is no fallback: if the shared library is
missing, or it reports an error, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import pathlib

_LIB_PATH = pathlib.Path(__file__).resolve().parent / "libodeintr_b200.so"

c_dbl_p = C.POINTER(C.c_double)
c_ll_p = C.POINTER(C.c_longlong)
c_int_p = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every symbol include/odeintr_b200.h declares
SIGNATURES = {
    "odeintr_compile": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_uint,
                                  C.POINTER(C.c_void_p)]),
    "odeintr_compile_cubin": (C.c_int, [C.c_char_p, C.c_uint, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "odeintr_destroy": (C.c_int, [C.c_void_p]),
    "odeintr_last_error": (C.c_char_p, []),
    "odeintr_compile_log": (C.c_char_p, [C.c_void_p]),
    "odeintr_set_device": (C.c_int, [C.c_int]),
    "odeintr_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "odeintr_set_max_tries": (C.c_int, [C.c_void_p, C.c_longlong]),
    "odeintr_set_instance_order": (C.c_int, [C.c_void_p, C.c_int, c_int_p, C.c_size_t]),
    "odeintr_get_instance_order": (C.c_int, [C.c_void_p, c_int_p, C.c_size_t]),
    "odeintr_synchronize": (C.c_int, [C.c_void_p]),
    "odeintr_set_state": (C.c_int, [C.c_void_p, c_dbl_p, C.c_size_t, C.c_size_t]),
    "odeintr_get_state": (C.c_int, [C.c_void_p, c_dbl_p, C.c_size_t]),
    "odeintr_set_params": (C.c_int, [C.c_void_p, c_dbl_p, C.c_size_t, C.c_size_t]),
    "odeintr_get_params": (C.c_int, [C.c_void_p, c_dbl_p, C.c_size_t]),
    "odeintr_n_instances": (C.c_size_t, [C.c_void_p]),
    "odeintr_n_param_sets": (C.c_size_t, [C.c_void_p]),
    "odeintr_fill_state_uniform": (C.c_int, [C.c_void_p, C.c_size_t, C.c_ulonglong, C.c_ulonglong, c_dbl_p, c_dbl_p]),
    "odeintr_fill_params_linspace": (C.c_int, [C.c_void_p, C.c_size_t, C.c_ulonglong, C.c_ulonglong, c_dbl_p,
                                               c_dbl_p]),
    "odeintr_integrate_const": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_longlong, C.c_int]),
    "odeintr_integrate_times": (C.c_int, [C.c_void_p, c_dbl_p, C.c_size_t, C.c_double]),
    "odeintr_integrate_adaptive": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int]),
    "odeintr_integrate_const_host": (C.c_int, [C.c_void_p, c_dbl_p, C.c_size_t, c_dbl_p, C.c_size_t, C.c_double,
                                               C.c_double, C.c_double, C.c_longlong, c_dbl_p, C.c_size_t, c_dbl_p,
                                               C.c_size_t]),
    "odeintr_const_rows": (C.c_size_t, [C.c_double, C.c_double, C.c_double, C.c_longlong]),
    "odeintr_lattice_set_halo": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int]),
    "odeintr_lattice_get_halo": (C.c_longlong, [C.c_void_p]),
    "odeintr_lattice_set_exchange_interval": (C.c_int, [C.c_void_p, C.c_int]),
    "odeintr_lattice_shard": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, C.c_int,
                                        C.c_void_p]),
    "odeintr_lattice_ghost_width": (C.c_longlong, [C.c_void_p, C.c_int]),
    "odeintr_lattice_pad_width": (C.c_longlong, [C.c_void_p, C.c_int]),
    "odeintr_lattice_pitch": (C.c_longlong, [C.c_void_p, C.c_int, C.c_longlong]),
    "odeintr_lattice_shard_step": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "odeintr_nccl_unique_id": (C.c_int, [C.c_void_p, C.c_size_t]),
    "odeintr_lattice_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]),
    "odeintr_lattice_shard_run": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_double, C.c_double]),
    "odeintr_lattice_exchange": (C.c_int, [C.c_void_p]),
    "odeintr_lattice_shard_get_output": (C.c_int, [C.c_void_p, c_dbl_p, C.c_size_t]),
    "odeintr_reset_observer": (C.c_int, [C.c_void_p]),
    "odeintr_release_record": (C.c_int, [C.c_void_p]),
    "odeintr_output_rows": (C.c_size_t, [C.c_void_p]),
    "odeintr_get_output_times": (C.c_int, [C.c_void_p, c_dbl_p, C.c_size_t]),
    "odeintr_get_output": (C.c_int, [C.c_void_p, c_dbl_p, C.c_size_t, C.c_size_t, C.c_size_t]),
    "odeintr_get_output_soa": (C.c_int, [C.c_void_p, c_dbl_p, C.c_size_t]),
    "odeintr_output_is_ragged": (C.c_int, [C.c_void_p]),
    "odeintr_get_output_rows_per_instance": (C.c_int, [C.c_void_p, c_int_p, C.c_size_t]),
    "odeintr_get_output_times_per_instance": (C.c_int, [C.c_void_p, c_dbl_p, C.c_size_t, C.c_size_t]),
    "odeintr_get_stats": (C.c_int, [C.c_void_p, c_ll_p, c_ll_p, c_int_p, C.c_size_t]),
    "odeintr_last_steps": (C.c_longlong, [C.c_void_p]),
    "odeintr_state_device_ptr": (C.c_void_p, [C.c_void_p]),
    "odeintr_output_device_ptr": (C.c_void_p, [C.c_void_p]),
    "odeintr_params_device_ptr": (C.c_void_p, [C.c_void_p]),
    "odeintr_last_kernel_ms": (C.c_double, [C.c_void_p]),
    "odeintr_last_launches": (C.c_longlong, [C.c_void_p]),
    "odeintr_total_launches": (C.c_longlong, []),
    "odeintr_kernel_attributes": (C.c_int, [C.c_void_p, C.c_char_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "odeintr_host_alloc": (C.c_void_p, [C.c_size_t]),
    "odeintr_host_free": (C.c_int, [C.c_void_p]),
    "odeintr_selftest_exact_division": (C.c_int, [C.c_int, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_ulonglong)]),
    "odeintr_fp64_probe": (C.c_int, [C.c_int, C.c_int, c_dbl_p, c_dbl_p]),
    "odeintr_hbm_probe": (C.c_int, [C.c_int, C.c_size_t, c_dbl_p]),
    "odeintr_pcie_probe": (C.c_int, [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, c_dbl_p]),
    "odeintr_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t, c_int_p, c_int_p, c_int_p,
                                      C.POINTER(C.c_size_t)]),
}

ORDER_INDEX, ORDER_COST_PROXY, ORDER_PERMUTATION = 0, 1, 2

FLAG_FMAD = 1
FLAG_KEEP_ZERO_TERMS = 2
FLAG_DENSE_OUTPUT = 4
FLAG_LATTICE_EULER = 8
FLAG_LATTICE_ADAPTIVE = 16
FLAG_LATTICE_CK54 = 64  # large-system TU stepped with Cash-Karp (rk54 / rk54_a): set by the generator
FLAG_LATTICE_RKF78 = 128  # large-system TU stepped with Fehlberg 7(8) (rk78 / rk78_a): set by the generator
FLAG_LATTICE_BS = 512  # large-system TU stepped with Bulirsch-Stoer (bs; bsd with DENSE_OUTPUT): set by the generator
FLAG_LATTICE_RK5 = 256  # large-system TU stepped with dopri5 as a plain fixed-step stepper (rk5): set by the generator
FLAG_RECIPROCAL = 32  # rosenbrock4 fast mode: reciprocals instead of repeated divisions (not bit-compatible)

_lib = None


def lib_path() -> pathlib.Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load libodeintr_b200.so and type every entry point.  Fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise ImportError(
            "odeintr_b200: %s is missing -- build it with `make -C odeintr_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback." % _LIB_PATH)
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = load().odeintr_last_error()
    return msg.decode("utf-8", "replace") if msg else ""
