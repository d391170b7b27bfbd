"""Code generator: user snippet -> CUDA translation unit.

Host-side mirror of the reference's generator (``/root/reference/R/odeintr.R:293-357,418-470,559-627`` and the
helpers in ``R/utils.R:3-72,123-240``).  Same argument surface, same parameter-text formatting, same
``sys_dim`` inference / 1-D vectorisation / invalid-method errors -- but the template that is filled in
is ``templates/compile_*_template.cu`` (a ``__device__`` functor plus the hand-written stepper kernels)
instead of an Rcpp translation unit.  Nothing here touches the GPU.
"""
from __future__ import annotations

import math
import pathlib
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Union

_TEMPLATE_DIR = pathlib.Path(__file__).resolve().parent / "templates"


class OdeintrError(RuntimeError):
    """Raised where the reference calls ``stop()`` / ``Rcpp::stop()``."""


# --------------------------------------------------------------------------------------------------
# number -> text, as R's paste()/as.character() prints a double (15 significant digits, fixed
# notation unless scientific is narrower).  R/utils.R:181-187 pastes parameter values into the C++
# source this way, so `b = 8/3` is compiled as 2.66666666666667.
# --------------------------------------------------------------------------------------------------
def r_num(x) -> str:
    if isinstance(x, bool):
        return "TRUE" if x else "FALSE"
    if isinstance(x, int):
        return str(x)
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == 0:
        return "0"
    mant, exp = ("%.14e" % abs(x)).split("e")
    kp = int(exp)
    digits = mant.replace(".", "").rstrip("0") or "0"
    nsig = len(digits)
    neg = 1 if x < 0 else 0
    # widths as in R's formatReal (src/main/format.c): fixed vs scientific
    left = kp + 1 if kp >= 0 else 1
    rgt = max(0, nsig - kp - 1)
    w_fixed = neg + left + (rgt + 1 if rgt > 0 else 0)
    e_digits = 3 if abs(kp) >= 100 else 2
    w_sci = neg + (nsig + 1 if nsig > 1 else nsig) + 2 + e_digits
    sign = "-" if neg else ""
    if w_fixed <= w_sci:
        if kp >= 0:
            whole = digits[: kp + 1].ljust(kp + 1, "0")
            frac = digits[kp + 1:]
        else:
            whole = "0"
            frac = "0" * (-kp - 1) + digits
        return sign + whole + ("." + frac if frac else "")
    m = digits[0] + ("." + digits[1:] if nsig > 1 else "")
    return "%s%se%s%0*d" % (sign, m, "+" if kp >= 0 else "-", e_digits, abs(kp))


def read_template(name: str) -> str:
    """R/utils.R:3-12."""
    return (_TEMPLATE_DIR / (name + ".cu")).read_text()


def vectorize_1d_sys(sys: str) -> str:
    """Bare ``x`` / ``dxdt`` -> ``x[0]`` / ``dxdt[0]`` (R/utils.R:14-20, incl. its never-matching guard)."""
    if re.search(r"\[\.*\]", sys):
        return sys
    sys = re.sub(r"\bx\b", "x[0]", sys)
    sys = re.sub(r"\bdxdt\b", "dxdt[0]", sys)
    return sys


def get_sys_dim(sys: str) -> Optional[int]:
    """1 + the largest literal index of ``dxdt[k]`` (R/utils.R:62-72); None where R returns NA."""
    idx = [int(m) for m in re.findall(r"dxdt\[\s*(\d+)\s*\]", sys)]
    return max(idx) + 1 if idx else None


# --------------------------------------------------------------------------------------------------
# method table (R/utils.R:22-60)
# --------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class StepperSpec:
    method: str
    base: str            # euler | rk4 | rk54 | rk5 | rk78 | bs | bsd
    mode: str            # "plain" | "controlled" | "dense"
    boost_type: str      # what the reference would have instantiated
    boost_constr: str
    kernel_header: str   # csrc/kernels header that instantiates the device kernel
    device_stepper: str  # template name used by the header


_BOOST_TYPES = {
    "euler": "euler<state_type>",
    "rk4": "runge_kutta4<state_type>",
    "rk54": "runge_kutta_cash_karp54<state_type>",
    "rk5": "runge_kutta_dopri5<state_type>",
    "rk78": "runge_kutta_fehlberg78<state_type>",
    "bs": "bulirsch_stoer<state_type>",
    "bsd": "bulirsch_stoer_dense_out<state_type>",
}
_PLAIN_DEVICE = {"euler": "euler_stepper", "rk4": "rk4_stepper", "rk54": "rk54_stepper", "rk5": "rk5_stepper",
                 "rk78": "rk78_stepper"}


def make_stepper_constr(method: str, atol, rtol) -> str:
    """R/utils.R:22-32."""
    if re.search(r"euler_|rk4_|rk54_i|rk78_i|bs_|bsd_", method):
        raise OdeintrError("Invalid integration method")
    constr = "stepper_type()"
    if method.endswith("_a"):
        constr = "odeint::make_controlled(%s, %s, stepper_type())" % (r_num(atol), r_num(rtol))
    if method.endswith("_i"):
        constr = "odeint::make_dense_output(%s, %s, stepper_type())" % (r_num(atol), r_num(rtol))
    return constr


def make_stepper_type(method: str) -> str:
    """R/utils.R:40-60 (the adams_* multistep family is out of scope: SURVEY.md §2.2)."""
    base = re.sub(r"_i$|_a$", "", method)
    if re.fullmatch(r"(ab|am|abm)\d+", base):
        raise OdeintrError("multistep method '%s' is not provided by odeintr_b200" % method)
    return _BOOST_TYPES.get(base, base + "<state_type>")


def stepper_spec(method: str, atol=1e-6, rtol=1e-6) -> StepperSpec:
    constr = make_stepper_constr(method, atol, rtol)
    boost_type = make_stepper_type(method)
    base = re.sub(r"_i$|_a$", "", method)
    if base not in _BOOST_TYPES:
        raise OdeintrError("Invalid integration method")
    if method.endswith("_a"):
        mode = "controlled"
    elif method.endswith("_i") or base == "bsd":
        mode = "dense"
    elif base == "bs":
        mode = "controlled"
    else:
        mode = "plain"
    if mode == "plain":
        if base not in _PLAIN_DEVICE:
            raise OdeintrError("method '%s' is not yet provided by odeintr_b200" % method)
        return StepperSpec(method, base, mode, boost_type, constr, "ensemble_plain.cuh", _PLAIN_DEVICE[base])
    if base == "rk5":
        dev = "dopri5_dense" if mode == "dense" else "dopri5_controlled"
        return StepperSpec(method, base, mode, boost_type, constr, "ensemble_adaptive.cuh", dev)
    if base == "rk54" and mode == "controlled":
        return StepperSpec(method, base, mode, boost_type, constr, "ensemble_adaptive.cuh", "rk54_controlled")
    if base == "bs":
        return StepperSpec(method, base, mode, boost_type, constr, "ensemble_adaptive.cuh", "bs_controlled")
    if base == "bsd":
        return StepperSpec(method, base, mode, boost_type, constr, "ensemble_adaptive.cuh", "bsd_dense")
    if base == "rk78" and mode == "controlled":
        return StepperSpec(method, base, mode, boost_type, constr, "ensemble_adaptive.cuh", "rk78_controlled")
    raise OdeintrError("method '%s' is not yet provided by odeintr_b200" % method)


# --------------------------------------------------------------------------------------------------
# parameters (R/utils.R:156-240)
# --------------------------------------------------------------------------------------------------
@dataclass
class ParsSpec:
    kind: str                     # "none" | "named" | "vector"
    names: List[str] = field(default_factory=list)
    defaults: List[float] = field(default_factory=list)
    const: bool = False
    globals: str = ""             # declaration text pasted in front of the user's globals
    load: str = ""                # body of load_params()
    has_defaults: bool = False    # named numeric pars: set_params() arguments default to these

    @property
    def n_runtime(self) -> int:
        return 0 if (self.const or self.kind == "none") else len(self.names)


def _load_named(names: Sequence[str]) -> str:
    return "\n      ".join("%s = odeintr_p[%d * odeintr_ld + odeintr_i];" % (n, k) for k, n in enumerate(names))


def handle_pars(pars, const: bool = False) -> ParsSpec:
    """R/utils.R:156-179.  ``pars`` is a dict (R named numeric vector), a str / list of str (R character
    vector: run-time parameters without defaults) or an int (R unnamed numeric: ``pars[k]`` vector)."""
    if pars is None:
        return ParsSpec("none")
    if isinstance(pars, dict):
        names = list(pars.keys())
        if not all(isinstance(n, str) and n for n in names):
            raise OdeintrError("All parameters must be named")
        vals = list(pars.values())
        decl = "double " + ", ".join("%s = %s" % (n, r_num(v)) for n, v in zip(names, vals)) + ";"
        # the run-time default of a parameter is what the pasted text compiles to (15 significant digits),
        # e.g. b = 8/3 defaults to 2.66666666666667 until name_set_params(b = 8/3) stores the exact double
        vals = [float(r_num(v)) for v in vals]
        if const:
            # compile-time constants: R/utils.R:181-187 with const = TRUE; no setter / getter
            return ParsSpec("named", names, [float(v) for v in vals], True, "const " + decl, "", True)
        return ParsSpec("named", names, [float(v) for v in vals], False, decl, _load_named(names), True)
    if isinstance(pars, str) or (isinstance(pars, (list, tuple)) and pars and all(isinstance(p, str) for p in pars)):
        names = [pars] if isinstance(pars, str) else list(pars)
        if not all(names):
            raise OdeintrError("All parameters must be named")
        decl = "double " + ", ".join(names) + ";"
        return ParsSpec("named", names, [0.0] * len(names), False, decl, _load_named(names))
    if isinstance(pars, (int, float)) and not isinstance(pars, bool):
        n = int(pars)
        if n < 1:
            raise OdeintrError("Invalid parameter specification")
        names = ["pars[%d]" % k for k in range(n)]
        return ParsSpec("vector", names, [0.0] * n, False, "odeintr_b200::vec<%d> pars;" % n, _load_named(names))
    if isinstance(pars, (list, tuple)) and pars and all(isinstance(p, (int, float)) for p in pars):
        # unnamed numeric vector: R uses pars[[1]] as the length
        return handle_pars(int(pars[0]), const)
    raise OdeintrError("Invalid parameter specification")


# --------------------------------------------------------------------------------------------------
# symbolic Jacobian (R/utils.R:123-154, R/odeintr.R:613-627) -- sympy instead of stats::D
# --------------------------------------------------------------------------------------------------
_MATH_FUNCS = {"exp", "log", "sin", "cos", "tan", "sinh", "cosh", "tanh", "sqrt", "asin", "acos", "atan", "pow",
               "fabs"}


def Jacobian2(code: Union[str, Sequence[str]], sys_dim: int = -1) -> str:
    import sympy
    from sympy.parsing.sympy_parser import parse_expr

    if not isinstance(code, str):
        code = ";".join(code)
    if sys_dim is None or sys_dim < 1:
        sys_dim = get_sys_dim(code)
    if sys_dim is None:
        code = vectorize_1d_sys(code)
        sys_dim = 1
    stmts = [s.strip() for s in code.split(";")]
    stmts = [s for s in stmts if s]
    rows: Dict[int, str] = {}
    for s in stmts:
        m = re.match(r"dxdt\[\s*(\d+)\s*\]\s*=\s*(.*)", s, flags=re.S)
        if not m:
            raise OdeintrError("cannot differentiate statement: " + s)
        rows[int(m.group(1))] = m.group(2)
    xs = [sympy.Symbol("x__%d" % j, real=True) for j in range(sys_dim)]
    local = {"x__%d" % j: xs[j] for j in range(sys_dim)}
    local["pow"] = sympy.Pow
    local["fabs"] = sympy.Abs
    out = []
    for i in range(sys_dim):
        if i not in rows:
            raise OdeintrError("no expression for dxdt[%d]" % i)
        rhs = re.sub(r"\bstd::", "", rows[i])
        rhs = re.sub(r"\bx\s*\[\s*(\d+)\s*\]", r"x__\1", rhs)
        # every identifier that is not a math function is a parameter (R's D() treats them the same way);
        # without this sympy would read names such as beta / gamma as its own special functions
        names = dict(local)
        for ident in set(re.findall(r"[A-Za-z_]\w*", rhs)):
            if ident not in names and ident not in _MATH_FUNCS:
                names[ident] = sympy.Symbol(ident, real=True)
        expr = parse_expr(rhs, local_dict=names, evaluate=True)
        for j in range(sys_dim):
            d = sympy.diff(expr, xs[j])
            txt = sympy.ccode(sympy.nsimplify(d, rational=False) if d.is_number else d)
            if d.is_number:
                txt = r_num(float(d))
            txt = re.sub(r"\bx__(\d+)\b", r"x[\1]", txt)
            out.append("J(%d, %d) = %s;" % (i, j, txt))
    return "\n".join(out)


def JacobianCpp(sys: Union[str, Sequence[str]], sys_dim: int = -1) -> str:
    """R/odeintr.R:613-627.  (The reference passes the un-vectorised snippet to Jacobian2 and therefore
    fails on bare 1-D syntax; here the vectorised text is used.)"""
    if not isinstance(sys, str):
        sys = "; ".join(sys)
    if sys_dim is None or sys_dim < 1:
        sys_dim = get_sys_dim(sys)
    if sys_dim is None:
        sys = vectorize_1d_sys(sys)
        sys_dim = 1
    jac = Jacobian2(sys, sys_dim)
    dfdt = "\n".join("dfdt[%d] = 0.0;" % i for i in range(sys_dim))
    return jac + "\n" + dfdt


# --------------------------------------------------------------------------------------------------
# translation-unit assembly
# --------------------------------------------------------------------------------------------------
@dataclass
class Generated:
    code: str
    name: str
    sys_dim: int
    pars: ParsSpec
    stepper: Optional[StepperSpec]
    family: str          # "plain" | "adaptive" | "implicit" | "lattice"
    atol: float
    rtol: float
    n_fields: int = 1    # large systems: distinct dxdt[i + offset] targets of the cell loop


def _fill(template: str, subs: Dict[str, str]) -> str:
    # the reference uses sub() (first match) for user text and gsub() for names; user text is inserted
    # last and literally, so a snippet containing "__SYS__" cannot trigger another substitution
    code = template
    for key in ("__FUNCNAME__", "__SYS_SIZE__", "__NPARS__", "__STEPPER_TYPE__", "__KERNEL_HEADER__", "__ATOL__",
                "__RTOL__"):
        if key in subs:
            code = code.replace(key, subs[key])
    for key in ("__HEADERS__", "__GLOBALS__", "__LOAD_PARS__", "__JACOBIAN__", "__SYS__", "__FOOTERS__"):
        if key in subs:
            code = code.replace(key, subs[key], 1)
    return code


def _join(sys) -> str:
    return sys if isinstance(sys, str) else "; ".join(sys)


def generate_sys(name: str, sys, pars=None, const: bool = False, method: str = "rk5_i", sys_dim: int = -1,
                 atol: float = 1e-6, rtol: float = 1e-6, globals: str = "", headers: str = "",
                 footers: str = "") -> Generated:
    """Text generation half of compile_sys (R/odeintr.R:308-336)."""
    p = handle_pars(pars, const)
    spec = stepper_spec(method, atol, rtol)
    sys = _join(sys)
    if sys_dim is None or math.ceil(sys_dim) < 1:
        sys_dim = get_sys_dim(sys)
    if sys_dim is None:
        sys = vectorize_1d_sys(sys)
        sys_dim = 1
    sys_dim = int(math.ceil(sys_dim))
    code = _fill(read_template("compile_sys_template"), {
        "__FUNCNAME__": name,
        "__SYS_SIZE__": str(sys_dim),
        "__NPARS__": str(p.n_runtime),
        "__STEPPER_TYPE__": spec.device_stepper,
        "__KERNEL_HEADER__": spec.kernel_header,
        "__HEADERS__": headers,
        "__GLOBALS__": p.globals + globals,
        "__LOAD_PARS__": p.load,
        "__SYS__": sys,
        "__FOOTERS__": footers,
    })
    family = "plain" if spec.mode == "plain" else "adaptive"
    return Generated(code, name, sys_dim, p, spec, family, float(atol), float(rtol))


def generate_implicit(name: str, sys, pars=None, const: bool = False, sys_dim: int = -1, jacobian=None,
                      atol: float = 1e-6, rtol: float = 1e-6, globals: str = "", headers: str = "",
                      footers: str = "") -> Generated:
    """Text generation half of compile_implicit (R/odeintr.R:573-599).  ``jacobian`` defaults to the
    symbolic JacobianCpp(sys, sys_dim), as the reference's default argument does (R/odeintr.R:563)."""
    p = handle_pars(pars, const)
    sys = _join(sys)
    if jacobian is None:
        jacobian = JacobianCpp(sys, sys_dim)
    jacobian = _join(jacobian)
    if sys_dim is None or math.ceil(sys_dim) < 1:
        sys_dim = get_sys_dim(sys)
    if sys_dim is None:
        sys = vectorize_1d_sys(sys)
        sys_dim = 1
    sys_dim = int(math.ceil(sys_dim))
    code = _fill(read_template("compile_implicit_template"), {
        "__FUNCNAME__": name,
        "__SYS_SIZE__": str(sys_dim),
        "__NPARS__": str(p.n_runtime),
        "__ATOL__": r_num(atol),
        "__RTOL__": r_num(rtol),
        "__HEADERS__": headers,
        "__GLOBALS__": p.globals + globals,
        "__LOAD_PARS__": p.load,
        "__JACOBIAN__": jacobian,
        "__SYS__": sys,
        "__FOOTERS__": footers,
    })
    return Generated(code, name, sys_dim, p, None, "implicit", float(atol), float(rtol))


# --------------------------------------------------------------------------------------------------
# large systems (compile_sys_openmp, R/odeintr.R:418-470): the snippet is a sequence of loops over the
# cells of the state, `for (int i = 0; i < BOUND; ++i) BODY`, optionally under `#pragma omp` lines
# (R/odeintr.R:396-401).  On the GPU every loop iteration becomes one thread's work item: the loop
# header is stripped and BODY becomes a per-cell __device__ function.
# --------------------------------------------------------------------------------------------------
@dataclass
class CellLoops:
    var_type: str
    start: str
    bound: str
    bodies: List[str]            # per-cell bodies, dxdt[...] replaced by odeintr_k[slot]
    var_names: List[str]
    offsets: List[str]           # index expression of every distinct dxdt[...] target (slot order)


def _strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def _match(src: str, pos: int, open_ch: str, close_ch: str) -> int:
    depth = 0
    for k in range(pos, len(src)):
        if src[k] == open_ch:
            depth += 1
        elif src[k] == close_ch:
            depth -= 1
            if depth == 0:
                return k
    raise OdeintrError("unbalanced '%s' in system definition" % open_ch)


def parse_cell_loops(sys_src: str) -> CellLoops:
    src = _strip_comments(sys_src)
    src = "\n".join(l for l in src.split("\n") if not l.strip().startswith("#pragma"))
    # the reference joins snippet lines with "; " (R/odeintr.R:317): drop the empty statements that leaves
    pos, n = 0, len(src)
    loops = []
    while True:
        while pos < n and (src[pos].isspace() or src[pos] == ";"):
            pos += 1
        if pos >= n:
            break
        # `laplace4(x, dxdt, D);` -- the helper the reference documents for the method of lines
        # (R/odeintr.R:383-388,398) but never shipped: dxdt = D * (periodic 4-neighbour Laplacian of x)
        ml = re.match(r"laplace4\s*\(\s*x\s*,\s*dxdt\s*,", src[pos:])
        if ml:
            p_open = pos + src[pos:].index("(")
            p_close = _match(src, p_open, "(", ")")
            coeff = src[pos + ml.end():p_close].strip()
            loops.append(("int", "odeintr_i", "0", "N",
                          "dxdt[odeintr_i] = (%s) * laplace2D(x, odeintr_i / M, odeintr_i %% M);" % coeff))
            pos = p_close + 1
            continue
        m = re.match(r"for\s*\(", src[pos:])
        if not m:
            raise OdeintrError("large-system snippets must be loops over the cells of the state "
                               "(`for (int i = 0; i < N; ++i) ...`); found: " + src[pos:pos + 40].strip())
        p_open = pos + m.end() - 1
        p_close = _match(src, p_open, "(", ")")
        header = src[p_open + 1:p_close]
        parts = header.split(";")
        if len(parts) != 3:
            raise OdeintrError("cannot parse loop header: for (" + header + ")")
        mi = re.match(r"\s*((?:unsigned\s+|const\s+|std::)?[A-Za-z_][\w:]*(?:\s+long)?)\s+([A-Za-z_]\w*)\s*=\s*(.+?)\s*$",
                      parts[0], flags=re.S)
        if not mi:
            raise OdeintrError("cannot parse loop initialiser: " + parts[0])
        vtype, var, start = mi.group(1), mi.group(2), mi.group(3)
        mc = re.match(r"\s*%s\s*(<|!=)\s*(.+?)\s*$" % re.escape(var), parts[1], flags=re.S)
        if not mc:
            raise OdeintrError("loop condition must be `%s < bound`: %s" % (var, parts[1]))
        bound = mc.group(2)
        if not re.match(r"\s*(\+\+\s*%s|%s\s*\+\+|%s\s*\+=\s*1)\s*$" % ((re.escape(var),) * 3), parts[2]):
            raise OdeintrError("loop increment must be ++%s: %s" % (var, parts[2]))
        pos = p_close + 1
        while pos < n and src[pos].isspace():
            pos += 1
        if pos < n and src[pos] == "{":
            b_close = _match(src, pos, "{", "}")
            body = src[pos + 1:b_close]
            pos = b_close + 1
        else:
            b_end = src.find(";", pos)
            if b_end < 0:
                b_end = n
            body = src[pos:b_end] + ";"
            pos = b_end + 1
        loops.append((vtype, var, start, bound, body))
    if not loops:
        raise OdeintrError("empty system definition")
    vtype, _, start, bound, _ = loops[0]
    norm = lambda s: re.sub(r"\s+", "", s)
    for l in loops[1:]:
        if norm(l[2]) != norm(start) or norm(l[3]) != norm(bound):
            raise OdeintrError("all cell loops of a large system must run over the same range")
    offsets: List[str] = []
    keys: List[str] = []
    bodies, names = [], []
    for (_, var, _, _, body) in loops:
        out, k = [], 0
        while True:
            m = re.search(r"\bdxdt\s*\[", body[k:])
            if not m:
                out.append(body[k:])
                break
            i_open = k + m.end() - 1
            i_close = _match(body, i_open, "[", "]")
            expr = body[i_open + 1:i_close]
            # canonical key: the index expression with the loop variable renamed, whitespace removed
            key = norm(re.sub(r"\b%s\b" % re.escape(var), "odeintr_c", expr))
            if key not in keys:
                keys.append(key)
                offsets.append(re.sub(r"\b%s\b" % re.escape(var), "odeintr_c", expr))
            out.append(body[k:k + m.start()] + "odeintr_k[%d]" % keys.index(key))
            k = i_close + 1
        bodies.append("".join(out))
        names.append(var)
    return CellLoops(vtype, start, bound, bodies, names, offsets)


_LATTICE_STEPPERS = {"rk4": "4", "euler": "1", "rk5_a": "7", "rk5_i": "7", "rk54": "6", "rk54_a": "6", "rk78": "13", "rk78_a": "13", "rk5": "7",
                     "bs": "2", "bsd": "2"}  # (the stage count selects the kernels a TU instantiates; bs / bsd use the combo kernel alone)


def generate_sys_openmp(name: str, sys, pars=None, const: bool = False, method: str = "rk5_i", sys_dim: int = -1,
                        atol: float = 1e-6, rtol: float = 1e-6, globals: str = "", headers: str = "",
                        footers: str = "") -> Generated:
    """Text generation half of compile_sys_openmp (R/odeintr.R:433-456)."""
    p = handle_pars(pars, const)
    make_stepper_constr(method, atol, rtol)  # same validity errors as the reference
    boost_type = make_stepper_type(method)
    base = re.sub(r"_i$|_a$", "", method)
    if method not in _LATTICE_STEPPERS:
        raise OdeintrError("method '%s' is not yet provided for large systems by odeintr_b200 "
                           "(euler, rk4, rk5, rk54, rk78, rk5_a, rk54_a, rk78_a, bs, bsd and the default rk5_i are)" % method)
    sys = _join(sys)
    if sys_dim is None or math.ceil(sys_dim) < 1:
        sys_dim = get_sys_dim(sys)
    if sys_dim is None:
        raise OdeintrError("large systems must pass sys_dim (loop-indexed dxdt has no literal index; "
                           "R/odeintr.R:403)")
    sys_dim = int(math.ceil(sys_dim))
    mode = "dense" if method.endswith("_i") else ("controlled" if method.endswith("_a") else "plain")
    if base in ("bs", "bsd"):  # a controlled / a dense-output stepper by type (R/utils.R:57-58)
        mode = "controlled" if base == "bs" else "dense"
    try:
        loops = parse_cell_loops(sys)
    except OdeintrError:
        # Not a set of loops over the cells (nested loops, a reduction ahead of the loop, statements outside any
        # loop ...): the reference compiles any C++ into sys() (compile_sys_openmp_template.cpp:39-43).  So does the
        # general form: the snippet verbatim, evaluated by one thread; only the steppers' vector algebra runs in
        # parallel.  Correct for everything the reference accepts, fast for nothing.
        body = "\n".join(l for l in _strip_comments(sys).split("\n") if not l.strip().startswith("#pragma"))
        code = _fill(read_template("compile_sys_openmp_generic_template"), {
            "__FUNCNAME__": name,
            "__SYS_SIZE__": str(sys_dim),
            "__NPARS__": str(p.n_runtime),
            "__STEPPER_TYPE__": _LATTICE_STEPPERS[method],
            "__HEADERS__": headers,
            "__GLOBALS__": p.globals + globals,
            "__LOAD_PARS__": p.load,
            "__SYS__": body,
            "__FOOTERS__": footers,
        })
        spec = StepperSpec(method, base, mode, boost_type, make_stepper_constr(method, atol, rtol), "lattice_rk.cuh",
                           "lattice_" + base)
        return Generated(code, name, sys_dim, p, spec, "lattice", float(atol), float(rtol), 1)
    cell = []
    for var, body in zip(loops.var_names, loops.bodies):
        cell.append("{ const auto %s = odeintr_b200::loop_index<%s>(odeintr_c); %s }" % (var, loops.var_type, body))
    offs = ", ".join("static_cast<long long>(%s)" % o for o in loops.offsets)
    code = _fill(read_template("compile_sys_openmp_template"), {
        "__FUNCNAME__": name,
        "__SYS_SIZE__": str(sys_dim),
        "__NPARS__": str(p.n_runtime),
        "__STEPPER_TYPE__": _LATTICE_STEPPERS[method],
        "__HEADERS__": headers,
        "__GLOBALS__": p.globals + globals,
        "__LOAD_PARS__": p.load,
        "__SYS__": "\n      ".join(cell),
        "__FOOTERS__": footers,
    })
    # an M x M lattice (laplace2D & friends, or M itself) is walked in column strips by the stage kernel
    if re.search(r"\b(laplace4|laplace2D|d_left|d_right|d_up|d_down|M)\b", _strip_comments(sys + " " + globals)):
        code = code.replace("#define ODEINTR_NFIELDS", "#define ODEINTR_LATTICE_2D 1\n#define ODEINTR_NFIELDS", 1)
    code = code.replace("__NFIELDS__", str(len(loops.offsets)))
    code = code.replace("__CELL_START__", loops.start).replace("__CELL_BOUND__", loops.bound)
    code = code.replace("__FIELD_OFFSETS__", offs)
    spec = StepperSpec(method, base, mode, boost_type, make_stepper_constr(method, atol, rtol), "lattice_rk.cuh",
                       "lattice_" + base)
    return Generated(code, name, sys_dim, p, spec, "lattice", float(atol), float(rtol), len(loops.offsets))
