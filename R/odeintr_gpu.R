# R/odeintr_gpu.R -- what changes on the R side of thk686/odeintr when the compiled system runs on the B200 engine.
#
# compile_sys(), compile_implicit() and compile_sys_openmp() keep their arguments and their text processing
# (handle_pars, get_sys_dim, vectorize_1d_sys, make_stepper_constr's validity checks: R/odeintr.R:293-357,
# 418-470, 559-627; R/utils.R:3-72, 123-240).  Two things differ:
#   1. the templates read by read_template() are the CUDA ones of odeintr_b200/templates/ (same placeholders);
#   2. Rcpp::sourceCpp(code = code, env = env) (R/odeintr.R:340) becomes odeintr_gpu_compile() of src/gpu_glue.cpp,
#      which returns an external pointer, and the functions the generated module used to export
#      (inst/templates/compile_sys_template.cpp:58-171, pars_*_template.cpp) become the closures installed by
#      gpu_install_functions() below.
#
# R is not available in the image this repository is built in: this file is the committed form of the change, it is
# not executed here.  The executed mirror (same names, arguments, defaults, warnings and errors) is
# odeintr_b200/codegen.py + api.py + system.py; tests/test_codegen.py pins the generated text.

# flag bits of include/odeintr_b200.h
ODEINTR_FLAG_DENSE_OUTPUT = 4L
ODEINTR_FLAG_LATTICE_EULER = 8L
ODEINTR_FLAG_LATTICE_ADAPTIVE = 16L
ODEINTR_FLAG_LATTICE_CK54 = 64L
ODEINTR_FLAG_LATTICE_RKF78 = 128L
ODEINTR_FLAG_LATTICE_RK5 = 256L
ODEINTR_FLAG_LATTICE_BS = 512L

# device stepper named by a method string (replaces make_stepper_type / make_stepper_constr, R/utils.R:22-60:
# the Boost type and constructor strings become the name of a hand-written device stepper)
gpu_stepper = function(method)
{
  make_stepper_constr(method, 0, 0)  # for its "Invalid integration method" checks
  base = sub("_i$|_a$", "", method)
  if (!(base %in% c("euler", "rk4", "rk54", "rk5", "rk78", "bs", "bsd")))
    stop("Invalid integration method")
  mode = if (grepl("_i$", method) || base == "bsd") "dense" else if (grepl("_a$", method) || base == "bs") "controlled" else "plain"
  name = switch(paste(base, mode),
                "euler plain" = "euler_stepper", "rk4 plain" = "rk4_stepper", "rk54 plain" = "rk54_stepper",
                "rk5 plain" = "rk5_stepper", "rk78 plain" = "rk78_stepper",
                "rk5 controlled" = "dopri5_controlled", "rk5 dense" = "dopri5_dense",
                "rk54 controlled" = "rk54_controlled", "rk78 controlled" = "rk78_controlled",
                "bs controlled" = "bs_controlled", "bsd dense" = "bsd_dense",
                stop("Invalid integration method"))
  list(name = name, base = base, mode = mode, method = method)
}

# number of run-time parameters of a pars specification (0 when const = TRUE: they are compiled into the text)
gpu_n_runtime_pars = function(pars, const)
{
  if (is.null(pars) || const) return(0L)
  if (length(pars) == 1 && is.numeric(pars) && is.null(names(pars))) return(as.integer(pars))
  length(pars)
}

# The generated-function surface of a system called `name` (compile_sys_template.cpp:58-171), as closures over the
# engine handle.  `init` may also be an N x M matrix: M trajectories integrated at once (the ensemble extension).
gpu_install_functions = function(name, h, sys_dim, pars, const, env)
{
  n_inst = function(x) if (is.matrix(x)) ncol(x) else 1L
  as_soa = function(x) if (is.matrix(x)) as.numeric(t(x)) else as.numeric(x)  # instance index fastest
  fn = list()
  fn[[name]] = function(init, duration, step_size = 1.0, start = 0.0)
    odeintr_gpu_run(h, init, duration, step_size, start)
  fn[[paste0(name, "_adap")]] = function(init, duration, step_size = 1.0, start = 0.0)
    odeintr_gpu_adap(h, init, duration, step_size, start)
  fn[[paste0(name, "_at")]] = function(init, times, step_size = 1.0, start = 0.0)
    odeintr_gpu_at(h, init, times, step_size, start)
  fn[[paste0(name, "_continue_at")]] = function(times, step_size = 1.0)
    odeintr_gpu_continue_at(h, sys_dim, times, step_size)
  fn[[paste0(name, "_no_record")]] = function(init, duration, step_size = 1.0, start = 0.0)
    odeintr_gpu_no_record(h, init, duration, step_size, start)
  fn[[paste0(name, "_reset_observer")]] = function() odeintr_gpu_reset_observer(h)
  fn[[paste0(name, "_get_state")]] = function() odeintr_gpu_get_state(h, sys_dim)
  fn[[paste0(name, "_set_state")]] = function(new_state) odeintr_gpu_set_state(h, as_soa(new_state), n_inst(new_state))
  fn[[paste0(name, "_get_output")]] = function() odeintr_gpu_get_output(h, sys_dim)
  np = gpu_n_runtime_pars(pars, const)
  if (np > 0)
  {
    if (is.null(names(pars)))
    {
      # vector form (pars_vec_setter_template.cpp): NAME_set_params(pars) / NAME_get_params()
      fn[[paste0(name, "_set_params")]] = function(pars)
      {
        if (length(pars) %% np != 0) stop("Invalid parameter vector")
        odeintr_gpu_set_params(h, as_soa(pars), length(pars) %/% np)
      }
      fn[[paste0(name, "_get_params")]] = function() odeintr_gpu_get_params(h, np)
    } else {
      # named form (pars_setter_template.cpp:1-5): NAME_set_params(a = dflt, b = dflt, ...); a vector of length M for
      # any of them sweeps that parameter over M instances
      dflt = if (is.character(pars)) rep(NA_real_, np) else as.numeric(pars)
      pnames = if (is.character(pars)) pars else names(pars)
      setter = function()
      {
        vals = mget(pnames, envir = environment())
        m = max(vapply(vals, length, 1L))
        odeintr_gpu_set_params(h, as.numeric(t(vapply(vals, function(v) rep_len(as.numeric(v), m), numeric(m)))), m)
      }
      formals(setter) = setNames(as.list(dflt), pnames)
      fn[[paste0(name, "_set_params")]] = setter
      fn[[paste0(name, "_get_params")]] = function()
        setNames(as.list(odeintr_gpu_get_params(h, np)), pnames)
    }
    if (!is.character(pars)) odeintr_gpu_set_params(h, as.numeric(pars), 1L)
  }
  for (n in names(fn)) assign(n, fn[[n]], envir = env)
  invisible(names(fn))
}

# compile_sys with the engine behind it: the body of R/odeintr.R:293-357 with the two changes marked
compile_sys_gpu = function(name, sys, pars = NULL, const = FALSE, method = "rk5_i", sys_dim = -1L,
                           atol = 1e-6, rtol = 1e-6, globals = "", headers = "", footers = "",
                           compile = TRUE, observer = NULL, env = new.env(), ...)
{
  if (!is.null(pars))
  {
    pcode = handle_pars(pars, const)
    globals = paste0(pcode$globals, globals, collapse = ";\n")
  }
  sys = paste0(sys, collapse = "; ")
  stepper = gpu_stepper(method)                                   # CHANGE 1a: device stepper instead of a Boost type
  if (ceiling(sys_dim) < 1) sys_dim = get_sys_dim(sys)
  if (is.na(sys_dim))
  {
    sys = vectorize_1d_sys(sys)
    sys_dim = 1L
  }
  code = read_template("compile_sys_template")                    # CHANGE 1b: odeintr_b200/templates/*.cu
  code = gsub("__STEPPER_TYPE__", stepper$name, code)
  code = sub("__SYS_SIZE__", ceiling(sys_dim), code)
  code = sub("__NPARS__", gpu_n_runtime_pars(pars, const), code)
  code = sub("__GLOBALS__", globals, code)
  code = sub("__SYS__", sys, code)
  code = sub("__HEADERS__", headers, code)
  code = sub("__FOOTERS__", footers, code)
  code = gsub("__FUNCNAME__", name, code)
  if (compile)
  {
    flags = if (stepper$mode == "dense") ODEINTR_FLAG_DENSE_OUTPUT else 0L
    h = try(odeintr_gpu_compile(code, name, ceiling(sys_dim), gpu_n_runtime_pars(pars, const), atol, rtol, flags))
    if (!inherits(h, "try-error"))                                # CHANGE 2: closures over the handle
    {
      gpu_install_functions(name, h, ceiling(sys_dim), pars, const, env)
      if (!identical(env, globalenv()))
      {
        if (name %in% search())
          detach(pos = match(name, search()))
        do.call("attach", list(what = env, name = name))
      }
    }
  }
  return(invisible(code))
}

# compile_sys_openmp (R/odeintr.R:418-470): same change; the large-system template, the method's flag bits, and the
# stencil radius of the cell loop measured by the engine unless `halo` is given (0 = stage-per-launch kernels)
compile_sys_openmp_gpu = function(name, sys, pars = NULL, const = FALSE, method = "rk5_i", sys_dim = -1L,
                                  atol = 1e-6, rtol = 1e-6, globals = "", headers = "", footers = "",
                                  compile = TRUE, env = new.env(), halo = NULL, ...)
{
  if (ceiling(sys_dim) < 1) stop("sys_dim must be given for large systems")
  if (!is.null(pars))
  {
    pcode = handle_pars(pars, const)
    globals = paste0(pcode$globals, globals, collapse = ";\n")
  }
  sys = paste0(sys, collapse = "; ")
  stepper = gpu_stepper(method)
  code = read_template("compile_sys_openmp_template")
  code = sub("__SYS_SIZE__", ceiling(sys_dim), code)
  code = sub("__NPARS__", gpu_n_runtime_pars(pars, const), code)
  code = sub("__GLOBALS__", globals, code)
  code = sub("__SYS__", sys, code)   # (the engine's generator also splits the cell loops: odeintr_b200/codegen.py)
  code = sub("__HEADERS__", headers, code)
  code = sub("__FOOTERS__", footers, code)
  code = gsub("__FUNCNAME__", name, code)
  if (compile)
  {
    flags = 0L
    if (stepper$mode == "dense") flags = bitwOr(flags, ODEINTR_FLAG_DENSE_OUTPUT)
    if (stepper$mode != "plain") flags = bitwOr(flags, ODEINTR_FLAG_LATTICE_ADAPTIVE)
    if (stepper$base == "euler") flags = bitwOr(flags, ODEINTR_FLAG_LATTICE_EULER)
    if (stepper$method == "rk5") flags = bitwOr(flags, ODEINTR_FLAG_LATTICE_RK5)
    if (stepper$base == "rk54") flags = bitwOr(flags, ODEINTR_FLAG_LATTICE_CK54)
    if (stepper$base == "rk78") flags = bitwOr(flags, ODEINTR_FLAG_LATTICE_RKF78)
    if (stepper$base %in% c("bs", "bsd")) flags = bitwOr(flags, ODEINTR_FLAG_LATTICE_BS)
    h = try(odeintr_gpu_compile(code, name, ceiling(sys_dim), gpu_n_runtime_pars(pars, const), atol, rtol, flags))
    if (!inherits(h, "try-error"))
    {
      if (method == "rk4") odeintr_gpu_set_halo(h, if (is.null(halo)) -1L else as.integer(halo))
      gpu_install_functions(name, h, ceiling(sys_dim), pars, const, env)
      if (!identical(env, globalenv()))
      {
        if (name %in% search())
          detach(pos = match(name, search()))
        do.call("attach", list(what = env, name = name))
      }
    }
  }
  return(invisible(code))
}

# compile_implicit (R/odeintr.R:559-627): the symbolic Jacobian text (JacobianCpp, R/utils.R:123-154) is pasted into
# the implicit CUDA template exactly as it is pasted into compile_implicit_template.cpp:49-56
compile_implicit_gpu = function(name, sys, pars = NULL, const = FALSE, sys_dim = -1L, jacobian = JacobianCpp(sys, sys_dim),
                                atol = 1e-6, rtol = 1e-6, globals = "", headers = "", footers = "",
                                compile = TRUE, env = new.env(), ...)
{
  if (!is.null(pars))
  {
    pcode = handle_pars(pars, const)
    globals = paste0(pcode$globals, globals, collapse = ";\n")
  }
  sys = paste0(sys, collapse = "; ")
  if (ceiling(sys_dim) < 1) sys_dim = get_sys_dim(sys)
  if (is.na(sys_dim))
  {
    sys = vectorize_1d_sys(sys)
    sys_dim = 1L
  }
  code = read_template("compile_implicit_template")
  code = sub("__SYS_SIZE__", ceiling(sys_dim), code)
  code = sub("__NPARS__", gpu_n_runtime_pars(pars, const), code)
  code = sub("__GLOBALS__", globals, code)
  code = sub("__SYS__", sys, code)
  code = sub("__JACOBIAN__", jacobian, code)
  code = sub("__HEADERS__", headers, code)
  code = sub("__FOOTERS__", footers, code)
  code = gsub("__FUNCNAME__", name, code)
  if (compile)
  {
    h = try(odeintr_gpu_compile(code, name, ceiling(sys_dim), gpu_n_runtime_pars(pars, const), atol, rtol,
                                ODEINTR_FLAG_DENSE_OUTPUT))
    if (!inherits(h, "try-error"))
    {
      gpu_install_functions(name, h, ceiling(sys_dim), pars, const, env)
      if (!identical(env, globalenv()))
      {
        if (name %in% search())
          detach(pos = match(name, search()))
        do.call("attach", list(what = env, name = name))
      }
    }
  }
  return(invisible(code))
}
