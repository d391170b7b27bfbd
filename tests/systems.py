"""Systems shared by the tests: the reference's own documentation examples."""

# /root/reference/R/odeintr.R:282-288
LORENZ = """
  dxdt[0] = sigma * (x[1] - x[0]);
  dxdt[1] = R * x[0] - x[1] - x[0] * x[2];
  dxdt[2] = -b * x[2] + x[0] * x[1];
"""
LORENZ_PARS = {"sigma": 10, "R": 28, "b": 8 / 3}

# /root/reference/README.md:151-154
VDP = """
dxdt[0] = x[1];
dxdt[1] = mu * (1 - x[0] * x[0]) * x[1] - x[0];
"""

# /root/reference/README.md:173-176
STIFF = """
  dxdt[0] = -101.0 * x[0] - 100.0 * x[1];
  dxdt[1] = x[0];
"""

# /root/reference/R/odeintr.R:534-541
ROBERTSON = """
dxdt[0] = -alpha * x[0] + beta * x[1] * x[2];
dxdt[1] = alpha * x[0] - beta * x[1] * x[2] - gamma * x[1] * x[1];
dxdt[2] = gamma * x[1] * x[1];
"""
ROBERTSON_PARS = {"alpha": 0.04, "beta": 1e4, "gamma": 3e7}

# /root/reference/tests/testthat/test_odeintr.R:33
LOGISTIC = "dxdt = x * (1 - x)"
