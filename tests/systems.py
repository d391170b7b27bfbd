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

# 1-D periodic bistable reaction-diffusion chain: the reference's own large-system example
# (/root/reference/R/odeintr.R:396-401) with the Laplacian written out (its laplace4 helper does not exist)
BISTABLE_1D = """
#pragma omp parallel for
for (int i = 0; i < N; ++i)
  dxdt[i] = D * (x[(i + N - 1) % N] + x[(i + 1) % N] - 2.0 * x[i]) + a * x[i] * (1 - x[i]) * (x[i] - b);
"""
BISTABLE_PARS = {"D": 0.1, "a": 1.0, "b": 0.5}

# the same on the periodic M x M lattice through the reference's laplace2D helper (inst/include/utils.h:29-31),
# written as two loops like the reference's example (Laplacian first, reaction term added by a second loop)
BISTABLE_2D = """
#pragma omp parallel for
for (int i = 0; i < N; ++i)
  dxdt[i] = D * laplace2D(x, i / M, i % M);
#pragma omp parallel for
for (int i = 0; i < N; ++i)
  dxdt[i] += a * x[i] * (1 - x[i]) * (x[i] - b);
"""

# periodic chain of L = N/2 anharmonic oscillators, SoA state [q_0..q_{L-1}, p_0..p_{L-1}]
CHAIN = """
#pragma omp parallel for
for (int i = 0; i < L; ++i) {
  const double q = x[i];
  dxdt[i] = x[L + i];
  dxdt[L + i] = -q - beta * q * q * q + K * (x[(i + L - 1) % L] - 2.0 * q + x[(i + 1) % L]);
}
"""
CHAIN_PARS = {"beta": 1.0, "K": 0.5}
CHAIN_GLOBALS = "static const int L = N / 2;"

# the reference's documented large-system example, verbatim (R/odeintr.R:396-401); `laplace4` is promised by the
# documentation but missing from the reference's sources, here it is provided
BISTABLE_DOC = """
 laplace4(x, dxdt, D);  // parallel 4-point discrete laplacian
 #pragma omp parallel for
 for (int i = 0; i < N; ++i)
   dxdt[i] += a * x[i] * (1 - x[i]) * (x[i] - b);
"""
