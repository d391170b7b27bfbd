// odeintr_b200/csrc/lattice_adaptive.h — included by abi.cu (inside its anonymous namespace).
//
// Adaptive Dormand-Prince 5(4) for large systems (compile_sys_openmp with the reference's default method
// rk5_i, or rk5_a): the state and every stage derivative are device-resident vectors of N doubles; each
// Runge-Kutta stage is one launch of odeintr_lattice_combo (RHS + Boost-ordered vector algebra, the last one
// also reducing default_error_checker's max norm).  The stepper classes below expose the interfaces the shared
// integrate loops expect (kernels/integrate_loops.cuh), so the control flow is the same code that every GPU
// thread runs for ensembles; the controller arithmetic (pow) runs on the host.
// Reference semantics: controlled_runge_kutta / dense_output_runge_kutta over runge_kutta_dopri5 with
// openmp_range_algebra (R/utils.R:34-38,52), SURVEY.md Appendix A.2-A.5.
#pragma once

struct lat_sys {};  // the system lives in the NVRTC module; the loops only pass it through

// handle of one device-resident vector; assignment copies the values (like std::vector assignment does)
struct lat_state {
  odeintr_system_t h = nullptr;
  double* p = nullptr;
  lat_state() {}
  lat_state(odeintr_system_t h_, double* p_) : h(h_), p(p_) {}
  lat_state(const lat_state&) = default;
  lat_state& operator=(const lat_state& o) {
    if (p != o.p && p && o.p) {
      cudaMemcpyAsync(p, o.p, h->lat_vec_len * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
      if (h->sharded) slab_mark(h, p, slab_is_fresh(h, o.p));  // ghost cells travel with the copy
    }
    return *this;
  }
};

struct lat_counters {
  long long accepted = 0, rejected = 0, budget = 0x7fffffffffffffffLL;
  int status = ODEINTR_STATUS_OK;
  bool exhausted() {
    if (accepted + rejected < budget) return false;
    status |= ODEINTR_STATUS_STEP_BUDGET;
    return true;
  }
};

struct lattice_dopri5 {
  odeintr_system_t h;
  double *k2, *k3, *k4, *k5, *k6, *xtA, *xtB;
  int rc = 0;  // first launch error, reported by the caller
  double eps_abs_o = -1.0, eps_rel_o = -1.0;  // tolerances of the error norm when they are not the handle's (bs / bsd: 1e-6)
  int all = 0;                                 // combos also cover the entries no user loop writes (bs / bsd)

  int combo(const double* xs, const double* base, int nin, const double* const* in, const double* c, double ck,
            double* out, double* k_out, double t, const double* err_x = nullptr, const double* err_d = nullptr,
            double adt = 0.0) {
    odeintr_lattice_combo_args a;
    std::memset(&a, 0, sizeof(a));
    if (xs && h->k_serial) {
      // general snippet: k = f(xs) by the serial kernel into k_out (or a scratch vector), then the same sum with k as
      // its last stored term -- the order of the additions is unchanged
      double* kdst = k_out;
      if (!kdst) {
        if (h->lat_negzero.reserve((size_t)h->N) != cudaSuccess) { if (!rc) rc = fail(ODEINTR_E_CUDA, "out of device memory"); return rc; }
        kdst = h->lat_negzero.p;
        cudaMemsetAsync(kdst, 0, (size_t)h->N * sizeof(double), h->stream);
      }
      const double* pars = h->P ? h->pars.p : nullptr;
      void* params[4] = {(void*)&xs, (void*)&kdst, (void*)&pars, (void*)&t};
      if (cudaLaunchKernel((const void*)h->k_serial, dim3(1), dim3(32), params, 0, h->stream) != cudaSuccess) {
        if (!rc) rc = fail(ODEINTR_E_CUDA, "launch of odeintr_lattice_serial_rhs failed");
        return rc;
      }
      h->last_launches += 1;
      g_total_launches += 1;
      const double* in2[12];
      double c2[12];
      for (int j = 0; j < nin; ++j) { in2[j] = in[j]; c2[j] = c[j]; }
      int n2 = nin;
      if (ck != 0.0) { in2[n2] = kdst; c2[n2] = ck; ++n2; }  // (a zero weight is a skipped tableau entry)
      if (!out && !err_x) return 0;  // the derivative alone was asked for
      xs = nullptr; in = in2; c = c2; nin = n2; ck = 0.0; k_out = nullptr;
      a.xs = nullptr; a.base = base; a.nin = nin;
      for (int j = 0; j < nin; ++j) { a.in[j] = in[j]; a.c[j] = c[j]; }
      a.ck = 0.0; a.out = out; a.k_out = nullptr;
    } else {
    a.xs = xs; a.base = base; a.nin = nin;
    for (int j = 0; j < nin; ++j) { a.in[j] = in[j]; a.c[j] = c[j]; }
    a.ck = ck; a.out = out; a.k_out = k_out;
    }
    a.err_x = err_x; a.err_d = err_d; a.adt = adt;
    a.eps_abs = eps_abs_o >= 0.0 ? eps_abs_o : h->atol; a.eps_rel = eps_rel_o >= 0.0 ? eps_rel_o : h->rtol;
    a.all = all;
    a.err_max = err_x ? h->lat_err.p : nullptr;
    a.pars = h->P ? h->pars.p : nullptr;
    a.n_total = h->N;
    a.t = t;
    const int r = launch(h, h->k_combo, h->stream, (unsigned)h->sm_count * 8u, 256u, &a);
    if (r && !rc) rc = r;
    return r;
  }

  // --- temporally blocked attempt (kernels/lattice_dopri5_tiled.cuh) ---------------------------------------------
  // The stage derivatives k3..k6 never reach HBM on this path: the dense output of a step is evaluated by repeating
  // the accepted attempt with the interpolation folded into its last stage (same arithmetic, same bits).
  bool k_valid = true;  // k3..k6 hold the last step's stage derivatives (stage-per-launch path)
  const double *last_in = nullptr, *last_din = nullptr;
  double last_t = 0.0, last_dt = 0.0;

  // returns the max-norm error, or a negative number when the kernel could not be used
  double tiled_attempt(const double* in, const double* d_in, double t, double* out, double* d_out, double dt,
                       double* dense_out = nullptr, const double* dcoef = nullptr) {
    odeintr_lattice_dopri5_args a;
    std::memset(&a, 0, sizeof(a));
    a.x_in = in; a.d_in = d_in; a.x_out = out; a.d_out = d_out;
    a.dense_out = dense_out;
    if (dense_out) for (int j = 0; j < 6; ++j) a.dcoef[j] = dcoef[j];
    a.pars = h->P ? h->pars.p : nullptr;
    a.halo = h->tiled5_halo;
    a.t = t; a.dt = dt;
    a.eps_abs = h->atol; a.eps_rel = h->rtol; a.adt = 1.0 * std::fabs(dt);
    a.err_max = h->lat_err.p;
    a.flag = reinterpret_cast<int*>(h->lat_err.p + 1);
    if (slab) {  // the attempt reads 6 * halo ghost cells of x and k1 on either side: bring them up to date
      const int erc = slab_refresh(h, in, d_in);
      if (erc) { if (!rc) rc = erc; return -1.0; }
      if (out) slab_mark(h, out, false);
      if (d_out) slab_mark(h, d_out, false);
      if (dense_out) slab_mark(h, dense_out, false);
    }
    cudaMemsetAsync(h->lat_err.p, 0, 2 * sizeof(unsigned long long), h->stream);
    if (h->tiled5 == 2) {  // M x M lattice: the streaming kernel, one warp per (strip of 52 columns, segment of rows)
      a.own_lo = 0; a.own_hi = (long long)h->N; a.c0 = 0; a.pitch = (long long)h->N; a.sharded = 0;
      a.seg_rows = h->stream5_rows;
      {  // coefficient * dt products (left to right sums use them as they are: the same doubles the other paths form)
        const double a2 = 1.0 / 5.0, a3 = 3.0 / 10.0, a4 = 4.0 / 5.0, a5 = 8.0 / 9.0;
        const double b21 = 1.0 / 5.0;
        const double b31 = 3.0 / 40.0, b32 = 9.0 / 40.0;
        const double b41 = 44.0 / 45.0, b42 = -56.0 / 15.0, b43 = 32.0 / 9.0;
        const double b51 = 19372.0 / 6561.0, b52 = -25360.0 / 2187.0, b53 = 64448.0 / 6561.0, b54 = -212.0 / 729.0;
        const double b61 = 9017.0 / 3168.0, b62 = -355.0 / 33.0, b63 = 46732.0 / 5247.0, b64 = 49.0 / 176.0,
                     b65 = -5103.0 / 18656.0;
        const double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0, c5 = -2187.0 / 6784.0, c6 = 11.0 / 84.0;
        const double dc1 = c1 - 5179.0 / 57600.0, dc3 = c3 - 7571.0 / 16695.0, dc4 = c4 - 393.0 / 640.0,
                     dc5 = c5 - (-92097.0 / 339200.0), dc6 = c6 - 187.0 / 2100.0, dc7 = -1.0 / 40.0;
        volatile double vdt = dt;  // (keeps the host compiler from folding or contracting anything: plain IEEE products)
        const double d = vdt;
        double* cf = a.cf;
        cf[ODEINTR_CF_B21] = d * b21;
        cf[ODEINTR_CF_B31] = d * b31; cf[ODEINTR_CF_B32] = d * b32;
        cf[ODEINTR_CF_B41] = d * b41; cf[ODEINTR_CF_B42] = d * b42; cf[ODEINTR_CF_B43] = d * b43;
        cf[ODEINTR_CF_B51] = d * b51; cf[ODEINTR_CF_B52] = d * b52; cf[ODEINTR_CF_B53] = d * b53; cf[ODEINTR_CF_B54] = d * b54;
        cf[ODEINTR_CF_B61] = d * b61; cf[ODEINTR_CF_B62] = d * b62; cf[ODEINTR_CF_B63] = d * b63; cf[ODEINTR_CF_B64] = d * b64;
        cf[ODEINTR_CF_B65] = d * b65;
        cf[ODEINTR_CF_C1] = d * c1; cf[ODEINTR_CF_C3] = d * c3; cf[ODEINTR_CF_C4] = d * c4; cf[ODEINTR_CF_C5] = d * c5;
        cf[ODEINTR_CF_C6] = d * c6;
        if (dense_out) {
          cf[ODEINTR_CF_E1] = dcoef[0]; cf[ODEINTR_CF_E3] = dcoef[1]; cf[ODEINTR_CF_E4] = dcoef[2];
          cf[ODEINTR_CF_E5] = dcoef[3]; cf[ODEINTR_CF_E6] = dcoef[4]; cf[ODEINTR_CF_E7] = dcoef[5];
        } else {
          cf[ODEINTR_CF_E1] = d * dc1; cf[ODEINTR_CF_E3] = d * dc3; cf[ODEINTR_CF_E4] = d * dc4;
          cf[ODEINTR_CF_E5] = d * dc5; cf[ODEINTR_CF_E6] = d * dc6; cf[ODEINTR_CF_E7] = d * dc7;
        }
        cf[ODEINTR_CF_T2] = t + d * a2; cf[ODEINTR_CF_T3] = t + d * a3; cf[ODEINTR_CF_T4] = t + d * a4;
        cf[ODEINTR_CF_T5] = t + d * a5; cf[ODEINTR_CF_T6] = t + d;
      }
      const long long side = h->lat_side;
      const long long tasks = ((side + 51) / 52) * ((side + h->stream5_rows - 1) / h->stream5_rows);
      const long long wpb = h->stream5_threads / 32;
      unsigned grid = (unsigned)std::min<long long>((tasks + wpb - 1) / wpb, (long long)h->sm_count * h->stream5_blocks);
      if (const char* c = std::getenv("ODEINTR_STREAM_MAX_BLOCKS")) grid = std::max(1u, std::min(grid, (unsigned)std::atoi(c)));  // (tests)
      void* params[1] = {&a};
      if (cudaLaunchKernel((const void*)h->k_stream5, dim3(grid), dim3((unsigned)h->stream5_threads), params, 0, h->stream) != cudaSuccess) {
        if (!rc) rc = fail(ODEINTR_E_CUDA, "launch of odeintr_lattice_dopri5_stream2d failed");
        return -1.0;
      }
      h->last_launches += 1;
      g_total_launches += 1;
      if (dense_out) return 0.0;
      unsigned long long words2[2] = {0, 0};
      cudaMemcpyAsync(words2, h->lat_err.p, sizeof(words2), cudaMemcpyDeviceToHost, h->stream);
      cudaStreamSynchronize(h->stream);
      if (words2[1] & 0xffffffffull) return -1.0;  // a read beyond radius 1: not this kernel's system
      double err2;
      std::memcpy(&err2, &words2[0], sizeof(err2));
      return err2;
    }
    const long long L = h->lat_cells_ct;
    const long long tile = 256LL * tile5_cpt((int)((long long)h->N / L));
    long long cells = L;
    if (slab) {  // sharded run: this rank's cells, addressed through its slab (ghost cells hold the neighbours')
      a.own_lo = h->sh_c0; a.own_hi = h->sh_c0 + h->sh_n; a.c0 = h->sh_c0 - h->sh_pad; a.pitch = h->sh_pitch; a.sharded = 1;
      cells = h->sh_n;
    } else {
      a.own_lo = 0; a.own_hi = L; a.c0 = 0; a.pitch = L; a.sharded = 0;
    }
    // one launch of the block-tiled kernel over the cells [lo, hi)
    auto launch_tiled = [&](long long lo, long long hi) -> bool {
      if (hi <= lo) return true;
      odeintr_lattice_dopri5_args at = a;
      at.own_lo = lo; at.own_hi = hi;
      void* params[1] = {&at};
      if (cudaLaunchKernel((const void*)h->k_dopri5_tiled, dim3((unsigned)((hi - lo + tile - 1) / tile)), dim3(256), params,
                           h->tiled5_smem, h->stream) != cudaSuccess) {
        if (!rc) rc = fail(ODEINTR_E_CUDA, "launch of odeintr_lattice_dopri5_tiled failed");
        return false;
      }
      h->last_launches += 1;
      g_total_launches += 1;
      return true;
    };
    (void)cells;
    // Interior cells -- the whole window of 6 * halo (+ halo: the warp kernel evaluates its outermost cells too) around
    // them inside the user loop's range, away from the periodic seam -- go to the warp-autonomous kernel.
    bool warp_done = false;
    if (h->k_warp5 && h->warp5_blocks > 0) {
      const long long hh = h->warp5_halo, m6 = 6 * hh + h->tiled5_halo;
      const long long glo = h->lat_start > 0 ? h->lat_start : 0, ghi = h->lat_bound < L ? h->lat_bound : L;
      long long ilo = std::max(a.own_lo, glo + m6), ihi = std::min(a.own_hi, ghi - m6);
      const int nf = (int)((long long)h->N / L);
      const double* outs[2] = {dense_out ? dense_out : out, dense_out ? dense_out : d_out};
      bool tma = !std::getenv("ODEINTR_NO_TMA") && (nf == 1 || a.pitch % 2 == 0) &&
                 reinterpret_cast<uintptr_t>(in) % 16 == 0 && reinterpret_cast<uintptr_t>(d_in) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(outs[0]) % 16 == 0 && reinterpret_cast<uintptr_t>(outs[1]) % 16 == 0;
      if (tma) {
        if ((ilo - a.c0) & 1) ++ilo;
        if ((ihi - a.c0) & 1) --ihi;
      }
      const long long P = 32LL * h->warp5_cpt - 12LL * hh;
      if (ihi - ilo >= P) {
        odeintr_lattice_dopri5_args aw = a;
        aw.own_lo = ilo; aw.own_hi = ihi; aw.tma = tma ? 1 : 0;
        const long long warp_tiles = (ihi - ilo + P - 1) / P;
        const unsigned grid = (unsigned)std::min<long long>((warp_tiles + 7) / 8, (long long)h->sm_count * h->warp5_blocks);
        void* params[1] = {&aw};
        if (cudaLaunchKernel((const void*)h->k_warp5, dim3(grid), dim3(256), params, h->warp5_smem, h->stream) != cudaSuccess) {
          if (!rc) rc = fail(ODEINTR_E_CUDA, "launch of odeintr_lattice_dopri5_warp failed");
          return -1.0;
        }
        h->last_launches += 1;
        g_total_launches += 1;
        if (!launch_tiled(a.own_lo, ilo) || !launch_tiled(ihi, a.own_hi)) return -1.0;
        warp_done = true;
      }
    }
    if (!warp_done && !launch_tiled(a.own_lo, a.own_hi)) return -1.0;
    if (dense_out) return 0.0;  // nothing to read back: the attempt it repeats has already been vetted
    if (slab) {  // error norm over the whole lattice: max over the ranks' slabs (one small all-reduce per attempt)
      const int erc = slab_allreduce_err(h);
      if (erc) { if (!rc) rc = erc; return -1.0; }
    }
    unsigned long long words[2] = {0, 0};
    cudaMemcpyAsync(words, h->lat_err.p, sizeof(words), cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    if (words[1] & 0xffffffffull) return -1.0;  // a read beyond the measured radius: not this kernel's system
    double err;
    std::memcpy(&err, &words[0], sizeof(err));
    return err;
  }

  bool slab = false;  // the vectors are slabs of a sharded run (odeintr_lattice_shard): [field][pitch] with ghost cells

  // runge_kutta_dopri5::do_step_impl(sys, in, dxdt_in, t, out, dxdt_out, dt, xerr) + default_error_checker::error
  // returns the max-norm error (A.2); out / dxdt_out must not alias in / dxdt_in
  double step_with_error(const double* in, const double* d_in, double t, double* out, double* d_out, double dt) {
    if (h->tiled5 >= 1) {
      const double e = tiled_attempt(in, d_in, t, out, d_out, dt);
      if (e >= 0.0) {
        k_valid = false;
        last_in = in; last_din = d_in; last_t = t; last_dt = dt;
        return e;
      }
      if (rc) return 0.0;
      if (slab) {
        rc = fail(ODEINTR_E_INVALID, "the system definition reads x further than the measured stencil radius; sharded "
                                     "adaptive stepping has no stage-per-launch path");
        return 0.0;
      }
      h->tiled5 = 0;  // the stage-per-launch path below has no reach limit
    }
    if (slab) {
      rc = fail(ODEINTR_E_STATE, "sharded adaptive stepping needs the temporally blocked dopri5 kernel (stencil radius <= 16)");
      return 0.0;
    }
    k_valid = true;
    const double a2 = 1.0 / 5.0, a3 = 3.0 / 10.0, a4 = 4.0 / 5.0, a5 = 8.0 / 9.0;
    const double b21 = 1.0 / 5.0;
    const double b31 = 3.0 / 40.0, b32 = 9.0 / 40.0;
    const double b41 = 44.0 / 45.0, b42 = -56.0 / 15.0, b43 = 32.0 / 9.0;
    const double b51 = 19372.0 / 6561.0, b52 = -25360.0 / 2187.0, b53 = 64448.0 / 6561.0, b54 = -212.0 / 729.0;
    const double b61 = 9017.0 / 3168.0, b62 = -355.0 / 33.0, b63 = 46732.0 / 5247.0, b64 = 49.0 / 176.0,
                 b65 = -5103.0 / 18656.0;
    const double c1 = 35.0 / 384.0, c3 = 500.0 / 1113.0, c4 = 125.0 / 192.0, c5 = -2187.0 / 6784.0, c6 = 11.0 / 84.0;
    const double dc1 = c1 - 5179.0 / 57600.0, dc3 = c3 - 7571.0 / 16695.0, dc4 = c4 - 393.0 / 640.0,
                 dc5 = c5 - (-92097.0 / 339200.0), dc6 = c6 - 187.0 / 2100.0, dc7 = -1.0 / 40.0;
    {  // x_tmp = in + dt*b21*dxdt
      const double* v[] = {d_in};
      const double c[] = {dt * b21};
      combo(nullptr, in, 1, v, c, 0.0, xtA, nullptr, t);
    }
    {  // k2 = f(x_tmp, t + dt*a2);  x_tmp = in + dt*b31*dxdt + dt*b32*k2
      const double* v[] = {d_in};
      const double c[] = {dt * b31};
      combo(xtA, in, 1, v, c, dt * b32, xtB, k2, t + dt * a2);
    }
    {  // k3
      const double* v[] = {d_in, k2};
      const double c[] = {dt * b41, dt * b42};
      combo(xtB, in, 2, v, c, dt * b43, xtA, k3, t + dt * a3);
    }
    {  // k4
      const double* v[] = {d_in, k2, k3};
      const double c[] = {dt * b51, dt * b52, dt * b53};
      combo(xtA, in, 3, v, c, dt * b54, xtB, k4, t + dt * a4);
    }
    {  // k5
      const double* v[] = {d_in, k2, k3, k4};
      const double c[] = {dt * b61, dt * b62, dt * b63, dt * b64};
      combo(xtB, in, 4, v, c, dt * b65, xtA, k5, t + dt * a5);
    }
    {  // k6 = f(x_tmp, t + dt);  out = in + dt*c1*dxdt + dt*c3*k3 + dt*c4*k4 + dt*c5*k5 + dt*c6*k6
      const double* v[] = {d_in, k3, k4, k5};
      const double c[] = {dt * c1, dt * c3, dt * c4, dt * c5};
      combo(xtA, in, 4, v, c, dt * c6, out, k6, t + dt);
    }
    cudaMemsetAsync(h->lat_err.p, 0, sizeof(unsigned long long), h->stream);
    {  // dxdt_out = f(out, t + dt);  xerr = dt*dc1*dxdt + dt*dc3*k3 + ... + dt*dc6*k6 + dt*dc7*dxdt_out, and its norm
      const double* v[] = {d_in, k3, k4, k5, k6};
      const double c[] = {dt * dc1, dt * dc3, dt * dc4, dt * dc5, dt * dc6};
      combo(out, nullptr, 5, v, c, dt * dc7, nullptr, d_out, t + dt, in, d_in, 1.0 * std::fabs(dt));
    }
    unsigned long long bits = 0;
    cudaMemcpyAsync(&bits, h->lat_err.p, sizeof(bits), cudaMemcpyDeviceToHost, h->stream);
    cudaStreamSynchronize(h->stream);
    double err;
    std::memcpy(&err, &bits, sizeof(err));
    return err;
  }

  // dense output (A.2) between (t_old, x_old, d_old) and (t_new, ., d_new) with the k3..k6 of the last step
  void calc_state(double t, double* x, const double* x_old, const double* d_old, double t_old, const double* d_new,
                  double t_new) {
    const double b1 = 35.0 / 384.0, b3 = 500.0 / 1113.0, b4 = 125.0 / 192.0, b5 = -2187.0 / 6784.0, b6 = 11.0 / 84.0;
    const double dt = (t_new - t_old);
    const double theta = (t - t_old) / dt;
    const double X1 = 5.0 * (2558722523.0 - 31403016.0 * theta) / 11282082432.0;
    const double X3 = 100.0 * (882725551.0 - 15701508.0 * theta) / 32700410799.0;
    const double X4 = 25.0 * (443332067.0 - 31403016.0 * theta) / 1880347072.0;
    const double X5 = 32805.0 * (23143187.0 - 3489224.0 * theta) / 199316789632.0;
    const double X6 = 55.0 * (29972135.0 - 7076736.0 * theta) / 822651844.0;
    const double X7 = 10.0 * (7414447.0 - 829305.0 * theta) / 29380423.0;
    const double theta_m_1 = theta - 1.0;
    const double theta_sq = theta * theta;
    const double A = theta_sq * (3.0 - 2.0 * theta);
    const double B = theta_sq * theta_m_1;
    const double C = theta_sq * theta_m_1 * theta_m_1;
    const double D = theta * theta_m_1 * theta_m_1;
    const double b1_theta = A * b1 - C * X1 + D;
    const double b3_theta = A * b3 + C * X3;
    const double b4_theta = A * b4 - C * X4;
    const double b5_theta = A * b5 + C * X5;
    const double b6_theta = A * b6 - C * X6;
    const double b7_theta = B + C * X7;
    const double* v[] = {d_old, k3, k4, k5, k6, d_new};
    const double c[] = {dt * b1_theta, dt * b3_theta, dt * b4_theta, dt * b5_theta, dt * b6_theta, dt * b7_theta};
    if (!k_valid && last_in) {
      // tiled path: repeat the accepted attempt (x_old, d_old, its t and dt) with the interpolation in its last stage
      tiled_attempt(last_in, last_din, last_t, nullptr, nullptr, last_dt, x, c);
      return;
    }
    combo(nullptr, x_old, 6, v, c, 0.0, x, nullptr, t);
  }

  // dxdt = f(x, t)
  void rhs(const double* x, double* dxdt, double t) {
    if (!slab) { combo(x, nullptr, 0, nullptr, nullptr, 1.0, nullptr, dxdt, t); return; }
    // slab: the sharded stage kernel with acc = (-0.0) + 1.0 * k, i.e. k itself bit for bit, on the owned cells
    int erc = slab_refresh(h, x, nullptr);
    if (erc) { if (!rc) rc = erc; return; }
    odeintr_lattice_args a;
    std::memset(&a, 0, sizeof(a));
    a.pars = h->P ? h->pars.p : nullptr;
    a.n = h->sh_n; a.n_total = h->sh_total; a.c0 = h->sh_c0; a.ghost = h->sh_ghost; a.pitch = h->sh_pitch;
    a.xs = x + h->sh_pad; a.x = x + h->sh_pad; a.xt = nullptr;
    a.acc_in = h->lat_negzero.p + h->sh_pad; a.acc_out = dxdt + h->sh_pad;
    a.t = t; a.a_dt = 0.0; a.b_dt = 1.0;
    a.lo = h->sh_c0; a.hi = h->sh_c0 + h->sh_n; a.halo = h->sh_halo;
    erc = launch(h, h->k_lattice_sharded, h->stream, (unsigned)h->sm_count * 8u, 256u, &a);
    if (erc && !rc) rc = erc;
    slab_mark(h, dxdt, false);
  }
};

// default_step_adjuster on the host: glibc pow, i.e. exactly what the CPU reference evaluates
inline double lat_decrease_step(double dt, double error, int error_order) {
  dt *= std::max(9.0 / 10.0 * std::pow(error, -1.0 / (error_order - 1)), 1.0 / 5.0);
  return dt;
}
inline double lat_increase_step(double dt, double error, int stepper_order) {
  if (error < 0.5) {
    error = std::max(std::pow(5.0, -static_cast<double>(stepper_order)), error);
    dt *= 9.0 / 10.0 * std::pow(error, -1.0 / stepper_order);
  }
  return dt;
}

// runge_kutta_cash_karp54 over device vectors: the six stages as combo launches, in the operation order of the
// per-trajectory stepper (kernels/steppers_plain.cuh rk54_stepper, adaptive_steppers.cuh rk54_controlled)
struct lattice_ck54 {
  lattice_dopri5 core;  // combo() and the work vectors k2..k6, xtA, xtB
  double* k1 = nullptr;
  // x -> xnew (5th-order solution); with_error: returns default_error_checker's max norm of the embedded estimate
  double stages(const double* x, double t, double dt, double* xnew, bool with_error) {
    const double a21 = 1.0 / 5.0;
    const double a31 = 3.0 / 40.0, a32 = 9.0 / 40.0;
    const double a41 = 3.0 / 10.0, a42 = -9.0 / 10.0, a43 = 6.0 / 5.0;
    const double a51 = -11.0 / 54.0, a52 = 5.0 / 2.0, a53 = -70.0 / 27.0, a54 = 35.0 / 27.0;
    const double a61 = 1631.0 / 55296.0, a62 = 175.0 / 512.0, a63 = 575.0 / 13824.0,
                 a64 = 44275.0 / 110592.0, a65 = 253.0 / 4096.0;
    const double b1 = 37.0 / 378.0, b3 = 250.0 / 621.0, b4 = 125.0 / 594.0, b6 = 512.0 / 1771.0;
    const double db1 = b1 - 2825.0 / 27648.0, db3 = b3 - 18575.0 / 48384.0, db4 = b4 - 13525.0 / 55296.0,
                 db5 = -277.0 / 14336.0, db6 = b6 - 1.0 / 4.0;
    const double c2 = 1.0 / 5.0, c3 = 3.0 / 10.0, c4 = 3.0 / 5.0, c5 = 1.0, c6 = 7.0 / 8.0;
    double *k2 = core.k2, *k3 = core.k3, *k4 = core.k4, *k5 = core.k5, *k6 = core.k6, *xtA = core.xtA, *xtB = core.xtB;
    core.rhs(x, k1, t);
    {
      const double* v[] = {k1};
      const double c[] = {a21 * dt};
      core.combo(nullptr, x, 1, v, c, 0.0, xtA, nullptr, t);
    }
    {  // k2; x + a31 k1 + a32 k2
      const double* v[] = {k1};
      const double c[] = {a31 * dt};
      core.combo(xtA, x, 1, v, c, a32 * dt, xtB, k2, t + c2 * dt);
    }
    {  // k3
      const double* v[] = {k1, k2};
      const double c[] = {a41 * dt, a42 * dt};
      core.combo(xtB, x, 2, v, c, a43 * dt, xtA, k3, t + c3 * dt);
    }
    {  // k4
      const double* v[] = {k1, k2, k3};
      const double c[] = {a51 * dt, a52 * dt, a53 * dt};
      core.combo(xtA, x, 3, v, c, a54 * dt, xtB, k4, t + c4 * dt);
    }
    {  // k5
      const double* v[] = {k1, k2, k3, k4};
      const double c[] = {a61 * dt, a62 * dt, a63 * dt, a64 * dt};
      core.combo(xtB, x, 4, v, c, a65 * dt, xtA, k5, t + c5 * dt);
    }
    {  // k6; xnew = x + b1 k1 + b3 k3 + b4 k4 + b6 k6
      const double* v[] = {k1, k3, k4};
      const double c[] = {b1 * dt, b3 * dt, b4 * dt};
      core.combo(xtA, x, 3, v, c, b6 * dt, xnew, k6, t + c6 * dt);
    }
    if (!with_error) return 0.0;
    cudaMemsetAsync(core.h->lat_err.p, 0, sizeof(unsigned long long), core.h->stream);
    {  // xerr = db1 k1 + db3 k3 + db4 k4 + db5 k5 + db6 k6 and its norm against (x, k1)
      const double* v[] = {k1, k3, k4, k5, k6};
      const double c[] = {db1 * dt, db3 * dt, db4 * dt, db5 * dt, db6 * dt};
      core.combo(nullptr, nullptr, 5, v, c, 0.0, nullptr, nullptr, t, x, k1, 1.0 * std::fabs(dt));
    }
    unsigned long long bits = 0;
    cudaMemcpyAsync(&bits, core.h->lat_err.p, sizeof(bits), cudaMemcpyDeviceToHost, core.h->stream);
    cudaStreamSynchronize(core.h->stream);
    double err;
    std::memcpy(&err, &bits, sizeof(err));
    return err;
  }
};

// controlled_runge_kutta< runge_kutta_cash_karp54 > over device vectors (method "rk54_a"); not FSAL: f(x) is
// evaluated at the start of every attempt
struct lattice_ck54_controlled {
  static constexpr bool is_dense = false;
  lattice_dopri5 core;
  double *k1 = nullptr, *xnew = nullptr;
  lat_counters cnt;
  lat_counters& counters() { return cnt; }
  void restart() {}
  bool try_step(lat_sys&, lat_state& x, double& t, double& dt) {
    lattice_ck54 ck;
    ck.core = core;
    ck.k1 = k1;
    const double err = ck.stages(x.p, t, dt, xnew, true);
    if (ck.core.rc && !core.rc) core.rc = ck.core.rc;
    if (err > 1.0) {
      dt = lat_decrease_step(dt, err, 4);
      cnt.rejected++;
      return false;
    }
    t += dt;
    dt = lat_increase_step(dt, err, 5);
    cnt.accepted++;
    cudaMemcpyAsync(x.p, xnew, core.h->lat_vec_len * sizeof(double), cudaMemcpyDeviceToDevice, core.h->stream);
    return true;
  }
};

// runge_kutta_fehlberg78 over device vectors, tableau driven like kernels/rk78.cuh: zero entries are skipped, what
// remains is Boost's left-to-right scale_sum with coefficient * dt formed first.  Launch s computes k_s = f(xt_s) and
// forms the next stage state from the stored k's and (when its weight is non-zero) the fresh one.
struct lattice_rkf78 {
  lattice_dopri5 core;  // combo()
  double* k[13] = {nullptr};
  double* xt[2] = {nullptr, nullptr};
  double stages(const double* x, double t, double dt, double* xnew, bool with_error) {
    static const double a[13][12] = {
        {0},
        {2.0 / 27.0},
        {1.0 / 36.0, 1.0 / 12.0},
        {1.0 / 24.0, 0.0, 1.0 / 8.0},
        {5.0 / 12.0, 0.0, -25.0 / 16.0, 25.0 / 16.0},
        {1.0 / 20.0, 0.0, 0.0, 1.0 / 4.0, 1.0 / 5.0},
        {-25.0 / 108.0, 0.0, 0.0, 125.0 / 108.0, -65.0 / 27.0, 125.0 / 54.0},
        {31.0 / 300.0, 0.0, 0.0, 0.0, 61.0 / 225.0, -2.0 / 9.0, 13.0 / 900.0},
        {2.0, 0.0, 0.0, -53.0 / 6.0, 704.0 / 45.0, -107.0 / 9.0, 67.0 / 90.0, 3.0},
        {-91.0 / 108.0, 0.0, 0.0, 23.0 / 108.0, -976.0 / 135.0, 311.0 / 54.0, -19.0 / 60.0, 17.0 / 6.0, -1.0 / 12.0},
        {2383.0 / 4100.0, 0.0, 0.0, -341.0 / 164.0, 4496.0 / 1025.0, -301.0 / 82.0, 2133.0 / 4100.0, 45.0 / 82.0,
         45.0 / 164.0, 18.0 / 41.0},
        {3.0 / 205.0, 0.0, 0.0, 0.0, 0.0, -6.0 / 41.0, -3.0 / 205.0, -3.0 / 41.0, 3.0 / 41.0, 6.0 / 41.0, 0.0},
        {-1777.0 / 4100.0, 0.0, 0.0, -341.0 / 164.0, 4496.0 / 1025.0, -289.0 / 82.0, 2193.0 / 4100.0, 51.0 / 82.0,
         33.0 / 164.0, 12.0 / 41.0, 0.0, 1.0}};
    static const double c[13] = {0.0, 2.0 / 27.0, 1.0 / 9.0, 1.0 / 6.0, 5.0 / 12.0, 1.0 / 2.0, 5.0 / 6.0,
                                 1.0 / 6.0, 2.0 / 3.0, 1.0 / 3.0, 1.0, 0.0, 1.0};
    static const double b[13] = {0.0, 0.0, 0.0, 0.0, 0.0, 34.0 / 105.0, 9.0 / 35.0, 9.0 / 35.0, 9.0 / 280.0,
                                 9.0 / 280.0, 0.0, 41.0 / 840.0, 41.0 / 840.0};
    static const double db[13] = {0.0 - 41.0 / 840.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
                                  0.0 - 41.0 / 840.0, 41.0 / 840.0, 41.0 / 840.0};
    // weights of the state formed after k_s is known: row s + 1 of a, or b after the last stage
    auto next_state = [&](int s, const double* xs, double* out) {
      const double* w = s + 1 < 13 ? a[s + 1] : b;
      const double* v[12];
      double cc[12];
      int n = 0;
      for (int j = 0; j < s; ++j)
        if (w[j] != 0.0) { v[n] = k[j]; cc[n] = w[j] * dt; ++n; }
      const double ck = w[s] != 0.0 ? w[s] * dt : 0.0;
      core.combo(xs, x, n, v, cc, ck, out, k[s], t + c[s] * dt);
    };
    core.rhs(x, k[0], t);
    {  // xt_1 = x + a[1][0] * dt * k_0
      const double* v[] = {k[0]};
      const double cc[] = {a[1][0] * dt};
      core.combo(nullptr, x, 1, v, cc, 0.0, xt[1], nullptr, t);
    }
    for (int s = 1; s < 13; ++s) next_state(s, xt[s & 1], s + 1 < 13 ? xt[(s + 1) & 1] : xnew);
    if (!with_error) return 0.0;
    cudaMemsetAsync(core.h->lat_err.p, 0, sizeof(unsigned long long), core.h->stream);
    {
      const double* v[12];
      double cc[12];
      int n = 0;
      for (int j = 0; j < 13; ++j)
        if (db[j] != 0.0) { v[n] = k[j]; cc[n] = db[j] * dt; ++n; }
      core.combo(nullptr, nullptr, n, v, cc, 0.0, nullptr, nullptr, t, x, k[0], 1.0 * std::fabs(dt));
    }
    unsigned long long bits = 0;
    cudaMemcpyAsync(&bits, core.h->lat_err.p, sizeof(bits), cudaMemcpyDeviceToHost, core.h->stream);
    cudaStreamSynchronize(core.h->stream);
    double err;
    std::memcpy(&err, &bits, sizeof(err));
    return err;
  }
};

// controlled_runge_kutta< runge_kutta_fehlberg78 > (method "rk78_a"): error order 7, stepper order 8
struct lattice_rkf78_controlled {
  static constexpr bool is_dense = false;
  lattice_rkf78 rk;
  double* xnew = nullptr;
  lat_counters cnt;
  lat_counters& counters() { return cnt; }
  void restart() {}
  bool try_step(lat_sys&, lat_state& x, double& t, double& dt) {
    const double err = rk.stages(x.p, t, dt, xnew, true);
    if (err > 1.0) {
      dt = lat_decrease_step(dt, err, 7);
      cnt.rejected++;
      return false;
    }
    t += dt;
    dt = lat_increase_step(dt, err, 8);
    cnt.accepted++;
    cudaMemcpyAsync(x.p, xnew, rk.core.h->lat_vec_len * sizeof(double), cudaMemcpyDeviceToDevice, rk.core.h->stream);
    return true;
  }
};

// controlled_runge_kutta< runge_kutta_dopri5 > over device vectors (method "rk5_a")
struct lattice_dopri5_controlled {
  static constexpr bool is_dense = false;
  lattice_dopri5 core;
  double *dxdt, *xnew, *dnew;
  bool first = true;
  lat_counters cnt;
  lat_counters& counters() { return cnt; }
  void restart() { first = true; }

  bool try_step_full(const double* in, const double* d_in, double& t, double* out, double* d_out, double& dt) {
    const double err = core.step_with_error(in, d_in, t, out, d_out, dt);
    if (err > 1.0) {
      dt = lat_decrease_step(dt, err, 4);
      cnt.rejected++;
      return false;
    }
    t += dt;
    dt = lat_increase_step(dt, err, 5);
    cnt.accepted++;
    return true;
  }
  bool try_step(lat_sys&, lat_state& x, double& t, double& dt) {
    if (first) { core.rhs(x.p, dxdt, t); first = false; }
    const bool ok = try_step_full(x.p, dxdt, t, xnew, dnew, dt);
    if (ok) {
      // the new state and its derivative trade places with the old ones (no copy: the caller moves the final state
      // home when the loops are over); ghost-cell freshness is tracked per vector and travels with the pointers
      std::swap(x.p, xnew);
      std::swap(dxdt, dnew);
    }
    return ok;
  }
};

// dense_output_runge_kutta< controlled_runge_kutta< runge_kutta_dopri5 > > over device vectors (method "rk5_i")
struct lattice_dopri5_dense {
  static constexpr bool is_dense = true;
  lattice_dopri5_controlled ctl;
  lat_state x_cur, x_old;
  double *d_cur, *d_old;
  double t = 0, t_old = 0, dt = 0;
  bool deriv_ready = false;
  lat_counters& counters() { return ctl.cnt; }
  void initialize(const lat_state& x0, double t0, double dt0) {
    x_cur = x0;  // value copy (no-op when x0 is the current state itself)
    t = t0; dt = dt0; deriv_ready = false;
  }
  const lat_state& current_state() const { return x_cur; }
  bool do_step(lat_sys&) {
    if (!deriv_ready) { ctl.core.rhs(x_cur.p, d_cur, t); deriv_ready = true; }
    t_old = t;
    int fails = 0;
    if (ctl.cnt.exhausted()) return false;
    while (!ctl.try_step_full(x_cur.p, d_cur, t, x_old.p, d_old, dt)) {
      if (++fails >= ODEINTR_MAX_FAILED_STEPS) { ctl.cnt.status |= ODEINTR_STATUS_STEP_OVERFLOW; return false; }
      if (ctl.core.rc) return false;
    }
    std::swap(x_cur.p, x_old.p);  // toggle current / old like Boost's m_current_state_x1 flag
    std::swap(d_cur, d_old);
    return ctl.core.rc == 0;
  }
  void calc_state(double tq, lat_state& x) {
    ctl.core.calc_state(tq, x.p, x_old.p, d_old, t_old, d_cur, t);
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Bulirsch-Stoer on large systems (methods "bs" and "bsd", R/utils.R:57-58): Boost's bulirsch_stoer /
// bulirsch_stoer_dense_out over device vectors.  Every vector operation of modified_midpoint, the extrapolation
// tables, the finite differences and the Taylor polynomial of the dense output is one launch of odeintr_lattice_combo
// (left-to-right sums, Boost's order); the midpoint steps fuse the right-hand side with the update that uses it.  The
// order / step-size controller runs on the host (glibc pow, like the oracle).  The reference default-constructs these
// steppers: tolerances 1e-6 / 1e-6 whatever atol / rtol say.  Extrapolating an entry no user loop writes does not
// reproduce it bit for bit ((1 + c) x - c x), so these combos cover every entry of the vectors (`all`).
// Same control flow as the per-trajectory steppers kernels/bulirsch_stoer.cuh, bulirsch_stoer_dense.cuh.
struct lattice_bs_common {
  static constexpr int K_MAX = 8;
  lattice_dopri5 core;
  lat_counters cnt;
  bool dense_sequence = false;  // interval sequence 2, 6, 10, ... (bsd) instead of 2, 4, 6, ... (bs)
  bool last_step_rejected = false, first = true;
  int k_opt = 4;
  double coeff[K_MAX + 1][K_MAX + 1];
  long long seq[K_MAX + 1], cost[K_MAX + 1];
  double facmin_table[K_MAX + 1];
  double* table[K_MAX] = {nullptr};
  double* mm[3] = {nullptr, nullptr, nullptr};

  // a work vector of lat_vec_len doubles from the handle's pool (kept across calls)
  double* vec() {
    odeintr_system_t h = core.h;
    // handed out zeroed: a general snippet's right-hand side leaves the entries it does not write alone, and in the
    // reference those are the zeros of a freshly resized vector
    if (h->lat_pool_used < h->lat_pool.size()) {
      double* q = h->lat_pool[h->lat_pool_used++];
      cudaMemsetAsync(q, 0, h->lat_pool_len * sizeof(double), h->stream);
      return q;
    }
    double* p = nullptr;
    if (cudaMalloc(&p, h->lat_pool_len * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      if (!core.rc) core.rc = fail(ODEINTR_E_CUDA, "out of device memory for the Bulirsch-Stoer work vectors");
      return h->lat_pool.empty() ? nullptr : h->lat_pool[0];  // keep the launches harmless; the call fails with core.rc
    }
    cudaMemsetAsync(p, 0, h->lat_pool_len * sizeof(double), h->stream);
    h->lat_pool.push_back(p);
    h->lat_pool_used++;
    return p;
  }
  void setup(bool dense) {
    dense_sequence = dense;
    core.eps_abs_o = 1e-6; core.eps_rel_o = 1e-6; core.all = 1;
    for (int i = 0; i <= K_MAX; ++i) {
      seq[i] = dense ? 2 + 4 * i : 2 * (i + 1);
      cost[i] = i == 0 ? seq[i] : cost[i - 1] + seq[i];
      for (int k = 0; k < i; ++k) {
        const double r = static_cast<double>(seq[i]) / static_cast<double>(seq[k]);
        coeff[i][k] = 1.0 / (r * r - 1.0);
      }
      facmin_table[i] = std::pow(0.02, 1.0 / static_cast<double>(2 * i + 1));
    }
    for (int i = 0; i < 3; ++i) mm[i] = vec();
  }
  double* table_row(int j) {
    if (!table[j]) table[j] = vec();
    return table[j];
  }
  void copy(double* dst, const double* src) {
    cudaMemcpyAsync(dst, src, core.h->lat_vec_len * sizeof(double), cudaMemcpyDeviceToDevice, core.h->stream);
  }
  // out = c0 * a + c1 * b  (out may be a or b: every entry is read before it is written, by the same thread)
  void lin2(double* out, double c0, const double* a, double c1, const double* b) {
    const double* v[] = {a, b};
    const double c[] = {c0, c1};
    core.combo(nullptr, nullptr, 2, v, c, 0.0, out, nullptr, 0.0);
  }
  // modified_midpoint(_dense_out)::do_step: `steps` substeps from (in, dxdt) over dt; derivs / x_mp only for bsd
  void midpoint(const double* in, const double* dxdt, double t, double* out, double dt, int steps, double* x_mp,
                double* const* derivs) {
    const double h = dt / static_cast<double>(steps);
    const double h2 = 2.0 * h;
    double th = t + h;
    {
      const double* v[] = {dxdt};
      const double c[] = {h};
      core.combo(nullptr, in, 1, v, c, 0.0, mm[0], nullptr, t);  // x1 = in + h * dxdt
    }
    const double* x0 = in;
    double* x1 = mm[0];
    int i1 = 0, i0 = -1;
    if (x_mp && steps == 2) copy(x_mp, x1);
    for (int i = 1; i != steps; ++i) {
      int in_ = 0;
      while (in_ == i1 || in_ == i0) ++in_;
      // d = f(x1, th); x1' = x0 + h2 * d; x0' = x1 -- the derivative and the update it feeds in one launch
      core.combo(x1, x0, 0, nullptr, nullptr, h2, mm[in_], derivs ? derivs[i - 1] : nullptr, th);
      x0 = x1; i0 = i1;
      x1 = mm[in_]; i1 = in_;
      if (x_mp && i == steps / 2 - 1) copy(x_mp, x1);
      th += h;
    }
    const double* v[] = {x0, x1};
    const double c[] = {1.0 / 2.0, 1.0 / 2.0};
    core.combo(x1, nullptr, 2, v, c, (1.0 / 2.0) * h, out, derivs ? derivs[steps - 1] : nullptr, th);
  }
  void extrapolate(int k, double* xest) {
    for (int j = k - 1; j > 0; --j) lin2(table[j - 1], 1.0 + coeff[k][j], table[j], -coeff[k][j], table[j - 1]);
    lin2(xest, 1.0 + coeff[k][0], table[0], -coeff[k][0], xest);
  }
  // default_error_checker::error(in, dxdt, out - table[0], dt)
  double error_of(const double* in, const double* dxdt, const double* out, double dt) {
    cudaMemsetAsync(core.h->lat_err.p, 0, sizeof(unsigned long long), core.h->stream);
    const double* v[] = {out, table[0]};
    const double c[] = {1.0, -1.0};
    core.combo(nullptr, nullptr, 2, v, c, 0.0, nullptr, nullptr, 0.0, in, dxdt, 1.0 * std::fabs(dt));
    unsigned long long bits = 0;
    cudaMemcpyAsync(&bits, core.h->lat_err.p, sizeof(bits), cudaMemcpyDeviceToHost, core.h->stream);
    cudaStreamSynchronize(core.h->stream);
    double err;
    std::memcpy(&err, &bits, sizeof(err));
    return err;
  }
  double calc_h_opt(double h, double error, int k) const {
    const double STEPFAC1 = 0.65, STEPFAC2 = 0.94, STEPFAC4 = 4.0;
    const double expo = 1.0 / (2 * k + 1);
    const double facmin = facmin_table[k];
    double fac;
    if (error == 0.0) fac = 1.0 / facmin;
    else {
      fac = STEPFAC2 / std::pow(error / STEPFAC1, expo);
      fac = std::max(facmin / STEPFAC4, std::min(1.0 / facmin, fac));
    }
    return h * fac;
  }
  bool should_reject(double error, int k) const {
    if (k == k_opt - 1) {
      const double d = static_cast<double>(seq[k_opt] * seq[k_opt + 1] / (seq[0] * seq[0]));  // (integer arithmetic, as in Boost)
      return error > d * d;
    } else if (k == k_opt) {
      const double d = static_cast<double>(seq[k_opt] / seq[0]);
      return error > d * d;
    }
    return error > 1.0;
  }
  // the order / step-size decisions after column k of the tableau (identical in bs and bsd); true = leave the k loop
  bool decide(int k, double error, const double* h_opt, const double* work, bool& reject, double& new_h) {
    const double KFAC2 = 0.9;
    if ((k == k_opt - 1) || first) {  // convergence before k_opt
      if (error < 1.0) {
        reject = false;
        if ((work[k] < KFAC2 * work[k - 1]) || (k_opt <= 2)) {
          k_opt = std::min(K_MAX - 1, std::max(2, k + 1));
          // (Boost writes the two steppers' products differently: bs scales by the ratio, bsd multiplies, then divides)
          if (dense_sequence) new_h = h_opt[k] * static_cast<double>(cost[k + 1]) / static_cast<double>(cost[k]);
          else { new_h = h_opt[k]; new_h *= static_cast<double>(cost[k + 1]) / static_cast<double>(cost[k]); }
        } else {
          k_opt = std::min(K_MAX - 1, std::max(2, k));
          new_h = h_opt[k];
        }
        return true;
      } else if (should_reject(error, k) && !first) {
        reject = true;
        new_h = h_opt[k];
        return true;
      }
    }
    if (k == k_opt) {  // convergence at k_opt
      if (error < 1.0) {
        reject = false;
        if (work[k - 1] < KFAC2 * work[k]) {
          k_opt = std::max(2, k_opt - 1);
          new_h = h_opt[k_opt];
        } else if ((work[k] < KFAC2 * work[k - 1]) && !last_step_rejected) {
          k_opt = std::min(K_MAX - 1, k_opt + 1);
          if (dense_sequence) new_h = h_opt[k] * static_cast<double>(cost[k_opt]) / static_cast<double>(cost[k]);
          else { new_h = h_opt[k]; new_h *= static_cast<double>(cost[k_opt]) / static_cast<double>(cost[k]); }
        } else
          new_h = h_opt[k_opt];
        return true;
      } else if (should_reject(error, k)) {
        reject = true;
        new_h = h_opt[k_opt];
        return true;
      }
    }
    if (k == k_opt + 1) {  // convergence at k_opt + 1
      if (error < 1.0) {
        reject = false;
        if (work[k - 2] < KFAC2 * work[k - 1]) k_opt = std::max(2, k_opt - 1);
        if ((work[k] < KFAC2 * work[k_opt]) && !last_step_rejected) k_opt = std::min(K_MAX - 1, k);
        new_h = h_opt[k_opt];
      } else {
        reject = true;
        new_h = h_opt[k_opt];
      }
      return true;
    }
    return false;
  }
};

// bulirsch_stoer< state_type, double, state_type, double, openmp_range_algebra > (controlled stepper, method "bs")
struct lattice_bs_controlled : lattice_bs_common {
  static constexpr bool is_dense = false;
  double *dxdt = nullptr, *xnew = nullptr;
  double dt_last = 0.0;
  lat_counters& counters() { return cnt; }
  void reset() { first = true; last_step_rejected = false; k_opt = 4; }
  void restart() { reset(); dt_last = 0.0; }  // integrate_const hands every interval a fresh copy of the stepper
  void init() { setup(false); dxdt = vec(); xnew = vec(); }
  bool try_step(lat_sys&, lat_state& x, double& t, double& dt) {
    core.rhs(x.p, dxdt, t);
    if (dt != dt_last) reset();  // step size changed from outside -> reset
    bool reject = true;
    double h_opt[K_MAX + 1], work[K_MAX + 1];
    double new_h = dt;
    for (int k = 0; k <= k_opt + 1; ++k) {
      if (k == 0) {
        midpoint(x.p, dxdt, t, xnew, dt, (int)seq[k], nullptr, nullptr);
      } else {
        midpoint(x.p, dxdt, t, table_row(k - 1), dt, (int)seq[k], nullptr, nullptr);
        extrapolate(k, xnew);
        const double error = error_of(x.p, dxdt, xnew, dt);
        if (core.rc) { cnt.budget = 0; return true; }  // the loops stop at their next budget check; the caller reports core.rc
        h_opt[k] = calc_h_opt(dt, error, k);
        work[k] = static_cast<double>(cost[k]) / h_opt[k];
        if (decide(k, error, h_opt, work, reject, new_h)) break;
      }
    }
    if (!reject) t += dt;
    if (!last_step_rejected || odeintr_b200::less_with_sign(new_h, dt, dt)) {
      dt_last = new_h;
      dt = new_h;
    }
    last_step_rejected = reject;
    first = false;
    if (reject) { cnt.rejected++; return false; }
    copy(x.p, xnew);
    cnt.accepted++;
    return true;
  }
};

// bulirsch_stoer_dense_out< ... > (dense-output stepper, method "bsd"): the derivatives of every midpoint run are kept,
// their finite differences at the interval centre extrapolated, the dense output is the Taylor polynomial about the
// centre in theta = -1 ... 1
struct lattice_bsd_dense : lattice_bs_common {
  static constexpr bool is_dense = true;
  lat_state x_cur, x_old;
  double *d_cur = nullptr, *d_old = nullptr;
  double t = 0, t_old = 0, dt = 0;
  int k_final = 0;
  double* mp_states[K_MAX + 1] = {nullptr};
  std::vector<double*> derivs[K_MAX + 1];
  std::vector<double*> diffs[2 * K_MAX + 2];
  lat_counters& counters() { return cnt; }
  void init() {
    setup(true);
    x_cur.h = x_old.h = core.h;
    x_cur.p = vec(); x_old.p = vec(); d_cur = vec(); d_old = vec();
    for (int i = 0; i <= K_MAX; ++i) derivs[i].assign((size_t)seq[i], nullptr);
    int num = 1;
    for (int i = 2 * K_MAX + 1; i >= 0; --i) {
      diffs[i].assign((size_t)num, nullptr);
      num += (i + 1) % 2;
    }
  }
  void reset() { first = true; last_step_rejected = false; }
  void initialize(const lat_state& x0, double t0, double dt0) {
    x_cur = x0;  // value copy (no-op when x0 is the current state itself)
    t = t0; dt = dt0;
    reset();
  }
  const lat_state& current_state() const { return x_cur; }
  double* mp(int k) { if (!mp_states[k]) mp_states[k] = vec(); return mp_states[k]; }
  double* dv(int k, int i) { if (!derivs[k][i]) derivs[k][i] = vec(); return derivs[k][i]; }
  double* df(int kappa, int j) { if (!diffs[kappa][j]) diffs[kappa][j] = vec(); return diffs[kappa][j]; }
  static double binomial(int n, int k) {
    double r = 1.0;
    for (int i = 1; i <= k; ++i) r = r * static_cast<double>(n - k + i) / static_cast<double>(i);
    return std::floor(r + 0.5);
  }
  // out = sum of c_j v_j, left to right, in launches of at most 12 terms (later launches continue from out)
  void sum_terms(double* out, const double* base, const std::vector<const double*>& v, const std::vector<double>& c) {
    size_t done = 0;
    const double* b = base;
    while (true) {
      const double* vv[12];
      double cc[12];
      int n = 0;
      if (!b && done > 0) { vv[n] = out; cc[n] = 1.0; ++n; }  // 1.0 * out: exact
      while (done < v.size() && n < 12) { vv[n] = v[done]; cc[n] = c[done]; ++n; ++done; }
      core.combo(nullptr, b, n, vv, cc, 0.0, out, nullptr, 0.0);
      if (done >= v.size()) break;
      b = nullptr;
      if (base) b = out;  // continue from what is there: base + ... (the base enters with weight one)
    }
  }
  void calculate_finite_difference(int j, int kappa, double fac, const double* dxdt) {
    const int m = static_cast<int>(seq[j] / 2) - 1;
    if (kappa == 0) {  // no calculation required for the 0th derivative of f
      const double* v[] = {dv(j, m)};
      const double c[] = {fac};
      core.combo(nullptr, nullptr, 1, v, c, 0.0, df(0, j), nullptr, 0.0);
      return;
    }
    const int j_diffs = j - kappa / 2;
    std::vector<const double*> v;
    std::vector<double> c;
    v.push_back(dv(j, m + kappa)); c.push_back(fac);
    double sign = -1.0;
    int cc = 1;
    for (int i = m + kappa - 2; i >= m - kappa; i -= 2) {
      if (i >= 0) { v.push_back(dv(j, i)); c.push_back(sign * fac * binomial(kappa, cc)); }
      else { v.push_back(dxdt); c.push_back(sign * fac); }
      sign *= -1;
      ++cc;
    }
    sum_terms(df(kappa, j_diffs), nullptr, v, c);
  }
  template <class Get>
  void extrapolate_dense_out(int k, Get get, int order_start_index) {
    for (int j = k; j > 1; --j) {
      const double cf = coeff[k + order_start_index][j + order_start_index - 1];
      lin2(get(j - 1), 1.0 + cf, get(j), -cf, get(j - 1));
    }
    const double cf = coeff[k + order_start_index][order_start_index];
    lin2(get(0), 1.0 + cf, get(1), -cf, get(0));
  }
  void prepare_dense_output(int k, const double* dxdt_start, double dt_) {
    for (int j = 0; j <= k; ++j) {
      const double d = static_cast<double>(seq[j]) / (2.0 * dt_);
      double f = 1.0;
      for (int kappa = 0; kappa <= 2 * j + 1; ++kappa) {
        calculate_finite_difference(j, kappa, f, dxdt_start);
        f *= d;
      }
      if (j > 0) extrapolate_dense_out(j, [&](int q) { return mp(q); }, 0);
    }
    double d = dt_ / 2;
    for (int kappa = 0; kappa <= 2 * k + 1; ++kappa) {
      for (int j = 1; j <= (k - kappa / 2); ++j) extrapolate_dense_out(j, [&](int q) { return df(kappa, q); }, kappa / 2);
      const double* v[] = {df(kappa, 0)};
      const double c[] = {d};
      core.combo(nullptr, nullptr, 1, v, c, 0.0, df(kappa, 0), nullptr, 0.0);  // (dt/2)^(kappa+1) / (kappa+1)!
      d *= dt_ / (2 * (kappa + 2));
    }
  }
  bool try_step(const double* in, const double* dxdt, double& tt, double* out, double* dxdt_new, double& dtt) {
    bool reject = true;
    double h_opt[K_MAX + 1], work[K_MAX + 1];
    k_final = 0;
    double new_h = dtt;
    for (int k = 0; k <= k_opt + 1; ++k) {
      const int steps = (int)seq[k];
      double* dptr[40];
      for (int i = 0; i < steps; ++i) dptr[i] = dv(k, i);
      if (k == 0) {
        midpoint(in, dxdt, tt, out, dtt, steps, mp(k), dptr);
      } else {
        midpoint(in, dxdt, tt, table_row(k - 1), dtt, steps, mp(k), dptr);
        extrapolate(k, out);
        const double error = error_of(in, dxdt, out, dtt);
        if (core.rc) return true;
        h_opt[k] = calc_h_opt(dtt, error, k);
        work[k] = static_cast<double>(cost[k]) / h_opt[k];
        k_final = k;
        if (decide(k, error, h_opt, work, reject, new_h)) break;
      }
    }
    if (!reject) {
      core.rhs(out, dxdt_new, tt + dtt);  // dxdt for the next step and the dense output
      prepare_dense_output(k_final, dxdt, dtt);  // (no interpolation control in the default-constructed stepper)
      tt += dtt;
    }
    if (!last_step_rejected || (new_h < dtt)) dtt = new_h;
    last_step_rejected = reject;
    if (reject) cnt.rejected++; else cnt.accepted++;
    return !reject;
  }
  bool do_step(lat_sys&) {
    if (first) core.rhs(x_cur.p, d_cur, t);
    t_old = t;
    int fails = 0;
    if (cnt.exhausted()) return false;
    while (true) {
      const bool ok = try_step(x_cur.p, d_cur, t, x_old.p, d_old, dt);
      first = false;
      if (core.rc) return false;
      if (ok) break;
      if (++fails >= ODEINTR_MAX_FAILED_STEPS) { cnt.status |= ODEINTR_STATUS_STEP_OVERFLOW; return false; }
      if (cnt.exhausted()) return false;
    }
    std::swap(x_cur.p, x_old.p);
    std::swap(d_cur, d_old);
    return true;
  }
  void calc_state(double tq, lat_state& x) {
    const double theta = 2 * ((tq - t_old) / (t - t_old)) - 1;  // interpolation polynomial on theta = -1 ... 1
    std::vector<const double*> v;
    std::vector<double> c;
    double theta_pow = theta;
    for (int i = 0; i <= 2 * k_final + 1; ++i) {
      v.push_back(df(i, 0)); c.push_back(theta_pow);
      theta_pow *= theta;
    }
    sum_terms(x.p, mp(0), v, c);
  }
};

// observer of a large system: one record row = the whole state
struct lattice_observer {
  odeintr_system_t h;
  std::vector<double>* rec_t;
  size_t cap_rows;
  int rc = 0;
  void operator()(const lat_state& x, double t) {
    const size_t N = h->lat_vec_len, row = rec_t->size();
    if (row >= cap_rows) {  // grow the record, keeping what is there
      const size_t new_cap = std::max<size_t>(16, cap_rows * 2);
      double* bigger = nullptr;
      if (cudaMalloc(&bigger, new_cap * N * sizeof(double)) != cudaSuccess) { rc = ODEINTR_E_CUDA; return; }
      if (h->out.p && row)
        cudaMemcpyAsync(bigger, h->out.p, row * N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
      cudaStreamSynchronize(h->stream);
      if (h->out.p) cudaFree(h->out.p);
      h->out.p = bigger; h->out.cap = new_cap * N;
      cap_rows = new_cap;
    }
    cudaMemcpyAsync(h->out.p + row * N, x.p, N * sizeof(double), cudaMemcpyDeviceToDevice, h->stream);
    rec_t->push_back(t);
  }
};
