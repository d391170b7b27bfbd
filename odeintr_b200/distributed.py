"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on GPUs, gloo on CPU for
the host-logic tests).

* Ensembles (parameter sweeps, initial-condition batches) shard by instance index -- contiguous ranges, no
  data-path collective.  ``shard_range`` / ``EnsembleShard`` only compute who owns what.
* Large lattices shard spatially into slabs.  Every rank keeps ``ghost = halo * stages`` extra cells on both
  sides of its slab and refreshes them ONCE per Runge-Kutta step from its ring neighbours
  (``exchange_halos``: batched isend/irecv, i.e. ncclSend/ncclRecv inside one group on NCCL); the fused stage
  kernels then recompute a window that shrinks by the stencil radius per stage
  (``odeintr_lattice_shard_step``), so no communication happens between stages.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _abi
from .codegen import OdeintrError


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `total` units owned by `rank`: sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise OdeintrError("invalid rank / world size")
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def exchange_halos(x, ghost: int, rank: int, world: int, group=None, pad: Optional[int] = None, n_local: Optional[int] = None):
    """Periodic ring exchange of the ghost cells of a slab-sharded lattice.

    x: torch tensor [fields, pitch] (CPU with gloo, CUDA with NCCL), updated in place.  The owned cells of a field
    sit at [pad, pad + n_local), the ghost cells directly around them (pad defaults to ghost, n_local to
    pitch - 2 * pad).  The left ghost cells receive the left neighbour's last `ghost` owned cells, the right ghost
    cells the right neighbour's first `ghost` owned cells.  With one rank the lattice wraps onto itself."""
    import torch
    import torch.distributed as dist

    if ghost == 0:
        return
    pad = ghost if pad is None else pad
    n_local = x.shape[1] - 2 * pad if n_local is None else n_local
    if n_local < ghost:
        raise OdeintrError("slab too small for the halo width")
    lo, hi = pad, pad + n_local  # owned cells
    if world == 1:
        x[:, lo - ghost:lo] = x[:, hi - ghost:hi]
        x[:, hi:hi + ghost] = x[:, lo:lo + ghost]
        return
    left, right = (rank - 1) % world, (rank + 1) % world
    send_l = x[:, lo:lo + ghost].contiguous()
    send_r = x[:, hi - ghost:hi].contiguous()
    recv_l = torch.empty_like(send_l)
    recv_r = torch.empty_like(send_r)
    ops = [dist.P2POp(dist.isend, send_r, right, group), dist.P2POp(dist.irecv, recv_l, left, group),
           dist.P2POp(dist.isend, send_l, left, group), dist.P2POp(dist.irecv, recv_r, right, group)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    x[:, lo - ghost:lo] = recv_l
    x[:, hi:hi + ghost] = recv_r


class ShardedLattice:
    """A compile_sys_openmp system whose cells are split into one slab per rank (GPU).

    The state buffer is a torch CUDA tensor (torch owns the memory and talks NCCL); the stepping is done by the
    C-ABI on that memory.  Fixed-step rk4, periodic lattice, stencil radius `halo`."""

    def __init__(self, system, cells_total: int, fields: int = 1, halo: int = 1, rank: int = 0, world: int = 1,
                 group=None, device=None, native: bool = True, steps_per_exchange: int = 8):
        """native=True: the C-ABI owns an NCCL communicator and runs exchange + step loops itself
        (odeintr_lattice_shard_run); torch.distributed only ships the 128-byte NCCL unique id.  Slabs that step
        through the temporally blocked rk4 kernels (radius <= 32) keep ghost zones `steps_per_exchange` steps wide
        and exchange them only that often (odeintr_lattice_set_exchange_interval).
        native=False: halos are exchanged by torch.distributed P2P ops between the C-ABI's steps."""
        import torch

        if system.gen.family != "lattice":
            raise OdeintrError("ShardedLattice needs a system compiled with compile_sys_openmp")
        self.system, self.fields, self.halo = system, fields, halo
        self.rank, self.world, self.group = rank, world, group
        self.cells_total = cells_total
        self.begin, self.end = shard_range(cells_total, rank, world)
        self.n_local = self.end - self.begin
        self.lib = _abi.load()
        if self.lib.odeintr_lattice_set_exchange_interval(system._h, steps_per_exchange if native else 1) != 0:
            raise OdeintrError(_abi.last_error())
        self.ghost = int(self.lib.odeintr_lattice_ghost_width(system._h, halo))
        self.pad = int(self.lib.odeintr_lattice_pad_width(system._h, halo))
        self.pitch = int(self.lib.odeintr_lattice_pitch(system._h, halo, self.n_local))
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.x = torch.zeros((fields, self.pitch), dtype=torch.float64, device=self.device)
        # One dedicated stream orders everything: the halo copies and NCCL send/recv are issued by torch with this
        # stream current, the stage kernels are launched on the same stream by the C-ABI.
        self.stream = torch.cuda.Stream(device=self.device)
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        if self.lib.odeintr_set_stream(system._h, C.c_void_p(self.stream.cuda_stream)) != 0:
            raise OdeintrError(_abi.last_error())
        rc = self.lib.odeintr_lattice_shard(system._h, cells_total, self.begin, self.n_local, fields, halo,
                                            C.c_void_p(self.x.data_ptr()))
        if rc != 0:
            raise OdeintrError(_abi.last_error())
        self.t = 0.0
        self.steps = 0
        self.native = native
        if native:
            ident = torch.zeros(128, dtype=torch.uint8)
            if world > 1:
                import torch.distributed as dist

                if rank == 0:
                    buf = (C.c_ubyte * 128)()
                    if self.lib.odeintr_nccl_unique_id(buf, 128) != 0:
                        raise OdeintrError(_abi.last_error())
                    ident = torch.tensor(list(buf), dtype=torch.uint8)
                # the id travels over whatever backend the group has (NCCL needs a device tensor)
                backend = dist.get_backend(group)
                ident = ident.to(self.device) if backend == "nccl" else ident
                dist.broadcast(ident, src=0, group=group)
                ident = ident.cpu()
            raw = (C.c_ubyte * 128)(*ident.tolist())
            if self.lib.odeintr_lattice_comm_init(system._h, world, rank, raw, 128) != 0:
                raise OdeintrError(_abi.last_error())

    def set_state_from(self, fn):
        """fn(field, global_cell_indices: np.ndarray) -> values; every rank fills only its own slab."""
        import torch

        cells = np.arange(self.begin, self.end)
        with torch.cuda.stream(self.stream):
            for f in range(self.fields):
                vals = np.asarray(fn(f, cells), dtype=np.float64)
                self.x[f, self.pad:self.pad + self.n_local] = torch.from_numpy(vals).to(self.device)

    def step(self, n_steps: int, dt: float, t0: Optional[float] = None, exchange: bool = True):
        """n_steps rk4 steps.  exchange=False leaves the ghost cells to the caller (tests emulate several ranks
        inside one process that way)."""
        if t0 is not None:
            self.t, self.steps = t0, 0
            self._t0 = t0
        import torch

        t0 = getattr(self, "_t0", 0.0)
        if self.native and exchange:
            rc = self.lib.odeintr_lattice_shard_run(self.system._h, self.steps, n_steps, t0, dt)
            if rc != 0:
                raise OdeintrError(_abi.last_error())
            self.steps += n_steps
            self.t = t0 + self.steps * dt
            return
        with torch.cuda.stream(self.stream):
            for _ in range(n_steps):
                if exchange:
                    exchange_halos(self.x, self.ghost, self.rank, self.world, self.group, self.pad, self.n_local)
                rc = self.lib.odeintr_lattice_shard_step(self.system._h, self.t, dt)
                if rc != 0:
                    raise OdeintrError(_abi.last_error())
                self.steps += 1
                self.t = t0 + self.steps * dt  # integrate_const time rule

    # -- adaptive steppers (rk5_i, the default of compile_sys_openmp, and rk5_a): collective calls --------------------------
    def _record(self):
        rows = int(self.lib.odeintr_output_rows(self.system._h))
        t = np.empty(rows, dtype=np.float64)
        if self.lib.odeintr_get_output_times(self.system._h, t.ctypes.data_as(_abi.c_dbl_p), rows) != 0:
            raise OdeintrError(_abi.last_error())
        x = np.empty((rows, self.fields, self.n_local), dtype=np.float64)
        if self.lib.odeintr_lattice_shard_get_output(self.system._h, x.ctypes.data_as(_abi.c_dbl_p), x.size) != 0:
            raise OdeintrError(_abi.last_error())
        return t, x

    def integrate_times(self, times, step_size=1.0):
        """NAME_at() on the sharded lattice (every rank calls it with the same arguments): returns (Time, x) with x
        [rows, fields, n_local] = this rank's cells of every recorded row.  The state stays in the slab."""
        tt = np.ascontiguousarray(np.asarray(times, dtype=np.float64).reshape(-1))
        self.stream.synchronize()
        if self.lib.odeintr_integrate_times(self.system._h, tt.ctypes.data_as(_abi.c_dbl_p), tt.size, step_size) != 0:
            raise OdeintrError(_abi.last_error())
        return self._record()

    def integrate_const(self, duration, step_size=1.0, start=0.0):
        """NAME() on the sharded lattice: observer every step_size."""
        self.stream.synchronize()
        if self.lib.odeintr_integrate_const(self.system._h, start, start + duration, step_size, 1, 1) != 0:
            raise OdeintrError(_abi.last_error())
        return self._record()

    def no_record(self, duration, step_size=1.0, start=0.0):
        """NAME_no_record() on the sharded lattice: integrate_adaptive without observer; the slab holds the result."""
        self.stream.synchronize()
        if self.lib.odeintr_integrate_adaptive(self.system._h, start, start + duration, step_size, 0) != 0:
            raise OdeintrError(_abi.last_error())

    def stats(self):
        """(accepted, rejected) step counts of the last adaptive call: global, identical on every rank."""
        acc, rej, st = (C.c_longlong * 1)(), (C.c_longlong * 1)(), (C.c_int * 1)()
        if self.lib.odeintr_get_stats(self.system._h, acc, rej, st, 1) != 0:
            raise OdeintrError(_abi.last_error())
        return int(acc[0]), int(rej[0])

    def local_state(self):
        return self.x[:, self.pad:self.pad + self.n_local]

    def gather_state(self):
        """Full state [fields, cells_total] on every rank (for checks; not part of the timed path)."""
        import torch
        import torch.distributed as dist

        self.stream.synchronize()
        loc = self.local_state().contiguous()
        if self.world == 1:
            return loc.clone()
        sizes = [shard_range(self.cells_total, r, self.world) for r in range(self.world)]
        widest = max(e - b for b, e in sizes)  # all_gather needs equal shapes: pad the short slabs
        padded = torch.zeros((self.fields, widest), dtype=loc.dtype, device=loc.device)
        padded[:, :loc.shape[1]] = loc
        parts = [torch.empty_like(padded) for _ in range(self.world)]
        dist.all_gather(parts, padded, group=self.group)
        return torch.cat([p[:, :e - b] for p, (b, e) in zip(parts, sizes)], dim=1)
