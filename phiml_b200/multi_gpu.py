"""
Multi-GPU drivers for the path.  One process per GPU (``torchrun``), ``torch.distributed`` for the plumbing
(NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests).  The reference has no multi-GPU concept at all
(README.md:78 "does not currently allow mapping operations onto multiple GPUs"), so this is new design:

``solve_sharded``  independent systems (a batch of right-hand sides sharing one matrix, BASELINE config 3): the batch axis is
                   cut into contiguous shards, every rank solves its shard with the ordinary device loop.  No collective
                   inside the loop.  The only coupling the reference's semantics require is the batch-wide MINIMUM of the
                   squared tolerances (the (batch,batch) broadcast of stop_on_l2, phiml/backend/_linalg.py:26-29): one scalar
                   all-reduce(min) after the initial residual.

``SlabCG``         ONE large system (north-star 1024^3 case): row partition in x-slabs.  Rank g owns planes [x_g, x_{g+1})
                   and keeps one halo plane per neighbour.  Per CG iteration: two all-reduces of the per-system dot products
                   (d.q and r.M^-1 r; the kernels leave the local sums in a device buffer and the scalar step runs identically on
                   every rank afterwards) and one halo exchange of the residual (2 planes per rank over NVLink); the direction's
                   halo planes are recomputed locally by the fused kernel, so they never travel.
"""
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


# ----------------------------------------------------------------------------------------------------------------------
# partitioning helpers (pure host logic; covered by the gloo tests)
# ----------------------------------------------------------------------------------------------------------------------

def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``range(total)``: the first ``total % world`` ranks get one extra item."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def slab_bounds(nx: int, world: int, rank: int) -> Tuple[int, int]:
    lo, hi = shard_bounds(nx, world, rank)
    if hi - lo < 1:
        raise ValueError(f"slab partition needs at least one plane per rank (nx={nx}, world={world})")
    return lo, hi


def exchange_halo_planes(ext: torch.Tensor, nxl: int, plane: int, rank: int, world: int, group=None):
    """
    ``ext`` is a (batch, (nxl+2)*plane) tensor: owned planes 1..nxl, halo planes 0 and nxl+1.  Sends the first / last owned
    plane to the lower / upper neighbour and receives their boundary planes into the halos.  Physical boundaries keep zeros.
    """
    if world == 1:
        return
    v = ext.view(ext.shape[0], nxl + 2, plane)
    ops, keep = [], []
    if rank > 0:
        lo_send = v[:, 1, :].contiguous()
        lo_recv = torch.empty_like(lo_send)
        ops += [dist.P2POp(dist.isend, lo_send, rank - 1, group), dist.P2POp(dist.irecv, lo_recv, rank - 1, group)]
        keep.append((0, lo_recv))
    if rank < world - 1:
        hi_send = v[:, nxl, :].contiguous()
        hi_recv = torch.empty_like(hi_send)
        ops += [dist.P2POp(dist.isend, hi_send, rank + 1, group), dist.P2POp(dist.irecv, hi_recv, rank + 1, group)]
        keep.append((nxl + 1, hi_recv))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for idx, buf in keep:
        v[:, idx, :].copy_(buf)


# ----------------------------------------------------------------------------------------------------------------------
# independent systems
# ----------------------------------------------------------------------------------------------------------------------

def solve_sharded(method: str, lin, y_local: torch.Tensor, x0_local: torch.Tensor, rtol, atol, max_iter: np.ndarray, pre=None, group=None, generic_rules: bool = False):
    """
    Every rank passes ITS shard of the right-hand sides (``shard_bounds`` gives the split).  Returns the local SolveResult.
    Identical to one ``linear_solve`` over the whole batch: the tolerance every system has to reach is the minimum over ALL
    ranks' systems, exactly as in the single-process reference.  ``generic_rules`` selects the generic loops of
    phiml/backend/_linalg.py (what the NumPy backend runs) instead of the torch fast-path rules for un-preconditioned CG.
    """
    from . import solve as host
    from . import _engine as E
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    code, order, name = _method_code(method, pre, max_iter)
    if generic_rules and code == E.CG_TORCH:
        code, name = E.CG, f"Φ-ML CG ({host.BACKEND_NAME}) {host._pre_str(pre)}"
    if generic_rules and code == E.CG_ADAPTIVE_TORCH:
        code, name = E.CG_ADAPTIVE, f"Φ-ML CG-adaptive ({host.BACKEND_NAME}) "
    matrix, y, x0, rtol_t, atol_t, max_iter = host._prepare(lin, y_local, x0_local, rtol, atol, max_iter)
    pre = host._device_pre(pre, matrix)
    sm = matrix.product_twin
    nb = y.shape[0]
    mi = torch.as_tensor(np.broadcast_to(np.asarray(max_iter)[-1, :], (nb,)).astype(np.int32).copy(), device=y.device)
    solver = host.DeviceSolver(sm, pre, code, order, nb)
    solver.start(y, x0, rtol_t, atol_t, mi)
    if world > 1:
        # batch-wide stopping rule: sum of |y|^2 over ALL ranks' systems (generic cg_adaptive / bicg), then the minimum tolerance;
        # the first progress check is repeated with it
        ops = {'sum': dist.ReduceOp.SUM, 'min': dist.ReduceOp.MIN}
        host.combine_tolerances([solver], lambda t, op: dist.all_reduce(t, op=ops[op], group=group))
    solver.run(int(np.max(max_iter)))
    host.LAST_RUN['resident'] = solver.resident_used
    its, fev, conv, div = solver.result()
    return host.SolveResult(name, solver.x, solver.r, its, fev, conv, div, [""] * nb)


def _method_code(method: str, pre, max_iter):
    from . import _engine as E
    from .solve import BACKEND_NAME, _pre_str
    single = np.asarray(max_iter).shape[0] == 1
    if method == 'CG':
        if pre or not single:
            return E.CG, 0, f"Φ-ML CG ({BACKEND_NAME}) {_pre_str(pre)}"
        return E.CG_TORCH, 0, "Φ-ML CG (PyTorch*)"
    if method in ('auto', 'CG-adaptive'):
        if pre:
            return E.CG, 0, f"Φ-ML CG ({BACKEND_NAME}) {_pre_str(pre)}"
        return (E.CG_ADAPTIVE_TORCH, 0, "Φ-ML CG (PyTorch*)") if single else (E.CG_ADAPTIVE, 0, f"Φ-ML CG-adaptive ({BACKEND_NAME}) ")
    if method == 'biCG-stab':
        return E.BICGSTAB, 1, f"Φ-ML biCG-stab {_pre_str(pre)}"
    if method.startswith('biCG-stab('):
        order = int(method[len('biCG-stab('):-1])
        return E.BICGSTAB, order, (f"Φ-ML biCG-stab {_pre_str(pre)}" if order == 1 else f"Φ-ML biCG-stab({order}) ({BACKEND_NAME}) {_pre_str(pre)}")
    raise NotImplementedError(f"Method '{method}' not supported for sharded solves.")


# ----------------------------------------------------------------------------------------------------------------------
# one system, x-slab row partition
# ----------------------------------------------------------------------------------------------------------------------

class SlabCG:
    """
    Jacobi- or un-preconditioned CG (the generic loop of phiml/backend/_linalg.py:52-90) for ONE star-stencil system whose grid
    is split into x-slabs over the ranks of ``group``.

    ``engine_factory(local_ext_grid, xb, xe)`` builds the per-rank engine; the default is the CUDA engine below, the CPU tests
    inject a NumPy engine with the same five steps to exercise partitioning, halo exchange and the collective sequence with gloo.
    """

    def __init__(self, grid: Sequence[int], offsets, coeffs, dtype: torch.dtype, jacobi: bool = True, device=None, group=None, engine_factory=None):
        assert len(grid) == 3, "slab partition is for 3-D grids (x is the partitioned axis)"
        self.grid = tuple(int(g) for g in grid)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.x_lo, self.x_hi = slab_bounds(self.grid[0], self.world, self.rank)
        self.nxl = self.x_hi - self.x_lo
        self.plane = self.grid[1] * self.grid[2]
        self.dtype = dtype
        self.device = device
        ext_grid = (self.nxl + 2, self.grid[1], self.grid[2])
        factory = engine_factory or (lambda g, xb, xe: CudaSlabEngine(g, xb, xe, offsets, coeffs, dtype, jacobi, device))
        self.engine = factory(ext_grid, 1, self.nxl + 1)

    # local (batch, nxl*plane) <-> extended (batch, (nxl+2)*plane) with zero halos
    def extend(self, v: torch.Tensor) -> torch.Tensor:
        out = torch.zeros((v.shape[0], (self.nxl + 2) * self.plane), dtype=v.dtype, device=v.device)
        out[:, self.plane:(self.nxl + 1) * self.plane] = v
        return out

    def interior(self, ext: torch.Tensor) -> torch.Tensor:
        return ext[:, self.plane:(self.nxl + 1) * self.plane]

    def _allreduce(self, t: torch.Tensor):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _halo(self, ext: torch.Tensor):
        exchange_halo_planes(ext, self.nxl, self.plane, self.rank, self.world, self.group)

    def begin(self, y_local: torch.Tensor, x0_local: torch.Tensor, rtol, atol, max_iter: int, peer_halos: Optional[bool] = None, peer_reduce: Optional[bool] = None):
        """Set-up of one solve: extended arrays, initial residual, tolerances, first halos, peer-memory mappings.  Follow with
        ``iterate`` (any number of times) and ``finish``.  See ``solve`` for the arguments."""
        eng = self.engine
        use_peer = self.world > 1 and hasattr(eng, 'open_peer_halos') and (peer_halos is None or peer_halos)
        if peer_halos and not use_peer and self.world > 1:
            raise RuntimeError("this engine has no peer-memory halo exchange")
        use_pr = self.world > 1 and self.world <= 8 and hasattr(eng, 'peer_reduce_cfg') and (peer_reduce is None or peer_reduce)
        if hasattr(eng, 'peer_reduce_cfg'):
            eng.peer_reduce_cfg = (self.rank, self.world, self.group) if use_pr else None
        self.peer_reduce_active = use_pr
        self._reduce = (lambda t: None) if use_pr else self._allreduce
        y_ext, x_ext = self.extend(y_local), self.extend(x0_local)
        self._halo(x_ext)                                   # A x0 needs the neighbours' boundary planes
        eng.start(y_ext, x_ext, rtol, atol, max_iter)       # local sums of r.s, y.y, r.r -> eng.totals
        self._reduce(eng.totals)
        eng.step(10)                                        # tolerances, first progress check (identical on all ranks)
        self._halo(eng.r)
        self._halo(eng.d0)                                  # d_0 = M^-1 r_0 including the halo planes
        if use_peer:
            eng.open_peer_halos(self.rank, self.world, self.nxl, self.plane, self.group)
        self.peer_halos_active = use_peer
        self.iterations_enqueued = 0

    def iterate(self, n: int):
        """Enqueues ``n`` CG iterations.  With peer-memory halos AND the in-kernel all-reduce an iteration needs nothing from the
        host: the engine queues all of them in one call (two kernels per iteration, no NCCL, no Python in between)."""
        eng = self.engine
        if self.peer_halos_active and self.peer_reduce_active and hasattr(eng, 'advance'):
            eng.advance(n)
        else:
            for _ in range(n):
                eng.step(0)                                 # d, x, q = A d, local d.q
                self._reduce(eng.totals)
                eng.step(1)                                 # alpha ; r -= alpha q, local r.M^-1 r
                self._reduce(eng.totals)
                eng.step(2)                                 # beta, delta, progress check
                if not self.peer_halos_active:
                    self._halo(eng.r)                       # (peer mode: step(1) already stored the planes at the neighbours)
        self.iterations_enqueued += n

    def finish(self):
        eng = self.engine
        eng.step(20)                                        # pending x += alpha d
        its, fev, conv, div = eng.result()
        if self.peer_halos_active or self.peer_reduce_active:
            eng.close_peer_halos(self.group)
        return self.interior(eng.x), self.interior(eng.r), its, fev, conv, div

    def solve(self, y_local: torch.Tensor, x0_local: torch.Tensor, rtol, atol, max_iter: int, poll_every: int = 16, fixed_iterations: Optional[int] = None,
              peer_halos: Optional[bool] = None, peer_reduce: Optional[bool] = None):
        """
        ``y_local``/``x0_local``: this rank's slab, shape (batch, nxl*ny*nz).  Returns (x_local, residual_local, iterations,
        function_evaluations, converged, diverged) — the scalars are identical on every rank.
        ``fixed_iterations`` runs exactly that many iterations without polling (throughput measurements).
        ``peer_halos``: the residual kernel writes its boundary planes straight into the neighbours' halo planes (CUDA IPC over
        NVLink) instead of a separate NCCL send/recv step per iteration; default: whenever the engine supports it.
        ``peer_reduce``: the kernels all-reduce their dot products themselves through mailboxes in peer memory (no NCCL call and no
        finalize launch inside the loop: an iteration is two kernels); default: whenever the engine supports it.
        """
        self.begin(y_local, x0_local, rtol, atol, max_iter, peer_halos, peer_reduce)
        limit = fixed_iterations if fixed_iterations is not None else max_iter
        done = 0
        chunk = poll_every
        while done < limit:
            n = min(chunk, limit - done)
            self.iterate(n)
            done += n
            if fixed_iterations is None and not self.engine.any_continue():
                break
            chunk = min(2 * chunk, 64)
        return self.finish()


class CudaSlabEngine:
    """The five steps on libphb200 (phb200_matrix_set_slab + phb200_solver_set_dist / _dist_step)."""

    def __init__(self, ext_grid, xb, xe, offsets, coeffs, dtype, jacobi, device):
        import ctypes
        from . import _engine as E
        from .sparse import Matrix
        from .precond import JacobiPreconditioner
        self.E = E
        self.device = torch.device(device if device is not None else 'cuda')
        self.matrix = Matrix.stencil_matrix(ext_grid, offsets, coeffs, dtype, self.device)
        E.check(E.lib().phb200_matrix_set_slab(self.matrix.handle, xb, xe))
        self.pre = JacobiPreconditioner(self.matrix) if jacobi else None
        self.solver = None
        self.peer_reduce_cfg = None        # (rank, world, group): in-kernel all-reduce through peer-memory mailboxes
        self._peer_bases = []

    def _setup_peer_reduce(self, nb, dev):
        """Mailboxes: allocate + zero, exchange CUDA IPC handles, map every rank's, hand the table to the solver (before start)."""
        import ctypes
        E = self.E
        rank, world, group = self.peer_reduce_cfg
        nbytes = int(E.lib().phb200_peer_mailbox_bytes(world, nb))
        self.mailbox = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)                     # zero-filled before anybody learns the address
        handle = (ctypes.c_ubyte * 64)()
        offset = ctypes.c_int64()
        E.check(E.lib().phb200_ipc_export(E.ptr(self.mailbox), handle, ctypes.byref(offset)))
        everyone = [None] * world
        dist.all_gather_object(everyone, (bytes(handle), int(offset.value)), group=group)
        table = (ctypes.c_void_p * world)()
        for r in range(world):
            if r == rank:
                table[r] = self.mailbox.data_ptr()
                continue
            base = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(everyone[r][0])
            E.check(E.lib().phb200_ipc_open(buf, ctypes.byref(base)))
            self._peer_bases.append(base)
            table[r] = base.value + everyone[r][1]
        E.check(E.lib().phb200_solver_set_peer_reduce(self.solver._h, table, world, rank, E.stream_ptr(dev)))

    def start(self, y_ext, x_ext, rtol, atol, max_iter):
        from .solve import DeviceSolver
        E = self.E
        nb = y_ext.shape[0]
        dev = y_ext.device
        self.totals = torch.zeros(nb * 4, dtype=torch.float64, device=dev)
        self.solver = DeviceSolver(self.matrix, self.pre, E.CG, 0, nb)
        E.check(E.lib().phb200_solver_set_dist(self.solver._h, E.ptr(self.totals)))
        if self.peer_reduce_cfg is not None:
            self._setup_peer_reduce(nb, dev)
        rt = torch.as_tensor(rtol, device=dev).to(y_ext.dtype).reshape(-1).expand(nb).contiguous()
        at = torch.as_tensor(atol, device=dev).to(y_ext.dtype).reshape(-1).expand(nb).contiguous()
        mi = torch.full((nb,), int(max_iter), dtype=torch.int32, device=dev)
        self.solver.start(y_ext, x_ext, rt, at, mi, zero_init=True)
        n_ext = y_ext.shape[1]
        self.x, self.r = self.solver.x, self.solver.r
        ws = self.solver.workspace
        self.d0 = ws[: nb * n_ext * y_ext.element_size()].view(y_ext.dtype).view(nb, n_ext)   # work vector 0 = first direction buffer

    def step(self, code: int):
        E = self.E
        E.check(E.lib().phb200_solver_dist_step(self.solver._h, code, E.stream_ptr(self.x.device)))

    def advance(self, n: int):
        """n whole iterations in one library call (needs peer halos + in-kernel all-reduce)."""
        E = self.E
        E.check(E.lib().phb200_solver_dist_advance(self.solver._h, int(n), E.stream_ptr(self.x.device)))

    # --- fused halo exchange: map the neighbours' residual arrays (CUDA IPC) and let the residual kernel store into them ---
    def open_peer_halos(self, rank, world, nxl, plane, group=None):
        import ctypes
        E = self.E
        handle = (ctypes.c_ubyte * 64)()
        offset = ctypes.c_int64()
        E.check(E.lib().phb200_ipc_export(E.ptr(self.r), handle, ctypes.byref(offset)))
        mine = (bytes(handle), int(offset.value), int(nxl))
        everyone = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        w = self.r.element_size()
        ptrs, lds = [None, None], [0, 0]
        for side, nb in ((0, rank - 1), (1, rank + 1)):
            if nb < 0 or nb >= world:
                continue
            h, off, nxl_nb = everyone[nb]
            base = ctypes.c_void_p()
            buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
            E.check(E.lib().phb200_ipc_open(buf, ctypes.byref(base)))
            self._peer_bases.append(base)
            halo_plane = nxl_nb + 1 if side == 0 else 0         # the lower neighbour's upper halo plane / the upper neighbour's lower one
            ptrs[side] = ctypes.c_void_p(base.value + off + halo_plane * plane * w)
            lds[side] = (nxl_nb + 2) * plane
        E.check(E.lib().phb200_solver_set_peer_halos(self.solver._h, ptrs[0], lds[0], ptrs[1], lds[1]))

    def close_peer_halos(self, group=None):
        E = self.E
        torch.cuda.synchronize(self.x.device)
        if dist.is_initialized():
            dist.barrier(group=group)            # nobody unmaps while a neighbour's kernel may still store
        E.check(E.lib().phb200_solver_set_peer_halos(self.solver._h, None, 0, None, 0))
        for base in getattr(self, '_peer_bases', []):
            E.check(E.lib().phb200_ipc_close(base))
        self._peer_bases = []

    def any_continue(self) -> bool:
        return self.solver.any_continue()

    def result(self):
        return self.solver.result()


# ----------------------------------------------------------------------------------------------------------------------
# one system, x-slab row partition, ANY method of the device loop (BASELINE config 5: BiCGstab(2) across 8 GPUs)
# ----------------------------------------------------------------------------------------------------------------------

class SlabSolver:
    """
    ``Backend.linear_solve`` for ONE star-stencil system split into x-slabs over the ranks of ``group``: 'CG', 'CG-adaptive',
    'biCG-stab', 'biCG-stab(L)' with no or Jacobi preconditioner, generic (NumPy-flavour) rules.

    The Krylov loop stays inside the engine; the engine calls back into ``allreduce(totals)`` wherever the algorithm has a
    global reduction (rank-local sums -> sums over all ranks, then the scalar step runs identically everywhere) and into
    ``halo(vec)`` before every product (boundary planes of the input vector from the two slab neighbours).
    ``engine_factory(ext_grid, xb, xe)`` builds the per-rank engine: the CUDA engine below (libphb200's callback-driven
    distributed mode), or the CPU test double of tests/_slab_engine.py in the gloo tests.
    """

    def __init__(self, grid: Sequence[int], offsets, coeffs, dtype: torch.dtype, method: str = 'biCG-stab(2)', jacobi: bool = False,
                 device=None, group=None, engine_factory=None, native: Optional[bool] = None):
        assert len(grid) == 3, "slab partition is for 3-D grids (x is the partitioned axis)"
        self.grid = tuple(int(g) for g in grid)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.x_lo, self.x_hi = slab_bounds(self.grid[0], self.world, self.rank)
        self.nxl = self.x_hi - self.x_lo
        self.plane = self.grid[1] * self.grid[2]
        self.dtype = dtype
        self.method = method
        self.counts = {'allreduce': 0, 'halo': 0}
        ext_grid = (self.nxl + 2, self.grid[1], self.grid[2])
        # native: nothing on the host inside the loop (peer-memory halo pushes + in-kernel all-reduces, CudaPeerEngine); needs one process per GPU
        # with peer access.  Default: native on CUDA with more than one rank (PHB200_SLAB_NATIVE=0 restores the callbacks).
        if native is None:
            import os
            native = engine_factory is None and self.world > 1 and os.environ.get('PHB200_SLAB_NATIVE', '1') != '0'
        self.native = bool(native)
        if self.native:
            factory = lambda g, xb, xe: CudaPeerEngine(g, xb, xe, offsets, coeffs, dtype, method, jacobi, device, self.rank, self.world, self.nxl, self.plane, self.group)
        else:
            factory = engine_factory or (lambda g, xb, xe: CudaCallbackEngine(g, xb, xe, offsets, coeffs, dtype, method, jacobi, device))
        self.engine = factory(ext_grid, 1, self.nxl + 1)

    extend = SlabCG.extend
    interior = SlabCG.interior

    def _allreduce(self, t: torch.Tensor):
        self.counts['allreduce'] += 1
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _halo(self, ext: torch.Tensor):
        self.counts['halo'] += 1
        exchange_halo_planes(ext, self.nxl, self.plane, self.rank, self.world, self.group)

    def solve(self, y_local: torch.Tensor, x0_local: torch.Tensor, rtol, atol, max_iter: int, fixed_iterations: Optional[int] = None):
        """Returns (x_local, residual_local, iterations, function_evaluations, converged, diverged); scalars identical on all ranks."""
        y_ext, x_ext = self.extend(y_local), self.extend(x0_local)
        x, r, its, fev, conv, div = self.engine.solve(y_ext, x_ext, rtol, atol, int(max_iter), self._allreduce, self._halo, fixed_iterations)
        return self.interior(x), self.interior(r), its, fev, conv, div


def _timed_advance(solver, n: int, dev, group=None) -> float:
    """``solver.advance(n)`` between two CUDA events on the launch stream: the time of the LOOP alone (set-up — allocation, IPC mapping, the
    first residual — stays outside), for throughput measurements with a fixed iteration count."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group=group)          # all ranks enter the loop together: a rank still in its set-up would be waited for inside the first reduction
        torch.cuda.synchronize(dev)
    e0.record()
    solver.advance(n)
    e1.record()
    torch.cuda.synchronize(dev)
    return float(e0.elapsed_time(e1))


class CudaCallbackEngine:
    """libphb200's callback-driven distributed mode (phb200_solver_set_dist_callbacks) behind the SlabSolver engine interface."""

    def __init__(self, ext_grid, xb, xe, offsets, coeffs, dtype, method, jacobi, device):
        from . import _engine as E
        from .sparse import Matrix
        from .precond import JacobiPreconditioner
        self.E = E
        self.device = torch.device(device if device is not None else 'cuda')
        self.matrix = Matrix.stencil_matrix(ext_grid, offsets, coeffs, dtype, self.device)
        E.check(E.lib().phb200_matrix_set_slab(self.matrix.handle, xb, xe))
        self.pre = JacobiPreconditioner(self.matrix) if jacobi else None
        self.method = method

    def solve(self, y_ext, x_ext, rtol, atol, max_iter, allreduce, halo, fixed_iterations=None):
        import ctypes
        from .solve import DeviceSolver
        E = self.E
        nb, n_ext = y_ext.shape
        dev = y_ext.device
        code, order, _ = _method_code(self.method, self.pre, np.array([[max_iter]]))
        if code == E.CG_TORCH:
            code = E.CG                    # generic rules: the torch fast path has no distributed meaning of its own
        if code == E.CG_ADAPTIVE_TORCH:
            code = E.CG_ADAPTIVE
        totals = torch.zeros(nb * 4, dtype=torch.float64, device=dev)
        solver = DeviceSolver(self.matrix, self.pre, code, order, nb)
        rt = torch.as_tensor(rtol, device=dev).to(y_ext.dtype).reshape(-1).expand(nb).contiguous()
        at = torch.as_tensor(atol, device=dev).to(y_ext.dtype).reshape(-1).expand(nb).contiguous()
        mi = torch.full((nb,), int(max_iter), dtype=torch.int32, device=dev)
        views = {}          # device address -> tensor view of one (batch, n_ext) vector
        errors = []

        def view_of(ptr):
            v = views.get(ptr)
            if v is None:
                for base in (solver.x, solver.r, x_ext):
                    if base.data_ptr() == ptr:
                        v = base
                ws = solver.workspace
                off = ptr - ws.data_ptr()
                if v is None and 0 <= off < ws.numel():
                    v = ws[off: off + nb * n_ext * y_ext.element_size()].view(y_ext.dtype).view(nb, n_ext)
                if v is None:
                    raise RuntimeError(f"halo callback for an unknown vector at {ptr:#x}")
                views[ptr] = v
            return v

        @ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p)
        def ar_cb(ptr, count, stream, user):
            try:
                allreduce(totals)
                return 0
            except Exception as exc:           # an exception must not unwind through C
                errors.append(exc)
                return 1

        @ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)
        def halo_cb(ptr, stream, user):
            try:
                halo(view_of(ptr))
                return 0
            except Exception as exc:
                errors.append(exc)
                return 1

        E.check(E.lib().phb200_solver_set_dist_callbacks(solver._h, E.ptr(totals), ar_cb, halo_cb, None))
        halo(x_ext)                                 # A x0 needs the neighbours' boundary planes (start() copies x0 incl. halos)
        try:
            solver.start(y_ext, x_ext, rt, at, mi, zero_init=True)
            if fixed_iterations is not None:
                self.last_loop_ms = _timed_advance(solver, int(fixed_iterations), dev, getattr(self, 'group', None))
            else:
                solver.run(int(max_iter))
        except Exception:
            if errors:
                raise errors[0]
            raise
        its, fev, conv, div = solver.result()
        self.last_solver = solver
        return solver.x, solver.r, its, fev, conv, div


class CudaPeerEngine:
    """libphb200's NATIVE distributed mode (phb200_solver_set_peer_native) behind the SlabSolver engine interface: every rank maps its slab
    neighbours' X / R / workspace arrays and all ranks' mailboxes through CUDA IPC once per solve; after that a pass of ANY method is only
    kernels — halo planes are pushed into the neighbours' arrays by k_halo_push, dot products are all-reduced inside the kernels that form
    them.  The callbacks of the engine interface are ignored (counts of collectives stay at what the algorithm needs: see `counts`)."""

    def __init__(self, ext_grid, xb, xe, offsets, coeffs, dtype, method, jacobi, device, rank, world, nxl, plane, group):
        from . import _engine as E
        from .sparse import Matrix
        from .precond import JacobiPreconditioner
        self.E = E
        self.device = torch.device(device if device is not None else 'cuda')
        self.matrix = Matrix.stencil_matrix(ext_grid, offsets, coeffs, dtype, self.device)
        E.check(E.lib().phb200_matrix_set_slab(self.matrix.handle, xb, xe))
        self.pre = JacobiPreconditioner(self.matrix) if jacobi else None
        self.method = method
        self.rank, self.world, self.nxl, self.plane, self.group = rank, world, nxl, plane, group

    def _export(self, t: torch.Tensor):
        import ctypes
        handle = (ctypes.c_ubyte * 64)()
        offset = ctypes.c_int64()
        self.E.check(self.E.lib().phb200_ipc_export(self.E.ptr(t), handle, ctypes.byref(offset)))
        return bytes(handle), int(offset.value)

    def _open(self, exported):
        import ctypes
        h, off = exported
        base = ctypes.c_void_p()
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
        self.E.check(self.E.lib().phb200_ipc_open(buf, ctypes.byref(base)))
        self._bases.append(base)
        return base.value + off

    def solve(self, y_ext, x_ext, rtol, atol, max_iter, allreduce, halo, fixed_iterations=None):
        import ctypes
        from .solve import DeviceSolver
        E = self.E
        nb, n_ext = y_ext.shape
        dev = y_ext.device
        w = y_ext.element_size()
        code, order, _ = _method_code(self.method, self.pre, np.array([[max_iter]]))
        if code == E.CG_TORCH:
            code = E.CG
        if code == E.CG_ADAPTIVE_TORCH:
            code = E.CG_ADAPTIVE
        solver = DeviceSolver(self.matrix, self.pre, code, order, nb)
        rt = torch.as_tensor(rtol, device=dev).to(y_ext.dtype).reshape(-1).expand(nb).contiguous()
        at = torch.as_tensor(atol, device=dev).to(y_ext.dtype).reshape(-1).expand(nb).contiguous()
        mi = torch.full((nb,), int(max_iter), dtype=torch.int32, device=dev)
        # every array a neighbour may store into is its own allocation (IPC handles name whole allocations), zero-filled: halo planes start as zeros
        x = torch.zeros_like(y_ext)
        r = torch.zeros_like(y_ext)
        ws = torch.zeros(solver.workspace_bytes(), dtype=torch.uint8, device=dev)
        mail_bytes = int(E.lib().phb200_peer_mailbox_bytes(self.world, nb))
        mail = torch.zeros(mail_bytes + 64, dtype=torch.uint8, device=dev)          # + two sequence words of the halo exchange
        totals = torch.zeros(nb * 4, dtype=torch.float64, device=dev)
        torch.cuda.synchronize(dev)
        self._bases = []
        mine = {'x': self._export(x), 'r': self._export(r), 'ws': self._export(ws), 'mail': self._export(mail), 'nxl': int(self.nxl)}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        boxes = (ctypes.c_void_p * self.world)()
        for p in range(self.world):
            boxes[p] = mail.data_ptr() if p == self.rank else self._open(everyone[p]['mail'])

        class PeerSlab(ctypes.Structure):
            _fields_ = [('X', ctypes.c_void_p), ('R', ctypes.c_void_p), ('ws', ctypes.c_void_p), ('ld', ctypes.c_int64), ('halo_off', ctypes.c_int64), ('flag', ctypes.c_void_p)]

        def neighbour(nbr, lower):
            if nbr < 0 or nbr >= self.world:
                return None
            info = everyone[nbr]
            nxl_nb = info['nxl']
            halo_plane = nxl_nb + 1 if lower else 0            # its upper halo plane (I am its upper neighbour) / its lower one
            flag_index = 1 if lower else 0                      # its my_flags[1] is written by ITS upper neighbour, [0] by its lower one
            return PeerSlab(self._open(info['x']), self._open(info['r']), self._open(info['ws']), (nxl_nb + 2) * self.plane, halo_plane * self.plane,
                            boxes[nbr] + mail_bytes + 8 * flag_index)
        lo, hi = neighbour(self.rank - 1, True), neighbour(self.rank + 1, False)
        E.check(E.lib().phb200_solver_set_peer_native(solver._h, E.ptr(totals), boxes, self.world, self.rank, ctypes.byref(lo) if lo else None,
                                                      ctypes.byref(hi) if hi else None, ctypes.c_void_p(mail.data_ptr() + mail_bytes), E.stream_ptr(dev)))
        self._keep = (x, r, ws, mail, totals, lo, hi)
        if dist.is_initialized():
            dist.barrier(group=self.group)        # every rank has mapped and zero-filled before anybody stores
        try:
            solver.start(y_ext, x_ext, rt, at, mi, zero_init=True, buffers=(x, r, ws))
            if fixed_iterations is not None:
                self.last_loop_ms = _timed_advance(solver, int(fixed_iterations), dev, self.group)
            else:
                solver.run(int(max_iter))
            its, fev, conv, div = solver.result()
            torch.cuda.synchronize(dev)
        finally:
            if dist.is_initialized():
                dist.barrier(group=self.group)    # nobody unmaps while a neighbour's kernel may still store
            for base in self._bases:
                E.check(E.lib().phb200_ipc_close(base))
            self._bases = []
        self.last_solver = solver
        return solver.x, solver.r, its, fev, conv, div
