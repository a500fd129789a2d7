"""Slab decomposition across GPUs: the role of hydro/solver/choppedup.lua, one process per GPU.

Reference: choppedup.lua:52-106 cuts the grid into ``multiSlices`` blocks, one sub-solver per OpenCL device, takes the
host ``math.min`` of the sub-solvers' dt (:344-356) and copies numGhost-thick rectangles between neighbours once per
full step (:193-233,390-400).  Here the grid is cut into slabs along the slowest used axis, every rank (process) owns
one slab on its own GPU, ghost planes are exchanged after EVERY Runge-Kutta stage (so the decomposed run equals the
single-device run bit for bit, which the reference's once-per-step sync does not -- SURVEY App. C #15) and dt is
min-reduced across ranks.  On the GPU path the exchange and the reduction are NCCL calls issued by the C library
(hb_fv_comm_init); this class only carries the geometry and the rendezvous, and offers the same exchange over
torch.distributed (gloo) for host-side tests.
"""
import numpy as np

NUM_GHOST = 2


class SlabComm:
    def __init__(self, nranks=1, rank=0, dist=None):
        """``dist``: an initialised torch.distributed module (any backend) used for rendezvous / host-side exchange."""
        self.nranks = int(nranks)
        self.rank = int(rank)
        self.dist = dist
        self._uid = None

    # ---- geometry
    @staticmethod
    def axis(dim):
        return dim - 1

    def localSize(self, solver):
        return self.split(solver.sizeWithoutBorder, solver.dim)[0]

    def split(self, globalN, dim):
        ax = self.axis(dim)
        n = list(globalN)
        if n[ax] % self.nranks:
            raise ValueError("grid size %d along axis %d is not divisible by %d ranks" % (n[ax], ax, self.nranks))
        n[ax] //= self.nranks
        off = [0, 0, 0]
        off[ax] = self.rank * n[ax]
        if n[ax] < NUM_GHOST:
            raise ValueError("slab thinner than the ghost width")
        return n, off

    def neighbours(self, periodic):
        lo, hi = self.rank - 1, self.rank + 1
        if lo < 0:
            lo = self.nranks - 1 if periodic else None
        if hi >= self.nranks:
            hi = 0 if periodic else None
        if self.nranks == 1:
            lo = hi = None
        return lo, hi

    def localBoundaryIds(self, bcGlobal, dim):
        """Faces owned by a neighbouring slab become 'none' (3): they are filled by the exchange."""
        bc = list(bcGlobal)
        if self.nranks == 1:
            return bc
        ax = self.axis(dim)
        periodic = bc[2 * ax] == 0 and bc[2 * ax + 1] == 0
        if self.rank > 0 or periodic:
            bc[2 * ax] = 3
        if self.rank < self.nranks - 1 or periodic:
            bc[2 * ax + 1] = 3
        return bc

    # ---- rendezvous for the C library's NCCL communicator
    def uniqueId(self):
        if self._uid is None:
            import ctypes as C
            from importlib import import_module
            hb = import_module("hydro-cl-lua_b200._lib")
            buf = C.create_string_buffer(128)
            if self.rank == 0:
                hb.check(hb.lib().hb_comm_unique_id(buf))
            if self.nranks > 1:
                objs = [buf.raw if self.rank == 0 else None]
                self.dist.broadcast_object_list(objs, src=0)
                self._uid = objs[0]
            else:
                self._uid = buf.raw
        return self._uid

    # ---- host-side exchange (torch.distributed); U is the local ghost-inclusive array [Sz, Sy, Sx, nS]
    def exchangeHost(self, U, dim, periodic):
        import torch
        if self.nranks == 1:
            return U
        g = NUM_GHOST
        ax = 2 - self.axis(dim)            # numpy axis of the decomposed grid axis
        lo, hi = self.neighbours(periodic)
        S = U.shape[ax]

        def take(a, b):
            sl = [slice(None)] * U.ndim
            sl[ax] = slice(a, b)
            return tuple(sl)
        ops, recvs = [], []
        if lo is not None:
            send = torch.from_numpy(np.ascontiguousarray(U[take(g, 2 * g)]))
            recv = torch.empty_like(send)
            ops += [self.dist.P2POp(self.dist.isend, send, lo), self.dist.P2POp(self.dist.irecv, recv, lo)]
            recvs.append((take(0, g), recv))
        if hi is not None:
            send = torch.from_numpy(np.ascontiguousarray(U[take(S - 2 * g, S - g)]))
            recv = torch.empty_like(send)
            ops += [self.dist.P2POp(self.dist.isend, send, hi), self.dist.P2POp(self.dist.irecv, recv, hi)]
            recvs.append((take(S - g, S), recv))
        if lo is not None and lo == hi:
            # two ranks, periodic: both messages go to the same peer; order them low-face-first on both sides by tag
            ops = []
            send_lo = torch.from_numpy(np.ascontiguousarray(U[take(g, 2 * g)]))
            send_hi = torch.from_numpy(np.ascontiguousarray(U[take(S - 2 * g, S - g)]))
            recv_lo = torch.empty_like(send_lo)
            recv_hi = torch.empty_like(send_hi)
            reqs = [self.dist.isend(send_lo, lo, tag=1), self.dist.isend(send_hi, hi, tag=2),
                    self.dist.irecv(recv_lo, lo, tag=2), self.dist.irecv(recv_hi, hi, tag=1)]
            for r in reqs:
                r.wait()
            U[take(0, g)] = recv_lo.numpy()
            U[take(S - g, S)] = recv_hi.numpy()
            return U
        for r in self.dist.batch_isend_irecv(ops):
            r.wait()
        for sl, t in recvs:
            U[sl] = t.numpy()
        return U

    def minAllReduce(self, x):
        import torch
        if self.nranks == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t[0])

    def gatherInterior(self, Uint, dim):
        """All ranks' interior slabs -> the whole interior array (every rank gets it)."""
        import torch
        if self.nranks == 1:
            return Uint
        t = torch.from_numpy(np.ascontiguousarray(Uint))
        parts = [torch.empty_like(t) for _ in range(self.nranks)]
        self.dist.all_gather(parts, t)
        return np.concatenate([p.numpy() for p in parts], axis=2 - self.axis(dim))
