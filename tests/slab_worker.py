"""Worker for the decomposed-run tests: one process per rank (gloo on CPU with the oracle backend, or NCCL on GPUs with
the CUDA backend).  Writes the gathered interior state of rank 0 to <out>.npy."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    mode, case, nsteps, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    import hydrob200
    from importlib import import_module
    SlabComm = import_module("hydro-cl-lua_b200.hydro.solver.choppedup").SlabComm
    from cases import ADM_CASES, CASES
    cfg, _ = dict(CASES, **ADM_CASES)[case]
    if mode == "cpu":
        import oracle
        dist.init_process_group("gloo", rank=rank, world_size=world)
        comm = SlabComm(world, rank, dist)
        S = hydrob200.FiniteVolumeSolver(dict(cfg, backend=oracle.OracleBackend, comm=comm))
        assert S.rkOrder == 0, "the host-side exchange test drives single-stage (forward Euler) steps"
        periodic = S.boundaryMethods["xyz"[S.dim - 1] + "min"] == "periodic"
        U = comm.exchangeHost(S.getState(), S.dim, periodic)
        S.setState(U)
        for _ in range(nsteps):
            dt = comm.minAllReduce(S.backend.calc_dt())
            S.backend.step(dt)
            S.t += dt
            U = comm.exchangeHost(S.getState(), S.dim, periodic)
            S.setState(U)
    else:
        import torch
        torch.cuda.set_device(rank % torch.cuda.device_count())
        dist.init_process_group("nccl", rank=rank, world_size=world)
        comm = SlabComm(world, rank, dist)
        S = hydrob200.FiniteVolumeSolver(dict(cfg, comm=comm, device=rank % torch.cuda.device_count(),
                                              strict_fp=(mode == "gpu_strict"), use_graph=False))
        S.update(nsteps)
        if rank == 0:
            with open(out + ".describe", "w") as f:
                f.write(S.backend.describe())
        # the gather below runs on the host through a second, gloo group
    if mode != "cpu":
        g = dist.new_group(backend="gloo")
        import torch
        Ui = np.ascontiguousarray(S.interior())
        t = torch.from_numpy(Ui)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=g)
        full = np.concatenate([p.numpy() for p in parts], axis=2 - (S.dim - 1))
    else:
        full = S.getGlobalInterior()
    if rank == 0:
        np.save(out, full)
        np.save(out + ".t", np.array([S.t]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
