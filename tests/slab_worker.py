"""Worker for the decomposed-run tests: one process per rank (gloo on CPU with the oracle backend, or NCCL on GPUs with
the CUDA backend).  Writes the gathered interior state of rank 0 to <out>.npy."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def run_case(mode, case, nsteps, out, dist, rank, world):
    import hydrob200
    from importlib import import_module
    SlabComm = import_module("hydro-cl-lua_b200.hydro.solver.choppedup").SlabComm
    from cases import ADM_CASES, CASES
    cfg, _ = dict(CASES, **ADM_CASES)[case]
    if mode == "cpu":
        import oracle
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
        full = S.getGlobalInterior()
    else:
        import torch
        comm = SlabComm(world, rank, dist)
        S = hydrob200.FiniteVolumeSolver(dict(cfg, comm=comm, device=rank % torch.cuda.device_count(),
                                              strict_fp=(mode == "gpu_strict"), use_graph=False))
        S.update(nsteps)
        if rank == 0:
            with open(out + ".describe", "w") as f:
                f.write(S.backend.describe())
        # the gather runs on the host through a second, gloo group
        Ui = np.ascontiguousarray(S.interior())
        t = torch.from_numpy(Ui)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=run_case.gloo)
        full = np.concatenate([p.numpy() for p in parts], axis=2 - (S.dim - 1))
        S.backend.close()
    if rank == 0:
        np.save(out, full)
        np.save(out + ".t", np.array([S.t]))
    dist.barrier()


def main():
    """argv: mode case nsteps out   |   multi "mode:case:nsteps:out,mode:case:nsteps:out,..." (several cases in one process group: one
    interpreter / torch / NCCL start-up per rank instead of one per case)"""
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    if sys.argv[1] == "multi":
        jobs = [j.split(":") for j in sys.argv[2].split(",")]
    else:
        jobs = [sys.argv[1:5]]
    if jobs[0][0] == "cpu":
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        import torch
        torch.cuda.set_device(rank % torch.cuda.device_count())
        dist.init_process_group("nccl", rank=rank, world_size=world)
        run_case.gloo = dist.new_group(backend="gloo")
    for mode, case, nsteps, out in jobs:
        run_case(mode, case, int(nsteps), out, dist, rank, world)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
