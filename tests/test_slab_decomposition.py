"""Slab decomposition (hydro/solver/choppedup.py): a run split over 2 ranks must equal the single-domain run exactly.

CPU: world_size-2 gloo, oracle backend per slab, ghost planes exchanged by SlabComm.exchangeHost (host-side logic).
GPU (needs >= 2 devices): CUDA backend per slab, exchange by ncclSend/ncclRecv inside libhydrob200, dt by ncclAllReduce(min)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(mode, case, nsteps, out, world=2):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "slab_worker.py"), mode, case, str(nsteps), out], env=env))
    for p in procs:
        assert p.wait(timeout=600) == 0


def single(hydrob200, case, nsteps, **kw):
    from cases import ADM_CASES, CASES
    cfg, _ = dict(CASES, **ADM_CASES)[case]
    S = hydrob200.FiniteVolumeSolver(dict(cfg, **kw))
    for _ in range(nsteps):
        S.update()
    return S.interior(), S.t


def test_slab_geometry(hydrob200):
    from importlib import import_module
    SlabComm = import_module("hydro-cl-lua_b200.hydro.solver.choppedup").SlabComm
    c = SlabComm(4, 1)
    assert c.split([16, 12, 8], 3) == ([16, 12, 2], [0, 0, 2])
    assert c.split([16, 12, 1], 2) == ([16, 3, 1], [0, 3, 0])
    assert c.neighbours(False) == (0, 2) and SlabComm(4, 0).neighbours(False) == (None, 1)
    assert SlabComm(4, 0).neighbours(True) == (3, 1) and SlabComm(4, 3).neighbours(True) == (2, 0)
    assert SlabComm(4, 0).localBoundaryIds([2, 2, 1, 1, 0, 0], 3) == [2, 2, 1, 1, 3, 3]       # periodic: both faces exchanged
    assert SlabComm(4, 0).localBoundaryIds([2, 2, 1, 1, 2, 1], 3) == [2, 2, 1, 1, 2, 3]       # physical low face kept
    assert SlabComm(4, 3).localBoundaryIds([2, 2, 1, 1, 2, 1], 3) == [2, 2, 1, 1, 3, 1]
    with pytest.raises(ValueError):
        c.split([16, 12, 6], 3)


@pytest.mark.parametrize("case", ["slab_fe_2d_periodic", "slab_fe_3d_mixed"])
def test_cpu_gloo_two_ranks_equal_single(hydrob200, oracle, tmp_path, case):
    out = str(tmp_path / "dec")
    launch("cpu", case, 6, out)
    got = np.load(out + ".npy")
    ref, tref = single(hydrob200, case, 6, backend=oracle.OracleBackend)
    assert np.load(out + ".t.npy")[0] == tref
    assert np.array_equal(got, ref)


TWO = [("C4_sphere_rk4", "gpu_strict"), ("C2_kh_rk4tvd_minmod", "gpu"), ("C3_ot_rk3tvd", "gpu"),
       ("C4_sphere_rk4_mirror_periodic", "gpu"), ("C5_gauge_wave_slab", "gpu"),
       ("C5_warp_bubble_rk4", "gpu_strict"), ("slab_march3d_overlap", "gpu_strict"),
       ("slab_march3d_overlap", "gpu"), ("slab_march3d_overlap_periodic", "gpu"),
       ("slab_march2d_overlap", "gpu_strict"), ("slab_thin3d", "gpu_strict"), ("slab_thin2d_mhd", "gpu"),
       # the ops on a decomposed grid: potential ghost planes exchanged per sweep, residual / maximum all-reduced
       ("F3_selfgrav_sphere_rk4_plm_3d", "gpu_strict"), ("F3_nodiv_ot_rk3_2d", "gpu"),
       ("F3_nodiv_selfgrav_ot_fe_3d", "gpu_strict"), ("F3_selfgrav_sphere_fe_2d", "gpu"),
       # useCTU on a decomposed grid: the face-state blocks' ghost planes travel with boundaryLR
       ("F4_ctu_sphere_rk2_3d", "gpu_strict"), ("F4_ctu_ot_mhd_fe_2d", "gpu")]


def _steps(case):
    return 3 if ("overlap" in case or "thin" in case) else 5


@pytest.fixture(scope="module")
def two_rank_runs(tmp_path_factory):
    """all two-rank cases inside ONE process group (one interpreter / torch / NCCL start-up per rank instead of one per case)"""
    import torch
    if torch.cuda.device_count() < 2:
        return {}
    d = tmp_path_factory.mktemp("two")
    jobs = [(mode, case, _steps(case), str(d / ("%s_%s" % (case, mode)))) for case, mode in TWO]
    launch_multi(jobs, 2)
    return {(case, mode): out for mode, case, _, out in jobs}


@pytest.mark.gpu
@pytest.mark.parametrize("case,mode", TWO)
def test_gpu_two_ranks_equal_single(hydrob200, two_rank_runs, case, mode):
    if (case, mode) not in two_rank_runs:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    out = two_rank_runs[(case, mode)]
    n = _steps(case)
    got = np.load(out + ".npy")
    if "overlap" in case or "thin" in case:
        assert "exchange=overlapped" in open(out + ".describe").read()
    ref, tref = single(hydrob200, case, n, strict_fp=(mode == "gpu_strict"), use_graph=False)
    assert np.load(out + ".t.npy")[0] == tref
    assert np.array_equal(got, ref)


# ---- more than two ranks: interior ranks have two neighbours; thin slabs (a few planes per rank: the rim / interior split of the overlapped
# exchange, hb_fv.cu); ADM.  Needs 4 (8) devices: gpurun --gpus 4 (8).
MANY = [("slab_march3d_overlap", "gpu_strict", 4), ("slab_march3d_overlap", "gpu", 4), ("slab_march3d_overlap_periodic", "gpu", 4),
        ("slab_march2d_overlap", "gpu_strict", 4), ("slab_thin3d", "gpu_strict", 4), ("slab_thin3d", "gpu", 4), ("slab_thin2d_mhd", "gpu_strict", 4),
        ("C5_gauge_wave_slab", "gpu_strict", 4), ("C4_sphere_rk4", "gpu_strict", 4), ("F3_selfgrav_sphere_rk4_plm_3d", "gpu_strict", 4),
        ("slab_march3d_overlap", "gpu_strict", 8), ("slab_march3d_overlap_periodic", "gpu", 8), ("slab_thin3d8", "gpu_strict", 8),
        ("slab_thin3d8", "gpu", 8)]


def launch_multi(jobs, world):
    port = free_port()
    procs = []
    arg = ",".join("%s:%s:%d:%s" % j for j in jobs)
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="2")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "slab_worker.py"), "multi", arg], env=env))
    for p in procs:
        assert p.wait(timeout=1200) == 0


@pytest.fixture(scope="module")
def many_rank_runs(tmp_path_factory):
    """every MANY case of one world size runs inside ONE process group (one interpreter / NCCL start-up per rank)"""
    import torch
    done = {}
    for world in sorted(set(w for _, _, w in MANY)):
        if torch.cuda.device_count() < world:
            continue
        d = tmp_path_factory.mktemp("many%d" % world)
        jobs = [(mode, case, 3, str(d / ("%s_%s" % (case, mode)))) for case, mode, w in MANY if w == world]
        launch_multi(jobs, world)
        for mode, case, _, out in jobs:
            done[(case, mode, world)] = out
    return done


@pytest.mark.gpu
@pytest.mark.parametrize("case,mode,world", MANY)
def test_gpu_many_ranks_equal_single(hydrob200, many_rank_runs, case, mode, world):
    if (case, mode, world) not in many_rank_runs:
        pytest.skip("needs %d GPUs (run with gpurun --gpus %d)" % (world, world))
    out = many_rank_runs[(case, mode, world)]
    n = 3
    got = np.load(out + ".npy")
    if "overlap" in case or "thin" in case:
        assert "exchange=overlapped" in open(out + ".describe").read()
    ref, tref = single(hydrob200, case, n, strict_fp=(mode == "gpu_strict"), use_graph=False)
    assert np.load(out + ".t.npy")[0] == tref
    assert np.array_equal(got, ref)
