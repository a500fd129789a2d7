#!/usr/bin/env python
"""bench.py -- cell-updates/s of the explicit finite-volume update (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C4|C4M|C2|C3|C5] [--precision float]
                    [--scaling weak|strong] [--per-gpu-planes P] [--no-parity]

Workload (default C4, the config BASELINE.json quotes "at 1/2/4/8 B200" on): 3-D Euler spherical blast, Roe + PLM
('plm cons', minmod) + classic RK4, double precision, freeflow boundaries, 512^3 interior cells PER GPU (weak
scaling: the grid is 512 x 512 x 512*N, slab-decomposed along z, ghost planes exchanged every RK stage over NCCL, dt
min-allreduced).  One "step" = one solver:update() = 4 fused stage kernels + 4 ghost fills + bookkeeping, dt device-resident.
Synthetic inputs: the 'sphere' initial condition of hydro/init/euler.lua:1274-1296 evaluated on the host.

The JSON line carries: value (resident-state throughput, all ranks), e2e (same metric through the C-ABI with HOST
buffers: pinned host -> device state upload, one update, device -> host download, every step), roofline of the fused
stage kernel (algorithmic bytes per launch / live CUDA-event duration, against MEASURED_PEAKS.json), cpu_baseline (the
oracle -- a CPU restatement of the reference's kernels, NOT the reference binary -- on a bounded sample), clocks, gpu_launches.

--impl reference times that CPU restatement with all host threads on a bounded sample of the same workload (the reference
itself is LuaJIT + OpenCL with un-vendored dependencies and cannot run here or on the GPU box: DESIGN.md).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

os.environ.setdefault("NCCL_DEBUG", "WARN")     # (NCCL's version banner goes to stdout: the contract is ONE JSON line there)
ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT,):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

# algorithmic HBM bytes per cell-update: words x nI x sizeof(double)
WORKLOADS = {
    "C4": dict(name="3D Euler spherical blast, Roe+PLM(minmod)+RK4, double, freeflow, 512^3 per GPU (z slabs)",
               cfg=dict(eqn="euler", dim=3, gridSize=[512, 512, 512], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                        usePLM="plm cons", slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1),
               words=16, nI=5, stages=4, cpu_sample=[256, 256, 256], parity=(2, 32)),
    # the north_star's second 3-D target: ideal MHD (hydro/eqn/mhd.cl:296-783) through the same fused stage kernel
    "C4M": dict(name="3D ideal-MHD Orszag-Tang (z-invariant data on a 3-D grid), Roe+PLM(minmod)+RK4, double, periodic, 384^3 per GPU (z slabs)",
                cfg=dict(eqn="mhd", dim=3, gridSize=[384, 384, 384], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="Orszag-Tang",
                         usePLM="plm cons", slopeLimiter="minmod", integrator="Runge-Kutta 4", cfl=.1),
                words=16, nI=8, stages=4, cpu_sample=[128, 128, 128], parity=(2, 16)),
    # the flux-limiter path with the marching structure (fv_march3 GEN; HB_MARCH_GEN=0 keeps it on the tile kernel)
    "C4FL": dict(name="3D Euler spherical blast, Roe + superbee FLUX limiter (no PLM), RK4, double, freeflow, 256^3 per GPU (z slabs)",
                 cfg=dict(eqn="euler", dim=3, gridSize=[256, 256, 256], mins=[-2, -2, -2], maxs=[2, 2, 2], initCond="sphere",
                          fluxLimiter="superbee", integrator="Runge-Kutta 4", cfl=.1),
                 words=16, nI=5, stages=4, cpu_sample=[128, 128, 128], parity=(1, 16)),
    "C2": dict(name="2D Euler Kelvin-Helmholtz, Roe+PLM(minmod)+RK4-TVD, double, periodic, 2048^2 per GPU (y slabs)",
               cfg=dict(eqn="euler", dim=2, gridSize=[2048, 2048], initCond="Kelvin-Helmholtz", usePLM="plm cons",
                        slopeLimiter="minmod", integrator="Runge-Kutta 4, TVD", cfl=.15),
               words=21, nI=5, stages=4, cpu_sample=[2048, 2048], parity=(4, 256)),
    "C3": dict(name="2D ideal-MHD Orszag-Tang, Roe+PLM(minmod)+RK3-TVD, double, periodic, 4096^2 per GPU (y slabs)",
               cfg=dict(eqn="mhd", dim=2, gridSize=[4096, 4096], initCond="Orszag-Tang", usePLM="plm cons",
                        slopeLimiter="minmod", integrator="Runge-Kutta 3, TVD", cfl=.15),
               words=8, nI=8, stages=3, cpu_sample=[2048, 2048], parity=(4, 256)),
    # 16 words x 37 integrated variables + 4 stages x 14 auxiliary reads = 5184 B per cell-update
    # domain +-1 (two wavelengths): at 256^3 on +-.5 the cell volume is 5.96e-8 and the reference's `volume > 1e-7` guard
    # (fvsolver.cl:97) switches the flux divergence off altogether -- the same reason C4 runs on +-2
    "C5": dict(name="3D ADM Bona-Masso gauge wave (A=.1, d=1, f=2/alpha), Roe + superbee flux limiter, RK4, double, periodic, domain +-1, 256^3 per GPU (z slabs)",
               cfg=dict(eqn="adm3d", dim=3, gridSize=[256, 256, 256], mins=[-1.] * 3, maxs=[1.] * 3, initCond="testbed - gauge wave",
                        fluxLimiter="superbee", integrator="Runge-Kutta 4", cfl=.1,
                        boundary=dict(xmin="periodic", xmax="periodic", ymin="periodic", ymax="periodic", zmin="periodic", zmax="periodic")),
               words=16, nI=37, stages=4, extra_bytes=4 * 14 * 8, cpu_sample=[64, 64, 64], parity=None),
}


def alg_bytes(w):
    return w["words"] * w["nI"] * 8 + w.get("extra_bytes", 0)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650., "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(.2)

    def summary(self):
        sm = [float(s[1]) for s in self.samples if len(s) > 8 and s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if len(s) > 8 and s[2].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            if len(s) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def cpu_oracle_rate(w, nthreads, steps, warmup=1):
    """cell-updates/s of the CPU restatement (oracle) on a bounded sample of the workload.  Test/baseline infrastructure:
    the one place besides tests/ and smoke() that executes oracle/."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hydrob200
    import oracle
    # the -O3 -march=native -ffp-contract=fast build, compiled on this machine (oracle.build_native)
    cfg = dict(w["cfg"], gridSize=w["cpu_sample"], backend=oracle.OracleBackendThreads(nthreads, native=True))
    S = hydrob200.FiniteVolumeSolver(cfg)
    cells = int(np.prod(w["cpu_sample"]))
    for _ in range(warmup):
        S.update()
    per = []
    for _ in range(steps):
        t0 = time.perf_counter()
        S.update()
        per.append(time.perf_counter() - t0)
    total = sum(per)
    return cells * steps / total, total / steps, cells


try:
    ORIG_AFFINITY = os.sched_getaffinity(0)
except AttributeError:
    ORIG_AFFINITY = None


def host_threads():
    """every hardware thread this process may run on (torchrun exports OMP_NUM_THREADS=1: not a property of the machine); the CPU legs
    run on all of them, whatever NUMA node the e2e leg pinned the process to"""
    if ORIG_AFFINITY:
        try:
            os.sched_setaffinity(0, ORIG_AFFINITY)
        except OSError:
            pass
        return max(1, len(ORIG_AFFINITY))
    return os.cpu_count() or 1


def pin_to_gpu_numa_node(index):
    """Run this rank (and first-touch its pinned staging buffers) on the CPUs of the NUMA node its GPU hangs off: with eight ranks
    staging through one node the e2e leg measures that node's memory controller, not PCIe."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read())
        if node < 0:
            return "gpu %d: no NUMA affinity reported" % index
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return "gpu %d: pinned to NUMA node %d (%d cpus)" % (index, node, len(allowed))
        return "gpu %d: NUMA node %d has no cpu this process may use" % (index, node)
    except Exception as e:       # no sysfs / nvidia-smi: leave the affinity alone
        return "gpu %d: not pinned (%s)" % (index, type(e).__name__)


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cores = host_threads()
    rate, sec, cells = cpu_oracle_rate(w, cores, args.steps, max(1, min(args.warmup, 2)))
    sample = "one update of the %s workload on a %s interior sample per step (CPU restatement of the reference kernels, g++ -O3 -march=native -ffp-contract=fast, OpenMP, %d threads)" % (
        args.workload, "x".join(map(str, w["cpu_sample"])), cores)
    line = {
        "impl": "reference", "metric": "cell-updates/sec", "value": rate, "unit": "cell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if args.precision == "float" else "f64", "data": "synthetic",
        "config": {"workload": w["name"], "sample": "x".join(map(str, w["cpu_sample"]))},
        "cpu_baseline": {"value": rate, "unit": "cell-updates/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C4", choices=list(WORKLOADS))
    ap.add_argument("--grid", default=None, help="override the per-GPU interior grid, e.g. 256,256,256")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity leg (oracle on a sub-slab of the benchmarked grid)")
    ap.add_argument("--precision", default="double", choices=["double", "float"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the per-GPU slab is fixed and the grid grows with N; strong: the workload's grid is split over the N GPUs")
    ap.add_argument("--per-gpu-planes", type=int, default=0, help="weak scaling with this many planes of the decomposed axis per GPU (SURVEY 8e: 64)")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.precision == "float":
        w["cfg"] = dict(w["cfg"], precision="float")
    if args.impl == "reference":
        return run_reference(args, w)
    if args.warmup < 3:
        args.warmup = 3

    import hydrob200
    from importlib import import_module
    hb = import_module("hydro-cl-lua_b200._lib")
    SlabComm = import_module("hydro-cl-lua_b200.hydro.solver.choppedup").SlabComm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world != args.gpus and world != 1:
        raise SystemExit("WORLD_SIZE (%d) != --gpus (%d)" % (world, args.gpus))
    dist = None
    comm = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        comm = SlabComm(world, rank, dist)

    numa_note = pin_to_gpu_numa_node(local)
    cfg = dict(w["cfg"])
    if args.grid:
        cfg["gridSize"] = [int(x) for x in args.grid.split(",")]
    perGpu = list(cfg["gridSize"])
    ax = cfg["dim"] - 1
    if args.per_gpu_planes:
        # same dx as the workload's grid: the domain shrinks with the plane count
        frac = args.per_gpu_planes / float(perGpu[ax])
        mins = list(cfg.get("mins", [-1.] * 3)); maxs = list(cfg.get("maxs", [1.] * 3))
        maxs[ax] = mins[ax] + (maxs[ax] - mins[ax]) * frac
        cfg["mins"], cfg["maxs"] = mins, maxs
        perGpu[ax] = args.per_gpu_planes
        cfg["gridSize"] = list(perGpu)
    if args.scaling == "strong":
        if perGpu[ax] % world:
            raise SystemExit("--scaling strong: %d planes do not split over %d GPUs" % (perGpu[ax], world))
        perGpu[ax] //= world          # the grid stays, each rank owns 1/N of its planes
    elif world > 1:                   # weak scaling: per-GPU slab fixed, the grid grows along the decomposed axis
        cfg["gridSize"] = list(perGpu)
        cfg["gridSize"][ax] = perGpu[ax] * world
        span = cfg.get("maxs", [1.] * 3)[ax] - cfg.get("mins", [-1.] * 3)[ax]
        mins = list(cfg.get("mins", [-1.] * 3)); maxs = list(cfg.get("maxs", [1.] * 3))
        maxs[ax] = mins[ax] + span * world       # same dx as the 1-GPU grid
        cfg["mins"], cfg["maxs"] = mins, maxs
    # the reference freezes the update when the cell volume is <= 1e-7 (fvsolver.cl:97): such a workload would time skipped work
    vol = float(np.prod([(cfg.get("maxs", [1.] * 3)[k] - cfg.get("mins", [-1.] * 3)[k]) / cfg["gridSize"][k] for k in range(cfg["dim"])]))
    if not vol > 1e-7:
        raise SystemExit("cell volume %.3g <= 1e-7: the reference's volume guard would switch the flux divergence off" % vol)
    S = hydrob200.FiniteVolumeSolver(dict(cfg, device=local, comm=comm, use_graph=True))
    B = S.backend
    L = B.L
    ctx = B.ctx
    cellsLocal = int(np.prod(perGpu))
    cellsAll = cellsLocal * world

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
        ctx.sync()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- parity of THIS run's state against the oracle (N = 1): the first steps of the benchmarked grid, compared on a central
    # sub-slab with the oracle advancing that slab + its domain of dependence from the same initial state with the same dt sequence
    # (oracle/subslab.py; the oracle is the checker here, never the thing timed).  The timed updates continue from that state.
    parity = None
    tol = 1e-5 if args.precision == "float" else 1e-12
    if world == 1 and not args.no_parity and w.get("parity"):
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle
        import subslab
        nst, planes = w["parity"]
        r = subslab.subslab_parity(hydrob200, oracle, cfg, nst, planes, G=S, nthreads=host_threads())
        parity = {k: r[k] for k in ("rel_linf", "per_group", "steps", "planes", "oracle_planes", "finite", "against")}
        parity["tolerance"] = tol
        parity["pass"] = bool(r["finite"] and r["rel_linf"] <= tol)

    # ---- resident-state throughput: W warm-up updates, then exactly K timed updates (CUDA events on the library's stream)
    hb.check(L.hb_fv_update(B.h, args.warmup))
    n0 = B.launch_count()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ctx.timerStart()
    hb.check(L.hb_fv_update(B.h, args.steps))
    ms = ctx.timerStop()
    barrier()
    sampler.stop_flag = True
    ms = max_over_ranks(ms)
    launches = B.launch_count() - n0
    value = cellsAll * args.steps / (ms * 1e-3)
    t_sim, dt_sim = B.get_time()
    finite = bool(np.isfinite(t_sim) and np.isfinite(dt_sim) and dt_sim > 0)

    # ---- the dominant kernel (fused stage kernel), timed live per launch with CUDA events (eager launches)
    hb.check(L.hb_fv_profile(B.h, 1))
    hb.check(L.hb_fv_update(B.h, max(2, min(args.steps, 5))))
    sms, sn = C.c_double(), C.c_longlong()
    hb.check(L.hb_fv_profile_read(B.h, C.byref(sms), C.byref(sn)))
    hb.check(L.hb_fv_profile(B.h, 0))
    stage_ms = sms.value / max(1, sn.value)
    wordScale = .5 if args.precision == "float" else 1.
    bytes_per_launch = alg_bytes(w) * wordScale / w["stages"] * cellsLocal
    peak, peak_src = peaks()
    achieved = bytes_per_launch / (stage_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f).get(args.workload)
        if tr and not args.grid:
            traffic = tr["bytes_per_launch"]
    except Exception:
        pass
    desc = B.describe().split("\n")[0]
    roofline = {"bound": "hbm", "kernel": desc.split()[0].replace("kernel=", ""), "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "stage_kernel_ms": stage_ms, "kernel_config": desc,
                "stage_share_of_step": stage_ms * w["stages"] / (ms / args.steps),
                "algorithmic_bytes_per_cell_update": alg_bytes(w) * wordScale}

    # ---- end to end through the C-ABI with HOST buffers: upload state (pinned host, AoS doubles) -> update -> download, every step.
    # Blocking calls first (hb_fv_set_state / hb_fv_get_state: what a time-stepping script does), then the non-blocking ones
    # (hb_fv_set_state_async / hb_fv_get_state_async): upload of step i+1, update of step i and download of step i-1 overlap on
    # separate copy streams -- independent problems streamed through one solver.  Every step still moves the whole state both ways.
    nS = B.nS
    nbytes = int(B.ncells) * nS * 8
    hp, hq0, hq1 = C.c_void_p(), C.c_void_p(), C.c_void_p()
    for h in (hp, hq0, hq1):
        hb.check(L.hb_host_alloc(nbytes, C.byref(h)))
    hb.check(L.hb_fv_get_state(B.h, hp))
    # one untimed round trip through each pair of calls: the library allocates its staging buffers on first use
    hb.check(L.hb_fv_set_state(B.h, hp)); hb.check(L.hb_fv_update(B.h, 1)); hb.check(L.hb_fv_get_state(B.h, hq0))
    hb.check(L.hb_fv_set_state_async(B.h, hp)); hb.check(L.hb_fv_update(B.h, 1)); hb.check(L.hb_fv_get_state_async(B.h, hq1))
    hb.check(L.hb_fv_wait_transfers(B.h))
    hb.check(L.hb_fv_get_state(B.h, hp))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        hb.check(L.hb_fv_set_state(B.h, hp))
        hb.check(L.hb_fv_update(B.h, 1))
        hb.check(L.hb_fv_get_state(B.h, hq0))
    ctx.sync()
    e2e_block_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    nasync = max(6, 3 * args.e2e_steps)
    t0 = time.perf_counter()
    for i in range(nasync):
        hb.check(L.hb_fv_set_state_async(B.h, hp))
        hb.check(L.hb_fv_update(B.h, 1))
        hb.check(L.hb_fv_get_state_async(B.h, hq1 if i & 1 else hq0))
    hb.check(L.hb_fv_wait_transfers(B.h))
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    for h in (hp, hq0, hq1):
        hb.check(L.hb_host_free(h))
    e2e = {"value": cellsAll * nasync / e2e_s, "unit": "cell-updates/s", "per_gpu": cellsAll * nasync / e2e_s / world,
           "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "host_numa": numa_note,
           "steps": nasync, "ms_per_step": e2e_s / nasync * 1e3,
           "mode": "non-blocking transfers (hb_fv_set_state_async / get_state_async), upload, update and download of consecutive steps overlap",
           "blocking": {"value": cellsAll * args.e2e_steps / e2e_block_s, "steps": args.e2e_steps,
                        "ms_per_step": e2e_block_s / args.e2e_steps * 1e3}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle
        cores = host_threads()
        rate, sec, cells = cpu_oracle_rate(w, cores, 4)
        cpu = {"value": rate, "unit": "cell-updates/s", "cores": cores, "kind": "port",
               "sample": "4 updates of the same workload on a %s interior sample (CPU restatement of the reference kernels, g++ -O3 -march=native -ffp-contract=fast, OpenMP, %d threads)" % (
                   "x".join(map(str, w["cpu_sample"])), cores)}

    if rank == 0:
        line = {
            "metric": "cell-updates/sec", "value": value, "unit": "cell-updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32" if args.precision == "float" else "f64", "data": "synthetic",
            "config": {"workload": w["name"] if not args.grid else w["name"] + " [grid override %s]" % args.grid,
                       "per_gpu_grid": perGpu, "global_grid": cfg["gridSize"], "parallelism": "slab-%s x%d" % ("xyz"[ax], world),
                       "l2": "state arrays (%.1f GB per buffer) exceed the 126 MB L2; no flush needed" % (B.ncells * nS * 8 / 1e9),
                       "timing": "CUDA events on the library stream, max over ranks", "finite": finite, "t": t_sim, "dt": dt_sim},
            "roofline": roofline, "parity": parity, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
