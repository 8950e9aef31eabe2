#!/usr/bin/env python
"""bench.py -- grid-point updates/s of the 3-D isotropic C-PML time loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload ...]

One "step" = one full time step of the loop body (stress + velocity + source + Dirichlet
+ seismogram sample + energy) over the rank's z-slab.  Workloads (BASELINE.json configs):
  cfg3  seismic_CPML_3D_isotropic_MPI_OpenMP default grid 101 x 641 x 640 per GPU
        (N = 1: exactly the reference grid; N > 1: weak scaling, NZ = 640 N)   [default]
  cfg3f the same grid in SINGLE precision (cpml_config.precision = 1, the build the reference endorses, 3D-iso :114-116;
        one GPU; its CPU arm is the double-precision restatement)
  cfg4  1024 x 1024 x 128 per GPU (N = 8: 1024^3), weak scaling
  cfg2  2-D fourth order 4096 x 4096 (single GPU only)
  cfg5  seismic_CPML_3D_viscoelastic_MPI (4th order, N_SLS = 2) 1024 x 1024 x 128 per GPU, weak scaling
  cfg5d the same program on its default grid 210 x 800 x 220 (single GPU; N > 1: 220 planes per GPU)
  cfg6  seismic_CPML_2D_velocity_and_stress_fourth_order_viscoelastic as shipped, 2001 x 2001 (single GPU only)
Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the
reference loop (oracle/, OpenMP, all host threads): the Fortran reference itself cannot be
built in this image (no Fortran compiler, no MPI).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "grid-point updates/s, 3-D isotropic C-PML time loop (FP64)"
UNIT = "Gpts/s"


def metric_name(kind):
    return {"3d": METRIC, "3dv": "grid-point updates/s, 3-D viscoelastic C-PML time loop (FP64)",
            "2d": "grid-point updates/s, 2-D isotropic C-PML time loop (FP64)",
            "2dv": "grid-point updates/s, 2-D viscoelastic C-PML time loop (FP64)"}[kind]


def measured_peak():
    """HBM roofline denominator: the driver-measured copy bandwidth, else the recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons, power = [], [], set(), []
        for ts, line in self.lines:
            if t0 is not None and not (t0 <= ts <= t1 + 0.2):
                continue
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# The workloads as plain numbers (BASELINE.json configs), shared by both arms: the CPU arm builds its inputs from
# these with oracle/workloads.py and never touches the product package.
WORKLOADS = {
    "cfg3": dict(kind="3d", nx=101, ny=641, nz_per_gpu=640, deltat=1.6e-3, npml=10),
    "cfg3f": dict(kind="3d", nx=101, ny=641, nz_per_gpu=640, deltat=1.6e-3, npml=10, precision=1),
    "cfg4": dict(kind="3d", nx=1024, ny=1024, nz_per_gpu=128, deltat=1.6e-3, npml=10),
    "cfg5": dict(kind="3dv", nx=1024, ny=1024, nz_per_gpu=128, deltat=4e-4, npml=10),
    "cfg5d": dict(kind="3dv", nx=210, ny=800, nz_per_gpu=220, deltat=4e-4, npml=10),
    "cfg2": dict(kind="2d", nx=4096, ny=4096, nz_per_gpu=1, deltat=1e-3, npml=10, order=4),
    "cfg6": dict(kind="2dv", nx=2001, ny=2001, nz_per_gpu=1, deltat=2.2e-4, npml=10, order=4),
}


def common_config(name, n_gpus):
    """The `config` object of the JSON line: identical in both arms (the driver compares them)."""
    w = WORKLOADS[name]
    kind = w["kind"]
    nz = w["nz_per_gpu"] * n_gpus if kind in ("3d", "3dv") else 1
    if kind == "2d":
        label = f"seismic_CPML_2D_isotropic_fourth_order {w['nx']}x{w['ny']}"
    elif kind == "2dv":
        label = f"seismic_CPML_2D_velocity_and_stress_fourth_order_viscoelastic (N_SLS=3) {w['nx']}x{w['ny']}"
    elif kind == "3dv":
        label = (f"seismic_CPML_3D_viscoelastic_MPI (4th order, N_SLS=2, reference NPROC={_visco_nproc(n_gpus, nz)} emulated): "
                 f"{w['nx']}x{w['ny']}x{nz} ({w['nx']}x{w['ny']}x{w['nz_per_gpu']} per GPU, z-slabs)")
    else:
        tag = ("default grid" if name == "cfg3" else "default grid, SINGLE PRECISION (3D-iso :114-116)" if name == "cfg3f"
               else "~1024^3 scaled grid")
        label = (f"seismic_CPML_3D_isotropic_MPI_OpenMP {tag}: {w['nx']}x{w['ny']}x{nz} "
                 f"({w['nx']}x{w['ny']}x{w['nz_per_gpu']} per GPU, z-slabs)")
    return {"workload": label, "grid": [w["nx"], w["ny"], nz], "npoints_pml": w["npml"], "deltat": w["deltat"],
            "parallelism": f"z-slabs x{n_gpus}" if n_gpus > 1 else "single GPU / whole grid",
            "l2": "the state streamed every step (0.5-10 GB per rank) is far larger than the 126 MB L2; no flush needed"}


def workload_params(name, n_gpus, nstep):
    from seismic_cpml_b200 import programs as P
    if name in ("cfg3", "cfg3f"):
        return P.Params3DIso(NZ=640 * n_gpus, NSTEP=nstep), "3d"
    if name == "cfg4":
        return P.Params3DIso(NX=1024, NY=1024, NZ=128 * n_gpus, NSTEP=nstep), "3d"
    if name == "cfg5":
        return P.Params3DVisco(NX=1024, NY=1024, NZ=128 * n_gpus, NSTEP=nstep, NPROC=_visco_nproc(n_gpus, 128 * n_gpus)), "3dv"
    if name == "cfg5d":
        return P.Params3DVisco(NZ=220 * n_gpus, NSTEP=nstep, NPROC=_visco_nproc(n_gpus, 220 * n_gpus)), "3dv"
    if name == "cfg6":
        if n_gpus != 1:
            raise SystemExit("cfg6 (2-D) runs on one GPU")
        return P.Params2DVisco(order=4, NSTEP=nstep), "2dv"
    if name == "cfg2":
        if n_gpus != 1:
            raise SystemExit("cfg2 (2-D) runs on one GPU")
        return P.Params2DIso(order=4, NX=4096, NY=4096, NSTEP=nstep), "2d"
    raise SystemExit(f"unknown workload {name}")


def _visco_nproc(n_gpus, nz):
    """Reference NPROC to emulate: its default 4 (3D-visco :158) where the grid allows, else the GPU count."""
    for n in (4, n_gpus, 2):
        if n > 1 and nz % n == 0 and nz // n >= 10:
            return n
    return 1


# ------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's OpenMP build on the host cores
# ------------------------------------------------------------------------------------

_CPU_THREADS = None      # None: every core this process may run on; an int: that many (the 1-thread figure)


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm asks for the cores it may run on."""
    from oracle import oracle as O
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    O.set_num_threads(n if _CPU_THREADS is None else _CPU_THREADS)


def _mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 16 << 30


def _timed(run, kw, steps, warmup, **extra):
    from oracle import oracle as O
    _use_all_host_threads()
    O.set_ftz(True)                 # cf. reference Makefile:18 (-ftz "critical for performance")
    O.set_warmup_steps(warmup)
    run(**kw, kind="timed", **extra)
    sec = O.last_loop_seconds()
    O.set_warmup_steps(0)
    return sec, O.num_threads()


def cpu_arm(name, n_gpus, steps, warmup, exact=True):
    """Times `steps` time steps (after `warmup`) of the CPU restatement of the reference loop on the workload's
    own grid (exact=True and the host has the memory for the reference's full-grid arrays), else on a bounded
    sample of it.  Inputs come from oracle/workloads.py.  Returns (Gpts/s, seconds, threads, sample text)."""
    from oracle import oracle as O
    from oracle import workloads as WL
    w = WORKLOADS[name]
    kind, nx, ny = w["kind"], w["nx"], w["ny"]
    nstep = steps + warmup
    avail = _mem_available_bytes()
    if kind == "3d":
        nz = w["nz_per_gpu"] * n_gpus
        need = lambda nz_: 8.0 * nx * ny * (nz_ + 4) * 27 * 1.15     # 9 fields + 18 full-grid memory variables (:275)
        nz_s = nz
        if not exact:
            nz_s = min(nz, 160)
        while need(nz_s) > 0.6 * avail and nz_s > 40:
            nz_s = max(40, nz_s // 2 // 2 * 2)
        kw = WL.iso3d(nx, ny, nz_s, nstep, dt=w["deltat"], npml=w["npml"])
        sec, cores = _timed(O.run_3d_iso, kw, steps, warmup, nproc=2)
        pts = float(nx) * ny * nz_s * steps
        what = (f"the workload grid itself, {nx}x{ny}x{nz_s}" if nz_s == nz else
                f"{nx}x{ny}x{nz_s} z-reduced sample of the workload grid")
        sample = (f"{what} (2 emulated MPI slabs), {steps} timed steps after {warmup}; full-grid memory variables and "
                  "separate Dirichlet/energy passes as in the reference; FTZ/DAZ on; inputs from oracle/workloads.py")
    elif kind == "3dv":
        nz = w["nz_per_gpu"] * n_gpus
        need = lambda ny_, nz_: 8.0 * (nx + 2) * (ny_ + 2) * (nz_ + 16) * 45 * 1.15   # 45 full arrays
        ny_s, nz_s = ny, nz
        if not exact:
            ny_s, nz_s = min(ny, 256), min(nz, 80)
        while need(ny_s, nz_s) > 0.6 * avail and (nz_s > 40 or ny_s > 128):
            if nz_s > 40: nz_s = max(40, nz_s // 2 // 4 * 4)
            else: ny_s = max(128, ny_s // 2)
        nproc = _visco_nproc(n_gpus, nz) if nz_s == nz else 4
        kw = WL.visco3d(nx, ny_s, nz_s, nstep, dt=w["deltat"], npml=w["npml"])
        sec, cores = _timed(O.run_3d_visco, kw, steps, warmup, nproc=nproc)
        pts = float(nx) * ny_s * nz_s * steps
        what = (f"the workload grid itself, {nx}x{ny_s}x{nz_s}" if (ny_s, nz_s) == (ny, nz) else
                f"{nx}x{ny_s}x{nz_s} reduced sample of the workload grid")
        sample = (f"{what} ({nproc} emulated MPI slabs), {steps} timed steps after {warmup}; full-grid memory variables "
                  "and separate Dirichlet/energy passes as in the reference; FTZ/DAZ on; inputs from oracle/workloads.py")
    elif kind == "2dv":
        n_s = nx if exact else 1001
        kw = WL.visco2d(w["order"], n_s, n_s, nstep, dt=w["deltat"], npml=w["npml"])
        sec, cores = _timed(O.run_2d_visco, kw, steps, warmup)
        cores = 1                                               # the 2-D programs are serial
        pts = float(n_s) * n_s * steps
        sample = f"{n_s}x{n_s} grid, {steps} timed steps after {warmup}, serial like the reference; inputs from oracle/workloads.py"
    else:
        n_s = nx if exact else 1024
        kw = WL.iso2d(w["order"], n_s, n_s, nstep, npml=w["npml"])
        sec, cores = _timed(O.run_2d, kw, steps, warmup)
        cores = 1
        pts = float(n_s) * n_s * steps
        sample = f"{n_s}x{n_s} grid, {steps} timed steps after {warmup}, serial like the reference; inputs from oracle/workloads.py"
    return pts / sec / 1e9, sec, cores, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind = WORKLOADS[args.workload]["kind"]
    v, sec, cores, sample = cpu_arm(args.workload, args.gpus, args.steps, args.warmup, exact=True)
    line = {"impl": "reference", "metric": metric_name(kind), "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic (fields start at zero, analytic source; no RNG)",
            "config": common_config(args.workload, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "C/OpenMP restatement of the reference loops (oracle/cpml_oracle*.c, "
                                     "gcc -O3 -march=x86-64-v3 -fopenmp); the Fortran reference cannot be "
                                     "compiled in this image (no Fortran compiler, no MPI)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------

def run_b200(args):
    import torch
    import torch.distributed as dist
    from seismic_cpml_b200 import lib as L
    from seismic_cpml_b200 import programs as P
    from seismic_cpml_b200.slab import GpuSlab, SlabDriver, owner_of_plane

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                             "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    K, W = args.steps, max(args.warmup, 3)
    nstep_total = W + K + K + 8          # warm-up + device-timed + e2e-timed (+ slack)
    p, kind = workload_params(args.workload, world, nstep_total)
    s = P.setup_3d(p) if kind == "3d" else P.setup_3d_visco(p) if kind == "3dv" else P.setup_2d_visco(p) if kind == "2dv" else P.setup_2d(p)
    f32 = WORKLOADS[args.workload].get("precision", 0) == 1
    if f32 and world != 1:
        raise SystemExit("cfg3f (single precision) runs on one GPU")
    if kind == "3d":
        sol = P.make_solver_3d(p, s, nslabs=world, slab_rank=rank, device=local_rank, precision=1 if f32 else 0)
        pts_step_rank = float(p.NX) * p.NY * (p.NZ // world)
    elif kind == "3dv":
        sol = P.make_solver_3d_visco(p, s, nslabs=world, slab_rank=rank, device=local_rank)
        pts_step_rank = float(p.NX) * p.NY * (p.NZ // world)
    elif kind == "2dv":
        sol = P.make_solver_2d_visco(p, s, device=local_rank)
        pts_step_rank = float(p.NX) * p.NY
    else:
        sol = P.make_solver_2d(p, s, device=local_rank)
        pts_step_rank = float(p.NX) * p.NY
    is3d = kind in ("3d", "3dv")
    slab = GpuSlab(sol) if is3d else sol
    drv = SlabDriver(slab, rank, world, sol.nzl, halo=args.halo, visco=(kind == "3dv")) if (is3d and world > 1) else None
    if not is3d:
        sol.set_stream(torch.cuda.current_stream().cuda_stream)

    def do_steps(a, b):
        if drv is not None:
            for it in range(a, b + 1):
                drv.step(it)
        else:
            for it in range(a, b + 1):
                sol.step_stress(it); sol.step_velocity(it); sol.step_finish(it)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    do_steps(1, W)
    barrier()

    # ---- device-timed region: K steps, fields resident in HBM, CUDA events, max over ranks
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    sol.get_kernel_times(reset=True)
    sol.enable_kernel_timing(True)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    do_steps(W + 1, W + K)
    ev1.record()
    barrier()
    t1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    ms_stress, ms_velocity, n_launch = sol.get_kernel_times(reset=True)
    sol.enable_kernel_timing(False)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = pts_step_rank * world * K / (ms_max * 1e-3) / 1e9

    # ---- e2e: the same K steps through the public driver API with HOST buffers, structured
    # like the reference's `do it` body: every step the host evaluates the source term
    # (:1058-1071) and hands it over (pinned staging, H2D), runs the step, and queues the read
    # of that step's energy and receiver sample (D2H); on the reference's display schedule
    # (it == 5 and every IT_DISPLAY steps, :1183) the max norm, the seismograms and two snapshot
    # planes come back; at the end the full traces.
    h2d = d2h = 0
    snap_pending, snap_sum = False, 0.0
    # (first use of the snapshot path allocates its pinned / device buffers: done here, outside the timed region,
    # like the warm-up steps -- a driver pays it once per run, at its first display)
    if rank == (owner_of_plane(p.NZ // 2, p.NZ, world) if is3d else 0):
        for slot in (0, 1):
            sol.snapshot_begin(slot, slot, p.NZ // 2 if is3d else 0)
            sol.snapshot_end(slot, copy=False)
    barrier()
    te0 = time.perf_counter()
    a0 = W + K + 1
    for rel in range(1, K + 1):
        it = a0 + rel - 1
        sol.set_source_step(it, s.force_x[it - 1], s.force_y[it - 1])
        h2d += 16
        do_steps(it, it)
        sol.fetch_step(it)
        d2h += 32
        if rel == 5 or rel % p.IT_DISPLAY == 0:          # :1183
            vn = drv.maxnorm() if drv is not None else sol.get_maxnorm()
            d2h += 8
            if vn > P.STABILITY_THRESHOLD:
                raise SystemExit("code became unstable and blew up")
            # the two snapshot planes of the display (vx, vy of the cut plane / the 2-D fields): asynchronous pulls
            # (device-side copy + D2H into pinned memory on a side stream); the planes of the PREVIOUS display are
            # collected first -- the image of step it is written while the loop is already beyond it
            own = owner_of_plane(p.NZ // 2, p.NZ, world) if is3d else 0
            if rank == own:
                sx, sy = sol.get_seismograms()
                d2h += sx.nbytes + sy.nbytes
                kcut = p.NZ // 2 if is3d else 0
                for slot in (0, 1):
                    if snap_pending:
                        pv = sol.snapshot_end(slot, copy=False)
                        snap_sum += float(pv[pv.shape[0] // 2, pv.shape[1] // 2])
                    sol.snapshot_begin(slot, slot, kcut)
                    d2h += 8 * p.NX * p.NY
                snap_pending = True
    if snap_pending:
        for slot in (0, 1):
            pv = sol.snapshot_end(slot, copy=False)
            snap_sum += float(pv[pv.shape[0] // 2, pv.shape[1] // 2])
    sx, sy = sol.get_seismograms()
    en = sol.get_energy()
    d2h += sx.nbytes + sy.nbytes + 3 * en[0].nbytes
    barrier()
    te1 = time.perf_counter()
    last = sol.get_fetched_step(a0 + K - 1)
    if not np.isfinite(last).all() or abs(last[0] + last[1] - en[0][a0 + K - 2]) > 1e-9 * max(1.0, abs(en[0][a0 + K - 2])):
        raise SystemExit("per-step result read back through the pinned staging does not match the energy trace")
    te = torch.tensor([te1 - te0], dtype=torch.float64, device="cuda")
    td = torch.tensor([float(d2h)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(td, op=dist.ReduceOp.MAX)
    e2e_value = pts_step_rank * world * K / float(te.item()) / 1e9
    finite = bool(np.isfinite(en[0]).all() and np.isfinite(sx).all())

    # ---- roofline: the step (both update kernels) is the headline, the two kernels are sub-records
    b_stress, b_velocity = sol.algorithmic_bytes()
    peak, peak_src = measured_peak()
    step_ms = ms_max / K
    ach_s = b_stress / (ms_stress / K * 1e-3) / 1e9 if ms_stress > 0 else None
    ach_v = b_velocity / (ms_velocity / K * 1e-3) / 1e9 if ms_velocity > 0 else None
    ach_step = (b_stress + b_velocity) / (step_ms * 1e-3) / 1e9
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            traffic = json.load(f).get(args.workload, {})
    except Exception:
        pass
    names = {"3d": ("k_stress3d_ws", "k_velocity3d_ws"), "3dv": ("k_vstress3d", "k_vvelocity3d"),
             "2dv": ("k_vstress2d", "k_vvelocity2d"), "2d": ("k_stress2d_pair", "k_velocity2d_pair")}[kind]
    if kind == "3d":
        names = sol.kernel_names()
    elif kind == "3dv" and sol.launch_info()["tma"] == 2:
        names = ("k_vstress3d", "k_vvelocity3d_ws")
    elif kind == "2d" and os.environ.get("CPML_2D_KERNEL", "ws") != "pair":
        names = ("k_stress2d_ws", "k_velocity2d_ws")
    t_s, t_v = traffic.get("stress_dram_bytes_per_launch"), traffic.get("velocity_dram_bytes_per_launch")

    if rank == 0:
        cfg = common_config(args.workload, world)
        line = {
            "metric": metric_name(kind).replace("FP64", "FP32") if f32 else metric_name(kind), "value": value, "unit": UNIT,
            "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if f32 else "f64", "data": "synthetic (fields start at zero, analytic source; no RNG)",
            "config": cfg,
            "run": {"halo": (None if world == 1 else
                             "boundary planes stored straight into the neighbour GPU's halo planes by the update kernels "
                             "over NVLink (CUDA IPC), ordered by device-side flags" if args.halo == "p2p" else
                             "NCCL send/recv of the halo planes between the kernels"),
                    "fmad": False, "finite": finite, "launch": sol.launch_info() if is3d else None},
            "roofline": {"bound": "hbm", "kernel": f"time step = {names[0]} + {names[1]}",
                         "achieved": ach_step, "peak": peak, "unit": "GB/s", "frac": ach_step / peak,
                         "traffic": (t_s + t_v) if (t_s and t_v) else None, "peak_source": peak_src,
                         "algorithmic_bytes_per_step": b_stress + b_velocity, "step_ms": step_ms,
                         "slowest_kernel": names[1] if (ach_v or 0) < (ach_s or 0) else names[0],
                         "kernels": {
                             names[0]: {"achieved": ach_s, "frac": (ach_s / peak) if ach_s else None,
                                        "algorithmic_bytes_per_launch": b_stress, "avg_launch_ms": ms_stress / K,
                                        "traffic": t_s},
                             names[1]: {"achieved": ach_v, "frac": (ach_v / peak) if ach_v else None,
                                        "algorithmic_bytes_per_launch": b_velocity, "avg_launch_ms": ms_velocity / K,
                                        "traffic": t_v}}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / K,
                    "d2h_bytes_per_step": float(td.item()) / K,
                    "what": "per step: cpml_set_source_step (16 B pinned H2D) + the step + cpml_fetch_step (32 B D2H); "
                            "reference display schedule (max norm, seismograms, 2 snapshot planes through "
                            "cpml_snapshot_begin/_end at it==5 and every IT_DISPLAY) + final traces, host buffers"},
            "gpu_launches": int(n_launch),
            "clocks": clocks,
        }
        if args.cpu_baseline and world == 1:
            try:
                big = kind in ("3d", "3dv")
                v, sec, cores, sample = cpu_arm(args.workload, 1, 24 if big else 40, 2, exact=False)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                        "seconds": sec}
            except Exception as exc:   # the oracle is only the yardstick; never fail the GPU number on it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
        print(json.dumps(line), flush=True)
    sol.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg3f", "cfg4", "cfg2", "cfg5", "cfg5d", "cfg6"])
    ap.add_argument("--halo", default="p2p", choices=["p2p", "sendrecv"],
                    help="N > 1: peer stores from inside the kernels (default) or NCCL send/recv")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
