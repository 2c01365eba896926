#!/usr/bin/env python
"""bench.py -- MLUPS of the hybird LB hot path on B200 (fp64 D3Q19), driver contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

A "step" is one LB time step (LB::latticeBolzmannStep and, for free-surface workloads, the
free-surface step in front of it) over the whole lattice.  Workload at N=1: BASELINE.json
configs[1], the pure-fluid D3Q19 periodic channel 256^3 (BGK, Newtonian, body force, no
particles); at N>1 the channel is extended along z to 256x256x(254*N+2) and cut into N z-slabs
(weak scaling), one rank per GPU, halo planes exchanged every step.

Prints ONE JSON line (rank 0).  `value` = active-cell updates / s with the state resident in HBM,
timed with CUDA events on the engine's stream; `e2e` = the same metric for the whole job through
the C ABI with host buffers, wall clock: lbGpuInit from host arrays, every step's lbGpuStep +
lbGpuParticleForces, one lbGpuFetchFields (`e2e.step_value`: the per-step calls alone);
`roofline` = 304 B per
update x active cells / fused-kernel time against MEASURED_PEAKS.json; `cpu_baseline` = the
unmodified reference (oracle/_ref/ref_harness) timed on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_PER_UPDATE = 2 * 19 * 8  # 19 fp64 populations read + 19 written
METRIC = "MLUPS (fp64 D3Q19) at 1/2/4/8 B200 and % of HBM roofline vs ref CPU"


_REAL_STDOUT = None


def emit(line: dict):
    """The one JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            for ln in self.proc.stdout:
                self.lines.append((time.time(), ln.strip()))
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self, t0, t1):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [l for (t, l) in self.lines if t0 <= t <= t1] or [l for (_, l) in self.lines]
        for ln in rows:
            c = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except Exception:
                continue
            for k, nm in enumerate(names):
                if len(c) > 5 + k and c[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def workload_label(name, n_slabs=1):
    if name == "cfg2":
        return "cfg2: D3Q19 periodic channel 256x256x%d, BGK Newtonian tau=1, body force, no particles" % (
            256 if n_slabs == 1 else 254 * n_slabs + 2)
    return {"cfg1": "cfg1: the shipped lbmConfigDevisFluid.cfg case, 100x150x30, free surface + Smagorinsky, periodic in x",
            "cfg3": "cfg3: single sphere r=8 in a 128x128x256 no-slip box, particle coupling + force reduction",
            "cfg4": "cfg4: free-surface dam break 512x128x256, Bingham rheology",
            "cfg5": "cfg5: debris flow, 20000 spheres in a 1024x256x256 free-surface fluid (long axis stored as z), "
                    "%d z-slab(s)" % n_slabs,
            "cfg5_dem": "cfg5 with the DEM in the loop on the device: 20000 spheres advanced every cycle (contacts, ghost particles "
                        "of the periodic y faces), %d z-slab(s)" % n_slabs}.get(name, name)


def workload_case(name, n_slabs=1):
    from hybird_b200 import workloads
    case = workloads.materialise(dict(workloads.catalogue()[name]))
    if n_slabs > 1 and name == "cfg2":
        case["lbSizeZ"] = 254 * n_slabs + 2  # weak scaling: every rank keeps 254 interior planes of 256x256
    return case


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (unmodified, oracle/_ref/ref_harness)
# ---------------------------------------------------------------------------------------------
def reference_sample_case(name, full=False):
    """The workload for the CPU runs.  full: the lattice of the GPU arm itself (the reference arm); else a bounded sample
    -- same physics and boundaries, the periodic channel cropped in y (the flow is invariant along y) so that a step
    takes ~0.1 s on 16 threads and the 1-thread run of BASELINE.md section 4 stays within seconds."""
    case = workload_case(name)
    if name == "cfg2" and not full:
        case["lbSizeY"] = 34
        sample = "cfg2 channel cropped to 256x34x256 (2.2 M cells, same physics/boundaries)"
    else:
        sample = "%s at full size (%sx%sx%s)" % (name, case["lbSizeX"], case["lbSizeY"], case["lbSizeZ"])
    case["name"] = name + "_cpu_sample"
    return case, sample


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(name, steps, warmup, threads=None, full=False):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cases
    threads = threads or (os.cpu_count() or 1)
    case, sample = reference_sample_case(name, full)
    kind = "reference"
    phases = None
    with tempfile.TemporaryDirectory() as wd:
        if os.path.exists(cases.REF_HARNESS):
            _, so = cases.run_reference(case, wd, steps, time_mode=True, warmup=warmup, threads=threads)
            rec = json.loads([l for l in so.splitlines() if l.startswith("{")][-1])
            mlups, ms, act = rec["mlups_active"], rec["ms_per_step"], rec["active_mean"]
            threads = rec["threads"]
            tot = sum(rec["phase_ms"].values()) or 1.0
            phases = {k: round(100.0 * v / tot, 1) for k, v in rec["phase_ms"].items()}  # per cent of the step
        else:  # the C restatement (port) when the compiled reference did not travel
            kind = "port"
            mlups, ms, act = run_port(case, steps, warmup, threads)
    return dict(value=mlups, unit="MLUPS", cores=int(threads), kind=kind, sample=sample + ", %d steps after %d warm-up" % (steps, warmup),
                ms_per_step=ms, active_cells=act, phase_pct=phases, cpu=cpu_model(), host_cores=os.cpu_count(),
                lattice=[int(round(float(case["lbSize" + a]))) for a in "XYZ"])


def run_port(case, steps, warmup, threads):
    import numpy as np
    import lbo
    from hybird_b200 import lattice_init as li
    st = li.build_state(case)
    o = lbo.Oracle(st.params, st.type_flags, st.solidIndex, st.n, st.u, st.mass, st.visc, threads=threads)
    fs = st.params["freeSurface"]
    def one():
        if fs:
            o.latticeBoltzmannFreeSurfaceStep()
        o.latticeBolzmannStep()
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    act = float(np.count_nonzero(np.isin(o.type_flags & 15, (0, 3))))
    o.close()
    return act * steps / dt / 1e6, 1e3 * dt / steps, act


def main_reference(args, rank, world):
    if rank != 0:
        return 0
    # the GPU arm's own lattice (same config): cfg2 at 256^3 takes ~1-3 s per step on the box's host cores plus ~15 s
    # of the reference's serial initialisation, so the step count is bounded to keep the run within a few minutes
    full = not args.ref_crop
    steps = max(1, min(args.steps, 20 if full else 100))
    warmup = max(1, min(args.warmup, 2 if full else 3))
    r = run_reference(args.workload, steps, warmup, full=full)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "MLUPS", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_label(args.workload, args.gpus), "sample": r["sample"]},
            "cpu_baseline": {"value": r["value"], "unit": "MLUPS", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                             "cpu": r["cpu"], "host_cores": r["host_cores"], "lattice": r["lattice"],
                             "active_cells": r["active_cells"], "phase_pct": r["phase_pct"]},
            "e2e": {"value": r["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def resident_run(name, args, rank, world, local_rank, dist, K, W, with_clocks):
    """Device-resident throughput of one workload: W warm-up steps, then exactly K steps back to back (lbGpuRun; the
    free-surface update, particle-flag update, LB step and force reduction of every cycle), CUDA events on the engine's
    stream, barrier + synchronize on both sides, maximum over the ranks.  Returns (engine, info, result dict)."""
    import torch
    from hybird_b200 import slabs
    case = workload_case(name, world)
    t_init = time.time()
    lb, info = slabs.build_engine(case, rank, world, device=local_rank, dist=dist if world > 1 else None,
                                  device_init=not args.host_init)
    t_init = time.time() - t_init

    def barrier():
        lb.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parts, elmts, comps = info["parts"], info["elmts"], info["comps"]
    fs = bool(info["params"]["freeSurface"])
    from hybird_b200 import dem_init
    run = lb.run
    info["dem"] = "none"
    if len(parts) and case.get("motion") == "dem" and dem_init.covered(case, info["params"]):
        # the configuration as the reference runs it: the DEM advances the spheres every cycle -- here on the device, inside
        # the same call (lbGpuRunDem: DEM sub-steps, coupling step, LB step; no host round trip)
        lb.demInit(dem_init.dem_from_case(case, info["params"]))
        run = lb.runDem
        info["dem"] = "device (lbGpuRunDem)"
    elif len(parts):
        # particle state becomes resident with the first coupling step (rescan, as dem.newNeighborList does); the particles
        # then stay where they are (the device-side DEM does not cover periodic boundaries / clusters)
        if fs:
            lb.latticeBoltzmannFreeSurfaceStep()
        lb.latticeBoltzmannCouplingStep(True, elmts, parts, comps)
        lb.latticeBolzmannStep(elmts, parts)
        info["dem"] = "particles fixed (lbGpuRun)"
    run(W)
    barrier()
    sampler = ClockSampler(local_rank) if (rank == 0 and with_clocks) else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = lb.launch_count()
    barrier()
    t0w = time.time()
    run(K)
    lb.synchronize()
    ms_dev = lb.last_step_ms()             # CUDA events on the engine's stream around the K steps
    kern_ms, kern_n = lb.last_kernel_ms()  # CUDA events around each step-kernel group (last <=512 of the K)
    kern_ms = kern_ms / max(kern_n, 1)
    t1w = time.time()
    launches = lb.launch_count() - l0
    barrier()
    clocks = {}
    if sampler:
        time.sleep(0.2)
        sampler.stop()
        clocks = sampler.summary(t0w, t1w)
    if world > 1:
        t = torch.tensor([ms_dev, kern_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, kern_ms = float(t[0]), float(t[1])
    peak, peak_src = load_peaks()
    active_total, active_local = info["active_total"], info["active_local"]
    achieved = BYTES_PER_UPDATE * active_local / (kern_ms * 1e-3) / 1e9
    size = info["params"]["size"]
    res = dict(value=active_total * K / (ms_dev * 1e-3) / 1e6, ms_per_step=ms_dev / K, launches=int(launches), clocks=clocks,
               init_s=round(t_init, 2), barrier=barrier,
               lattice=[int(size[0]), int(size[1]), int(info["global_z"])], active_cells=int(active_total), dem=info["dem"],
               roofline={"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "kernel": info["kernel"], "bytes_per_update": BYTES_PER_UPDATE,
                         "updates_per_launch": int(active_local), "kernel_ms": kern_ms,
                         "whole_step_frac": BYTES_PER_UPDATE * active_local / (ms_dev / K * 1e-3) / 1e9 / peak})
    return lb, info, res


def traffic_of(name):
    """DRAM bytes per step-kernel launch from the committed ncu capture (profiles/traffic.json): a citation of that
    capture, not a measurement of this run."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        rec = json.load(open(tp)).get(name, {})
        return rec.get("dram_bytes_per_launch"), "profiles/traffic.json (%s)" % rec.get("source", "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum")
    except Exception:
        return None, None


def main_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist

    from hybird_b200 import LB, lattice_init as li  # noqa: F401
    from hybird_b200 import slabs

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        slabs.init_comm(rank, world, local_rank, dist)  # the engine's own NCCL communicator (set-up, sums, in-cycle planes)

    K, W = args.steps, max(args.warmup, 3)
    lb, info, res = resident_run(args.workload, args, rank, world, local_rank, dist, K, W, True)
    barrier = res.pop("barrier")
    active_total = info["active_total"]
    parts, elmts, comps = info["parts"], info["elmts"], info["comps"]
    fs = bool(info["params"]["freeSurface"])
    case = workload_case(args.workload, world)

    # ---- end to end through the C ABI with host buffers: lbGpuStep + lbGpuParticleForces per step ----
    Ke = min(K, 200)
    h2d = int(parts.nbytes + elmts.nbytes + comps.nbytes)
    d2h = int(8 * (7 * len(elmts) + 3 * lb.nWalls))
    x0_elmt = parts["x0"].copy() if len(parts) else None
    dem_on_device = str(info.get("dem", "")).startswith("device")
    if dem_on_device:
        h2d = 0  # the particles never leave the device; the forces still come back every cycle
    def e2e_step():
        if dem_on_device:  # the DEM step, the coupling step and the LB step of one cycle, then the forces on the host
            lb.runDem(1)
            return lb.forces()
        if fs:
            lb.latticeBoltzmannFreeSurfaceStep()
        if len(parts):
            li.advance_kinematic(parts, elmts, x0_elmt, 1.0)  # the host moves the spheres (stand-in for the DEM step)
            lb.latticeBoltzmannCouplingStep(False, elmts, parts, comps)
        return lb.latticeBolzmannStep(elmts, parts)  # returns host F, M, V, wallF (synchronous)
    for _ in range(3):
        e2e_step()
    barrier()
    te = time.perf_counter()
    for _ in range(Ke):
        e2e_step()
    lb.synchronize()
    te = time.perf_counter() - te
    if world > 1:
        t = torch.tensor([te], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t[0])
    e2e_mlups = active_total * Ke / te / 1e6

    # ---- the whole job through the C ABI with host buffers: lbGpuInit from the caller's host arrays (what the drop-in
    # shim does after the reference's host initialisation), K steps as above, one lbGpuFetchFields into the caller's
    # host arrays (what an export step of the reference's IO reads; the caller's mirrors exist before the job, as the
    # shim's do).  One process: a second engine, one contiguous timed region ("measured").  Several processes: this
    # rank's own slab uploaded from host arrays + steps + fetch, composed from separately timed legs, maximum over the
    # ranks ("composed"). ----
    FIELDS = ("type_flags", "n", "u", "mass")
    if world == 1:
        Kj = min(K, 1000)
        # the job is a run of its own, from host arrays to host arrays: the resident engine is closed first (its device memory
        # stays with the library for the next engine of the same shape, see DevCache in lbgpu.cu)
        lb.close()
        st_host = li.build_state(case, parts if len(parts) else None)  # the caller's data: not timed
        # ... resident in host memory like a driver's own arrays: numpy hands out untouched zero pages for np.zeros (solidIndex of
        # a lattice without particles), whose first read inside lbGpuInit would be timed as page faults, not as a transfer
        host_arrays = []
        for name in ("type_flags", "solidIndex", "n", "u", "mass", "visc"):
            a = getattr(st_host, name)
            if name == "solidIndex" and not a.any():
                a = np.array(a, copy=True)  # np.zeros, never written: touch it (one array at a time, the lattice is 0.9 GB)
            host_arrays.append(a)
        host_params, n_cells = st_host.params, st_host.type_flags.size
        upload_bytes = int(sum(a.nbytes for a in host_arrays))
        xj = parts["x0"].copy() if len(parts) else None
        pj = parts.copy()
        mirrors = LB.host_mirrors(n_cells, FIELDS)  # the caller's (touched) output arrays: not timed
        tj = time.perf_counter()
        lbj = LB(host_params, device=local_rank)
        lbj.latticeBolzmannInit(*host_arrays)
        t_up = time.perf_counter() - tj
        if dem_on_device:
            from hybird_b200 import dem_init as _di
            lbj.demInit(_di.dem_from_case(case, host_params))
        for k in range(Kj):
            if dem_on_device:  # one cycle with the DEM on the device, the forces back on the host
                lbj.runDem(1)
                lbj.forces()
                continue
            if fs:
                lbj.latticeBoltzmannFreeSurfaceStep()
            if len(pj):
                li.advance_kinematic(pj, elmts, xj, 1.0)
                lbj.latticeBoltzmannCouplingStep(k == 0, elmts, pj, comps)
            lbj.latticeBolzmannStep(elmts, pj)
        tf = time.perf_counter()
        fields = lbj.fetch(FIELDS, out=mirrors)
        job_s = time.perf_counter() - tj
        fetch_ms = 1e3 * (time.perf_counter() - tf)
        init_ms = 1e3 * t_up
        fetch_bytes = int(sum(v.nbytes for v in fields.values()))
        del fields, mirrors
        lbj.close()
        del st_host, host_arrays
        e2e_kind = "measured"
    else:
        Kj = K
        st_host, _ = slabs.build_slab_state(case, rank, world, parts if len(parts) else None, dist)  # not timed
        upload_bytes = int(sum(a.nbytes for a in (st_host.type_flags, st_host.solidIndex, st_host.n, st_host.u, st_host.mass, st_host.visc)))
        mirrors = LB.host_mirrors(lb.N, FIELDS)
        tf = time.perf_counter()
        fields = lb.fetch(FIELDS, out=mirrors)
        fetch_ms = 1e3 * (time.perf_counter() - tf)
        fetch_bytes = int(sum(v.nbytes for v in fields.values()))
        del fields, mirrors
        barrier()
        lb.close()
        tj = time.perf_counter()
        lbj = LB(st_host.params, device=local_rank)
        lbj.latticeBolzmannInit(st_host.type_flags, st_host.solidIndex, st_host.n, st_host.u, st_host.mass, st_host.visc)
        lbj.synchronize()
        init_ms = 1e3 * (time.perf_counter() - tj)
        lb = lbj
        del st_host
        t = torch.tensor([1e-3 * init_ms + K * te / Ke + 1e-3 * fetch_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        job_s = float(t[0])
        e2e_kind = "composed"
    job_mlups = active_total * Kj / job_s / 1e6
    if world > 1:
        dist.barrier()
    lb.close()

    # ---- the other configurations, device-resident, so that they are measured by the same driver run ----
    extra = {}
    if args.workload == "cfg2" and not args.no_extra:
        for nm in (("cfg1", "cfg3", "cfg4", "cfg5", "cfg5_dem") if world == 1 else ("cfg5",)):
            try:
                Kx = min(K, 100)
                lbx, infx, rx = resident_run(nm, args, rank, world, local_rank, dist, Kx, W, False)
                rx.pop("barrier")
                rx.pop("clocks")
                rx["steps"] = Kx
                rx["scaling"] = "strong"
                rx["launches_per_step"] = round(rx.pop("launches") / Kx, 1)
                rx["workload"] = workload_label(nm, world)
                rx["roofline"]["traffic"], rx["roofline"]["traffic_source"] = traffic_of(nm)
                rx["peer_halo"] = lbx.peer_halo() if world > 1 else None
                # where the device time of a cycle goes (CUDA events at the phase boundaries, a separate short run: the
                # trace launches every cycle eagerly); rank 0's view
                lbx.phase_trace(True)
                (lbx.runDem if rx["dem"].startswith("device") else lbx.run)(min(Kx, 32))
                ph, ncyc = lbx.phase_ms()
                lbx.phase_trace(False)
                rx["phase_ms"] = {k: round(v, 4) for k, v in ph.items()}
                rx["phase_ms"]["cycles"] = ncyc
                if world > 1:
                    dist.barrier()
                lbx.close()
                extra[nm] = rx
            except Exception as e:  # noqa: BLE001
                extra[nm] = {"error": str(e)[:300]}
                break  # the ranks may no longer be in step

    if rank != 0:
        if world > 1:
            slabs.finalize_comm()
            dist.destroy_process_group()
        return 0

    roof = res["roofline"]
    roof["traffic"], roof["traffic_source"] = traffic_of(args.workload)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = run_reference(args.workload, 60, 3)    # ~8 s of CPU work on 16 threads
            r1 = run_reference(args.workload, 4, 1, threads=1)  # the same at one thread
            cpu = {"value": r["value"], "unit": "MLUPS", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                   "cpu": r["cpu"], "host_cores": r["host_cores"], "lattice": r["lattice"], "active_cells": r["active_cells"],
                   "phase_pct": r["phase_pct"], "value_1_thread": r1["value"], "phase_pct_1_thread": r1["phase_pct"]}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "MLUPS", "cores": 0, "kind": "reference", "sample": "failed: %s" % str(e)[:200]}
    line = {
        "metric": METRIC, "value": res["value"], "unit": "MLUPS", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak" if args.workload == "cfg2" else "strong",
        "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_label(args.workload, world),
                   "lattice": res["lattice"], "active_cells": res["active_cells"],
                   "parallelism": info["parallelism"], "particles": info.get("dem", "none"),
                   "l2": "working set %.1f GB per GPU >> 126 MB L2 (no flush needed)" % (info["bytes_resident"] / 1e9),
                   "init_s": res["init_s"], "init": info["init"]},
        "roofline": roof,
        # value: the whole job through the C ABI with host buffers -- state upload, every step's particle arrays in and
        # forces out, one field fetch; step_value: the per-step calls alone (what the reference arm's per-step time is
        # the counterpart of; a pure-fluid step has no host input)
        "e2e": {"value": job_mlups, "unit": "MLUPS", "kind": e2e_kind,
                "h2d_bytes_per_step": int(h2d + upload_bytes / Kj), "d2h_bytes_per_step": int(d2h + fetch_bytes / Kj),
                "steps": Kj,
                "call": ("lbGpuInit(host arrays) + lbGpuDemInit + %d x [lbGpuRunDem(1) + lbGpuParticleForces(host)] + lbGpuFetchFields(host arrays)" % Kj) if dem_on_device else "lbGpuInit(host arrays) + %d x [lbGpuStep(host particle/element arrays) + lbGpuParticleForces(host)] + "
                        "lbGpuFetchFields(host arrays)" % Kj,
                "step_value": e2e_mlups, "step_h2d_bytes": h2d, "step_d2h_bytes": d2h, "step_steps": Ke,
                "init_upload_bytes": upload_bytes, "fetch_fields_bytes": fetch_bytes,
                "init_ms": init_ms, "fetch_fields_ms": fetch_ms},
        "gpu_launches": res["launches"],
        "clocks": res["clocks"],
    }
    if world > 1:
        line["config"]["halo"] = info.get("halo")
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if extra:
        line["extra"] = extra
    emit(line)
    if world > 1:
        slabs.finalize_comm()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the device-resident lines of the other configurations (extra block)")
    ap.add_argument("--ref-crop", action="store_true", help="reference arm on the cropped 256x34x256 sample instead of the GPU arm's lattice")
    ap.add_argument("--host-init", action="store_true", help="build the initial state on the host and upload it (lbGpuInit) "
                    "instead of initialising the lattice on the device (lbGpuInitBox; one process only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: whatever libraries print there (NCCL's version banner ...) goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        os.dup2(_REAL_STDOUT, 1)  # the ranks inherit the real stdout and keep it clean themselves
        return subprocess.call(cmd)
    if args.impl == "reference":
        return main_reference(args, rank, world)
    return main_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
