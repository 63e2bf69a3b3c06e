#!/usr/bin/env python
"""Benchmark of the DEM step hot path (BASELINE.json: particle-steps/s at 1M 3D spheres).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path

Workload (BASELINE.json configs[1], SURVEY.md §8d "C2"): N = 2**20 monodisperse spheres
(r = 0.5) in a periodic cube at packing fraction 0.5, jittered simple-cubic packing,
spring contacts (young_eff 1e4), velocity Verlet, cell-list collider, float32/int32.
A "step" is one System.step (one _step_once, jaxdem/system.py:60-82) over all particles.

Prints ONE JSON line (rank 0).  Keys follow the driver contract; additionally
``roofline`` (dominant kernel), ``step_roofline`` (whole step, 168 algorithmic bytes per
particle-step), ``cpu_baseline`` and ``kernels`` (per-kernel device time of one step).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PARTICLES = 1 << 20
B_ALG_STEP = 168.0  # algorithmic bytes per particle-step, config 2
# algorithmic bytes per particle of each kernel family
B_ALG_KERNEL = {
    "k_pair_force": 28.0,   # R pos 12 + rad 4, W force 12
    # fused sphere driver: + R (vel, mass) 16 of the sorted shadow record, W vel 12 (the after-force kick)
    "k_pair_flat": 56.0,
    "k_hash": 105.0,        # R pos 12 vel 12 force 12 rad 4 mass 4 fixed 1 + external buffers 36 + pos_p_rot 12 + bond 4; W pos 12
    "k_finalize": 44.0,
    "k_scatter": 20.0,
}
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def make_workload(n=N_PARTICLES, seed=1, phi=0.5, dtype=np.float32, packing="grid", stack=1):
    """jittered simple-cubic ("grid") or uniform random ("random") packing.
    ``stack`` > 1 (multi-GPU weak scaling): ``stack`` copies of the C2 cube on top of each other
    along z — ONE periodic system of stack * n spheres in a box (L, L, stack * L)."""
    rng = np.random.default_rng(seed)
    r = 0.5
    L = (n * (4.0 / 3.0) * np.pi * r**3 / phi) ** (1.0 / 3.0)
    nt = n * stack
    if packing == "grid":
        g = int(np.ceil(n ** (1.0 / 3.0)))
        sites = rng.permutation(g**3 * stack)[:nt]
        sites.sort()
        ijk = np.stack(np.unravel_index(sites, (g * stack, g, g)), axis=1)[:, ::-1].astype(np.float64)  # (x, y, z)
        pos = (ijk + 0.5) * (L / g) + rng.uniform(-0.1, 0.1, (nt, 3)) * r
        order = rng.permutation(nt)  # particle index carries no spatial order
        pos = pos[order]
    else:
        pos = rng.uniform(0, 1, (nt, 3)) * np.array([L, L, L * stack])
    vel = rng.uniform(-1, 1, (nt, 3))
    box = np.array([L, L, L * stack])
    return dict(pos=pos.astype(dtype), vel=vel.astype(dtype), rad=np.full(nt, r, dtype),
                mass=np.ones(nt, dtype), box=box.astype(dtype))


# ---------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed region
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        if not self.rows:  # timed region shorter than the sampling period: one synchronous sample
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=20).stdout
                self.rows = [[x.strip() for x in ln.split(",")] for ln in out.splitlines() if ln.strip()]
            except (OSError, subprocess.SubprocessError):
                pass
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gbps", "hbm_GBs"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except (OSError, ValueError):
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------
# CPU restatement of the reference path (oracle/c, OpenMP)
# ---------------------------------------------------------------------------
def cpu_steps_per_s(wl, steps, warmup=1, threads=None):
    """Time `steps` System.step's of the SAME workload on the host cores with the C/OpenMP
    restatement of the reference algorithm (oracle/c): stable sort of (hash, iota), 27
    binary searches per particle, run walk, spring law, velocity Verlet."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import build_oracle
    from oracle import c_oracle
    if threads:
        c_oracle.set_num_threads(threads)
    dtype = wl["pos"].dtype.type
    # default material of System.create: elastic(young 1e4, poisson 0.3) harmonic -> young_eff 1e4
    import oracle
    st = oracle.create_state(wl["pos"], vel=wl["vel"], rad=wl["rad"], mass=wl["mass"], dtype=dtype)
    sy = oracle.create_system(st, dt=1e-3, linear_integrator_type="verlet", rotation_integrator_type="",
                              collider_type="celllist", domain_type="periodic",
                              domain_kw=dict(box_size=wl["box"]), force_model_type="spring")
    cs = c_oracle.CStep(st, sy)
    cs.step(warmup)
    t0 = time.perf_counter()
    cs.step(steps)
    dt = time.perf_counter() - t0
    return st.N * steps / dt, dt, c_oracle.num_threads()


def try_jax_reference(wl, steps):
    """BASELINE.md §3 plan 1: the UNMODIFIED reference (baseline/_ref) on JAX CPU, if JAX is
    importable (it is not in this image; then None is returned and the C port is timed)."""
    try:
        os.environ.setdefault("JAX_PLATFORMS", "cpu")
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        import jax  # noqa: F401
        import jax.numpy as jnp
        import jaxdem as jdem
    except Exception:
        return None
    state = jdem.State.create(pos=jnp.asarray(wl["pos"]), vel=jnp.asarray(wl["vel"]),
                              rad=jnp.asarray(wl["rad"]), mass=jnp.asarray(wl["mass"]))
    system = jdem.System.create(state.shape, dt=1e-3, linear_integrator_type="verlet",
                                rotation_integrator_type="", collider_type="CellList",
                                collider_kw=dict(state=state), domain_type="periodic",
                                domain_kw=dict(box_size=jnp.asarray(wl["box"])), force_model_type="spring")
    state, system = jdem.System.step(state, system, n=1)
    jax.block_until_ready(state.pos_c)
    t0 = time.perf_counter()
    state, system = jdem.System.step(state, system, n=steps)
    jax.block_until_ready(state.pos_c)
    dt = time.perf_counter() - t0
    return wl["pos"].shape[0] * steps / dt, dt, os.cpu_count()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = make_workload(n=args.n_per_gpu, packing=args.packing)
    n = wl["pos"].shape[0]
    jr = try_jax_reference(wl, max(1, min(args.steps, 20)))
    if jr is not None:
        rate, secs, cores = jr
        steps = max(1, min(args.steps, 20))
        print(json.dumps({
            "impl": "reference", "metric": "particle-steps/sec at 1M 3D spheres", "value": rate,
            "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1,
            "ms_per_step": 1e3 * secs / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, n),
            "cpu_baseline": {"value": rate, "unit": "particle-steps/s", "cores": cores, "kind": "reference",
                             "sample": f"{steps} steps, jaxdem System.step on JAX CPU"},
            "e2e": {"value": rate, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return
    # bounded sample: each "step" of this arm is one full 1M-particle step on the host cores
    steps = max(1, min(args.steps, 20))
    rate, secs, cores = cpu_steps_per_s(wl, steps, warmup=max(1, min(args.warmup, 2)))
    line = {
        "impl": "reference", "metric": "particle-steps/sec at 1M 3D spheres", "value": rate,
        "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": steps, "warmup": max(1, min(args.warmup, 2)),
        "ms_per_step": 1e3 * secs / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n),
        "cpu_baseline": {"value": rate, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} full steps of the 1M-sphere workload, C/OpenMP restatement of the "
                                   "reference cell-list path (JAX is not installed in this image: the reference "
                                   "itself cannot run)"},
        "e2e": {"value": rate, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, n):
    if args.gpus > 1 and args.mode == "slab":
        par = (f"ONE periodic system of {args.gpus} x {n} spheres (box L x L x {args.gpus}L), z-slab decomposition, "
               "one slab per GPU, NCCL halo + migration exchange every step")
    elif args.gpus > 1:
        par = "1 system per GPU (replicas)"
    else:
        par = "single GPU"
    return {"workload": f"C2: {n} monodisperse 3D spheres r=0.5, periodic cube phi=0.5, {args.packing} packing, "
                        "cell-list collider (27-cell stencil), spring contact young_eff=1e4, velocity Verlet, dt=1e-3",
            "n_particles": n, "l2": "flushed between timed steps (256 MiB write, then 256 MiB read to drain the dirty lines)",
            "parallelism": par}


# ---------------------------------------------------------------------------
# CUDA arm
# ---------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist

    import jaxdem_b200 as jd
    from jaxdem_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: jaxdem_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if world > 1 and args.mode == "slab":
        return run_cuda_slab(args, world, rank, local, dev)
    wl = make_workload(n=args.n_per_gpu, seed=1 + rank, packing=args.packing)
    n = wl["pos"].shape[0]
    st = jd.State.create(wl["pos"], vel=wl["vel"], rad=wl["rad"], mass=wl["mass"], dtype=torch.float32, device=dev)
    sy = jd.System.create(st.shape, dt=1e-3, linear_integrator_type="verlet", rotation_integrator_type="",
                          collider_type="CellList", collider_kw=dict(state=st), domain_type="periodic",
                          domain_kw=dict(box_size=wl["box"]), force_model_type="spring",
                          dtype=torch.float32, device=dev)
    lib = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        # write a buffer larger than L2, then read another one: the write evicts everything the
        # step left behind, the read drains the flush's own dirty lines so that their
        # write-back is not billed to the first kernel of the timed step
        flush.fill_(1)
        flush_rd.sum()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step = jd.System.compile_step(st, sy, n=1) if args.graph else (lambda: jd.System.step(st, sy, n=1))
    jd.System.step(st, sy, n=1)  # first force evaluation (loop-carried state.force)
    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed: K steps, each bracketed by events, L2 flushed in between ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = lib.jdb200_launch_count()
    barrier()
    for a, b in ev:
        flush_l2()
        a.record()
        step()
        b.record()
    barrier()
    launches = lib.jdb200_launch_count() - l0
    ms = float(sum(a.elapsed_time(b) for a, b in ev))
    if args.graph:  # graph replays do not pass through the counter: one capture's worth per replay
        launches = args.steps * step_launches(jd, st, sy, lib)
    # steady state without the flush (what a real rollout sees): K steps in ONE call
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    jd.System.step(st, sy, n=args.steps)
    b.record()
    torch.cuda.synchronize()
    ms_warm = a.elapsed_time(b)
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * args.steps / (ms_max * 1e-3)

    # ---- e2e: host buffers in, host buffers out, every step ----
    fields = ("pos_c", "vel", "force")
    host_in = {k: getattr(st, k).detach().cpu().pin_memory() for k in fields}
    host_out = {k: torch.empty_like(v).pin_memory() for k, v in host_in.items()}
    h2d = sum(v.numel() * v.element_size() for v in host_in.values())
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_step():
        for k in fields:
            getattr(st, k).copy_(host_in[k], non_blocking=True)
        jd.System.step(st, sy, n=1)
        for k in fields:
            host_out[k].copy_(getattr(st, k), non_blocking=True)
        torch.cuda.synchronize()
        for k in fields:  # next step's input is this step's output (a rollout driven from the host)
            host_in[k], host_out[k] = host_out[k], host_in[k]

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(t.item())

    # ---- per-kernel device time of the step (diagnostic events; separate pass) ----
    _lib.kernel_timing(True)
    prof_steps = 5
    for _ in range(prof_steps):
        flush_l2()
        jd.System.step(st, sy, n=1)
    torch.cuda.synchronize()
    kt = _lib.kernel_timing_collect()
    _lib.kernel_timing(False)
    kernels = {k: {"us_per_launch": 1e3 * v[0] / v[1], "launches_per_step": v[1] / prof_steps,
                   "us_per_step": 1e3 * v[0] / prof_steps} for k, v in kt.items()}
    tot = sum(v["us_per_step"] for v in kernels.values())
    for v in kernels.values():
        v["share"] = v["us_per_step"] / tot if tot else 0.0
    peak, peak_src = measured_peak()
    dom = max(kernels, key=lambda k: kernels[k]["us_per_step"]) if kernels else None
    roof = None
    if dom is not None:
        bpp = B_ALG_KERNEL.get(dom)
        dur_s = kernels[dom]["us_per_launch"] * 1e-6
        ach = (bpp * n / dur_s / 1e9) if bpp else None
        roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": (ach / peak) if ach else None, "traffic": None,
                "algorithmic_bytes_per_launch": (bpp * n) if bpp else None,
                "us_per_launch": kernels[dom]["us_per_launch"], "peak_source": peak_src}
        tr = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the last ncu --set full
        if os.path.exists(tr):
            try:
                roof["traffic"] = json.load(open(tr)).get(dom)
            except (OSError, ValueError):
                pass
    step_gbs = B_ALG_STEP * n * args.steps / (ms * 1e-3) / 1e9

    if rank == 0:
        line = {
            "metric": "particle-steps/sec at 1M 3D spheres", "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, n), launch="cuda-graph replay" if args.graph else "stream launches"),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": h2d, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "roofline": roof,
            "step_roofline": {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s",
                              "frac": step_gbs / peak, "algorithmic_bytes_per_particle_step": B_ALG_STEP,
                              "peak_source": peak_src},
            "l2_warm": {"value": n * args.steps / (ms_warm * 1e-3), "unit": "particle-steps/s",
                        "ms_per_step": ms_warm / args.steps,
                        "note": "same K steps back to back in one jdb200_system_step call, no L2 flush"},
            "kernels": kernels,
        }
        if world == 1 and not args.no_cpu:
            cpu_steps = args.cpu_steps
            rate, secs, cores = cpu_steps_per_s(wl, cpu_steps)
            line["cpu_baseline"] = {"value": rate, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                                    "sample": f"{cpu_steps} full steps of the same 1M-sphere workload "
                                              f"({secs:.1f} s), C/OpenMP restatement (oracle/c)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_cuda_slab(args, world, rank, local, dev):
    """N > 1 GPUs: ONE periodic system of N x 2**20 spheres, z-slab decomposition (jaxdem_b200/slab.py).
    Per-GPU work is fixed (weak scaling); every step exchanges halos and migrants over NCCL."""
    import torch
    import torch.distributed as dist

    from jaxdem_b200 import _lib
    from jaxdem_b200.slab import create_slab_system

    lib = _lib.lib()
    # every rank generates the C2 cube of its own slab (seed 1 + rank) and shifts it to its place in
    # the stack; rows that fall into a neighbour's boundary layer migrate in the first exchange
    n = args.n_per_gpu
    wl = make_workload(n=n, seed=1 + rank, packing=args.packing)
    L = float(wl["box"][0])
    wl["pos"][:, 2] += np.float32(rank * L)
    box = np.array([L, L, L * world], dtype=np.float32)
    n_total = n * world
    slab = create_slab_system(dict(pos=wl["pos"], vel=wl["vel"], rad=wl["rad"], mass=wl["mass"]),
                              box_size=box, dt=1e-3, force_model_type="spring", rotation_integrator_type="",
                              dtype=torch.float32, device=dev, capacity_factor=1.35,
                              local_gid=rank * n + np.arange(n), n_total=n_total, rad_range=(0.5, 0.5))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        flush.fill_(1)
        flush_rd.sum()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    slab.compute_force()  # first force evaluation (loop-carried state.force)
    for _ in range(args.warmup):
        slab.step(1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = lib.jdb200_launch_count()
    barrier()
    for a, b in ev:
        flush_l2()
        a.record()
        slab.step(1)
        b.record()
    barrier()
    launches = lib.jdb200_launch_count() - l0
    per_step = [a.elapsed_time(b) for a, b in ev]
    ms = float(sum(per_step))
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = n_total * args.steps / (ms_max * 1e-3)

    # ---- e2e: every rank's owned rows host -> device before the step, device -> host after it ----
    fields = ("pos_c", "vel", "force")
    host = {k: torch.empty((slab.cap, 3), dtype=torch.float32).pin_memory() for k in fields}
    e2e_steps = max(3, min(args.steps, 20))
    h2d = d2h = 0

    def e2e_step(first=False):
        nonlocal h2d, d2h
        m = slab.n_own
        if not first:
            for k in fields:
                slab.buf[k][:m].copy_(host[k][:m], non_blocking=True)
        slab.step(1)
        m2 = slab.n_own
        for k in fields:
            host[k][:m2].copy_(slab.buf[k][:m2], non_blocking=True)
        torch.cuda.synchronize()
        h2d += 0 if first else 3 * m * 12
        d2h += 3 * m2 * 12

    e2e_step(first=True)
    h2d = d2h = 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s, float(h2d), float(d2h)], dtype=torch.float64, device=dev)
    tm = t.clone()
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    e2e_value = n_total * e2e_steps / float(tm[0].item())
    peak, peak_src = measured_peak()
    step_gbs = B_ALG_STEP * n_total * args.steps / (ms_max * 1e-3) / 1e9
    own = torch.tensor([slab.n_own, slab.n_ghost], dtype=torch.int64, device=dev)
    owns = [torch.zeros_like(own) for _ in range(world)]
    dist.all_gather(owns, own)
    if rank == 0:
        line = {
            "metric": "particle-steps/sec at 1M 3D spheres", "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, n), launch="stream launches, hook by hook",
                           n_particles_total=n_total,
                           owned_ghost_rows_per_rank=[[int(x) for x in o.tolist()] for o in owns]),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "particle-steps/s",
                    "h2d_bytes_per_step": int(t[1].item() / e2e_steps), "d2h_bytes_per_step": int(t[2].item() / e2e_steps),
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "step_ms_rank0": {"min": min(per_step), "median": float(np.median(per_step)), "max": max(per_step)},
            "roofline": None,
            "step_roofline": {"bound": "hbm", "achieved": step_gbs, "peak": peak * world, "unit": "GB/s",
                              "frac": step_gbs / (peak * world), "algorithmic_bytes_per_particle_step": B_ALG_STEP,
                              "peak_source": peak_src + f" x {world} GPUs"},
        }
        print(json.dumps(line))
    dist.destroy_process_group()


def step_launches(jd, st, sy, lib):
    """Kernels one step launches (counted on a throw-away stream-launched step)."""
    import torch
    c0 = lib.jdb200_launch_count()
    jd.System.step(st, sy, n=1)
    torch.cuda.synchronize()
    return lib.jdb200_launch_count() - c0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--packing", default="grid", choices=["grid", "random"])
    ap.add_argument("--mode", default="slab", choices=["slab", "replicas"],
                    help="N > 1 GPUs: one slab-decomposed system of N x 2^20 spheres (default) or N independent replicas")
    ap.add_argument("--n-per-gpu", type=int, default=N_PARTICLES,
                    help="spheres per GPU (default 2**20 = BASELINE config 2; 2**23 x 8 GPUs = the 64M system of config 3)")
    ap.add_argument("--graph", action="store_true", help="replay the step from a CUDA graph (System.compile_step)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-steps", type=int, default=10)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # the contract is ONE JSON line on stdout: libraries that print banners there (NCCL's version
    # line) are sent to stderr for the duration of the run
    sys.stdout.flush()
    real_out = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_out, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
