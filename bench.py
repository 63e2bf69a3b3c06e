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
particle-step), ``cpu_baseline``, ``kernels`` (per-kernel device time of one step) and
``parity`` (a small instance of the SAME configuration stepped on the GPU(s) and checked
against the CPU oracle before anything is timed; the run fails if it does not agree;
``parity.reference_golden``: the same CUDA path against stored outputs of the reference itself,
tests/golden/ref_*.npz — permutation / hashes bit for bit, forces and a 3-step trajectory).

Other workloads of BASELINE.json (``--config``; the default c2 is the headline):
    c3    config 3's physics: cundallstrack + velocity Verlet + verletspiral (352 B / particle-step);
          with --gpus N: slab-decomposed; ``--n-total 67108864`` = the 64 M system, strong scaling
    c5    config 5: polydisperse rigid clumps of 3-6 spheres, cundallstrack + verletspiral (436 B)
    c4    config 4: ensemble of B x 1000 bidisperse 2D discs, f64, FIRE minimiser iterations
          (192 B / particle-iteration); --gpus N shards the batch axis, no collective
    c2nl  config 2 through the Verlet NeighborList collider (amortised rebuilds)
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
B_ALG = {"c2": 168.0, "c2nl": 168.0, "c3": 352.0, "c5": 436.0, "c4": 192.0}  # SURVEY.md §8d, per config
METRIC = {"c2": "particle-steps/sec at 1M 3D spheres",
          "c2nl": "particle-steps/sec at 1M 3D spheres (Verlet NeighborList collider)",
          "c3": "particle-steps/sec, 3D spheres with tangential friction + rotation integrator",
          "c5": "particle-steps/sec, polydisperse rigid clumps (friction + torque accumulation)",
          "c4": "particle-iterations/sec, ensemble of 2D 1k-disc FIRE minimisations (f64)"}
C3_MATERIAL = dict(young=1.0e4, poisson=0.3, density=1.0, mu=0.5, e=0.8, mu_r=0.05)  # benchmarks/force_model.py:15-28
# algorithmic bytes per particle of each kernel family
B_ALG_KERNEL = {
    "k_pair_force": 28.0,   # R pos 12 + rad 4, W force 12
    # fused sphere driver: + R (vel, mass) 16 of the sorted shadow record, W vel 12 (the after-force kick)
    "k_pair_flat": 56.0,
    "k_pair_rows": 32.0,    # row kernel: R sorted (x, y, z, rad) 16, W sorted force sum 16 (k_after4 un-permutes and kicks)
    "k_after4": 60.0,       # R sorted force 16 + inv 4 + (vel, mass) 16, W force 12 + vel 12
    "k_hash4": 64.0,        # R pos 12 vel 12 force 12 rad 4 mass 4 (promised-away streams not read), W pos 12 + key 4 + rank 4
    "k_nl_force": 16.0 + 12.0,  # + the list rows actually read (data dependent)
    "k_hash": 105.0,        # R pos 12 vel 12 force 12 rad 4 mass 4 fixed 1 + external buffers 36 + pos_p_rot 12 + bond 4; W pos 12
    "k_finalize": 44.0,
    "k_scatter": 20.0,
}
# config 3 / 5 (cundallstrack + rotation): the pair kernel also reads (vel, mass) + ang_vel records and writes torque
B_ALG_KERNEL_C3 = {"k_pair_rows": 28.0 + 40.0, "k_pair_force": 28.0 + 40.0, "k_rotation": 72.0, "k_linear": 32.0}
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


def make_clump_workload(n_spheres, seed=1, phi=0.3, dtype=np.float32):
    """SURVEY.md §8d C5 (recipe of the reference's benchmarks/base.py:41-65): rigid clumps of 3, 4, 5, 6 spheres
    (cycling), clump centres uniform in a periodic cube at sphere-volume fraction ``phi``, member offsets
    |pos_p| ~ U(0.05, 0.45) * 0.2 in random directions, radii ~ U(0.06, 0.14) (r_max / r_min < 2.5); mass and
    inertia are clump totals replicated on every member (jaxdem/state.py:169-181)."""
    rng = np.random.default_rng(seed)
    sizes = np.resize(np.array([3, 4, 5, 6]), n_spheres // 3 + 4)
    ends = np.cumsum(sizes)
    nc = int(np.searchsorted(ends, n_spheres, side="left")) + 1
    sizes = sizes[:nc].copy()
    sizes[-1] -= ends[nc - 1] - n_spheres
    first = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    cid = np.repeat(first, sizes)                      # clump id = index of its first member
    rad = rng.uniform(0.06, 0.14, n_spheres)
    vol = 4.0 / 3.0 * np.pi * rad**3
    L = (vol.sum() / phi) ** (1.0 / 3.0)
    rep = lambda a: np.repeat(a, sizes, axis=0)
    com = rep(rng.uniform(0, L, (nc, 3)))
    d = rng.normal(size=(n_spheres, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pos_p = d * rng.uniform(0.05, 0.45, (n_spheres, 1)) * 0.2
    q = rng.normal(size=(nc, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    cmass = np.add.reduceat(vol, first)                # density 1
    cin = 0.4 * cmass[:, None] * (0.12**2) * rng.uniform(0.8, 1.2, (nc, 3))
    f = lambda a: a.astype(dtype)
    return dict(pos=f(com), pos_p=f(pos_p), rad=f(rad), mass=f(rep(cmass)), inertia=f(rep(cin)), q=f(rep(q)),
                clump_id=cid, vel=f(rep(rng.uniform(-1, 1, (nc, 3)))), ang_vel=f(rep(rng.uniform(-1, 1, (nc, 3)))),
                box=f(np.full(3, L)))


def make_ensemble_workload(batch, n=1000, seed0=0, phi=0.4):
    """SURVEY.md §8d C4 (examples/jam_spheres.py:43-88): ``batch`` independent systems of ``n`` bidisperse discs
    (radii 0.5 / 0.7, 50:50) at packing fraction 0.4 in a periodic square, positions U(0, L)^2 with seed = system
    index; float64."""
    rad = np.where(np.arange(n) < n // 2, 0.5, 0.7)
    L = float((np.sum(np.pi * rad**2) / phi) ** 0.5)
    pos = np.stack([np.random.default_rng(seed0 + b).uniform(0, L, (n, 2)) for b in range(batch)])
    return dict(pos=pos, rad=np.broadcast_to(rad, (batch, n)).copy(), mass=np.ones((batch, n)),
                box=np.full((batch, 2), L))


def build_gpu_config(jd, torch, cfg, wl, dev, n_check=None):
    """(State, System) of one ``--config`` on ``dev`` from a workload dict."""
    F = torch.float64 if cfg == "c4" else torch.float32
    if cfg == "c4":
        st = jd.State.create(wl["pos"], rad=wl["rad"], mass=wl["mass"], dtype=F, device=dev)
        probe = jd.State.create(wl["pos"][0], rad=wl["rad"][0], mass=wl["mass"][0], dtype=F, device=dev)
        mt = jd.MaterialTable.from_materials([jd.Material.create("elastic", young=1.0, poisson=0.5, density=1.0)])
        sy = jd.System.create(st.shape, dt=1e-2, minimizer=jd.minimizers.fire, minimizer_kw=dict(dt=1e-2),
                              collider_type="CellList", collider_kw=dict(state=probe), domain_type="periodic",
                              domain_kw=dict(box_size=wl["box"]), force_model_type="spring", mat_table=mt,
                              dtype=F, device=dev)
        return st, sy
    kw = {k: wl[k] for k in ("vel", "rad", "mass", "ang_vel", "pos_p", "inertia", "q", "clump_id") if k in wl}
    st = jd.State.create(wl["pos"], dtype=F, device=dev, **kw)
    mt, law, rot = None, "spring", ""
    if cfg in ("c3", "c5"):
        mt = jd.MaterialTable.from_materials([jd.Material.create("elasticfrict", **C3_MATERIAL)],
                                             matcher=jd.MaterialMatchmaker.create("harmonic"))
        law, rot = "cundallstrack", "verletspiral"
    col, ckw = "CellList", dict(state=st)
    if cfg == "c2nl":
        col, ckw = "NeighborList", dict(state=st, cutoff=1.0, skin=0.1, number_density=0.96)
    sy = jd.System.create(st.shape, dt=1e-3, linear_integrator_type="verlet", rotation_integrator_type=rot,
                          collider_type=col, collider_kw=ckw, domain_type="periodic",
                          domain_kw=dict(box_size=wl["box"]), force_model_type=law, mat_table=mt, dtype=F, device=dev)
    return st, sy


def build_oracle_config(cfg, wl):
    """The same configuration for the CPU oracle (numpy restatement; oracle/c where it covers the config)."""
    import oracle
    dtype = np.float64 if cfg == "c4" else np.float32
    kw = {k: wl[k] for k in ("vel", "rad", "mass", "ang_vel", "pos_p", "inertia", "q", "clump_id") if k in wl}
    st = oracle.create_state(wl["pos"], dtype=dtype, **kw)
    mt, law, rot = None, "spring", ""
    if cfg in ("c3", "c5"):
        mt = oracle.make_material_table([C3_MATERIAL], "harmonic")
        law, rot = "cundallstrack", "verletspiral"
    if cfg == "c4":
        mt = oracle.make_material_table([dict(young=1.0, poisson=0.5, density=1.0)], "harmonic")
    col, ckw = "celllist", dict(state=st)
    if cfg == "c2nl":
        col, ckw = "neighborlist", dict(cutoff=1.0, skin=0.1, number_density=0.96)
    sy = oracle.create_system(st, dt=1e-2 if cfg == "c4" else 1e-3, linear_integrator_type="verlet",
                              rotation_integrator_type=rot, collider_type=col, collider_kw=ckw,
                              domain_type="periodic", domain_kw=dict(box_size=wl["box"]), force_model_type=law,
                              mat_table=mt)
    return st, sy


def workload_for(cfg, n, seed, packing="grid", stack=1, batch=None):
    if cfg == "c5":
        return make_clump_workload(n, seed=seed)
    if cfg == "c4":
        return make_ensemble_workload(batch, n=1000, seed0=seed)
    wl = make_workload(n=n, seed=seed, packing=packing, stack=stack)
    if cfg == "c3":
        wl["ang_vel"] = np.random.default_rng(seed + 77).uniform(-1, 1, wl["pos"].shape).astype(np.float32)
    return wl


def _max_rel(got, want):
    want = np.asarray(want, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - want).max() / max(float(np.abs(want).max()), 1e-30))


def parity_check(jd, torch, cfg, dev, world=1, rank=0):
    """Before anything is timed: a small instance of the SAME configuration, stepped through the same public
    call the timed region uses, against the CPU oracle on rank 0 (oracle/c for config 2, the numpy restatement
    otherwise).  Bounds (field scale, a few steps without re-synchronisation): 1e-4 positions / velocities,
    1e-3 forces (f32); 1e-9 for the f64 ensemble.  Returns the dict that goes into the JSON line."""
    steps = 4
    if cfg == "c4":
        wl = make_ensemble_workload(2, n=1000, seed0=900)
        st, sy = build_gpu_config(jd, torch, cfg, wl, dev)
        jd.System.minimize(st, sy, max_steps=6, pe_tol=0.0, pe_diff_tol=0.0, force_tol=-1.0)
        from oracle import minimizers as omin
        errs = {}
        for b in range(2):
            one = dict(pos=wl["pos"][b], rad=wl["rad"][b], mass=wl["mass"][b], box=wl["box"][b])
            ost, osy = build_oracle_config(cfg, one)
            omin.minimize(ost, osy, omin.FireConfig(1e-2), max_steps=6, pe_tol=0.0, pe_diff_tol=0.0, force_tol=-1.0)
            for f in ("pos_c", "force"):
                errs[f] = max(errs.get(f, 0.0), _max_rel(getattr(st, f)[b].cpu().numpy(), getattr(ost, f)))
        ok = all(v <= 1e-9 for v in errs.values())
        return dict(parity_checked=bool(ok), against="oracle (numpy), 2 systems x 1000 discs, 6 FIRE iterations",
                    max_rel_err=errs, bound=1e-9)
    n = 16384 if cfg in ("c3", "c5", "c2nl") else 65536
    wl = workload_for(cfg, n, seed=4242)
    st, sy = build_gpu_config(jd, torch, cfg, wl, dev)
    jd.System.step(st, sy, n=steps)
    torch.cuda.synchronize()
    ost, osy = build_oracle_config(cfg, wl)
    against = "oracle (numpy)"
    if cfg == "c2":
        from oracle import c_oracle
        c_oracle.CStep(ost, osy).step(steps)
        against = "oracle/c (C restatement)"
    else:
        import oracle
        oracle.step(ost, osy, steps)
    errs = {f: _max_rel(getattr(st, f).cpu().numpy(), getattr(ost, f)) for f in ("pos_c", "vel", "force")}
    ok = errs["pos_c"] <= 1e-4 and errs["vel"] <= 1e-4 and errs["force"] <= 1e-3
    return dict(parity_checked=bool(ok), against=f"{against}, {n} particles, {steps} steps", max_rel_err=errs,
                bound={"pos_c": 1e-4, "vel": 1e-4, "force": 1e-3})


def reference_golden_check(jd, torch, cfg):
    """Next to the live oracle run: the CUDA path against STORED OUTPUTS OF THE REFERENCE ITSELF for this
    configuration family (tests/golden/ref_*.npz — the unmodified reference sources run on the numpy stand-in for
    JAX in the build container): cell permutation and sorted hashes bit for bit, contact forces and the
    3-step trajectory within the float32 bounds of the tests.  A mismatch fails the run; a missing fixture is
    reported as skipped."""
    name = {"c2": "ref_c2_like_f32", "c2nl": "ref_c2_like_f32", "c3": "ref_c3_like_f32", "c5": "ref_c5_like_f32"}.get(cfg)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", f"{name}.npz")
    if name is None or not os.path.exists(path):
        return dict(checked=False, skipped=f"no stored reference output for {cfg}")
    tdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests")
    if tdir not in sys.path:
        sys.path.insert(0, tdir)
    from helpers import build_gpu, load_golden
    inp, kw, want, meta = load_golden(path)
    st, sy = build_gpu(inp, **kw)
    perm, sh, _, _ = sy.collider.partition(st, sy)
    ints = bool(np.array_equal(perm.cpu().numpy(), want["perm"]) and np.array_equal(sh.cpu().numpy(), want["sorted_hash"]))
    sy.collider.compute_force(st, sy)
    errs = {"force0": _max_rel(st.force.cpu().numpy(), want["force0"])}
    jd.System.step(st, sy, n=meta["steps"])
    torch.cuda.synchronize()
    for f in ("pos_c", "vel", "force"):
        errs[f] = _max_rel(getattr(st, f).cpu().numpy(), want[f + "_after"])
    ok = ints and errs["pos_c"] <= 1e-4 and errs["vel"] <= 1e-4 and errs["force0"] <= 1e-4 and errs["force"] <= 1e-3
    return dict(checked=bool(ok), file=f"tests/golden/{name}.npz", permutation_and_hashes_bit_exact=ints,
                max_rel_err=errs, steps=int(meta["steps"]))


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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
            # nvidia-smi takes 0.1 - 1 s to attach to the driver, during which kernel launches of this process can
            # stall for milliseconds: wait for its first sample so the timed region only sees the steady 100 ms polls
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < 5.0 and self.proc.poll() is None:
                time.sleep(0.01)
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
    from oracle import c_oracle
    c_oracle.set_num_threads(threads or os.cpu_count() or 1)  # all host cores, whatever OMP_NUM_THREADS torchrun exported
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


def cpu_numpy_rate(cfg, seed=1):
    """Configs the C restatement does not cover (friction + rotation, clumps, neighbour list, minimiser): the numpy
    restatement, single thread, on a BOUNDED sample of the same workload.  -> (rate, seconds, cores, sample)."""
    import oracle
    if cfg == "c4":
        from oracle import minimizers as omin
        wl = make_ensemble_workload(2, n=1000, seed0=seed)
        iters, t_all = 10, 0.0
        for b in range(2):
            one = dict(pos=wl["pos"][b], rad=wl["rad"][b], mass=wl["mass"][b], box=wl["box"][b])
            ost, osy = build_oracle_config(cfg, one)
            t0 = time.perf_counter()
            omin.minimize(ost, osy, omin.FireConfig(1e-2), max_steps=iters, pe_tol=0.0, pe_diff_tol=0.0, force_tol=-1.0)
            t_all += time.perf_counter() - t0
        return 2 * 1000 * iters / t_all, t_all, 1, f"2 systems x 1000 discs x {iters} FIRE iterations, numpy restatement"
    n, steps = 32768, 2
    wl = workload_for(cfg, n, seed=seed)
    ost, osy = build_oracle_config(cfg, wl)
    oracle.step(ost, osy, 1)
    t0 = time.perf_counter()
    oracle.step(ost, osy, steps)
    secs = time.perf_counter() - t0
    return n * steps / secs, secs, 1, f"{steps} steps of a {n}-particle instance of the same recipe, numpy restatement"


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
    """--impl reference: the reference's own CPU path.  JAX is not installed in this image, so the reference itself
    cannot run; what is timed is the restatement of its algorithm on ALL host cores — oracle/c for
    config 2 (exactly `--steps` steps after `--warmup` warm-up steps of the same total particle count the CUDA arm
    steps at this --gpus), the numpy restatement on a bounded sample for the other configs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config
    n_total = total_particles(args)
    base = {"impl": "reference", "metric": METRIC[cfg], "n_gpus": args.gpus, "higher_is_better": True,
            "scaling": scaling_of(args), "vs_baseline": None, "dtype": "f64" if cfg == "c4" else "f32",
            "data": "synthetic", "config": workload_config(args, args.n_per_gpu), "gpu_launches": 0}
    unit = "particle-iterations/s" if cfg == "c4" else "particle-steps/s"
    if cfg == "c2":
        wl = make_workload(n=args.n_per_gpu, packing=args.packing, stack=args.gpus if args.mode == "slab" else 1)
        jr = try_jax_reference(wl, args.steps)
        if jr is not None:
            rate, secs, cores = jr
            kind, sample = "reference", f"{args.steps} steps, jaxdem System.step on JAX CPU"
        else:
            rate, secs, cores = cpu_steps_per_s(wl, args.steps, warmup=args.warmup)
            kind = "port"
            sample = (f"{args.steps} full steps of the {n_total}-sphere workload after {args.warmup} warm-up steps, "
                      "C/OpenMP restatement of the reference cell-list path (JAX is not installed in this image: "
                      "the reference itself cannot run)")
        steps, warm = args.steps, args.warmup
    else:
        rate, secs, cores, sample = cpu_numpy_rate(cfg)
        kind, steps, warm = "port", args.steps, args.warmup
        secs = secs  # bounded sample; ms_per_step below is scaled to the full workload
    ms_step = 1e3 * n_total / rate
    base.update({"value": rate, "unit": unit, "steps": steps, "warmup": warm, "ms_per_step": ms_step,
                 "cpu_baseline": {"value": rate, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
                 "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(base))


def total_particles(args):
    if args.config == "c4":
        return args.batch * 1000
    return args.n_per_gpu * (args.gpus if args.mode == "slab" or args.gpus == 1 else args.gpus)


def scaling_of(args):
    return "strong" if (args.n_total or args.config == "c4") else "weak"


def workload_config(args, n):
    cfg = args.config
    if cfg == "c4":
        par = "single GPU" if args.gpus == 1 else f"batch axis sharded over {args.gpus} GPUs, {args.batch // args.gpus} systems each, no collective"
        return {"workload": f"C4: ensemble of {args.batch} independent systems x 1000 bidisperse 2D discs (r 0.5 / 0.7), "
                            "periodic square phi=0.4, cell-list collider, spring k=1, FIRE minimiser dt=1e-2, f64; one step = "
                            "one minimiser iteration of every system (tolerances disabled inside the timed region)",
                "n_particles": args.batch * 1000, "l2": "working set (1 GB of State + workspace) exceeds L2; no flush",
                "parallelism": par}
    if args.gpus > 1 and args.mode == "slab":
        par = (f"ONE periodic system of {args.gpus} x {n} spheres (box L x L x {args.gpus}L), z-slab decomposition, "
               "one slab per GPU; halo + migration exchange every step by peer-memory stores over NVLink with "
               "neighbour-to-neighbour flags and device-side row counts, no host synchronisation inside a step "
               "(symmetric memory; NCCL only sets up the group, send/recv is the fallback transport)")
    elif args.gpus > 1:
        par = "1 system per GPU (replicas)"
    else:
        par = "single GPU"
    desc = {
        "c2": "monodisperse 3D spheres r=0.5, periodic cube phi=0.5, {p} packing, cell-list collider (27-cell stencil), "
              "spring contact young_eff=1e4, velocity Verlet, dt=1e-3",
        "c2nl": "monodisperse 3D spheres r=0.5, periodic cube phi=0.5, {p} packing, Verlet NeighborList collider (cutoff 1.0, "
                "skin 0.1, rebuilt through the cell list), spring contact young_eff=1e4, velocity Verlet, dt=1e-3",
        "c3": "monodisperse 3D spheres r=0.5, periodic cube phi=0.5, {p} packing, cell-list collider, cundallstrack "
              "(young 1e4, poisson 0.3, mu 0.5, e 0.8, mu_r 0.05), velocity Verlet + verletspiral, dt=1e-3",
        "c5": "rigid clumps of 3-6 spheres (radii U(0.06, 0.14)), periodic cube phi=0.3, cell-list collider, cundallstrack, "
              "velocity Verlet + verletspiral, clump force / torque reduction, dt=1e-3",
    }[cfg].format(p=args.packing)
    return {"workload": f"{cfg.upper()}: {n} {desc}", "n_particles": n,
            "l2": "flushed between timed steps (256 MiB write, then 256 MiB read to drain the dirty lines)",
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
    cfg = args.config

    if world > 1 and args.mode == "slab" and cfg in ("c2", "c3"):
        return run_cuda_slab(args, world, rank, local, dev)
    parity = parity_check(jd, torch, cfg, dev) if not args.no_parity else {"parity_checked": False, "skipped": True}
    if not args.no_parity and not parity["parity_checked"]:
        raise RuntimeError(f"bench.py: the CUDA path does not agree with the CPU oracle: {parity}")
    if not args.no_parity and rank == 0:
        parity["reference_golden"] = reference_golden_check(jd, torch, cfg)
        if world == 1 and not parity["reference_golden"]["checked"] and "skipped" not in parity["reference_golden"]:
            raise RuntimeError(f"bench.py: the CUDA path does not agree with the stored reference outputs: {parity}")
    if cfg == "c4":
        return run_cuda_ensemble(args, world, rank, local, dev, parity)
    wl = workload_for(cfg, args.n_per_gpu, seed=1 + rank, packing=args.packing)
    n = wl["pos"].shape[0]
    st, sy = build_gpu_config(jd, torch, cfg, wl, dev)
    b_alg = B_ALG[cfg]
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
        flush_l2()  # the warm-up steps run exactly what the timed steps run (first use of the flush kernels included)
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
    fields = ("pos_c", "vel", "force") if cfg in ("c2", "c2nl") else ("pos_c", "vel", "force", "ang_vel", "torque")
    qfields = () if cfg in ("c2", "c2nl") else ("w", "xyz")
    leaf = lambda k: getattr(st.q, k) if k in qfields else getattr(st, k)
    names = fields + qfields
    host_in = {k: leaf(k).detach().cpu().pin_memory() for k in names}
    host_out = {k: torch.empty_like(v).pin_memory() for k, v in host_in.items()}
    h2d = sum(v.numel() * v.element_size() for v in host_in.values())
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_step():
        for k in names:
            leaf(k).copy_(host_in[k], non_blocking=True)
        jd.System.step(st, sy, n=1)
        for k in names:
            host_out[k].copy_(leaf(k), non_blocking=True)
        torch.cuda.synchronize()
        for k in names:  # next step's input is this step's output (a rollout driven from the host)
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
    roof = dominant_roofline(kernels, n, peak, peak_src, cfg)
    step_gbs = b_alg * n * args.steps / (ms * 1e-3) / 1e9

    if rank == 0:
        line = {
            "metric": METRIC[cfg], "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, n),
            "launch": "cuda-graph replay" if args.graph else "stream launches",
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": h2d, "steps": e2e_steps},
            "gpu_launches": int(launches),
            "parity": parity,
            "roofline": roof,
            "step_roofline": {"bound": "hbm", "achieved": step_gbs, "peak": peak, "unit": "GB/s",
                              "frac": step_gbs / peak, "algorithmic_bytes_per_particle_step": b_alg,
                              "peak_source": peak_src},
            "l2_warm": {"value": n * args.steps / (ms_warm * 1e-3), "unit": "particle-steps/s",
                        "ms_per_step": ms_warm / args.steps,
                        "note": "same K steps back to back in one jdb200_system_step call, no L2 flush"},
            "kernels": kernels,
        }
        if cfg == "c2nl":
            line["neighbor_list"] = {"n_build_times": int(sy.collider.n_build_times), "max_neighbors": sy.collider.max_neighbors,
                                     "steps_taken": int(sy.step_count)}
        if world == 1 and not args.no_cpu:
            if cfg == "c2":
                cpu_steps = args.cpu_steps
                rate, secs, cores = cpu_steps_per_s(wl, cpu_steps)
                sample = (f"{cpu_steps} full steps of the same 1M-sphere workload ({secs:.1f} s), "
                          "C/OpenMP restatement (oracle/c)")
            else:
                rate, secs, cores, sample = cpu_numpy_rate(cfg)
            line["cpu_baseline"] = {"value": rate, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                                    "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def dominant_roofline(kernels, n, peak, peak_src, cfg):
    """`roofline` of the JSON line: the kernel with the largest share of the step, its algorithmic bytes per launch
    (null for kernels without a stated figure) over its mean launch time in this run."""
    if not kernels:
        return None
    dom = max(kernels, key=lambda k: kernels[k]["us_per_step"])
    bpp = B_ALG_KERNEL.get(dom) if cfg in ("c2", "c2nl") else B_ALG_KERNEL_C3.get(dom, B_ALG_KERNEL.get(dom))
    dur_s = kernels[dom]["us_per_launch"] * 1e-6
    ach = (bpp * n / dur_s / 1e9) if bpp else None
    roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": (ach / peak) if ach else None, "traffic": None,
            "algorithmic_bytes_per_launch": (bpp * n) if bpp else None,
            "us_per_launch": kernels[dom]["us_per_launch"], "peak_source": peak_src}
    tr = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the last ncu --set full
    if os.path.exists(tr) and cfg in ("c2", "c3"):
        try:
            roof["traffic"] = json.load(open(tr)).get(cfg, {}).get(dom)
        except (OSError, ValueError, AttributeError):
            pass
    return roof


def run_cuda_ensemble(args, world, rank, local, dev, parity):
    """Config 4: B independent 2D systems, FIRE minimiser iterations.  --gpus N shards the batch axis (no collective);
    a "step" is one minimiser iteration of every system, timed K at a time in one jdb200_minimize_fire call."""
    import torch
    import torch.distributed as dist

    import jaxdem_b200 as jd
    from jaxdem_b200 import _lib
    from jaxdem_b200.minimizers import minimize

    lib = _lib.lib()
    bl = args.batch // world
    wl = make_ensemble_workload(bl, n=1000, seed0=rank * bl)
    st, sy = build_gpu_config(jd, torch, "c4", wl, dev)
    n_units = bl * 1000
    off = dict(pe_tol=0.0, pe_diff_tol=0.0, force_tol=-1.0)  # every system iterates for the whole timed region

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    minimize(st, sy, max_steps=args.warmup, poll=args.warmup, **off)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.jdb200_launch_count()
    barrier()
    a.record()
    minimize(st, sy, max_steps=args.steps, poll=args.steps, **off)  # init evaluation + K iterations
    b.record()
    barrier()
    launches = lib.jdb200_launch_count() - l0
    ms = a.elapsed_time(b) * args.steps / (args.steps + 1)  # K iterations + the initial evaluation ~ K + 1 evaluations
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n_units * args.steps / (ms_max * 1e-3)
    # e2e: positions in from pinned host memory, minimise K iterations, positions + energies out
    host_in = st.pos_c.detach().cpu().pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    pe_out = torch.empty(bl, dtype=torch.float64).pin_memory()
    e2e_steps = max(3, min(args.steps, 20))

    def e2e_step():
        st.pos_c.copy_(host_in, non_blocking=True)
        _, _, _, pe = minimize(st, sy, max_steps=1, poll=1, **off)
        host_out.copy_(st.pos_c, non_blocking=True)
        pe_out.copy_(pe, non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n_units * e2e_steps / float(t.item())
    # per-kernel times
    _lib.kernel_timing(True)
    minimize(st, sy, max_steps=5, poll=5, **off)
    torch.cuda.synchronize()
    kt = _lib.kernel_timing_collect()
    _lib.kernel_timing(False)
    kernels = {k: {"us_per_launch": 1e3 * v[0] / v[1], "launches_per_step": v[1] / 6.0, "us_per_step": 1e3 * v[0] / 6.0}
               for k, v in kt.items()}
    tot = sum(v["us_per_step"] for v in kernels.values())
    for v in kernels.values():
        v["share"] = v["us_per_step"] / tot if tot else 0.0
    peak, peak_src = measured_peak()
    gbs = B_ALG["c4"] * world * n_units * args.steps / (ms_max * 1e-3) / 1e9
    if rank == 0:
        dom = max(kernels, key=lambda k: kernels[k]["us_per_step"]) if kernels else None
        line = {
            "metric": METRIC["c4"], "value": value, "unit": "particle-iterations/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args, n_units),
            "launch": "stream launches, one jdb200_minimize_fire call per K iterations", "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "particle-iterations/s", "h2d_bytes_per_step": host_in.numel() * 8,
                    "d2h_bytes_per_step": host_in.numel() * 8 + bl * 8, "steps": e2e_steps},
            "gpu_launches": int(launches), "parity": parity,
            "roofline": None if dom is None else {"kernel": dom, "bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s",
                                                  "frac": None, "traffic": None, "us_per_launch": kernels[dom]["us_per_launch"],
                                                  "peak_source": peak_src},
            "step_roofline": {"bound": "hbm", "achieved": gbs, "peak": peak * world, "unit": "GB/s", "frac": gbs / (peak * world),
                              "algorithmic_bytes_per_particle_step": B_ALG["c4"], "peak_source": peak_src},
            "kernels": kernels,
        }
        if world == 1 and not args.no_cpu:
            rate, secs, cores, sample = cpu_numpy_rate("c4")
            line["cpu_baseline"] = {"value": rate, "unit": "particle-iterations/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def slab_parity_check(jd, torch, dist, cfg, dev, world, rank):
    """Multi-GPU self-check (VERDICT r1 #1a): a small slab-decomposed system (16384 spheres per rank) is stepped
    by all ranks through the same exchange / kernels the timed region uses, gathered by global id, and compared ON
    RANK 0 WITH THE CPU ORACLE of the whole system (oracle/c for config 2, numpy restatement for config 3) —
    not with the single-GPU CUDA path."""
    from jaxdem_b200.slab import create_slab_system
    n, steps = 16384 * world, 4
    wl = workload_for(cfg, 16384, seed=4242, stack=world)
    if cfg == "c3":
        wl["ang_vel"] = np.random.default_rng(4319).uniform(-1, 1, wl["pos"].shape).astype(np.float32)
    mt, law, rot = None, "spring", ""
    arrays = dict(pos=wl["pos"], vel=wl["vel"], rad=wl["rad"], mass=wl["mass"])
    if cfg == "c3":
        mt = jd.MaterialTable.from_materials([jd.Material.create("elasticfrict", **C3_MATERIAL)],
                                             matcher=jd.MaterialMatchmaker.create("harmonic"))
        law, rot = "cundallstrack", "verletspiral"
        arrays["ang_vel"] = wl["ang_vel"]
    slab = create_slab_system(arrays, box_size=wl["box"], dt=1e-3, force_model_type=law, rotation_integrator_type=rot,
                              mat_table=mt, dtype=torch.float32, device=dev)
    slab.compute_force()
    slab.step(steps)
    torch.cuda.synchronize()
    res = slab.gather(("pos_c", "vel", "force"))
    out = None
    if rank == 0:
        import oracle
        ost, osy = build_oracle_config(cfg, wl)
        if cfg == "c2":
            from oracle import c_oracle
            c_oracle.set_num_threads(os.cpu_count() or 1)
            cs = c_oracle.CStep(ost, osy)
            cs.compute_force()
            cs.step(steps)
            against = "oracle/c (C restatement)"
        else:
            from oracle import colliders as ocol
            ocol.compute_force(ost, osy)
            oracle.step(ost, osy, steps)
            against = "oracle (numpy)"
        ok_ids = bool(np.array_equal(res["gid"], np.arange(n)))
        errs = {f: _max_rel(res[f], getattr(ost, f)) for f in ("pos_c", "vel", "force")}
        tight = {"pos_c": 1e-4, "vel": 1e-4, "force": 1e-3}
        # cundallstrack: the tangential direction vt / |vt| is discontinuous at vt = 0, so after a few f32 steps a
        # handful of sliding pairs may differ by O(mu Fn): every element within 10x the bound, and all but 1e-4
        # of the elements within the bound itself
        slack = 10.0 if cfg == "c3" else 1.0
        beyond = {f: float((np.abs(np.asarray(res[f], np.float64) - getattr(ost, f)) >
                            tight[f] * np.abs(getattr(ost, f)).max()).mean()) for f in tight}
        ok = ok_ids and all(errs[f] <= slack * tight[f] for f in tight) and all(v <= 1e-4 for v in beyond.values())
        out = dict(parity_checked=bool(ok), against=f"{against}, ONE system of {n} spheres over {world} slabs, {steps} steps, "
                   "gathered by global id", max_rel_err=errs, bound=tight, max_slack=slack, fraction_beyond_bound=beyond)
    box = [out]
    dist.broadcast_object_list(box, src=0)
    del slab
    torch.cuda.empty_cache()
    return box[0]


def run_cuda_slab(args, world, rank, local, dev):
    """N > 1 GPUs: ONE periodic system of N x n_per_gpu spheres, z-slab decomposition (jaxdem_b200/slab.py);
    every step exchanges halos and migrants over NVLink.  Per-GPU work is fixed (weak scaling) unless --n-total
    fixes the total (strong scaling: n_per_gpu = n_total / N)."""
    import torch
    import torch.distributed as dist

    import jaxdem_b200 as jd
    from jaxdem_b200 import _lib
    from jaxdem_b200.slab import create_slab_system

    lib = _lib.lib()
    cfg = args.config
    parity = ({"parity_checked": False, "skipped": True} if args.no_parity
              else slab_parity_check(jd, torch, dist, cfg, dev, world, rank))
    if not args.no_parity and not parity["parity_checked"]:
        raise RuntimeError(f"bench.py: the slab-decomposed CUDA path does not agree with the CPU oracle: {parity}")
    # every rank generates the cube of its own slab (seed 1 + rank) and shifts it to its place in
    # the stack; rows that fall into a neighbour's boundary layer migrate in the first exchange
    n = args.n_per_gpu
    wl = workload_for(cfg, n, seed=1 + rank, packing=args.packing)
    L = float(wl["box"][0])
    wl["pos"][:, 2] += np.float32(rank * L)
    box = np.array([L, L, L * world], dtype=np.float32)
    n_total = n * world
    b_alg = B_ALG[cfg]
    arrays = dict(pos=wl["pos"], vel=wl["vel"], rad=wl["rad"], mass=wl["mass"])
    mt, law, rot = None, "spring", ""
    if cfg == "c3":
        mt = jd.MaterialTable.from_materials([jd.Material.create("elasticfrict", **C3_MATERIAL)],
                                             matcher=jd.MaterialMatchmaker.create("harmonic"))
        law, rot = "cundallstrack", "verletspiral"
        arrays["ang_vel"] = wl["ang_vel"]
    slab = create_slab_system(arrays, box_size=box, dt=1e-3, force_model_type=law, rotation_integrator_type=rot,
                              mat_table=mt, dtype=torch.float32, device=dev, capacity_factor=1.35,
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
    launches_per_step = None
    if args.graph:
        c0 = lib.jdb200_launch_count()
        slab.step(2)
        launches_per_step = (lib.jdb200_launch_count() - c0) // 2  # graph replays do not pass through the counter
        slab.compile_step()  # the whole decomposed step replayed from a CUDA graph
    for _ in range(args.warmup):
        flush_l2()  # the warm-up steps run exactly what the timed steps run (first use of the flush kernels included)
        slab.step(1)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = lib.jdb200_launch_count()
    barrier()
    t_host0 = time.perf_counter()
    for a, b in ev:
        flush_l2()
        a.record()
        slab.step(1)
        b.record()
    t_host1 = time.perf_counter()
    barrier()
    t_host2 = time.perf_counter()
    launches = lib.jdb200_launch_count() - l0
    if launches_per_step is not None:
        launches = launches_per_step * args.steps
    slab.sync_counts()  # raises if any exchange of the timed region set a status bit (stray / capacity / timeout)
    per_step = [a.elapsed_time(b) for a, b in ev]
    ms = float(sum(per_step))
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    # per-rank view of the timed loop (diagnostic): median step, host time to enqueue the loop, wall time of the
    # loop incl. L2 flushes — ranks are coupled only through the neighbour flags, so a rank that starts its step
    # early waits for its neighbour INSIDE its own events
    top = sorted(per_step)[-3:] if len(per_step) >= 3 else [0.0, 0.0, 0.0]
    mine = torch.tensor([float(np.median(per_step)), 1e3 * (t_host1 - t_host0) / args.steps,
                         1e3 * (t_host2 - t_host0) / args.steps, float(np.mean(per_step)), *top,
                         float(int(np.argmax(per_step)))], dtype=torch.float64, device=dev)
    per_rank = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(per_rank, mine)
    per_rank = [[round(float(x), 4) for x in r.tolist()] for r in per_rank]
    value = n_total * args.steps / (ms_max * 1e-3)

    # ---- e2e: every rank's owned rows host -> device before the step, device -> host after it ----
    fields = ("pos_c", "vel", "force")
    host = {k: torch.empty((slab.cap, 3), dtype=torch.float32).pin_memory() for k in fields}
    e2e_steps = max(3, min(args.steps, 20))
    h2d = d2h = 0

    def e2e_step(first=False):
        nonlocal h2d, d2h
        slab.sync_counts()  # device protocol: the owned-row count lives on the device (after the previous step's sync)
        m = slab.n_own
        if not first:
            for k in fields:
                slab.buf[k][:m].copy_(host[k][:m], non_blocking=True)
        slab.step(1)
        slab.sync_counts()
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
    # per-kernel device time of rank 0's step (diagnostic events; separate pass, all ranks step together)
    _lib.kernel_timing(rank == 0)
    for _ in range(5):
        flush_l2()
        slab.step(1)
    torch.cuda.synchronize()
    kt = _lib.kernel_timing_collect() if rank == 0 else {}
    _lib.kernel_timing(False)
    kernels = {k: {"us_per_launch": 1e3 * v[0] / v[1], "launches_per_step": v[1] / 5.0, "us_per_step": 1e3 * v[0] / 5.0}
               for k, v in kt.items()}
    barrier()
    peak, peak_src = measured_peak()
    step_gbs = b_alg * n_total * args.steps / (ms_max * 1e-3) / 1e9
    own = torch.tensor([slab.n_own, slab.n_ghost], dtype=torch.int64, device=dev)
    owns = [torch.zeros_like(own) for _ in range(world)]
    dist.all_gather(owns, own)
    if rank == 0:
        line = {
            "metric": METRIC[cfg], "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": scaling_of(args), "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, n),
            "launch": ("CUDA-graph replay of the whole decomposed step (SlabSystem.compile_step); " if args.graph else "") +
                      ("stream launches, hook by hook; row counts, exchange flags and error bits stay on the device "
                       "(no host synchronisation inside a step)" if slab.device_protocol else
                       "stream launches, hook by hook; one host read of the exchange headers per step"),
            "n_particles_total": n_total,
            "owned_ghost_rows_per_rank": [[int(x) for x in o.tolist()] for o in owns],
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "particle-steps/s",
                    "h2d_bytes_per_step": int(t[1].item() / e2e_steps), "d2h_bytes_per_step": int(t[2].item() / e2e_steps),
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "parity": parity,
            "step_ms_rank0": {"min": min(per_step), "median": float(np.median(per_step)), "max": max(per_step)},
            "per_rank_ms": {"columns": ["median step (events)", "host enqueue per step", "wall per step incl. L2 flush",
                                        "mean step", "3rd slowest step", "2nd slowest", "slowest", "index of the slowest"],
                            "rows": per_rank},
            "roofline": dominant_roofline(kernels, n, peak, peak_src, cfg),
            "kernels_rank0": kernels,
            "step_roofline": {"bound": "hbm", "achieved": step_gbs, "peak": peak * world, "unit": "GB/s",
                              "frac": step_gbs / (peak * world), "algorithmic_bytes_per_particle_step": b_alg,
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
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5", "c2nl"],
                    help="workload (BASELINE.json configs): c2 = the headline (default), see the module docstring")
    ap.add_argument("--n-total", type=int, default=0,
                    help="fix the TOTAL particle count (strong scaling): n_per_gpu = n_total / gpus (67108864 = config 3's 64 M)")
    ap.add_argument("--batch", type=int, default=4096, help="config 4: number of independent systems (sharded over --gpus)")
    ap.add_argument("--no-parity", action="store_true", help="skip the pre-timing self-check against the CPU oracle")
    ap.add_argument("--graph", action="store_true", help="replay the step from a CUDA graph (System.compile_step)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-steps", type=int, default=10)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.n_total:
        args.n_per_gpu = args.n_total // max(args.gpus, 1)
    if args.config == "c5" and args.n_per_gpu == N_PARTICLES:
        args.n_per_gpu = 1 << 22  # config 5: 4 M spheres
    if args.config in ("c4", "c5", "c2nl") and args.gpus > 1:
        args.mode = "replicas"  # no slab path for ensembles / clumps / the neighbour list: independent shards
    sys.path.insert(0, os.path.join(ROOT, "tests"))
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
