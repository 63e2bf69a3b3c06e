"""Regenerate tests/golden/ref_*.npz:  python tests/golden/make_reference_golden.py   (build container only)

REFERENCE OUTPUTS.  Unlike make_golden.py (oracle outputs), every array written here is produced by the
UNMODIFIED reference sources under /root/reference — ``jaxdem.State.create``, ``jaxdem.System.create``,
``colliders/_partition.py:_get_spatial_partition``, ``DynamicCellList.compute_force`` /
``compute_potential_energy`` / ``create_neighbor_list``, ``System.step`` — imported and run through their own
public API.  JAX itself is not installed in the image, so the sources execute on ``tests/golden/jaxshim``: a
numpy stand-in for the part of the JAX API the step path uses (identity ``jit``, Python-loop ``vmap`` /
``while_loop`` / ``scan``, functional ``.at`` updates, JAX's index clamping), emulating
``jax_enable_x64=True``.  What that does and does not pin:
  * the ALGORITHM, operation for operation, in float64 and int64 — the same Python statements the reference
    executes under XLA, evaluated with IEEE arithmetic by numpy;
  * not XLA's instruction selection (fusion / FMA contraction / reduction trees), which real JAX does not fix
    across backends either, and not the x64-disabled (float32) mode — f32 goldens stay oracle outputs.
The file layout equals make_golden.py's, so tests/test_host_cpu.py::test_oracle_matches_golden (CPU) pins the
oracle on these arrays and tests/test_gpu_parity.py::test_cuda_matches_golden (GPU) compares the CUDA path with
them; /root/reference is not needed to run either."""
import os
import sys
import time

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

CASES = [  # name (its suffix is the dtype), n, dim, domain, law, rot, clumps
    ("ref_c2_like_f32", 160, 3, "periodic", "spring", "", False),
    ("ref_c3_like_f32", 140, 3, "periodic", "cundallstrack", "verletspiral", False),
    ("ref_c5_like_f32", 140, 3, "periodic", "cundallstrack", "verletspiral", True),
    ("ref_hertz_free_f32", 140, 2, "free", "hertz", "spiral", False),
    ("ref_c2_like_f64", 160, 3, "periodic", "spring", "", False),
    ("ref_c3_like_f64", 140, 3, "periodic", "cundallstrack", "verletspiral", False),
    ("ref_c4_like_f64", 160, 2, "periodic", "spring", "verletspiral", False),
    ("ref_c5_like_f64", 140, 3, "periodic", "cundallstrack", "verletspiral", True),
    ("ref_readme_like_f64", 120, 3, "reflect", "spring", "verletspiral", False),
    ("ref_hertz_free_f64", 140, 2, "free", "hertz", "spiral", False),
    ("ref_spiral3d_free_f64", 120, 3, "free", "cundallstrack", "spiral", True),  # 3-D spiral + Euler on clumps, free box
]
STEPS = 3


def build_reference(jd, jnp, inp, *, domain, law, lin, rot, dt, nmat, collider="celllist", collider_kw=None,
                    gravity=None, **_):
    from helpers import MATS
    mats = [jd.Material.create("elasticfrict", **m) for m in MATS[:nmat]]
    mt = jd.MaterialTable.from_materials(mats, matcher=jd.MaterialMatchmaker.create("harmonic"))
    kw = {k: jnp.asarray(inp[k]) for k in ("vel", "ang_vel", "rad", "mass", "clump_id", "pos_p", "q", "inertia",
                                           "mat_id", "fixed") if k in inp}
    st = jd.State.create(jnp.asarray(inp["pos"]), **kw)
    ckw = dict(collider_kw or {})
    if collider in ("celllist", "neighborlist", "multicelllist"):
        ckw["state"] = st
    dkw = dict(box_size=jnp.asarray(inp["box"]))
    sy = jd.System.create(st.shape, dt=dt, linear_integrator_type=lin, rotation_integrator_type=rot,
                          collider_type=collider, collider_kw=ckw, domain_type=domain, domain_kw=dkw,
                          force_model_type=law, mat_table=mt,
                          force_manager_kw=None if gravity is None else dict(gravity=gravity))
    return st, sy


def run_case(jd, jax, jnp, name, n, dim, domain, law, rot, clumps):
    from helpers import make_inputs
    from jaxdem.colliders.cell_list import _get_spatial_partition
    from jaxdem.colliders.cell_list import _dedup_stencil_hashes
    dtype = np.float32 if name.endswith("_f32") else np.float64
    jax.config.update("jax_enable_x64", dtype == np.float64)  # float32 = JAX's default mode: no 64-bit arrays at all
    nmat = 2 if law != "spring" else 1
    inp = make_inputs(n, dim, seed=sum(map(ord, name)), dtype=dtype, phi=0.55, clumps=clumps, poly=1.3, nmat=nmat)
    lin = "euler" if domain == "free" else "verlet"
    kw = dict(domain=domain, law=law, lin=lin, rot=rot, dt=1e-3, nmat=nmat)
    st, sy = build_reference(jd, jnp, inp, **kw)
    col = sy.collider
    iota = jax.lax.iota(dtype=int, size=st.N)
    perm, sh, nh, ovf = _get_spatial_partition(st.pos, sy, col.cell_size, col.neighbor_mask, iota)
    if sy.domain.periodic:
        nh = jax.vmap(_dedup_stencil_hashes)(nh)
    st, sy = col.compute_force(st, sy)
    out = dict(perm=perm, sorted_hash=sh, nbr_hash=nh, force0=st.force, torque0=st.torque)
    out["energy0"] = col.compute_potential_energy(st, sy)[2]
    res = col.create_neighbor_list(st, sy, 1.1, 24)
    out["nlist"], out["nlist_overflow"] = res[2], res[3]
    st, sy = jd.System.step(st, sy, n=STEPS)
    for f in ("pos_c", "vel", "force", "torque", "ang_vel"):
        out[f"{f}_after"] = getattr(st, f)
    out["q_after"] = np.concatenate([np.asarray(st.q.w), np.asarray(st.q.xyz)], axis=1)
    out = {k: np.array(np.asarray(v)) for k, v in out.items()}
    assert all(v.dtype in (np.dtype(dtype), np.dtype(np.int32 if dtype == np.float32 else np.int64), np.dtype(bool))
               for v in out.values()), {k: v.dtype for k, v in out.items()}
    meta = dict(n=n, dim=dim, domain=domain, law=law, lin=lin, rot=rot, clumps=clumps, dtype=np.dtype(dtype).name,
                steps=STEPS, nmat=nmat, source="reference sources on tests/golden/jaxshim")
    inputs = {f"in_{k}": np.asarray(v) for k, v in inp.items() if not isinstance(v, list)}
    return dict(**inputs, **out, meta=np.array(repr(meta)))


def main():
    from _ref_import import import_reference
    jd = import_reference()
    import jax
    import jax.numpy as jnp
    only = sys.argv[1:]
    for case in CASES:
        if only and case[0] not in only:
            continue
        t = time.time()
        data = run_case(jd, jax, jnp, *case)
        np.savez_compressed(os.path.join(HERE, case[0] + ".npz"), **data)
        print(case[0], f"{time.time() - t:.1f} s", {k: v.shape for k, v in data.items() if k != "meta"}, flush=True)


if __name__ == "__main__":
    main()
