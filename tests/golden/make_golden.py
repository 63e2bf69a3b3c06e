"""Regenerate tests/golden/*.npz:  python tests/golden/make_golden.py

What these fixtures are — and are not.  The six files this script writes (c2_like_f32 ... hertz_free_f32) are
ORACLE outputs (oracle/*.py, the line-by-line restatement of the reference path) on small seeded inputs, in the
dtypes of the BASELINE configurations (float32 included).  They are committed so that
  * the oracle itself cannot drift unnoticed (tests/test_host_cpu.py::test_oracle_matches_golden, CPU), and
  * the CUDA path is compared with stored numbers as well as with a live oracle run
    (tests/test_gpu_parity.py::test_cuda_matches_golden, GPU).
REFERENCE outputs live next to them as ref_*.npz and extras/*.npz (make_reference_golden.py /
make_reference_extras.py: the unmodified reference sources on a numpy stand-in for JAX, float64); both tests above
run over those files too, which is what pins the oracle and the CUDA path on the reference itself.
Every case stores its inputs and its outputs; integer results are compared bit for bit, floats to rel 1e-12
(f64) / 1e-5 (f32) of the field scale."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

CASES = [  # name, n, dim, domain, law, rot, clumps, dtype
    ("c2_like_f32", 400, 3, "periodic", "spring", "", False, np.float32),
    ("c3_like_f64", 300, 3, "periodic", "cundallstrack", "verletspiral", False, np.float64),
    ("c4_like_f64", 300, 2, "periodic", "spring", "verletspiral", False, np.float64),
    ("c5_like_f64", 300, 3, "periodic", "cundallstrack", "verletspiral", True, np.float64),
    ("readme_like_f64", 250, 3, "reflect", "spring", "verletspiral", False, np.float64),
    ("hertz_free_f32", 300, 2, "free", "hertz", "spiral", False, np.float32),
]
STEPS = 3


def run_case(name, n, dim, domain, law, rot, clumps, dtype):
    import oracle
    from oracle import colliders as ocol
    from helpers import build_oracle, make_inputs
    inp = make_inputs(n, dim, seed=sum(map(ord, name)), dtype=dtype, phi=0.55,
                      clumps=clumps, poly=1.3, nmat=2 if law != "spring" else 1)
    lin = "euler" if domain == "free" else "verlet"
    kw = dict(dtype=dtype, domain=domain, law=law, lin=lin, rot=rot, dt=1e-3, nmat=2 if law != "spring" else 1)
    ost, osy = build_oracle(inp, **kw)
    perm, sh, nh, ovf, _ = ocol.get_spatial_partition(ost.pos, osy, osy.collider.cell_size, osy.collider.neighbor_mask,
                                                       ost.idtype)
    if osy.domain.periodic:
        nh = ocol.dedup_stencil_hashes(nh)
    ocol.celllist_compute_force(ost, osy)
    out = dict(perm=perm, sorted_hash=sh, nbr_hash=nh, force0=ost.force.copy(), torque0=ost.torque.copy())
    out["energy0"] = np.asarray(ocol.celllist_compute_potential_energy(ost, osy))
    nl, nl_ovf = ocol.celllist_create_neighbor_list(ost, osy, 1.1, 24)
    out["nlist"], out["nlist_overflow"] = nl, np.asarray(nl_ovf)
    oracle.step(ost, osy, STEPS)
    for f in ("pos_c", "vel", "force", "torque", "ang_vel"):
        out[f"{f}_after"] = getattr(ost, f).copy()
    out["q_after"] = np.concatenate([ost.q_w, ost.q_xyz], axis=1)
    meta = dict(n=n, dim=dim, domain=domain, law=law, lin=lin, rot=rot, clumps=clumps, dtype=np.dtype(dtype).name,
                steps=STEPS, nmat=kw["nmat"])
    inputs = {f"in_{k}": np.asarray(v) for k, v in inp.items() if not isinstance(v, list)}
    return dict(**inputs, **out, meta=np.array(repr(meta)))


def main():
    for case in CASES:
        data = run_case(*case)
        np.savez_compressed(os.path.join(HERE, case[0] + ".npz"), **data)
        print(case[0], {k: v.shape for k, v in data.items() if k != "meta"})


if __name__ == "__main__":
    main()
