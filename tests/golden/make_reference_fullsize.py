"""Regenerate tests/golden/ref_fullsize_c2.json:  python tests/golden/make_reference_fullsize.py   (build container)

FULL-SIZE reference output for BASELINE config 2 (bench.py's 2**20-sphere workload, float32): the reference's own
``State.create`` / ``System.create`` / ``_get_spatial_partition`` (colliders/cell_list.py:37-88) run on the numpy
stand-in for JAX in its x64-disabled mode (these functions are array code, no per-particle Python loop, so a
million particles take seconds).  The permutation, the sorted hashes and the de-duplicated stencil hashes are too
large to commit (4 + 4 + 108 MB), so their SHA-256 digests are — "a checksum of checksums": the tests hash what the
C oracle (CPU) and the CUDA path (GPU) produce on the same seeded workload and compare digests."""
import hashlib
import json
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def digest(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def main():
    from _ref_import import import_reference
    jd = import_reference()
    import jax
    import jax.numpy as jnp
    import bench
    from jaxdem.colliders.cell_list import _dedup_stencil_hashes, _get_spatial_partition
    jax.config.update("jax_enable_x64", False)
    wl = bench.make_workload()
    st = jd.State.create(jnp.asarray(wl["pos"]), vel=jnp.asarray(wl["vel"]), rad=jnp.asarray(wl["rad"]),
                         mass=jnp.asarray(wl["mass"]))
    sy = jd.System.create(st.shape, dt=1e-3, linear_integrator_type="verlet", rotation_integrator_type="",
                          collider_type="celllist", collider_kw=dict(state=st), domain_type="periodic",
                          domain_kw=dict(box_size=jnp.asarray(wl["box"])), force_model_type="spring")
    col = sy.collider
    iota = jax.lax.iota(dtype=int, size=st.N)
    perm, sh, nh, ovf = _get_spatial_partition(st.pos, sy, col.cell_size, col.neighbor_mask, iota)
    perm, sh, nh = (np.asarray(x) for x in (perm, sh, nh))
    assert perm.dtype == np.int32 and sh.dtype == np.int32 and nh.dtype == np.int32 and st.pos.dtype == np.float32
    # _dedup_stencil_hashes is vmapped per particle in the reference; it only acts on tiny grids (a stencil that
    # wraps onto itself) — at this size every row is duplicate-free, which is asserted instead of looping
    srt = np.sort(nh, axis=1)
    assert not (srt[:, 1:] == srt[:, :-1]).any()
    out = dict(n=int(st.N), dtype="float32", cell_size=float(np.asarray(col.cell_size)),
               perm_sha256=digest(perm), sorted_hash_sha256=digest(sh), nbr_hash_sha256=digest(nh),
               perm_head=[int(v) for v in perm[:8]], sorted_hash_tail=[int(v) for v in sh[-8:]],
               hash_overflow=bool(np.asarray(ovf)),
               source="reference sources on tests/golden/jaxshim (x64 disabled), bench.make_workload()")
    with open(os.path.join(HERE, "ref_fullsize_c2.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out)


if __name__ == "__main__":
    main()
