"""CPU oracle for the JaxDEM per-timestep hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-numpy restatement (plus a C/OpenMP restatement under
``oracle/c``) of the reference algorithm behind ``System.step`` in
cdelv/JaxDEM.  Every function cites the reference ``file:line`` it follows
(paths relative to the reference checkout, e.g. ``jaxdem/system.py:60-82``).

It is NOT part of the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it,
and only as the checker / CPU baseline.  ``jaxdem_b200`` never imports it.

Parity status: PINNED on reference outputs (float64)
----------------------------------------------------
The reference is pure JAX and JAX is not installed in this image.  Its sources under
``/root/reference`` nevertheless RUN in the build container, unmodified and through their own
public API, on a numpy stand-in for the part of the JAX API they use
(``tests/golden/jaxshim``: identity ``jit``, Python-loop ``vmap`` / ``scan`` / ``while_loop``,
functional ``.at`` updates, JAX's index clamping, ``jax_enable_x64=True``).  Their outputs are
committed as ``tests/golden/ref_*.npz`` and ``tests/golden/extras/*.npz`` together with the
generating scripts (``make_reference_golden.py``, ``make_reference_extras.py``), and the oracle is
checked against them on CPU (``tests/test_host_cpu.py::test_oracle_matches_golden``,
``tests/test_reference_golden.py``): cell permutation, sorted hashes, stencil hashes, neighbour
lists and rebuild counts bit for bit; forces, torques, energies, multi-step trajectories, FIRE
iterates and stop iteration, rollout frames to a few ulp (measured 0 .. 2e-15 relative).  That
pins the ALGORITHM in float64 / int64.  Not pinned: XLA's own instruction selection (fusion, FMA
contraction, reduction trees) and the x64-disabled float32 mode — the oracle is dtype-generic, so
its float32 runs are the same statements in narrower arithmetic.

Next to that, the closed-form / scalar pins the reference's own tests hold for this path
(``tests/test_oracle_pins.py``):

* periodic min-image spring force ``-0.3071067811865475`` per component
  (``tests/test_clump_pair_friction.py:167-186``),
* overlap-0.2 force ``[-0.2, 0]`` (``:190-217``),
* bond exclusion F1 = 0, F0 = -F2 for naive and cell list
  (``tests/test_excluded_pairs.py:11-61``),
* analytic free asymmetric top, log10 rel err < -4 for ``spiral`` and
  ``verletspiral`` (``tests/test_rotation_integrators.py:24-97``),
* ``_pos_p_rot`` cache consistency (``tests/test_state_cache.py:65-92``),
* cell list == naive (``tests/test_colliders_invariance.py``),
* energy-drift slopes (``tests/test_energy_conservation.py:34-88``).
"""

from .state import OState, create_state, grid_state  # noqa: F401
from .materials import MaterialTable, make_material_table  # noqa: F401
from .system import OSystem, create_system, step, step_once  # noqa: F401
