"""CPU oracle for the JaxDEM per-timestep hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-numpy restatement (plus a C/OpenMP restatement under
``oracle/c``) of the reference algorithm behind ``System.step`` in
cdelv/JaxDEM.  Every function cites the reference ``file:line`` it follows
(paths relative to the reference checkout, e.g. ``jaxdem/system.py:60-82``).

It is NOT part of the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it,
and only as the checker / CPU baseline.  ``jaxdem_b200`` never imports it.

Parity status
-------------
The reference is pure JAX and JAX is not installed in this image, so the
reference cannot be executed here and it stores no golden arrays.  The oracle
is pinned against every closed-form / scalar pin the reference's own tests
hold for this path (``tests/test_oracle_pins.py``):

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

At the bit-exact level (cell permutation, neighbour lists) and at the
rel-1e-5 / 1e-12 per-step level the reference has no stored vectors:
**parity unpinned** at that strictness; the oracle itself defines it.
"""

from .state import OState, create_state, grid_state  # noqa: F401
from .materials import MaterialTable, make_material_table  # noqa: F401
from .system import OSystem, create_system, step, step_once  # noqa: F401
