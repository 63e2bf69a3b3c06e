"""The numpy stand-in for JAX (tests/golden/jaxshim) obeys the JAX rules the reference relies on.  Runs in a
subprocess so that the stand-in's ``jax`` never becomes importable inside the test process itself."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

SCRIPT = r'''
import sys, dataclasses
sys.dont_write_bytecode = True
sys.path.insert(0, sys.argv[1])
import numpy as np
import jax, jax.numpy as jnp
assert "numpy-stand-in" in jax.__version__

# immutable arrays, functional updates, JAX's out-of-range rules
x = jnp.arange(5.0)
y = x.at[jnp.asarray([1, 7])].set(9.0)            # scatter: the out-of-range update is dropped
assert list(y) == [0, 9, 2, 3, 4] and list(x) == [0, 1, 2, 3, 4]
assert float(x[jnp.asarray(9)]) == 4.0             # gather: clamped
z = x
z += 1.0                                           # rebinds, never mutates the alias
assert list(x) == [0, 1, 2, 3, 4] and list(z) == [1, 2, 3, 4, 5]
m = jnp.zeros((2, 3)).at[jnp.asarray([[0], [1]]), jnp.asarray([[0, 3], [2, 5]])].set(1.0)
assert m.tolist() == [[1, 0, 0], [0, 0, 1]]
try:
    x[0] = 1.0
    raise SystemExit("in-place assignment must fail")
except TypeError:
    pass
a, b = jnp.asarray([[1, 2], [3, 4]])               # iteration by shape (the clamping gather never raises)
assert list(a) == [1, 2] and list(b) == [3, 4]

# pytrees with static fields, vmap with in_axes, fresh containers at jit / scan boundaries
@jax.tree_util.register_dataclass
@dataclasses.dataclass
class S:
    p: jax.Array
    k: int = jax.tree.static(default=3)

@jax.jit
def bump(s):
    s.p = s.p + s.k                                # the reference's hooks assign to fields of their argument
    return s

s0 = S(jnp.zeros(2))
s1 = bump(s0)
assert list(s0.p) == [0, 0] and list(s1.p) == [3, 3] and s1.k == 3
out = jax.vmap(lambda s, w: s.p.sum() * w, in_axes=(0, None))(S(jnp.ones((4, 2))), jnp.asarray(2.0))
assert out.shape == (4,) and list(out) == [4, 4, 4, 4]
carry, ys = jax.lax.scan(lambda c, _: (bump(c), c), s0, None, length=3)
assert ys.p.tolist() == [[0, 0], [3, 3], [6, 6]] and list(carry.p) == [9, 9]   # saved frames do not alias the carry
k, acc = jax.lax.while_loop(lambda v: v[0] < 4, lambda v: (v[0] + 1, v[1] + v[0]), (jnp.asarray(0), jnp.asarray(0)))
assert int(k) == 4 and int(acc) == 6
h, i = jax.lax.sort([jnp.asarray([3, 1, 3, 1]), jnp.asarray([0, 1, 2, 3])], num_keys=1)
assert list(h) == [1, 1, 3, 3] and list(i) == [1, 3, 0, 2]                     # stable
assert list(jax.ops.segment_sum(jnp.asarray([1.0, 2.0, 4.0]), jnp.asarray([1, 1, 5]), num_segments=3)) == [0, 3, 0]
assert list(jnp.bincount(jnp.asarray([0, 2, 2, 9]), length=3)) == [1, 0, 2]

# x64 disabled (JAX's default): nothing 64-bit survives an operation
jax.config.update("jax_enable_x64", False)
f = jnp.asarray([1.5, 2.5])
assert f.dtype == np.float32 and jnp.arange(3).dtype == np.int32 and jnp.zeros(2, dtype=float).dtype == np.float32
assert (f * jnp.arange(2)).dtype == np.float32      # numpy would promote int32 * float32 to float64
assert jnp.floor(f).astype(int).dtype == np.int32 and jnp.sum(jnp.arange(4)).dtype == np.int32
assert (f * 2.0).dtype == np.float32 and jnp.searchsorted(f, 2.0).dtype == np.int32
jax.config.update("jax_enable_x64", True)
assert jnp.asarray([1.5]).dtype == np.float64 and jnp.arange(3).dtype == np.int64
print("stand-in ok")
'''


def test_standin_obeys_the_jax_rules_the_reference_relies_on():
    shim = os.path.join(HERE, "golden", "jaxshim")
    r = subprocess.run([sys.executable, "-c", SCRIPT, shim], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "stand-in ok" in r.stdout, r.stdout + r.stderr
    assert "jax" not in sys.modules or "numpy-stand-in" not in getattr(sys.modules["jax"], "__version__", "")
