"""Run the reference's OWN tests for the step path on the numpy stand-in for JAX (build container only):

    python tests/golden/run_reference_tests.py [pytest args]

Evidence that the stand-in executes the reference faithfully enough to be a source of golden vectors: the test
files of /root/reference/tests that cover the path are collected from where they lie and run against
the reference package imported by _ref_import.py.  Nothing is written under /root/reference."""
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

# the files that finish in seconds on the stand-in; test_colliders_invariance.py (3-D fixture: minutes of Python-loop
# vmap), test_rotation_integrators.py (70 000 steps per case) and test_energy_conservation.py (needs --full) can be
# named on the command line (with a real JAX they run as usual)
FILES = ["test_clump_pair_friction.py", "test_excluded_pairs.py", "test_state_cache.py", "test_public_api.py"]


def main():
    from _ref_import import REFERENCE, import_reference
    import_reference()
    import pytest
    named = [a for a in sys.argv[1:] if a.endswith(".py")]
    args = [os.path.join(REFERENCE, "tests", f) for f in (named or FILES)]
    extra = [a for a in sys.argv[1:] if not a.endswith(".py")]
    return pytest.main(args + ["-p", "no:cacheprovider", "-q", "--rootdir", "/tmp", "-c", "/dev/null",
                               "--import-mode=importlib"] + extra)


if __name__ == "__main__":
    raise SystemExit(main())
