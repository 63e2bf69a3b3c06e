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

FILES = ["test_clump_pair_friction.py", "test_excluded_pairs.py", "test_colliders_invariance.py",
         "test_rotation_integrators.py", "test_state_cache.py", "test_energy_conservation.py", "test_public_api.py"]


def main():
    from _ref_import import REFERENCE, import_reference
    import_reference()
    import pytest
    args = [os.path.join(REFERENCE, "tests", f) for f in FILES]
    extra = sys.argv[1:]
    return pytest.main(args + ["-p", "no:cacheprovider", "-q", "--rootdir", "/tmp", "-c", "/dev/null",
                               "--import-mode=importlib"] + extra)


if __name__ == "__main__":
    raise SystemExit(main())
