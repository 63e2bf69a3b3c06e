"""jaxdem_b200 — B200-native DEM step engine behind JaxDEM's plugin surface.

Only the per-timestep hot path (System.step / trajectory_rollout) is implemented:
cell-list collider, spring / hertz / cundallstrack contact laws, force manager,
velocity-Verlet / Euler / SPIRAL integrators, free / periodic / reflect domains.
All compute runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/jaxdem_b200.h``; there is no CPU fallback.
"""

from .components import (Collider, CundallStrackForce, DirectEuler, Domain, DynamicCellList, DynamicMultiCellList, ForceManager,
                         ForceModel, FreeDomain, HertzianForce, Integrator, LinearIntegrator, NaiveSimulator, NeighborList,
                         PeriodicDomain, ReflectDomain, RotationIntegrator, Spiral, SpringForce,
                         VelocityVerlet, VelocityVerletSpiral)
from .factory import Factory
from .materials import Material, MaterialMatchmaker, MaterialTable
from .state import Quaternion, State, set_default_dtype
from .system import System
from . import minimizers, utils

__all__ = [
    "Collider", "CundallStrackForce", "DirectEuler", "Domain", "DynamicCellList", "DynamicMultiCellList", "Factory", "ForceManager",
    "ForceModel", "FreeDomain", "HertzianForce", "Integrator", "LinearIntegrator", "Material",
    "MaterialMatchmaker", "MaterialTable", "NaiveSimulator", "NeighborList", "PeriodicDomain", "Quaternion", "ReflectDomain",
    "RotationIntegrator", "Spiral", "SpringForce", "State", "System", "VelocityVerlet",
    "VelocityVerletSpiral", "set_default_dtype", "minimizers", "utils",
]
