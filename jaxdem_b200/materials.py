"""Material / MaterialMatchmaker / MaterialTable mirror (reference
jaxdem/materials/elastic_mats.py, material_table.py:77-128,
material_matchmakers/harmonic.py:29-39, linear.py:28-32).  Host-side, built once;
the kernels read the resulting (M,) and (M, M) device tables."""

from __future__ import annotations

from dataclasses import dataclass, fields
from typing import Sequence

import torch

from .factory import Factory


@dataclass
class Material(Factory):
    density: float = 1.0


@Material.register("elastic")
@dataclass
class Elastic(Material):
    young: float = 1.0e4
    poisson: float = 0.3


@Material.register("elasticfrict")
@dataclass
class ElasticFriction(Material):
    young: float = 1.0e4
    poisson: float = 0.3
    mu: float = 0.0
    e: float = 1.0
    mu_r: float = 0.0


class MaterialMatchmaker(Factory):
    @staticmethod
    def get_effective_property(p1: torch.Tensor, p2: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError


@MaterialMatchmaker.register("harmonic")
class HarmonicMaterialMatchmaker(MaterialMatchmaker):
    @staticmethod
    def get_effective_property(p1, p2):
        is_zero = (p1 == 0.0) | (p2 == 0.0)
        s1 = torch.where(is_zero, torch.ones_like(p1), p1)
        s2 = torch.where(is_zero, torch.ones_like(p2), p2)
        return torch.where(is_zero, torch.zeros_like(s1 * s2), 2.0 * s1 * s2 / (s1 + s2))


@MaterialMatchmaker.register("linear")
class LinearMaterialMatchmaker(MaterialMatchmaker):
    @staticmethod
    def get_effective_property(p1, p2):
        return (p1 + p2) / 2


_ALL_PROPS = ("density", "young", "poisson", "mu", "e", "mu_r")


class MaterialTable:
    """props[k]: (M,) tensors; pair[k_eff]: (M, M) tensors (float64 on the host;
    cast/moved by System.create)."""

    def __init__(self, props: dict, pair: dict, matcher: MaterialMatchmaker):
        self.props, self.pair, self.matcher = props, pair, matcher

    @staticmethod
    def from_materials(mats: Sequence[Material], *, matcher: MaterialMatchmaker | None = None,
                       fill: float = 0.0) -> "MaterialTable":
        # keys come from the materials' own fields only (material_table.py:112): a law that needs a property no
        # material defines fails in System.create with the reference's KeyError instead of running on zeros
        keys = {f.name for m in mats for f in fields(m)}
        props = {k: torch.tensor([float(getattr(m, k, fill)) for m in mats], dtype=torch.float64)
                 for k in sorted(keys)}
        if matcher is None:
            matcher = MaterialMatchmaker.create("harmonic")
        pair = {f"{k}_eff": matcher.get_effective_property(a[:, None], a[None, :]) for k, a in props.items()}
        return MaterialTable(props, pair, matcher)

    def __getattr__(self, item):
        if item in ("props", "pair", "matcher") or item.startswith("__"):
            raise AttributeError(item)
        if item in self.props:
            return self.props[item]
        if item in self.pair:
            return self.pair[item]
        raise AttributeError(item)

    def __len__(self) -> int:
        return int(next(iter(self.props.values())).shape[-1])  # (M,) or batched (B, M)

    def to(self, device=None, dtype=None) -> "MaterialTable":
        f = lambda t: t.to(device=device, dtype=dtype).contiguous()
        return MaterialTable({k: f(v) for k, v in self.props.items()},
                             {k: f(v) for k, v in self.pair.items()}, self.matcher)
