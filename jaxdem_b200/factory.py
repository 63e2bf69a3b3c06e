"""Plugin registry mirroring jaxdem.Factory (reference jaxdem/factory.py:28-35,240-320,371-503):
``@Root.register(key)`` / ``Root.create(key, **kw)`` with per-root registries, key
normalisation, ``Create`` classmethod preferred over the constructor, unknown
keyword arguments dropped with a warning."""

from __future__ import annotations

import warnings
from inspect import signature
from typing import Any, Callable, ClassVar


def _normalize_key(key: str) -> str:
    return key.replace(" ", "").replace("_", "").replace("-", "").lower()


class Factory:
    _registry: ClassVar[dict]
    __registry_name__: ClassVar[str] = ""

    def __init_subclass__(cls, **kw: Any) -> None:
        super().__init_subclass__(**kw)
        if Factory in cls.__bases__:  # each direct child of Factory is a registry root
            cls._registry = {}

    @classmethod
    def register(cls, key: str | None = None) -> Callable[[type], type]:
        def decorator(sub: type) -> type:
            k = _normalize_key(sub.__name__ if key is None else key)
            if k in cls._registry and cls._registry[k] is not sub:
                raise ValueError(f"{cls.__name__}: key '{k}' already registered for {cls._registry[k].__name__}")
            cls._registry[k] = sub
            if "__registry_name__" not in sub.__dict__:  # first key wins; later ones are aliases
                sub.__registry_name__ = k
            return sub

        return decorator

    @property
    def type_name(self) -> str:
        return type(self).__registry_name__

    @classmethod
    def registered(cls) -> list[str]:
        return list(cls._registry)

    @classmethod
    def create(cls, key: str, **kw: Any):
        try:
            sub = cls._registry[_normalize_key(key)]
        except KeyError as err:
            raise KeyError(f"Unknown {cls.__name__} '{key}'. Available: {list(cls._registry)}") from err
        fn = getattr(sub, "Create", None) or sub
        sig = signature(fn)
        params = sig.parameters
        if not any(p.kind == p.VAR_KEYWORD for p in params.values()):
            dropped = [k for k in kw if k not in params]
            for k in dropped:
                kw.pop(k)
            if dropped:
                warnings.warn(f"{cls.__name__}.create('{key}'): ignoring unknown keyword(s) {dropped}. "
                              f"Expected signature: {sub.__name__}{sig}", stacklevel=2)
        try:
            sig.bind_partial(**kw)
        except TypeError as err:
            raise TypeError(f"Invalid keyword(s) for {sub.__name__}: {err}. Expected signature: "
                            f"{sub.__name__}{sig}") from None
        return fn(**kw)
