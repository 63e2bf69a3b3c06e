from . import special, spatial  # noqa: F401
