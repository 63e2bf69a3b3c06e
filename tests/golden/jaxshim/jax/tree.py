import dataclasses
from ._core import tree_flatten as flatten, tree_unflatten as _unf, tree_leaves as leaves, tree_map as map, \
    tree_structure as structure


def unflatten(treedef, leaves):
    return _unf(treedef, leaves)


def static(*args, **kwargs):
    metadata = dict(kwargs.get("metadata", {}))
    metadata["static"] = True
    kwargs["metadata"] = metadata
    return dataclasses.field(*args, **kwargs)
