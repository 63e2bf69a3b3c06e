from ._core import (register_dataclass, register_pytree_node, register_pytree_node_class, tree_flatten,
                    tree_unflatten, tree_leaves, tree_map, tree_structure)
import functools


class Partial(functools.partial):
    pass
