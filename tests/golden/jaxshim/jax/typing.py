from typing import Any
ArrayLike = Any
DTypeLike = Any
