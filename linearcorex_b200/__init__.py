"""linearcorex_b200: B200-native Linear CorEx fit loop behind the `linearcorex.Corex` API."""
from .corex import Corex  # noqa: F401
from .hierarchy import fit_layers, transform_layers  # noqa: F401
from .sharding import Reducer, shard_rows  # noqa: F401

__all__ = ["Corex", "Reducer", "shard_rows", "fit_layers", "transform_layers"]
