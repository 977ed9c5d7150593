from . import graph_store, sampling  # noqa: F401
