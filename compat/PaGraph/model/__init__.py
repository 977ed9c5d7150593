from . import gcn_nssc, graphsage_nssc  # noqa: F401
