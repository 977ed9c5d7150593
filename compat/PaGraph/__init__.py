"""`PaGraph` name shim over pagraph_b200 (see ../README.md)."""
from . import data, model, parallel, partition, storage  # noqa: F401
