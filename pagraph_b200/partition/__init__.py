"""Offline partitioners producing the `{P}naive/` files the trainers load.
Reference: PaGraph/partition/{hash,dg,utils}.py."""
from .utils import get_sub_graph  # noqa: F401
