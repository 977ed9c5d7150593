"""Alias: this module IS pagraph_b200.data (see compat/README.md)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("pagraph_b200.data")
