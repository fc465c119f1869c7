"""Importable alias of the package directory `chinesechess-alphazero_b200/` (not an identifier)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("chinesechess-alphazero_b200")
