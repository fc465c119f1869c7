"""Importable alias of the package directory `chinesechess-alphazero_b200/` (not an identifier).

`cczero_b200` and every `cczero_b200.<sub>` name resolve to the SAME module objects as `chinesechess-alphazero_b200[.<sub>]`
(a second copy of e.g. `lib` would carry a second set of ctypes struct classes that the first copy's prototypes reject)."""
import importlib
import importlib.abc
import importlib.util
import sys

_REAL = "chinesechess-alphazero_b200"
_ALIAS = __name__


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.startswith(_ALIAS + "."):
            return importlib.util.spec_from_loader(fullname, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len(_ALIAS):])

    def exec_module(self, module):
        pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
sys.modules[_ALIAS] = importlib.import_module(_REAL)
