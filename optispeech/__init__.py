"""Alias package: `optispeech.*` -> `optispeech_b200.*`.

The reference's Hydra configs and pickled checkpoints name classes by dotted paths such as
`optispeech.model.OptiSpeech` or `optispeech.model.generator.modules.ConvNeXtBackbone`.  Importing any
`optispeech.X` module returns the very same module object as `optispeech_b200.X` (no second copy of the classes),
so those paths keep resolving when this repository replaces the reference on `sys.path`.
"""
import importlib
import importlib.abc
import importlib.machinery
import sys

_REAL = "optispeech_b200"


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if not fullname.startswith(__name__ + "."):
            return None
        real = _REAL + fullname[len(__name__):]
        try:
            importlib.import_module(real)
        except ImportError:
            return None
        return importlib.machinery.ModuleSpec(fullname, self, is_package=hasattr(sys.modules[real], "__path__"))

    def create_module(self, spec):
        return sys.modules[_REAL + spec.name[len(__name__):]]

    def exec_module(self, module):
        pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())

from optispeech_b200 import values  # noqa: E402,F401
from optispeech_b200.values import InferenceInputs, InferenceOutputs  # noqa: E402,F401
