"""scnym_b200 -- B200-native implementation of the scNym training / inference hot path.

Drop-in surface (same module and attribute names as the reference package `scnym`):
`scnym_b200.api.scnym_api`, `scnym_b200.trainer.{Trainer,SemiSupervisedTrainer,MultiTaskTrainer}`,
`scnym_b200.model.{CellTypeCLF,DANN,GradReverse}`, `scnym_b200.losses.*`,
`scnym_b200.predict.Predicter`.  `install_as_scnym()` registers the package under the name
`scnym` so that `import scnym` resolves to it.

All arithmetic runs in hand-written sm_100a CUDA behind the C ABI of `include/scnym_b200.h`
(`scnym_b200/libscnym_b200.so`); there is no CPU or eager-PyTorch fallback.
"""
import importlib
import sys

__version__ = "0.1.0"

_SUBMODULES = ("model", "dataprep", "losses", "trainer", "predict", "main", "api", "utils", "engine", "functional")


def __getattr__(name):
    if name in _SUBMODULES:
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)


def install_as_scnym():
    """Make `import scnym` (and `scnym.api`, `scnym.model`, ...) resolve to this package."""
    me = sys.modules[__name__]
    sys.modules["scnym"] = me
    for n in _SUBMODULES:
        try:
            sys.modules["scnym." + n] = importlib.import_module("." + n, __name__)
        except ImportError:
            pass
    return me
