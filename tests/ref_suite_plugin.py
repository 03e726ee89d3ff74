"""pytest plugin that runs the REFERENCE's own test files (installed, unmodified, under the git-ignored
`oracle/_ref/tests/` by `__graft_entry__.build()`) against THIS package on a GPU:

  * `import scnym` resolves to `scnym_b200` (`install_as_scnym()`): the drop-in name;
  * the reference's tests build CPU tensors and CPU models; this package has no CPU path, so the default torch
    device is set to CUDA for the run -- `torch.randn(...)`, `torch.zeros(...)` and freshly constructed modules then
    live on the GPU, and the test bodies run as written;
  * `anndata` / `scanpy` (not installed, no network) are stubbed for the module-level imports; tests that really
    need them are deselected by the caller.

Loaded with `-p tests.ref_suite_plugin` by tests/test_reference_suite.py (test infrastructure only)."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _stub(name, **attrs):
    try:
        __import__(name)
        return
    except Exception:
        pass
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


def pytest_configure(config):
    import torch
    import scnym_b200
    _stub("anndata", AnnData=type("AnnData", (), {}))
    _stub("scanpy")
    scnym_b200.install_as_scnym()
    if torch.cuda.is_available():
        torch.set_default_device("cuda")
