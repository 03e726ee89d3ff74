"""Import shim for the UNMODIFIED reference package at /root/reference.

TEST INFRASTRUCTURE ONLY.  Used in the build container (where /root/reference
exists) by `tests/golden/make_golden.py` and by the optional live-reference
tests.  Nothing on the product path, and nothing that runs on the GPU box,
may import this module.

The reference's `scnym/__init__.py` imports `anndata`, `scanpy` and
`tensorboardX`, none of which are installed.  The shim
registers the reference directory as a namespace package under the private
name `scnym_ref` so that it can coexist with `scnym_b200` in one process,
stubbing the three missing third-party modules.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
INSTALLED_ROOT = os.path.join(HERE, "_ref")      # `pip install --target oracle/_ref` of the unmodified reference (git-ignored)


def _pick_root():
    env = os.environ.get("SCNYM_REFERENCE_ROOT")
    if env:
        return env
    # the source tree exists in the build container only; the installed copy travels to the GPU box
    if os.path.isdir("/root/reference/scnym"):
        return "/root/reference"
    return INSTALLED_ROOT


REFERENCE_ROOT = _pick_root()


def install_reference(src="/root/reference", force=False):
    """`pip install --no-deps --target oracle/_ref` of the UNMODIFIED reference (run by `__graft_entry__.build()` in the build
    container).  /root/reference is read-only and the build writes an egg-info into the source tree, so the install runs from
    a scratch copy.  Returns True when oracle/_ref holds the package."""
    import shutil
    import subprocess
    import tempfile
    if os.path.isdir(os.path.join(INSTALLED_ROOT, "scnym")) and not force:
        return True
    if not os.path.isdir(os.path.join(src, "scnym")):
        return False
    tmp = tempfile.mkdtemp(prefix="scnym_ref_src_")
    try:
        work = os.path.join(tmp, "reference")
        shutil.copytree(src, work, ignore=shutil.ignore_patterns(".git", "notebooks", "assets"))
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", INSTALLED_ROOT, "--upgrade", work]
        rc = subprocess.call(cmd)
        return rc == 0 and os.path.isdir(os.path.join(INSTALLED_ROOT, "scnym"))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scnym"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference(modules=("model", "dataprep", "losses", "utils", "trainer", "predict", "main")):
    """Return the reference package (modules loaded unmodified from disk)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if "scnym" in sys.modules and getattr(sys.modules["scnym"], "__scnym_ref__", False):
        pkg = sys.modules["scnym"]
        for n in modules:                       # an earlier call may have loaded a subset only
            if not hasattr(pkg, n):
                setattr(pkg, n, importlib.import_module("scnym." + n))
        return pkg
    try:
        import anndata  # noqa: F401
    except Exception:
        _stub("anndata", AnnData=type("AnnData", (), {}))
    try:
        import scanpy  # noqa: F401
    except Exception:
        _stub("scanpy")
    try:
        import tensorboardX  # noqa: F401
    except Exception:
        _stub("tensorboardX", SummaryWriter=type("SummaryWriter", (), {}))
    try:
        import configargparse  # noqa: F401
    except Exception:
        _stub("configargparse")
    # the reference uses absolute-relative imports (`from .model import ...`)
    # so it must be registered under its own package name `scnym`.
    pkg = types.ModuleType("scnym")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "scnym")]
    pkg.__scnym_ref__ = True
    sys.modules["scnym"] = pkg
    for n in modules:
        setattr(pkg, n, importlib.import_module("scnym." + n))
    return pkg
