"""The reference's OWN tests, unmodified, against this package under the reference's name (`import scnym`).

`__graft_entry__.build()` places the reference's test files next to its installed package in the git-ignored
`oracle/_ref/tests/` (they travel to the GPU box with the snapshot; no reference source is committed).  The files are
run in a subprocess with `tests/ref_suite_plugin.py`, which aliases `scnym` to `scnym_b200` and makes CUDA the
default device.  Selected: the tests SURVEY.md §4 found runnable offline (no scanpy datasets, no network):
tests/test_model.py, tests/test_da.py, tests/test_utils.py (minus the AnnData one) and the MixUp / sharpen /
weight-schedule tests of tests/test_mixmatch.py."""
import os
import subprocess
import sys
import xml.etree.ElementTree as ET

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "oracle", "_ref", "tests")

SELECTED = {
    "test_model.py": ["test_model_build", "test_model_dan_build", "test_model_save_load"],
    "test_da.py": ["test_dan_pseudoconf", "test_dan_train"],
    "test_utils.py": ["test_build_classification_matrix_dense", "test_build_classification_matrix_sparse"],
    "test_mixmatch.py": ["test_samplemixup", "test_sharpen_labels", "test_sharpen_labels_t0", "test_weight_schedule"],
}


def test_install_as_scnym_registers_the_reference_module_names():
    """CPU: the alias itself (no kernels run)."""
    import importlib
    import scnym_b200
    saved = {k: v for k, v in sys.modules.items() if k == "scnym" or k.startswith("scnym.")}
    try:
        me = scnym_b200.install_as_scnym()
        import scnym
        assert scnym is me
        for n in ("model", "dataprep", "losses", "trainer", "predict", "main", "api", "utils"):
            assert importlib.import_module("scnym." + n) is getattr(scnym_b200, n)
        assert scnym.model.CellTypeCLF is scnym_b200.model.CellTypeCLF
        assert callable(scnym.api.scnym_api) and hasattr(scnym.trainer, "MultiTaskTrainer") and hasattr(scnym.trainer, "DANLoss")
    finally:
        for k in [k for k in sys.modules if k == "scnym" or k.startswith("scnym.")]:
            del sys.modules[k]
        sys.modules.update(saved)


@pytest.mark.gpu
def test_reference_own_tests_pass_against_this_package(tmp_path):
    if not os.path.isdir(REF_TESTS):
        pytest.skip("oracle/_ref/tests not present (build() installs it where /root/reference exists)")
    ids = [f"{os.path.join(REF_TESTS, f)}::{t}" for f, ts in SELECTED.items() for t in ts]
    xml = str(tmp_path / "ref.xml")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "pytest", "-q", "-x", "-p", "tests.ref_suite_plugin", "-p", "no:cacheprovider", "--rootdir", REF_TESTS,
           "-c", os.devnull, f"--junitxml={xml}"] + ids
    r = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    results = {}
    if os.path.exists(xml):
        for tc in ET.parse(xml).getroot().iter("testcase"):
            bad = [c.tag for c in tc if c.tag in ("failure", "error", "skipped")]
            results[tc.get("name")] = bad[0] if bad else "passed"
    want = [t for ts in SELECTED.values() for t in ts]
    not_ok = {t: results.get(t, "not run") for t in want if results.get(t) != "passed"}
    assert not not_ok, f"reference tests not passing against scnym_b200: {not_ok}\n{r.stdout[-3000:]}\n{r.stderr[-2000:]}"
