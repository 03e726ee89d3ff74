"""CPU checks of the C-ABI boundary: the library loads, exports every symbol the header declares,
and the ctypes signatures in scnym_b200/_lib.py agree with include/scnym_b200.h (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scnym_b200.h")

CTYPE = {"int": "I", "int32_t": "I", "float": "F", "uint32_t": "U32", "int64_t": "L", "long long": "L"}


def _declarations():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"(?:const\s+char\s*\*|int)\s+(scnb_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        parsed = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    parsed.append("P")
                else:
                    ty = " ".join(a.split()[:-1]).replace("const ", "")
                    parsed.append(CTYPE[ty])
        decls[name] = parsed
    return decls


def test_header_declares_entry_points():
    d = _declarations()
    assert len(d) >= 25
    for must in ("scnb_csr_linear_fwd", "scnb_csr_linear_bwd_dw", "scnb_bn_act_fwd", "scnb_bn_act_bwd", "scnb_pseudolabel",
                 "scnb_mixmatch_plan", "scnb_soft_ce_fwd_bwd", "scnb_softmax_mse_fwd_bwd", "scnb_optim_step", "scnb_predict_head"):
        assert must in d


def test_library_exports_every_declared_symbol():
    from scnym_b200 import _lib
    h = _lib.lib()
    for name in _declarations():
        assert hasattr(h, name), f"{name} declared in the header but not exported"
    assert h.scnb_version() >= 100
    assert isinstance(h.scnb_last_error(), bytes)


def test_ctypes_signatures_match_header():
    from scnym_b200 import _lib
    code = {ctypes.c_void_p: "P", ctypes.c_int: "I", ctypes.c_longlong: "L", ctypes.c_float: "F", ctypes.c_uint32: "U32"}
    decls = _declarations()
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in decls, f"{name} bound in _lib.py but missing from the header"
        got = [code[a] for a in argtypes]
        assert got == decls[name], f"{name}: ctypes {got} != header {decls[name]}"
    compute = {n for n in decls if n not in ("scnb_last_error", "scnb_version", "scnb_device_arch")}
    assert compute == set(_lib.SIGNATURES), compute ^ set(_lib.SIGNATURES)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from scnym_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.ScnbError):
        _lib.lib()


def test_cpu_tensors_are_rejected():
    import torch
    from scnym_b200 import _lib
    from scnym_b200.model import CellTypeCLF
    m = CellTypeCLF(n_genes=20, n_cell_types=3, n_hidden=8, n_hidden_init=8)
    with pytest.raises(_lib.ScnbError):
        m(torch.zeros(4, 20))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "scnym_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "/root/reference" not in src, f
