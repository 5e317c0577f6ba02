"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads without a GPU and exports every symbol
include/dmm_b200.h declares; the ctypes table mirrors the header; the product path has no CPU fallback."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from dmm_net_b200 import _lib, build


def header_functions():
    src = open(os.path.join(ROOT, "include", "dmm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(dmm_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = header_functions()
    assert len(names) >= 27
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dmm_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_header_arity_matches_ctypes_table():
    src = open(os.path.join(ROOT, "include", "dmm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_queries_work_without_a_device():
    lib = _lib.load()
    assert lib.dmm_b200_version() == 0x000200
    assert lib.dmm_b200_arch() == b"sm_100a"
    assert lib.dmm_b200_error_string(0) == b"ok" and b"workspace" in lib.dmm_b200_error_string(3)
    lim = (ctypes.c_int * 4)()
    assert lib.dmm_b200_limits(lim) == 0 and lim[0] == 16 and lim[1] == 128
    assert lib.dmm_mask_iou_workspace_bytes(4096, 50, 10, 256 * 448, 0) >= 4096 * 560 * 4
    assert lib.dmm_relax_saved_bytes(2, 10, 5) >= 2 * 50 * 32 * 8
    # argument validation happens before any CUDA call
    assert lib.dmm_mask_iou_pairwise(None, 0, None, 0, None, 0, 1, 2, 2, 16, None, None, None, None, None, 0.0, 0.0,
                                     None, None, None, 0, None) == 1


def test_no_cpu_fallback():
    from dmm_net_b200 import ops
    from dmm_net_b200.modules.match_model import MatchModel
    from dmm_net_b200.synth import default_cfg, make_problem
    pr = make_problem(4, 2, 8, 8, 16)
    layer = MatchModel(default_cfg(), is_test=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(pr.prop_feat, pr.prop_mask, [pr.tmpl_feat], pr.tmpl_mask, pr.prop_score)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.relax_solve(torch.zeros(1, 2, 3))


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dmm_net_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or "import oracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
