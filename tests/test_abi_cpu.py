"""CPU checks of the C-ABI boundary: the library builds for sm_100a, loads, and exports every symbol
include/scv.h declares; the ctypes struct mirrors match the header field-for-field."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header():
    return open(os.path.join(ROOT, "include", "scv.h")).read()


def test_library_builds_and_exports_all_declared_symbols():
    from scrubvae_b200 import build, _ops
    path = build.build()
    assert os.path.exists(path)
    lib = _ops.load_library(path)  # raises AttributeError on a missing symbol
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(scv_\w+)\s*\(", _header(), flags=re.M))
    assert declared, "no declarations parsed"
    assert declared == set(_ops.EXPORTS), declared ^ set(_ops.EXPORTS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.scv_version() >= 100


def test_struct_mirrors_match_header():
    from scrubvae_b200 import _ops
    hdr = _header()
    for cname, mirror in (("scv_gemm_t", _ops.GemmT), ("scv_wgrad_t", _ops.WgradT), ("scv_bnact_t", _ops.BnactT),
                          ("scv_bnact_bwd_t", _ops.BnactBwdT), ("scv_optim_t", _ops.OptimT)):
        body = [c for c in hdr.split("typedef struct {")[1:] if c.split("}")[1].strip().startswith(cname + ";")]
        body = body[0].split("}")[0]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(",")
            first = names[0].split()[-1].lstrip("*")
            fields.append(first)
            fields += [n.strip().lstrip("*") for n in names[1:]]
        assert fields == [f[0] for f in mirror._fields_], (cname, fields, [f[0] for f in mirror._fields_])


def test_missing_library_is_an_error_not_a_fallback(tmp_path):
    from scrubvae_b200 import _ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _ops.load_library(str(tmp_path / "libscv.so"))


def test_engine_refuses_cpu_model_with_cuda_ops():
    import torch
    import scrubvae_b200 as sv
    from scrubvae_b200.engine import Engine

    class FakeCuda:
        name = "cuda"
    mc = dict(type="rcnn", channel=[8, 16], kernel=5, z_dim=8, window=51, activation="prelu", diag=False,
              init_dilation=None, prior="gaussian", load_model=None, start_epoch=None)
    m = sv.get.model(mc, None, None, dict(method={}, features=[], alpha=1.0), 18, "midfwd",
                     arena_size=torch.tensor([[-1.0, -1, 0], [1, 1, 1]]), device="cpu", verbose=0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        Engine(m, ops=FakeCuda())
