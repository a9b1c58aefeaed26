"""`not gpu`: the C-ABI library builds, loads, and exports every symbol include/vdn_b200.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from video_depth_normal_v2_b200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    from video_depth_normal_v2_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "vdn_b200.h")).read()
    declared = set(re.findall(r"\b(vdn_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name


def test_host_only_calls(lib):
    assert lib.vdn_version() >= 100
    assert lib.vdn_set_operand_format(1) == 0 and lib.vdn_get_operand_format() == 1
    assert lib.vdn_set_operand_format(0) == 0
    assert lib.vdn_set_operand_format(7) != 0 and b"operand format" in lib.vdn_last_error()
    lib.vdn_reset_launch_count()
    assert lib.vdn_launch_count() == 0


def test_gemm_desc_matches_header():
    """The ctypes mirror must list the header's struct fields in the same order."""
    from video_depth_normal_v2_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "vdn_b200.h")).read()
    body = hdr[hdr.index("typedef struct vdn_gemm_desc {"):hdr.index("} vdn_gemm_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for stmt in body.split(";")[:-1]:
        stmt = stmt.replace("typedef struct vdn_gemm_desc {", "")
        parts = stmt.split(",")
        for i, part in enumerate(parts):
            names.append(re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*$", part.strip())[0])
    assert names == [f[0] for f in _lib.GemmDesc._fields_]


def test_ops_refuse_cpu_tensors(lib):
    import torch
    from video_depth_normal_v2_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.relu16(torch.zeros(8, dtype=ops.operand_dtype()), torch.zeros(8, dtype=ops.operand_dtype()))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from video_depth_normal_v2_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.load()


def test_torch_ops_are_registered_and_cuda_only():
    """The kernels are exposed as torch.ops.vdn.* (SURVEY.md §8b); no CPU implementation exists, so CPU tensors raise."""
    import torch
    import video_depth_normal_v2_b200 as pkg
    for name in pkg.torch_ops.REGISTERED:
        assert hasattr(torch.ops.vdn, name), name
    x = torch.zeros(4, 8)
    import pytest
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.vdn.layernorm(x, torch.ones(8), torch.zeros(8), torch.empty(4, 8, dtype=torch.float16), 1e-6)
