"""The C-ABI library loads on a CPU-only box, exports every symbol include/lc2is_b200.h declares,
and every compute entry fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "lc2is_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lc2is_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from lc2is_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/lc2is_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes SIGNATURES out of sync with the header"
    assert lib.lc2is_abi_version() == 5


def test_class_pad():
    from lc2is_b200 import _lib
    assert [_lib.class_pad(c) for c in (1, 16, 150, 151, 847)] == [16, 16, 160, 160, 848]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device error path")
def test_no_cpu_fallback():
    from lc2is_b200 import _lib, ops, metrics
    from lc2is_b200.model.loss import AuxiliaryLoss
    assert _lib.lib.lc2is_count_valid(None, 0, 5, 0, None, None) == -3
    assert "no CPU fallback" in _lib.last_error()
    with pytest.raises(_lib.Lc2isError):
        ops.count_valid(torch.zeros(8, dtype=torch.int64), 5, 0)
    with pytest.raises(_lib.Lc2isError):
        AuxiliaryLoss(ignore_index=0)(torch.zeros(1, 3, 2, 2), torch.zeros(1, 8, 8, dtype=torch.int64))
    with pytest.raises(_lib.Lc2isError):
        metrics.compute_mIOU_tensor(torch.zeros(1, 3, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64), 3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "lc2is_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert "from oracle" not in txt and "import oracle" not in txt, f
