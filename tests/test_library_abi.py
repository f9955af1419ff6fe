"""CPU: the C-ABI shared library loads and exports every symbol include/femus_b200.h declares;
compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import pytest

from femus_b200 import capi


def test_exports_every_declared_symbol():
    L = capi.lib()
    names = capi.header_symbols()
    assert len(names) > 60
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    undeclared = [n for n in capi._PROTOS if n not in names]
    assert not undeclared, undeclared


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.B2Error) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for dp, _, files in os.walk(os.path.join(root, "femus_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if "import oracle" in txt or "from oracle" in txt or "oracle/" in txt:
                    bad.append(f)
    assert not bad, bad
