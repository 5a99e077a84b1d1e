"""The C-ABI library loads on a CPU-only machine and exports every symbol that
include/*.h declares (no compute calls here — those need a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    inc = os.path.join(ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if not fn.endswith(".h"):
            continue
        src = open(os.path.join(inc, fn)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        for m in re.finditer(r"\b((?:xllm|tokenizers)_[a-z0-9_]+)\s*\(", src):
            syms.add(m.group(1))
    return sorted(syms)


@pytest.fixture(scope="module")
def cdll():
    import __graft_entry__ as ge
    ge.build()
    import xllm_service_b200 as x
    assert os.path.exists(x.lib_path())
    return ctypes.CDLL(x.lib_path())


def test_exports_every_declared_symbol(cdll):
    syms = declared_symbols()
    assert "xllm_ingest_create" in syms and "xllm_hash_blocks" in syms
    missing = [s for s in syms if not hasattr(cdll, s)]
    assert not missing, missing


def test_create_fails_loudly_without_gpu_or_bad_args():
    import torch
    import xllm_service_b200 as x
    with pytest.raises(x.IngestError):
        x.Ingest(block_size=252)  # hash_util.cpp:33 frame limit
    if not torch.cuda.is_available():
        with pytest.raises(x.IngestError) as ei:
            x.Ingest()
        assert ei.value.code == -2  # XLLM_ERR_CUDA: no CPU fallback
