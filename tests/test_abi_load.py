"""CPU checks of the C-ABI boundary: the library loads, exports every symbol the
header declares, and fails loudly (no CPU fallback) without a device."""
import os
import re

import pytest
import torch

from trajopt_grpo_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "trajopt_grpo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/trajopt_grpo.h but not exported"
    assert sorted(L.EXPORTS) == syms            # the ctypes table binds exactly the header
    assert lib.tg_abi_version() == 2


def test_host_side_queries_need_no_device():
    lib = L.load()
    import ctypes as C
    o, a = C.c_int(), C.c_int()
    assert lib.tg_env_dims(3, C.byref(o), C.byref(a)) == 0 and (o.value, a.value) == (20, 4)
    assert lib.tg_env_dims(9, C.byref(o), C.byref(a)) == -1
    assert b"unknown env kind" in lib.tg_last_error()
    cfg = L.mlp_cfg([20, 256, 256, 4])
    assert lib.tg_mlp_param_count(C.byref(cfg)) == 20 * 256 + 256 + 256 * 256 + 256 + 256 * 4 + 4


def test_time_thresholds_match_oracle():
    import restate as R
    for dt, T in ((0.05, 200), (0.02, 500), (0.02, 1000), (0.02, 120)):
        assert L.time_limit_step(dt, T) == R.time_limit_step(dt, T)
    assert L.balanced_limit_count(0.05) == R.balanced_limit_count(0.05) == 101


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(L.EngineError, match="no CPU fallback"):
        L.ctx()
    from trajopt_grpo_b200 import engine
    with pytest.raises(L.EngineError, match="CUDA tensor"):
        engine.adam_step(torch.zeros(4), torch.zeros(4), torch.zeros(4), torch.zeros(4), 1, 1e-3)
