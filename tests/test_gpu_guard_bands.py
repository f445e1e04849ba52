"""Out-of-bounds WRITE detection without compute-sanitizer (closed on this GPU pool, profiles/README_r2.md).

Every buffer the host layer hands to the C ABI as an output or workspace (trajopt_grpo_b200.engine allocates them with
torch.empty / zeros / empty_like) is placed between two 4 KB guard bands filled with a sentinel byte; the bands must be
intact after every kernel family has run at shapes that leave a partial last tile (N = 70 / 200 envs for 128-sample tiles),
ragged episode lengths and odd sample-id lists.  The guard after a buffer starts at its last byte + 1, so a one-element
overrun is caught.  The cases are the ones tools/sanitize_cases.py runs.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

PAD = 4096
SENTINEL = 0xA5


class GuardedTorch:
    """Stands in for the `torch` module inside trajopt_grpo_b200.engine: allocation calls return the middle of a
    sentinel-filled buffer; everything else is torch's."""

    def __init__(self):
        self.allocs = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def _alloc(self, shape, dtype, device, zero):
        dtype = dtype or torch.float32
        shape = tuple(int(s) for s in shape)
        item = torch.empty((), dtype=dtype).element_size()
        nbytes = int(np.prod(shape, dtype=np.int64)) * item if shape else item
        raw = torch.full((PAD + nbytes + PAD,), SENTINEL, dtype=torch.uint8, device=device)
        body = raw[PAD:PAD + nbytes].view(dtype).view(shape)
        if zero:
            body.zero_()
        self.allocs.append((raw, nbytes, shape, dtype))
        return body

    @staticmethod
    def _shape(size):
        if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)):
            return tuple(size[0])
        return tuple(size)

    def empty(self, *size, dtype=None, device=None):
        return self._alloc(self._shape(size), dtype, device, False)

    def zeros(self, *size, dtype=None, device=None):
        return self._alloc(self._shape(size), dtype, device, True)

    def empty_like(self, x):
        return self._alloc(x.shape, x.dtype, x.device, False)

    def zeros_like(self, x):
        return self._alloc(x.shape, x.dtype, x.device, True)

    def check(self):
        torch.cuda.synchronize()
        bad = []
        for raw, nbytes, shape, dtype in self.allocs:
            lo_ok = bool((raw[:PAD] == SENTINEL).all())
            hi_ok = bool((raw[PAD + nbytes:] == SENTINEL).all())
            if not (lo_ok and hi_ok):
                bad.append((shape, dtype, "before" if not lo_ok else "after"))
        return bad


def _cases():
    import sanitize_cases
    return sanitize_cases.CASES


@pytest.mark.parametrize("name", ["fp32_small", "fp32_deep128", "tc64", "tc128", "tc256", "ppo", "env"])
def test_no_kernel_writes_outside_its_buffers(monkeypatch, name):
    from trajopt_grpo_b200 import engine
    guard = GuardedTorch()
    monkeypatch.setattr(engine, "torch", guard)
    engine._ws_cache.clear()
    try:
        _cases()[name]()
        bad = guard.check()
    finally:
        engine._ws_cache.clear()
    assert guard.allocs, "the case allocated nothing through the engine"
    assert not bad, f"guard bands overwritten: {bad}"


def test_guard_detects_a_one_element_overrun():
    """The checker itself: a write one float past the end of a guarded buffer is reported."""
    guard = GuardedTorch()
    t = guard.empty((3, 5), dtype=torch.float32, device="cuda")
    assert guard.check() == []
    raw = guard.allocs[0][0]
    raw[PAD + 3 * 5 * 4] = 0
    assert guard.check() == [((3, 5), torch.float32, "after")]
    assert t.shape == (3, 5)
