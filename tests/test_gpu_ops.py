"""Kernel-level parity on the B200: every C-ABI op against a plain PyTorch fp32 reference of the same op
(the case lists live in tools/gpu_check.py so the bring-up tool and the gate share them).
Tolerance: 2e-3 (3e-3 attention) of the output's max magnitude -- fp16 output rounding (2^-11) plus
fp32-accumulated fp16 products; layout / copy kernels are exact."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _checks():
    spec = importlib.util.spec_from_file_location("gpu_check", os.path.join(ROOT, "tools", "gpu_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("group", ["tapgemm", "tapgemm_1cta", "tapgemm_2cta", "tapgemm_simt", "norm", "attention", "elementwise", "variants"])
def test_op_group(group):
    import torch
    mod = _checks()
    mod.RESULTS.clear()
    mod.GROUPS[group]()
    torch.cuda.synchronize()
    from vgen_b200 import ops
    ops.set_tapgemm_impl("sm100")
    assert mod.RESULTS, "no cases ran"
    bad = [r for r in mod.RESULTS if not r["ok"]]
    assert not bad, bad


def test_native_library_is_what_ran():
    from vgen_b200 import lib
    assert lib.launch_count() > 0
    maps = open("/proc/self/maps").read()
    assert "libvgen_b200.so" in maps
