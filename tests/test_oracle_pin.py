"""Direct pin of the oracle against the real reference classes (only where /root/reference is mounted)."""
import pytest
import torch

from oracle import refload, synth, vgen_oracle as vo
from oracle.cases import CASES, make_inputs

pytestmark = pytest.mark.skipif(not refload.available(), reason="reference not mounted (GPU box)")


def _maxrel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def test_unet_t2v_against_reference_class():
    torch.set_grad_enabled(False)
    ref = refload.load()
    case = CASES["t2v_tiny_b2"]
    m = ref.UNetSD_T2VBase(**case["ctor"]).eval()
    sd = synth.state_dict(synth.spec_of(m), seed=77)          # a seed the golden files do not use
    m.load_state_dict(sd, strict=True)
    inp = make_inputs(case)
    assert _maxrel(vo.unet_t2v_forward(sd, inp["x"], inp["t"], inp["y"]), m(inp["x"], inp["t"], y=inp["y"])) < 2e-5


@pytest.mark.parametrize("name", ["videolcm_tiny", "sr600_tiny", "higen_tiny", "higen_tiny_f1"])
def test_unet_variants_against_reference_classes(name):
    """a21 variants, with weights from a seed the golden files do not use."""
    from oracle.make_golden import build_variant
    from _helpers import oracle_call, product_call
    torch.set_grad_enabled(False)
    ref = refload.load()
    case = CASES[name]
    m = build_variant(ref, case["kind"], case["ctor"]).eval()
    sd = synth.state_dict(synth.spec_of(m), seed=78)
    m.load_state_dict(sd, strict=True)
    inp = make_inputs(case)
    assert _maxrel(oracle_call(case, sd, inp), product_call(case, m, inp)) < 2e-5


def test_default_init_is_degenerate_and_synth_is_not():
    """SURVEY.md section 8c hygiene: with the reference's default init the output is a per-channel constant."""
    torch.set_grad_enabled(False)
    ref = refload.load()
    case = CASES["t2v_tiny"]
    torch.manual_seed(0)
    m = ref.UNetSD_T2VBase(**case["ctor"]).eval()
    inp = make_inputs(case)
    out = m(inp["x"], inp["t"], y=inp["y"])
    assert float(out.std(dim=(2, 3, 4)).max()) < 1e-6
    m.load_state_dict(synth.state_dict(synth.spec_of(m), seed=case["seed"]), strict=True)
    assert float(m(inp["x"], inp["t"], y=inp["y"]).std()) > 0.1


def test_ddim_tables_against_reference_class():
    ref = refload.load()
    d = ref.DiffusionDDIM(schedule="cosine", schedule_param=dict(num_timesteps=1000, cosine_s=0.008, zero_terminal_snr=True),
                          mean_type="v", var_type="fixed_small")
    tab = vo.ddim_tables(vo.make_betas("cosine", 1000, True, cosine_s=0.008))
    for k, v in tab.items():
        assert torch.equal(v, getattr(d, k)), k


def test_higen_motion_cond_per_frame_branch():
    """get_motion_embedding with motion_cond.size(1) == f (no interpolation), unet_higen.py:393-394."""
    from oracle.make_golden import build_variant
    torch.set_grad_enabled(False)
    ref = refload.load()
    case = CASES["higen_tiny"]
    m = build_variant(ref, "higen", case["ctor"]).eval()
    sd = synth.state_dict(synth.spec_of(m), seed=79)
    m.load_state_dict(sd, strict=True)
    inp = make_inputs(case)
    b, f = inp["x"].shape[0], inp["x"].shape[2]
    mc = torch.tensor([[100 + 37 * i + 11 * j for j in range(f)] for i in range(b)], dtype=torch.long)
    kw = dict(spat_prior=inp["spat_prior"], motion_cond=mc, appearance_cond=inp["appearance_cond"])
    out = m(inp["x"], inp["t"], y=inp["y"], **kw)
    assert _maxrel(vo.unet_higen_forward(sd, inp["x"], inp["t"], inp["y"], **kw), out) < 2e-5


@pytest.mark.parametrize("kind", ["t2v", "videolcm"])
def test_fps_condition_branch(kind):
    """use_fps_condition=True adds fps_embedding(sinusoidal(fps)) to the time embedding (unet_t2v.py:246-249)."""
    from oracle.make_golden import build_variant
    torch.set_grad_enabled(False)
    ref = refload.load()
    case = CASES["t2v_tiny" if kind == "t2v" else "videolcm_tiny"]
    ctor = dict(case["ctor"], use_fps_condition=True)
    m = (ref.UNetSD_T2VBase(**ctor) if kind == "t2v" else build_variant(ref, "videolcm", ctor)).eval()
    sd = synth.state_dict(synth.spec_of(m), seed=80)
    m.load_state_dict(sd, strict=True)
    inp = make_inputs(case)
    fps = torch.tensor([8] * inp["x"].shape[0], dtype=torch.long)
    out = m(inp["x"], inp["t"], y=inp["y"], fps=fps)
    fn = vo.unet_t2v_forward if kind == "t2v" else vo.unet_videolcm_forward
    assert _maxrel(fn(sd, inp["x"], inp["t"], inp["y"], fps=fps, use_fps_condition=True), out) < 2e-5
    # and the product's parameter spec follows the flag
    from vgen_b200 import arch
    kk = dict(ctor, dim_mult=tuple(ctor["dim_mult"]), attn_scales=tuple(ctor["attn_scales"]))
    assert arch.unet_spec(arch.unet_plan(kind, **kk)) == [(k, tuple(s)) for k, s in synth.spec_of(m)]
