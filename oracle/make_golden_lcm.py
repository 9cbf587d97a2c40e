"""TEST INFRASTRUCTURE ONLY -- partial pins for the LCM sampler of BASELINE config 3.

`diffusers.LCMScheduler` itself is absent (see oracle/lcm_oracle.py), but the reference's own VideoLCM training script
restates three pieces of it in-tree (tools/train/train_videolcm_t2v_entrance.py):

  :129-133  scalings_for_boundary_conditions   "From LCMScheduler.get_scalings_for_boundary_condition_discrete"
  :136-150  predicted_origin                    "Compare LCMScheduler.step, Step 4"
  :159-176  DDIMSolver.__init__                 the 50-step DDIM grid the consistency model was distilled on

plus the zero-terminal-SNR rescale (arXiv:2305.08891 alg. 1), which the reference implements itself in
tools/modules/diffusions/schedules.py:143-165 (`rescale_zero_terminal_snr`): it is applied here to the scheduler's fp32
scaled-linear betas and the resulting alphas_cumprod frozen.

The script as a whole cannot be imported (it imports diffusers at the top), so this generator lifts exactly those
definitions out of the file with `ast`, executes them unmodified, and freezes their outputs on seeded inputs into
tests/golden/lcm_pins.npz.  What stays unpinned: diffusers' inference-time timestep selection, its re-noising
step.  Run in the build container:  python -m oracle.make_golden_lcm
"""
from __future__ import annotations

import ast
import os

import numpy as np
import torch

from . import lcm_oracle, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/tools/train/train_videolcm_t2v_entrance.py"
WANTED = ("scalings_for_boundary_conditions", "predicted_origin", "extract_into_tensor", "DDIMSolver")
TIMESTEPS = [999, 759, 499, 259, 19, 0]


def lift():
    tree = ast.parse(open(SRC).read())
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in WANTED]
    assert sorted(n.name for n in keep) == sorted(WANTED), [n.name for n in keep]
    ns = {"torch": torch, "np": np}
    exec(compile(ast.Module(body=keep, type_ignores=[]), SRC, "exec"), ns)
    return ns


def main():
    ns = lift()
    out = {"timesteps": np.asarray(TIMESTEPS, np.int64)}
    cs = [ns["scalings_for_boundary_conditions"](float(t)) for t in TIMESTEPS]
    out["c_skip"] = np.asarray([c[0] for c in cs], np.float64)
    out["c_out"] = np.asarray([c[1] for c in cs], np.float64)

    abar = lcm_oracle.alphas_cumprod(True).double()          # an INPUT here: the pinned part is the x0 formula
    alphas, sigmas = abar.sqrt(), (1 - abar).sqrt()
    shape = (6, 4, 2, 3, 5)
    x = synth.tensor("lcm_sample", shape, 1.0, 3).double()
    v = synth.tensor("lcm_model_output", shape, 1.0, 4).double()
    t = torch.as_tensor(TIMESTEPS, dtype=torch.long)
    out["x0_v"] = ns["predicted_origin"](v, t, x, "v_prediction", alphas, sigmas).numpy()
    safe = torch.as_tensor(TIMESTEPS[1:], dtype=torch.long)  # alphas[999] == 0 under zero terminal SNR: eps form divides by it
    out["x0_eps"] = ns["predicted_origin"](v[1:], safe, x[1:], "epsilon", alphas, sigmas).numpy()

    solver = ns["DDIMSolver"](abar.numpy(), timesteps=1000, ddim_timesteps=50)
    out["ddim_grid"] = solver.ddim_timesteps.numpy()
    out["ddim_abar"] = solver.ddim_alpha_cumprods.numpy()
    out["ddim_abar_prev"] = solver.ddim_alpha_cumprods_prev.numpy()
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ref_schedules", "/root/reference/tools/modules/diffusions/schedules.py")
    ref_sched = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_sched)
    betas = ref_sched.rescale_zero_terminal_snr(lcm_oracle.scaled_linear_betas())
    out["abar_zero_snr"] = torch.cumprod(1.0 - betas, dim=0).numpy()
    path = os.path.join(ROOT, "tests", "golden", "lcm_pins.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
