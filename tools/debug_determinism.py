"""Bitwise comparison of eager vs CUDA-graph-replayed forwards of the full-size UNetSD_I2VGen (diagnosis tool)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from oracle import synth  # noqa: E402
from oracle.cases import FULL_CASES, make_inputs  # noqa: E402
from _helpers import build_product, product_call  # noqa: E402
from vgen_b200 import graph  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "full_i2vgen"
tag = sys.argv[2] if len(sys.argv) > 2 else "x"
case = FULL_CASES[name]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = [(k, tuple(s)) for k, s in json.load(open(os.path.join(root, "tests", "golden", f"{case.get('spec', name)}.spec.json")))]
sd = synth.state_dict(spec, seed=case["seed"])
m = build_product(case)
m.load_state_dict(sd, strict=True)
m = m.cuda().eval()
del sd
inp = {k: v.cuda() for k, v in make_inputs(case).items()}


def diff(a, b):
    d = (a.float() - b.float()).abs()
    return {"equal": bool(torch.equal(a, b)), "max_abs": float(d.max()), "n_diff": int((d > 0).sum())}


os.environ["VGEN_CUDA_GRAPH"] = "0"
e1 = product_call(case, m, inp)
e2 = product_call(case, m, inp)
e3 = product_call(case, m, inp)
os.environ["VGEN_CUDA_GRAPH"] = "1"
g0 = product_call(case, m, inp)      # first sighting: eager
g1 = product_call(case, m, inp)      # capture + replay
g2 = product_call(case, m, inp)      # replay
g3 = product_call(case, m, inp)
torch.cuda.synchronize()
res = {"pdl": os.environ.get("VGEN_PDL", "default"), "name": name, "graph_stats": graph.stats(m),
       "eager_vs_eager": [diff(e1, e2), diff(e1, e3)], "eager_vs_first_sighting": diff(e1, g0),
       "eager_vs_replay": [diff(e1, g1), diff(e1, g2), diff(e1, g3)], "replay_vs_replay": [diff(g1, g2), diff(g2, g3)]}
print(json.dumps(res))
torch.save(e1.cpu(), os.path.join(root, "gpurun_out", f"det_{name}_{tag}.pt"))
