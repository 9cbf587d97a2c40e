set -x
VGEN_PDL=0 python tools/debug_determinism.py full_i2vgen pdl0 2>&1 | tail -1
VGEN_PDL=1 python tools/debug_determinism.py full_i2vgen pdl1 2>&1 | tail -1
python - <<'PY'
import torch
a=torch.load("gpurun_out/det_full_i2vgen_pdl0.pt"); b=torch.load("gpurun_out/det_full_i2vgen_pdl1.pt")
print("pdl0 vs pdl1 eager equal:", torch.equal(a,b), float((a.float()-b.float()).abs().max()))
PY
VGEN_ATTN_TCSUM=0 python tools/bench_attn.py > gpurun_out/r02f_attn_tcsum0.log 2>&1; head -3 gpurun_out/r02f_attn_tcsum0.log | cut -c1-330
VGEN_ATTN_TCSUM=1 python tools/bench_attn.py > gpurun_out/r02f_attn_tcsum1.log 2>&1; head -3 gpurun_out/r02f_attn_tcsum1.log | cut -c1-330
python -m pytest tests/test_gpu_ops.py -q -x -k "attention" 2>&1 | tail -3
