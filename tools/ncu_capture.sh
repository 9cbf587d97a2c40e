# `ncu --set full` captures of the kernels the bench line leans on, one launch each (run under gpurun; 1 GPU).
# Summaries: python tools/ncu_summary.py report gpurun_out/prof_*.ncu-rep > profiles/<tag>_ncu_summary.md
set -x
N="ncu --set full --clock-control none --import-source on"
$N -k regex:tapgemm_sm100_2cta -s 4 -c 1 -f -o gpurun_out/prof_conv3x3_88x160_c320 python tools/bench_shapes.py --only 7 > /dev/null 2>&1
$N -k regex:tapgemm_sm100_2cta -s 4 -c 1 -f -o gpurun_out/prof_linear_k320_n320_res python tools/bench_shapes.py --only 0 > /dev/null 2>&1
$N -k regex:tapgemm_sm100_2cta -s 4 -c 1 -f -o gpurun_out/prof_linear_k320_n960 python tools/bench_shapes.py --only 2 > /dev/null 2>&1
$N -k regex:tapgemm_sm100_2cta -s 4 -c 1 -f -o gpurun_out/prof_geglu_k320_n2560 python tools/bench_shapes.py --only 3 > /dev/null 2>&1
$N -k regex:attn_sm100 -s 2 -c 1 -f -o gpurun_out/prof_attn_14080 python tools/bench_attn.py --only 0 > /dev/null 2>&1
$N -k regex:attn_d512 -s 1 -c 1 -f -o gpurun_out/prof_attn_d512 python tools/bench_attn.py --only d512 > /dev/null 2>&1
$N -k regex:gn_ -s 3 -c 3 -f -o gpurun_out/prof_gn python tools/prof_gn.py --only 2 > /dev/null 2>&1
$N -k regex:layernorm -s 2 -c 1 -f -o gpurun_out/prof_ln python tools/prof_gn.py --only ln > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
