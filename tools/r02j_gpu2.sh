set -x
N=${1:-2}
for w in videolcm higen i2vgen; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-pass 0 > gpurun_out/r02j_bench_${w}_n$N.json 2> gpurun_out/r02j_bench_${w}_n$N.err; tail -c 400 gpurun_out/r02j_bench_${w}_n$N.err; grep -h '^{' gpurun_out/r02j_bench_${w}_n$N.json | cut -c1-200
done
