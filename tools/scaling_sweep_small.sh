# Mode A at N=2 and N=4 on a 4-GPU box (the full 1/2/4/8 sweep is tools/scaling_sweep.sh)
F="--no-ref-cuda --no-cpu-baseline --no-mode-b"
for n in 2 4; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n $F > gpurun_out/bench_r1b_n$n.json 2>> gpurun_out/sweep_err.log
done
tail -n 1 gpurun_out/bench_r1b_n*.json | cut -c1-300
