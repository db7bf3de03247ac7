set -x
mkdir -p gpurun_out
F="--no-ref-cuda --no-cpu-baseline --no-mode-b"
timeout 200 python bench.py --gpus 1 $F > gpurun_out/bench_r1_n1.json 2> gpurun_out/sweep_err.log
for n in 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n $F > gpurun_out/bench_r1_n$n.json 2>> gpurun_out/sweep_err.log
done
timeout 200 python bench.py --gpus 1 --mode b --steps 5 --warmup 3 $F > gpurun_out/bench_r1_modeb_n1.json 2>> gpurun_out/sweep_err.log
for n in 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --mode b --steps 5 --warmup 3 $F > gpurun_out/bench_r1_modeb_n$n.json 2>> gpurun_out/sweep_err.log
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --passes 64 --width 3840 --height 2160 --steps 3 --warmup 3 $F > gpurun_out/bench_r1_cfg5_n8.json 2>> gpurun_out/sweep_err.log
tail -n 1 gpurun_out/bench_r1_n*.json gpurun_out/bench_r1_modeb_n*.json gpurun_out/bench_r1_cfg5_n8.json | cut -c1-400
