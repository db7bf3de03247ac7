# N=8 points only (Mode A, Mode B, config 5); the full 1/2/4/8 sweep is tools/scaling_sweep.sh
F="--no-ref-cuda --no-cpu-baseline --no-mode-b"
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 120 $R --master-port 29511 bench.py --gpus 8 $F > gpurun_out/bench_r1b_n8.json 2>> gpurun_out/sweep_err.log
timeout 120 $R --master-port 29512 bench.py --gpus 8 --mode b --steps 5 --warmup 3 $F > gpurun_out/bench_r1b_modeb_n8.json 2>> gpurun_out/sweep_err.log
timeout 120 $R --master-port 29513 bench.py --gpus 8 --passes 64 --width 3840 --height 2160 --steps 3 --warmup 3 $F > gpurun_out/bench_r1b_cfg5_n8.json 2>> gpurun_out/sweep_err.log
tail -n 1 gpurun_out/bench_r1b_n8.json gpurun_out/bench_r1b_modeb_n8.json gpurun_out/bench_r1b_cfg5_n8.json | cut -c1-220
