# tools/run_n2.sh -- multi-GPU checks on a 2-GPU box (gpurun --gpus 2)
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/n2_topo.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/n2_tests.log 2>&1
tail -5 gpurun_out/n2_tests.log
for ex in peer nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 100 --warmup 5 --exchange $ex > gpurun_out/n2_bench_$ex.json 2> gpurun_out/n2_bench_$ex.err
  tail -c 1500 gpurun_out/n2_bench_$ex.json; tail -3 gpurun_out/n2_bench_$ex.err
done
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 --no-ref-cuda --no-cpu-baseline --no-mode-b > gpurun_out/n2_bench_n1.json 2> gpurun_out/n2_bench_n1.err
tail -c 1500 gpurun_out/n2_bench_n1.json
