# tools/run_n2.sh N -- multi-GPU checks on an N-GPU box (gpurun --gpus N)
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/n${N}_topo.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/n${N}_tests.log 2>&1
tail -5 gpurun_out/n${N}_tests.log
for ex in peer nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 100 --warmup 5 --exchange $ex --no-extras > gpurun_out/n${N}_bench_$ex.json 2> gpurun_out/n${N}_bench_$ex.err
  tail -3 gpurun_out/n${N}_bench_$ex.err | cut -c1-300
done
python tools/show_bench.py gpurun_out/n${N}_bench_peer.json gpurun_out/n${N}_bench_nccl.json
