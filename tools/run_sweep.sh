# tools/run_sweep.sh N "sms list" -- Mode A headline at N GPUs for several trace-kernel grid sizes (gpurun --gpus N)
N=${1:-4}
mkdir -p gpurun_out
for sms in $2; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 100 --warmup 5 --no-extras --trace-sms $sms > gpurun_out/n${N}_sms$sms.json 2> gpurun_out/n${N}_sms$sms.err
  python tools/show_bench.py gpurun_out/n${N}_sms$sms.json
done
