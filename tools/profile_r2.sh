# tools/profile_r2.sh TAG -- ncu --set full source-level captures of the two hot kernels (run under gpurun, one GPU)
set -x
TAG=${1:-r2}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 2 -c 1 -f -o gpurun_out/${TAG}_trace python tools/prof_trace.py 16777216 3 > gpurun_out/${TAG}_trace.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:knn_render_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_knn python tools/time_knn.py 16777216 50 1 > gpurun_out/${TAG}_knn.log 2>&1
python tools/time_trace.py > gpurun_out/${TAG}_time_trace.log 2>&1
python tools/time_knn.py 16777216 50 1 > gpurun_out/${TAG}_time_knn.log 2>&1
python tools/time_knn.py 4194304 100 0 >> gpurun_out/${TAG}_time_knn.log 2>&1
nvidia-smi topo -m > gpurun_out/${TAG}_topo.log 2>&1
tail -3 gpurun_out/${TAG}_*.log
