# tools/profile_r2.sh -- round-2 ncu evidence (run under gpurun, ONE GPU): launch list of the bench command, one --set full capture
# of every Mode A kernel of a frame, one of the Mode B gather kernel.  Numbers printed by runs under ncu are never bench values.
set -x
mkdir -p gpurun_out
F="--no-extras --no-ref-cuda --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 10 --warmup 5 $F > gpurun_out/r2_b_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:trace_kernel|fold_volume|build_map|build_tables|render_kernel' -s 25 -c 5 -f -o gpurun_out/r2_full python bench.py --steps 10 --warmup 5 $F > gpurun_out/r2_b_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:knn_render_kernel -s 1 -c 1 -f -o gpurun_out/r2_knn python tools/time_knn.py 16777216 50 1 > gpurun_out/r2_knn.log 2>&1
ls -la gpurun_out/r2_*
