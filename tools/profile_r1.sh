set -x
F="--no-ref-cuda --no-cpu-baseline --no-mode-b"
timeout 300 python bench.py > gpurun_out/bench_r1_full_n1.json 2> gpurun_out/bench_full_err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 10 --warmup 5 $F > gpurun_out/b_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:trace_kernel|fold_volume|build_map|build_tables|render_kernel' -s 25 -c 5 -f -o gpurun_out/r1_full python bench.py --steps 10 --warmup 5 $F > gpurun_out/b_full.log 2>&1
tail -c 600 gpurun_out/bench_r1_full_n1.json
