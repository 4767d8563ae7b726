ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2i_bench.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu > gpurun_out/launches_r2i_bench.log 2>&1
tail -2 gpurun_out/launches_r2i_bench.log | cut -c1-200
