tools/ncu_span.sh lattice.json 1024 64 r2c_lattice render_span_kernel
tools/ncu_span.sh pillar_array.json 4096 8 r2c_pillar render_span_kernel
tools/ncu_span.sh gyroid_example.json 1024 64 r2c_gyroid render_async_kernel deformation_sigmoid.json
tools/ncu_span.sh voxel1024 2048 8 r2c_voxel render_volume_tex_kernel
tools/ncu_span.sh cube_w_hole.json 512 1 r2c_cube render_span_kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > gpurun_out/b_ncu.log 2>&1
ls gpurun_out | grep r2c | head -40
