bash tools/ncu_span.sh lattice.json 1024 64 r2i_lattice "render_span_kernel<.int.1, .bool.0, .bool.0"
bash tools/ncu_span.sh pillar_array.json 4096 8 r2i_pillar "render_span_kernel<.int.1, .bool.0, .bool.0"
bash tools/ncu_span.sh cube_w_hole.json 512 1 r2i_cube "render_span_kernel<.int.1, .bool.0, .bool.0"
rm -f gpurun_out/prof_r2i_pillar.ncu-rep gpurun_out/cuda_r2i_*.csv gpurun_out/src_r2i_*.csv
