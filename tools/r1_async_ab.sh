#!/bin/bash
# Development aid (1 GPU): GPU test suite, then lockstep vs lane-asynchronous kernels on the one-primitive workloads.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --maxfail=12 --tb=short -p no:cacheprovider ) > gpurun_out/pytest_gpu.txt 2>&1
tail -8 gpurun_out/pytest_gpu.txt
one() {  # label, env assignments..., -- bench args
  label=$1; shift
  env "$@" python bench.py --steps 3 --warmup 3 --no-cpu --no-ref-cuda $BARGS 2>&1 | tail -1 > gpurun_out/ab_$label.json
  python - gpurun_out/ab_$label.json $label <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); r=d["roofline"]
    print(sys.argv[2], d["config"]["workload"][:16], "Gs/s", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms", round(d["ms_per_step"],2),
          "eval", int(r["evaluated_samples"]), "fb", int(r["fp64_fallbacks"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], (r.get("survey_8d_brute_force") or {}).get("frac"))
except Exception as e:
    print(sys.argv[2], "FAILED", e, open(sys.argv[1]).read()[-400:])
PY
}
# optional extra builds to compare: put them under xray_projection_render_b200/lib/variants/*.so (make EXTRA=-D..., then copy the .so)
V=xray_projection_render_b200/lib/variants
BARGS="--workload gyroid_sigmoid --views 8"; one gy_async A=1; one gy_lock XRAY_NO_ASYNC=1
for v in $V/*.so; do [ -e "$v" ] && one gy_$(basename $v .so) XRAY_CUDA_LIB=$PWD/$v; done
BARGS="--workload pillar_array --views 8"; one pil_async A=1; one pil_lock XRAY_NO_ASYNC=1
for v in $V/*.so; do [ -e "$v" ] && one pil_$(basename $v .so) XRAY_CUDA_LIB=$PWD/$v; done
BARGS="--workload lattice --views 60"; one lat A=1
for w in balls box_w_pped lattice_linear lattice_sigmoid cube_w_hole; do BARGS="--workload $w --views 8"; one $w A=1; done
