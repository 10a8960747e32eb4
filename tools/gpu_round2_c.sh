#!/usr/bin/env bash
# GPU batch C: full GPU tests after the cells-per-thread refactor; fp32-storage / all-fp32 variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_c.log
tail -4 gpurun_out/pytest_gpu_c.log
run() {  # name, env, args...
    name=$1; shift; envs=$1; shift
    env $envs python bench.py --no-e2e --no-cpu-baseline --no-also "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open("gpurun_out/%s.json" % name))
    print("%-28s %10.1f MLUPS  %9.4f ms/step  stepwise %9.4f  kernel %9.4f ms  frac %.3f  launches/step %.2f" % (
        name, d["value"], d["ms_per_step"], d["stepwise"]["ms_per_step"], d["roofline"]["launch_ms"],
        d["frac_of_roofline"], d["gpu_launches"] / d["steps"]))
except Exception as exc:
    print(name, "FAILED", exc, open("gpurun_out/%s.err" % name).read()[-400:])
PY
}
run c4_f64 X=1 --steps 30
run f32s_cpt2_mb4 X=1 --dtype float32 --steps 50
run f32s_cpt2_mb3 PYLBM_B200_MINBLOCKS=3 --dtype float32 --steps 50
run f32s_cpt2_mb5 PYLBM_B200_MINBLOCKS=5 --dtype float32 --steps 50
run f32a_cpt2_mb6 X=1 --dtype float32 --compute float32 --steps 50
run f32a_cpt2_mb5 PYLBM_B200_MINBLOCKS=5 --dtype float32 --compute float32 --steps 50
run f32a_cpt2_mb8 PYLBM_B200_MINBLOCKS=8 --dtype float32 --compute float32 --steps 50
run f32a_cpt1_mb8 "PYLBM_B200_CPT=1 PYLBM_B200_MINBLOCKS=8" --dtype float32 --compute float32 --steps 50
run c4_f64_cpt2 "PYLBM_B200_CPT=2 PYLBM_B200_MINBLOCKS=4" --steps 30
run c2 X=1 --workload d2q9_karman_4096x1024 --steps 400
run c2_cpt2 "PYLBM_B200_CPT=2 PYLBM_B200_MINBLOCKS=5" --workload d2q9_karman_4096x1024 --steps 400
run c3 X=1 --workload d2q4x3_shallow_water_4096 --steps 100
run c5 X=1 --workload d3q27_channel_512x256x256 --steps 40
run c1 X=1 --workload d2q9_lid_256 --steps 4000
