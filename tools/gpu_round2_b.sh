#!/usr/bin/env bash
# GPU batch B: task table on/off on the small and medium lattices; fp32-storage occupancy variants
mkdir -p gpurun_out
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
run c1_tasks1 PYLBM_B200_TASKS=1 --workload d2q9_lid_256 --steps 4000
run c1_tasks0 PYLBM_B200_TASKS=0 --workload d2q9_lid_256 --steps 4000
run c2_tasks1 PYLBM_B200_TASKS=1 --workload d2q9_karman_4096x1024 --steps 400
run c2_tasks0 PYLBM_B200_TASKS=0 --workload d2q9_karman_4096x1024 --steps 400
run c4s_tasks1 PYLBM_B200_TASKS=1 --workload d3q19_lid_256 --steps 100
run c4s_tasks0 PYLBM_B200_TASKS=0 --workload d3q19_lid_256 --steps 100
run c5_tasks1 PYLBM_B200_TASKS=1 --workload d3q27_channel_512x256x256 --steps 40
run c5_tasks0 PYLBM_B200_TASKS=0 --workload d3q27_channel_512x256x256 --steps 40
run c4_tasks1 PYLBM_B200_TASKS=1 --steps 20
run f32s_mb5 PYLBM_B200_TASKS=auto --dtype float32 --steps 50
run f32s_mb6 PYLBM_B200_MINBLOCKS=6 --dtype float32 --steps 50
run f32a_mb8 PYLBM_B200_TASKS=auto --dtype float32 --compute float32 --steps 50
