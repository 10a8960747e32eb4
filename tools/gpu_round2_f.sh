#!/usr/bin/env bash
# GPU batch F: programmatic dependent launch on/off: tests + small/medium/large lattices
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu_pdl.log; tail -3 gpurun_out/pytest_gpu_pdl.log
run() {  # name, env, args...
    name=$1; shift; envs=$1; shift
    env $envs python bench.py --no-e2e --no-cpu-baseline --no-also "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/%s.json" % name) if l.startswith("{")][-1])
    print("%-28s %10.1f MLUPS  %9.4f ms/step  stepwise %9.4f  kernel %9.4f ms  frac %.3f  launches/step %.2f" % (
        name, d["value"], d["ms_per_step"], d["stepwise"]["ms_per_step"], d["roofline"]["launch_ms"],
        d["frac_of_roofline"], d["gpu_launches"] / d["steps"]))
except Exception as exc:
    print(name, "FAILED", exc, open("gpurun_out/%s.err" % name).read()[-600:])
PY
}
run c1_pdl X=1 --workload d2q9_lid_256 --steps 4000
run c1_nopdl PYLBM_B200_NO_PDL=1 --workload d2q9_lid_256 --steps 4000
run c2_pdl X=1 --workload d2q9_karman_4096x1024 --steps 400
run c2_nopdl PYLBM_B200_NO_PDL=1 --workload d2q9_karman_4096x1024 --steps 400
run c5_pdl X=1 --workload d3q27_channel_512x256x256 --steps 40
run c5_nopdl PYLBM_B200_NO_PDL=1 --workload d3q27_channel_512x256x256 --steps 40
run c4_pdl X=1 --steps 30
run c4_nopdl PYLBM_B200_NO_PDL=1 --steps 30
run c1_pdl_tasks PYLBM_B200_TASKS=1 --workload d2q9_lid_256 --steps 4000
