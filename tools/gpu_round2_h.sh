#!/usr/bin/env bash
# GPU batch H: in-place streaming (AA pattern): parity tests + roofline check
mkdir -p gpurun_out
python -m pytest tests/test_gpu_aa.py -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu_aa.log; tail -25 gpurun_out/pytest_gpu_aa.log
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
    print(name, "FAILED", exc, open("gpurun_out/%s.err" % name).read()[-800:])
PY
}
run c4_in_place X=1 --in-place --steps 30
run c5_in_place X=1 --in-place --workload d3q27_channel_512x256x256 --steps 40
run c2_in_place X=1 --in-place --workload d2q9_karman_4096x1024 --steps 400
run c3_in_place X=1 --in-place --workload d2q4x3_shallow_water_4096 --steps 100
