#!/usr/bin/env bash
# GPU batch G: all demo dictionaries through pylbm.Simulation(generator='cuda'); occupancy of the D3Q27 kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_plugin.py -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu_plugin_all.log; tail -3 gpurun_out/pytest_gpu_plugin_all.log
run() {  # name, env, args...
    name=$1; shift; envs=$1; shift
    env $envs python bench.py --no-e2e --no-cpu-baseline --no-also --api b200 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
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
run c5_mb5 X=1 --workload d3q27_channel_512x256x256 --steps 40
run c5_mb4 PYLBM_B200_MINBLOCKS=4 --workload d3q27_channel_512x256x256 --steps 40
run c5_mb6 PYLBM_B200_MINBLOCKS=6 --workload d3q27_channel_512x256x256 --steps 40
run c2_mb8 PYLBM_B200_MINBLOCKS=8 --workload d2q9_karman_4096x1024 --steps 400
run c2_mb5 X=1 --workload d2q9_karman_4096x1024 --steps 400
