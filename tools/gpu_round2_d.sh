#!/usr/bin/env bash
# GPU batch D: fp32-storage / fp64-arithmetic kernel: occupancy and integer-pipe conversion variants
mkdir -p gpurun_out
run() {  # name, env, args...
    name=$1; shift; envs=$1; shift
    env $envs python bench.py --no-e2e --no-cpu-baseline --no-also --api b200 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
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
run f32s_mb5 X=1 --dtype float32 --steps 50
run f32s_mb6 PYLBM_B200_MINBLOCKS=6 --dtype float32 --steps 50
run f32s_mb7 PYLBM_B200_MINBLOCKS=7 --dtype float32 --steps 50
run f32s_mb8 PYLBM_B200_MINBLOCKS=8 --dtype float32 --steps 50
run f32s_int_mb5 "PYLBM_B200_F2D=int" --dtype float32 --steps 50
run f32s_int_mb6 "PYLBM_B200_F2D=int PYLBM_B200_MINBLOCKS=6" --dtype float32 --steps 50
run f32s_int_mb7 "PYLBM_B200_F2D=int PYLBM_B200_MINBLOCKS=7" --dtype float32 --steps 50
run f32a_mb8 X=1 --dtype float32 --compute float32 --steps 50
run f32a_mb10 PYLBM_B200_MINBLOCKS=10 --dtype float32 --compute float32 --steps 50
