#!/usr/bin/env bash
# final single-GPU validation: smoke, all GPU tests, headline bench + reference arm, in-place benches
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_final.log; tail -3 gpurun_out/pytest_gpu_final.log
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/final_reference.json 2>/dev/null
python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/final_bench.json") if l.startswith("{")][-1])
r = json.loads([l for l in open("gpurun_out/final_reference.json") if l.startswith("{")][-1])
print("bench: %.1f MLUPS %.4f ms/step frac %.3f kernel frac %.3f e2e %.0f | reference arm %.2f MLUPS | same config: %s" % (
    d["value"], d["ms_per_step"], d["frac_of_roofline"], d["roofline"]["frac"], d["e2e"]["value"], r["value"], d["config"] == r["config"]))
for a in d["also"]:
    print("   ", a.get("workload", "?")[:40], a.get("value"), a.get("ms_per_step"), a.get("frac_of_roofline"), a.get("cpu_baseline", {}).get("value"), a.get("error"))
PY
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
run c5 X=1 --workload d3q27_channel_512x256x256 --steps 40
run c2_in_place X=1 --in-place --workload d2q9_karman_4096x1024 --steps 400
