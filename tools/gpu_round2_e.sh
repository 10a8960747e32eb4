#!/usr/bin/env bash
# final single-GPU batch: smoke, full GPU tests, headline bench (+ reference arm), ncu launch list and full capture
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_final.log; tail -3 gpurun_out/pytest_gpu_final.log
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
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_d3q19_lid_512.csv \
    python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline --no-e2e > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lbmk_kernel_one_time_step -s 4 -c 1 -f -o gpurun_out/ncu_c4_fp64 \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-also > gpurun_out/ncu_c4_fp64.log 2>&1
ls -la gpurun_out | grep -i "ncu_c4\|launches\|final"
