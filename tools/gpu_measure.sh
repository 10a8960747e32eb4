#!/bin/bash
# One-GPU measurement pass (run under gpurun): bench lines of every workload, ncu launch lists and
# ncu --set full captures of the fused kernel.  Outputs go to gpurun_out/ (copied to profiles/ by hand).
set -u
TAG=${1:-v4}
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/bench_d3q19_lid_512_$TAG.json 2> $OUT/bench_$TAG.err
python bench.py --dtype float32 --no-cpu-baseline > $OUT/bench_d3q19_lid_512_f32storage_$TAG.json 2>> $OUT/bench_$TAG.err
python bench.py --dtype float32 --compute float32 --no-cpu-baseline > $OUT/bench_d3q19_lid_512_f32_$TAG.json 2>> $OUT/bench_$TAG.err
for w in d2q9_karman_4096x1024 d2q4x3_shallow_water_4096 d3q27_channel_512x256x256 d2q9_lid_256 d3q19_lid_256; do
  python bench.py --workload $w --steps 400 --warmup 10 > $OUT/bench_${w}_$TAG.json 2>> $OUT/bench_$TAG.err
done
python bench.py --workload d2q9_karman_4096x1024 --dtype float32 --compute float32 --steps 400 --warmup 10 --no-cpu-baseline > $OUT/bench_d2q9_karman_4096x1024_f32_$TAG.json 2>> $OUT/bench_$TAG.err
# launch lists (per-launch durations, cold-cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lbmk_kernel|k_bc|k_periodic|k_signal|k_wait' -c 40 --csv --log-file $OUT/launches_d3q19_lid_512_$TAG.csv \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lbmk_kernel|k_bc|k_periodic|k_signal|k_wait' -c 60 --csv --log-file $OUT/launches_d2q9_karman_$TAG.csv \
    python bench.py --workload d2q9_karman_4096x1024 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
# full captures of the fused kernel
ncu --set full --clock-control none --import-source on -k regex:lbmk_kernel_one_time_step -s 4 -c 1 -o $OUT/prof_d3q19_512_$TAG \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lbmk_kernel_one_time_step -s 4 -c 1 -o $OUT/prof_d3q19_512_f32_$TAG \
    python bench.py --dtype float32 --compute float32 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lbmk_kernel_one_time_step -s 4 -c 1 -o $OUT/prof_karman_$TAG \
    python bench.py --workload d2q9_karman_4096x1024 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
for f in $OUT/bench_*_$TAG.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], d["dtype"], "MLUPS %.0f" % d["value"], "ms %.4f" % d["ms_per_step"], "frac %.3f" % d["frac_of_roofline"],
          "kernel frac %.3f" % d["roofline"]["frac"], "launches", d["gpu_launches"], "e2e", d["e2e"] and round(d["e2e"]["value"]),
          "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 1))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
