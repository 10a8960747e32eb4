#!/bin/bash
# v5 pass (fused-kernel walls): bench lines of the workloads that changed + ncu captures
OUT=gpurun_out; TAG=v5; mkdir -p $OUT
python bench.py > $OUT/bench_d3q19_lid_512_$TAG.json 2> $OUT/bench_$TAG.err
python bench.py --dtype float32 --compute float32 --no-cpu-baseline > $OUT/bench_d3q19_lid_512_f32_$TAG.json 2>> $OUT/bench_$TAG.err
python bench.py --dtype float32 --no-cpu-baseline --no-e2e > $OUT/bench_d3q19_lid_512_f32storage_$TAG.json 2>> $OUT/bench_$TAG.err
for w in d3q27_channel_512x256x256 d3q19_lid_256; do
  python bench.py --workload $w --steps 400 --warmup 10 --no-e2e > $OUT/bench_${w}_$TAG.json 2>> $OUT/bench_$TAG.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lbmk_kernel|k_bc|k_periodic' -c 24 --csv --log-file $OUT/launches_d3q19_lid_512_$TAG.csv \
    python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'lbmk_kernel|k_bc|k_periodic' -c 30 --csv --log-file $OUT/launches_d3q27_channel_$TAG.csv \
    python bench.py --workload d3q27_channel_512x256x256 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lbmk_kernel_one_time_step -s 4 -c 1 -o $OUT/prof_d3q19_512_$TAG \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lbmk_kernel_one_time_step -s 4 -c 1 -o $OUT/prof_d3q19_512_f32_$TAG \
    python bench.py --dtype float32 --compute float32 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
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
