#!/bin/bash
# 8-GPU pass (run under gpurun --gpus 8): strong scaling fp64 and all-fp32, weak scaling D3Q27
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-v4}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:3}" > $OUT/$2 2> $OUT/$2.err; tail -c 600 $OUT/$2 | head -c 600; echo; python - $OUT/$2 <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["dtype"], "MLUPS %.0f"%d["value"], "ms %.4f"%d["ms_per_step"], "frac %.3f"%d["frac_of_roofline"], "e2e", d["e2e"] and round(d["e2e"]["value"]))
except Exception as e: print("unreadable", e)
PY
}
run 29521 scale_${TAG}_d3q19_lid_512_n8.json --steps 200 --warmup 10
run 29522 scale_${TAG}_d3q19_lid_512_f32_n8.json --steps 200 --warmup 10 --dtype float32 --compute float32
run 29523 weak_${TAG}_d3q27_channel_1024cubed_n8.json --steps 50 --warmup 5 --workload d3q27_channel_weak --no-e2e
