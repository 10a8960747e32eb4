#!/usr/bin/env bash
# the driver's scaling command on N GPUs with the final code (peer halo, parity block, e2e, also)
mkdir -p gpurun_out
N=${N:-4}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_final_n$N.json 2> gpurun_out/scale_final_n$N.err
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/scale_final_n%s.json" % n) if l.startswith("{")][-1])
    print("N=%s: %.1f MLUPS %.4f ms/step frac %.3f parity %s e2e %.0f launches/step %.1f" % (
        n, d["value"], d["ms_per_step"], d["frac_of_roofline"], d["parity"] and (d["parity"]["ok"], d["parity"]["max_rel_err"]),
        d["e2e"]["value"], d["gpu_launches"] / d["steps"]))
    for a in d.get("also") or []:
        print("   also:", json.dumps(a)[:300])
except Exception as exc:
    print("FAILED", exc, open("gpurun_out/scale_final_n%s.err" % n).read()[-1500:])
PY
