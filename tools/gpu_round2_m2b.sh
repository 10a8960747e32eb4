#!/usr/bin/env bash
mkdir -p gpurun_out
N=${N:-2}
bench() {  # name, env, args
    name=$1; shift; envs=$1; shift
    env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/%s.json" % name) if l.startswith("{")][-1])
    print("%-22s %10.1f MLUPS  %8.4f ms/step  fused %8.4f  frac %.3f  parity %s  e2e %s  launches/step %.1f" % (
        name, d["value"], d["ms_per_step"], d["roofline"]["launch_ms"], d["frac_of_roofline"],
        d["parity"] and (d["parity"]["ok"], d["parity"]["max_rel_err"]), d["e2e"] and round(d["e2e"]["value"]),
        d["gpu_launches"] / d["steps"]))
    for a in d.get("also") or []:
        print("    also:", json.dumps(a)[:420])
except Exception as exc:
    print(name, "FAILED", exc, open("gpurun_out/%s.err" % name).read()[-1500:])
PY
}
if [ "$N" = "2" ]; then
bench n2_nccl X=1 --steps 50 --halo nccl --no-e2e --no-also --no-parity
else
bench n${N}_peer X=1 --steps 100
bench n${N}_nccl X=1 --steps 100 --halo nccl --no-e2e --no-also
bench n${N}_nccl_noverlap PYLBM_B200_NO_OVERLAP=1 --steps 100 --halo nccl --no-e2e --no-also --no-parity
fi
