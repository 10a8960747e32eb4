#!/usr/bin/env bash
# 2-GPU batch: multi-GPU tests, bench with the parity block (peer and NCCL halos)
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_multi_n2.log
tail -5 gpurun_out/pytest_gpu_multi_n2.log
fi
bench() {  # name, env, args
    name=$1; shift; envs=$1; shift
    env $envs python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus 2 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
    python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads([l for l in open("gpurun_out/%s.json" % name) if l.startswith("{")][-1])
    print("%-22s %10.1f MLUPS  %8.4f ms/step  frac %.3f  parity %s  e2e %s" % (
        name, d["value"], d["ms_per_step"], d["frac_of_roofline"],
        d["parity"] and (d["parity"]["ok"], d["parity"]["max_rel_err"]), d["e2e"] and round(d["e2e"]["value"])))
    for a in d.get("also") or []:
        print("    also:", json.dumps(a)[:300])
except Exception as exc:
    print(name, "FAILED", exc, open("gpurun_out/%s.err" % name).read()[-1500:])
PY
}
bench n2_peer X=1 --steps 50
bench n2_nccl X=1 --steps 50 --halo nccl --no-e2e --no-also
bench n2_nccl_noverlap PYLBM_B200_NO_OVERLAP=1 --steps 50 --halo nccl --no-e2e --no-also --no-parity
