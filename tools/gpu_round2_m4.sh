#!/usr/bin/env bash
# 2-GPU validation of the final code: multi-GPU tests + bench with parity block (peer halo)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu_multi_n2.log; tail -3 gpurun_out/pytest_gpu_multi_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 50 --no-also > gpurun_out/n2_peer_final.json 2> gpurun_out/n2_peer_final.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/n2_peer_final.json") if l.startswith("{")][-1])
print("n2 peer final: %.1f MLUPS %.4f ms/step frac %.3f parity %s e2e %.0f" % (d["value"], d["ms_per_step"], d["frac_of_roofline"], d["parity"]["ok"], d["e2e"]["value"]))
PY
