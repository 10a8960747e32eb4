#!/usr/bin/env bash
# packed NCCL halo: 2-GPU tests of the NCCL path + bench; N from the environment
mkdir -p gpurun_out
N=${N:-2}
if [ "$N" = "2" ]; then
python -m pytest tests/test_gpu_multi.py -q -x -k "nccl or demos" 2>&1 | tail -6
fi
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
        d["parity"] and (d["parity"]["ok"], d["parity"]["max_rel_err"]), d["e2e"] and (round(d["e2e"]["value"]), round(d["e2e"]["h2d_gbs_rank0"], 1)),
        d["gpu_launches"] / d["steps"]))
except Exception as exc:
    print(name, "FAILED", exc, open("gpurun_out/%s.err" % name).read()[-1500:])
PY
}
bench n${N}_nccl_packed X=1 --steps 100 --halo nccl --no-e2e --no-also
if [ "$N" = "8" ]; then
bench n8_peer_20steps X=1 --steps 20 --no-also --no-parity
bench n8_peer_20steps_nonuma PYLBM_B200_NO_NUMA=1 --steps 20 --no-also --no-parity
echo "== NUMA probe"
nvidia-smi topo -m 2>&1 | head -24
for d in /sys/bus/pci/devices/*; do
  if [ -f $d/class ] && grep -q "^0x0302" $d/class 2>/dev/null; then echo "$(basename $d) numa_node=$(cat $d/numa_node)"; fi
done
ls /sys/devices/system/node/ | head; lscpu | grep -i "numa\|socket\|model name" | head
fi
