#!/usr/bin/env bash
# GPU batch: full GPU test suite, headline bench through the plugin API, fp32-storage bench + ncu capture
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log | tail -6
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -c 600 gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json | cut -c1-3000
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_c4_reference.json 2>/dev/null; cut -c1-400 gpurun_out/bench_c4_reference.json
python bench.py --dtype float32 --no-also --no-cpu-baseline --steps 50 > gpurun_out/bench_c4_f32storage.json 2>/dev/null; cut -c1-700 gpurun_out/bench_c4_f32storage.json
ncu --set full --clock-control none --import-source on -k regex:lbmk_kernel_one_time_step -s 4 -c 1 -f -o gpurun_out/ncu_f32storage \
    python bench.py --dtype float32 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-also --api b200 > gpurun_out/ncu_f32storage.log 2>&1
ls -la gpurun_out | tail -12
