"""dev tool: sweep PYLBM_B200_MINBLOCKS for a workload (each value = its own kernel library)."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = sys.argv[1] if len(sys.argv) > 1 else "d3q19_lid_256"
extra = os.environ.get("TUNE_ARGS", "").split()     # e.g. TUNE_ARGS="--dtype float32 --compute float32"
for mb in (sys.argv[2:] or ["1", "4", "5", "6", "8"]):
    env = dict(os.environ, PYLBM_B200_MINBLOCKS=mb)
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", wl, "--steps", "30", "--warmup", "5",
                          "--no-e2e", "--no-cpu-baseline"] + extra, env=env, capture_output=True, text=True)
    import json
    try:
        r = json.loads(out.stdout.strip().splitlines()[-1])
        print(wl, "minblocks", mb, "ms/step %.4f" % r["ms_per_step"], "MLUPS %.0f" % r["value"], "kernel ms %.4f" % r["roofline"]["launch_ms"], "frac %.3f" % r["roofline"]["frac"], flush=True)
    except Exception as e:
        print(wl, mb, "failed", out.stderr[-500:])
