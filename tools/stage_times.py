"""Scratch: print per-stage device times of the 1080p pipeline step (CUDA events), no CPU baseline."""
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py", "--steps", "5", "--warmup", "3", "--skip-e2e", "--no-cpu-baseline"] + sys.argv[1:],
                     capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    print("value %.0f faces/s  ms/step %.2f  stages %s" % (d["value"], d["ms_per_step"], d["stage_ms"]))
except Exception:
    print(out.stdout[-2000:], out.stderr[-3000:])
