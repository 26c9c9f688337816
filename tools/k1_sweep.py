"""K1 experiments (run on the GPU box): LANE_K1_EXPT bit0 = skip histogram, bit1 = skip the store (timing only)."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for expt in (0, 1, 2, 3):
    env = dict(os.environ, LANE_K1_EXPT=str(expt))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "4", "--warmup", "3", "--skip-cpu",
                          "--skip-e2e"], env=env, capture_output=True, text=True).stdout.strip().splitlines()
    d = json.loads(out[-1])
    print(f"expt={expt} k1={d['stage_ms_per_step']['blur_hist']:.4f} ms  frac={d['roofline']['frac']:.3f}", flush=True)
