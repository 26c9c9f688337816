"""K1 A/B on the GPU box: CTAs/SM (LANE_K1_MINB) x arithmetic variant (LANE_K1_VAR); parity for each via the K1 tests."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for minb in ("5", "6"):
    for var in ("0", "1", "2", "3"):
        env = dict(os.environ, LANE_K1_MINB=minb, LANE_K1_VAR=var)
        t = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "-x", "-q", "-k",
                            "blur or stage or k1"], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "20", "--warmup", "3", "--skip-cpu",
                              "--skip-e2e"], env=env, capture_output=True, text=True).stdout.strip().splitlines()
        d = json.loads(out[-1])
        print(f"minb={minb} var={var}: k1={d['stage_ms_per_step']['blur_hist']:.4f} ms frac={d['roofline']['frac']:.3f} | {t}", flush=True)
