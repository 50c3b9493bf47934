"""GPU-side diagnostic: run each parity area separately and write a JSON/text report under gpurun_out/."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
t0 = time.time()
r = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu", "-q", "--no-header", "-x" if "-x" in sys.argv else "-q",
                    "--tb=short", "-p", "no:cacheprovider"], cwd=ROOT, capture_output=True, text=True)
open(os.path.join(ROOT, "gpurun_out", "pytest_gpu.log"), "w").write(r.stdout[-60000:] + "\n--- stderr\n" + r.stderr[-5000:])
print(r.stdout[-6000:])
print("elapsed %.1fs rc=%d" % (time.time() - t0, r.returncode))
