"""Each debug switch set of the probe lifting kernel in its own process (a faulting variant must not poison the rest)."""
import os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
for bits in (0, 1, 2, 4, 8, 16, 30, 31):
    r = subprocess.run([sys.executable, os.path.join(here, "gpu_lift_tc_timeline.py"), "tc16", str(bits)], capture_output=True, text=True, timeout=120)
    print(r.stdout.strip().splitlines()[0] if r.stdout.strip() else "", "|", r.stderr.strip()[-300:].replace("\n", " "), flush=True)
