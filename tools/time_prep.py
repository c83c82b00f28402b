#!/usr/bin/env python
"""Op-level timing with the staged kernel excluded is not possible from Python; instead time the whole op at the
three stages for a few MDF_PREP_Z settings (the staged kernel's time is constant across them)."""
import os, subprocess, sys
for z in ("1", "2", "4", "8"):
    env = dict(os.environ, MDF_PREP_Z=z)
    out = subprocess.run([sys.executable, "tools/tune_staged.py"], env=env, capture_output=True, text=True).stdout
    print("z =", z, [l.split(":")[1].split("us")[0].strip() for l in out.splitlines() if "variant 0" in l])
