"""Timing of cc_rectify_f32c1_views: 64 1080p frames, 64 different views, one call (the reference's plot loop)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import json, subprocess
out = subprocess.run([sys.executable, "bench.py", "--steps", "3", "--warmup", "3", "--no-cpu"], capture_output=True, text=True).stdout
d = json.loads(out)
for k, v in d["extras"].items():
    if "views" in k:
        print(k, v)
