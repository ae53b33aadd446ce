#!/usr/bin/env python
"""Batches in flight (IZPI_RENDER_SLOTS build variants) x batch size, on render configs; scene built once per config."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
runs = [("", 26), ("variants/lib_slots3.so", 26), ("variants/lib_slots3.so", 25), ("variants/lib_slots4.so", 25), ("variants/lib_slots4.so", 24), ("", 26)]
for cfg, spp in (("3", 256), ("5", 64), ("4", 128)):
    for lib, bp in runs:
        env = dict(os.environ, IZPI_LIB_PATH=lib, IZPI_BATCH_PATHS=str(1 << (bp if cfg != "4" else bp - 2)))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "render_one.py"), "--config", cfg, "--spp", str(spp), "--repeat", "2"], env=env, capture_output=True, text=True).stdout.strip().split("\n")[-1]
        try:
            d = json.loads(out)
            print(json.dumps({"config": cfg, "lib": lib or "slots2", "batch_log2": bp if cfg != "4" else bp - 2, "best_msamples_per_s": d["best_msamples_per_s"]}), flush=True)
        except Exception:
            print("failed", cfg, lib, out[-300:], flush=True)
