"""Host-buffer closest hit (izpi_trace_closest, pinned buffers, copies inside the timed region) on the config-2 batch; used to
choose the slice schedule (IZPI_TRACE_SLICE_LOG2=20..23: 555 / 579 / 576 / 546 Mrays/s -> 2^21 is the default)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from izpi_b200 import cuda, scenes
sc, lo, hi = scenes.closest_hit_scene()
ctx = cuda.Context(0); ctx.upload(cuda.HostScene(sc))
n = 1 << 24
org, d = scenes.random_rays(n, lo, hi)
h_org = torch.from_numpy(org).pin_memory().numpy(); h_dir = torch.from_numpy(d).pin_memory().numpy()
ids = torch.empty(n, dtype=torch.int32).pin_memory().numpy(); t = torch.empty(n, dtype=torch.float64).pin_memory().numpy()
for _ in range(2): ctx.trace_closest(h_org, h_dir, out_ids=ids, out_t=t)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); ctx.trace_closest(h_org, h_dir, out_ids=ids, out_t=t); ts.append(time.perf_counter() - t0)
print(json.dumps({"slice_log2": os.environ.get("IZPI_TRACE_SLICE_LOG2", "21"), "ms_best": min(ts) * 1e3, "mrays_per_s": n / min(ts) / 1e6}))
ctx.close()
