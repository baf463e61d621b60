"""In-graph timeline of ONE MR-MT3 decode step at a small lane count (the regime of 8-GPU configs[3]):
N tracks x 1 segment as ONE lane group, per kernel: begin of CTA 0, dependency wait returned, operands
in shared memory, end of the last CTA.  Prints a JSON summary per kernel class."""
import importlib, sys, json, torch
sys.path.insert(0, '.')
syn = importlib.import_module("mr-mt3_b200.synthetic"); t5 = importlib.import_module("mr-mt3_b200.t5")
v2 = importlib.import_module("mr-mt3_b200.t5_segmem_v2_with_prev")
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
m = v2.T5SegMemV2WithPrev(t5.T5Config(), 1, 64); m.load_state_dict(syn.synthetic_state_dict(4322, segmem=True)); m = m.eval().cuda()
eng = m.engine(); eng.set_option("group_lanes", 0)
x = syn.synthetic_features(3, N).cuda()
eng.trace_enable(True)
eng.generate_segmem(x, [1] * N, max_length=T)
tr = eng.trace_read(512)
eng.trace_enable(False)
names = ["embed"]
for l in range(8): names += [f"L{l}.qkv", f"L{l}.self", f"L{l}.o", f"L{l}.cq", f"L{l}.cross", f"L{l}.co", f"L{l}.wi", f"L{l}.wff"]
names += ["lm_head", "argmax"]
t0 = min(b for b, e in tr[:len(names)] if b)       # fused greedy head: the embed / arg-max slots stay empty
cls = {}
prev_end = None
rows = []
for i, nm in enumerate(names):
    b, e = tr[i]
    if not b:
        continue
    w, ld = tr[i + 128]
    k = nm.split(".")[-1]
    d = {"name": nm, "begin_us": (b - t0) / 1e3, "end_us": (e - t0) / 1e3, "dur_us": (e - b) / 1e3}
    if prev_end is not None:
        d["gap_after_prev_end_us"] = (b - prev_end) / 1e3
        if w: d["wait_return_after_prev_end_us"] = (w - prev_end) / 1e3
        if w and ld: d["loaded_after_wait_us"] = (ld - w) / 1e3; d["end_after_loaded_us"] = (e - ld) / 1e3
        k0, kh = tr[i + 256]
        if w and k0: d["mark2_after_wait_us"] = (k0 - w) / 1e3
        if w and kh: d["mark3_after_wait_us"] = (kh - w) / 1e3
        c = cls.setdefault(k, {"n": 0, "span_us": 0.0})
        c["n"] += 1; c["span_us"] += (e - prev_end) / 1e3      # contribution to the critical path
    prev_end = e
    rows.append(d)
step_us = (max(e for b, e in tr[:len(names)]) - t0) / 1e3
print(json.dumps({"lanes": N, "position": T - 1, "step_us": step_us,
                  "critical_path_by_class_us": {k: round(v["span_us"], 2) for k, v in cls.items()},
                  "per_kernel_avg_us": {k: round(v["span_us"] / v["n"], 2) for k, v in cls.items()},
                  "layer3": [r for r in rows if r["name"].startswith("L3.")]}))
