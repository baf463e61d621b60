"""Timeline of one decode step as it runs inside the captured graph (globaltimer stamps)."""
import importlib, sys, json, torch
sys.path.insert(0, '.')
syn = importlib.import_module("mr-mt3_b200.synthetic"); t5 = importlib.import_module("mr-mt3_b200.t5")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = int(sys.argv[2]) if len(sys.argv) > 2 else 512
m = t5.T5ForConditionalGeneration(t5.T5Config()); m.load_state_dict(syn.synthetic_state_dict(1234)); m = m.eval().cuda()
eng = m.engine(); eng.set_option("group_lanes", 0)
x = syn.synthetic_features(3, B).cuda()
eng.trace_enable(True)
m.generate(x, max_length=T)
tr = eng.trace_read(67)
eng.trace_enable(False)
names = ["embed"]
for l in range(8): names += [f"L{l}.qkv", f"L{l}.self", f"L{l}.o", f"L{l}.cq", f"L{l}.cross", f"L{l}.co", f"L{l}.wi", f"L{l}.wff"]
names += ["lm_head", "argmax"]
t0 = tr[0][0]
rows = []
for i, (b, e) in enumerate(tr):
    nxt = tr[i + 1][0] if i + 1 < len(tr) else None
    rows.append({"k": names[i], "begin_us": (b - t0) / 1e3, "dur_us": (e - b) / 1e3, "gap_to_next_us": None if nxt is None else (nxt - e) / 1e3})
tot = (tr[-1][1] - t0) / 1e3
agg = {}
for r in rows:
    key = r["k"].split(".")[-1]
    a = agg.setdefault(key, [0.0, 0.0, 0]); a[0] += r["dur_us"]; a[1] += (r["gap_to_next_us"] or 0.0); a[2] += 1
print(f"B={B} position={T-1}: step = {tot:.1f} us, sum of kernel durations {sum(r['dur_us'] for r in rows):.1f} us, sum of gaps {sum((r['gap_to_next_us'] or 0) for r in rows):.1f} us")
for k, (d, g, n) in agg.items():
    print(f"  {k:8s} n={n:2d} dur/launch {d/n:7.2f} us  gap-after/launch {g/n:6.2f} us  total {d:7.1f} us")
for r in rows[:20]:
    print("   ", r)
json.dump(rows, open(f"gpurun_out/trace_B{B}_T{T}.json", "w"))
