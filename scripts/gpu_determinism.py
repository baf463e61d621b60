"""Run the full-size batch twice per configuration and report where the token rows differ."""
import importlib, sys, torch
sys.path.insert(0, '.')
syn = importlib.import_module("mr-mt3_b200.synthetic"); t5 = importlib.import_module("mr-mt3_b200.t5")
m = t5.T5ForConditionalGeneration(t5.T5Config()); m.load_state_dict(syn.synthetic_state_dict(1234, eos_scale=5.0)); m = m.eval().cuda()
eng = m.engine()
x = syn.synthetic_features(3, 256).cuda()
ref = None
for variant in (0, 1):
    for gl in (0, 32):
        for graphs in (1, 0):
            eng.set_option("attn_variant", variant); eng.set_option("group_lanes", gl); eng.set_option("use_graphs", graphs)
            outs = [m.generate(x, max_length=1024).cpu() for _ in range(3)]
            if ref is None: ref = outs[0]
            for i, o in enumerate(outs):
                n = min(o.shape[1], ref.shape[1])
                d = (o[:, :n] != ref[:, :n])
                rows = d.any(1).nonzero().flatten().tolist()
                first = [int(d[r].nonzero()[0]) for r in rows[:6]]
                print(f"variant {variant} group_lanes {gl} graphs {graphs} run {i}: shape {tuple(o.shape)} rows differing from ref: {len(rows)} {rows[:6]} first cols {first}", flush=True)
