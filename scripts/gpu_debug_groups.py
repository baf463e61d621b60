import importlib, sys, torch, numpy as np
sys.path.insert(0, '.')
syn = importlib.import_module("mr-mt3_b200.synthetic"); t5 = importlib.import_module("mr-mt3_b200.t5")
m = t5.T5ForConditionalGeneration(t5.T5Config()); m.load_state_dict(syn.synthetic_state_dict(1239, eos_scale=5.0)); m = m.eval().cuda()
eng = m.engine()
x = syn.synthetic_features(5, 40).cuda()
def run(gl, graphs=1, serial=0, xs=x):
    eng.set_option("group_lanes", gl); eng.set_option("use_graphs", graphs); eng.set_option("group_serial", serial)
    return eng.generate(xs, max_length=96).cpu().numpy()
one = run(0)
def cmp(tag, many, ref=one):
    n = min(many.shape[1], ref.shape[1])
    d = np.argwhere(many[:, :n] != ref[:, :n]); rows = sorted(set(d[:, 0].tolist()))
    print(tag, "shape", many.shape, "ndiff", len(d), "rows", rows[:10], "first cols", [int(d[d[:,0]==r][0,1]) for r in rows[:10]])
cmp("G1 eager", run(0, graphs=0))
cmp("gl3 graph conc", run(3))
cmp("gl3 graph serial", run(3, serial=1))
cmp("gl3 eager conc", run(3, graphs=0))
cmp("gl3 eager serial", run(3, graphs=0, serial=1))
# direct: row 17 alone and rows 15..17 as their own batch
cmp("rows15-17 as batch (G1)", run(0, xs=x[15:18]), one[15:18])
cmp("row17 alone (G1)", run(0, xs=x[17:18]), one[17:18])
cmp("rows 3-5 as batch", run(0, xs=x[3:6]), one[3:6])
