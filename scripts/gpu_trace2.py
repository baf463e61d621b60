"""Single-chain decode-step timeline with intermediate stamps: per kernel, when CTA 0 started, when
its dependency wait returned, when its operands were in shared memory, and when the last CTA ended."""
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
tr = eng.trace_read(512)
eng.trace_enable(False)
names = ["embed"]
for l in range(8): names += [f"L{l}.qkv", f"L{l}.self", f"L{l}.o", f"L{l}.cq", f"L{l}.cross", f"L{l}.co", f"L{l}.wi", f"L{l}.wff"]
names += ["lm_head", "argmax"]
t0 = tr[0][0]
prev_end = None
print(f"B={B} position={T-1}; times in us relative to the step start; d_* relative to the previous kernel's end")
for i, nm in enumerate(names):
    b, e = tr[i]
    w, ld = tr[i + 128]
    k0, kh = tr[i + 256]
    f = lambda v: "   -  " if v == 0 else f"{(v - t0) / 1e3:7.2f}"
    extra = ""
    if prev_end is not None and w:
        extra = f"  wait_return-prev_end {(w - prev_end) / 1e3:5.2f}" + (f"  loaded-wait {(ld - w) / 1e3:5.2f}  end-loaded {(e - ld) / 1e3:5.2f}" if ld else f"  end-wait {(e - w) / 1e3:5.2f}")
    if k0: extra += f"  | first k-tile {(k0 - w) / 1e3:5.2f}  half {(kh - w) / 1e3:5.2f}  last {(ld - w) / 1e3:5.2f} after the wait"
    if 16 < i < 34 or i > 63: print(f"{nm:9s} begin {f(b)} wait {f(w)} loaded {f(ld)} end {f(e)}{extra}")
    prev_end = e
