set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
( time timeout 1200 ncu --set full --clock-control none --import-source on -k regex:attn_decode_kernel -s 8000 -c 4 -o gpurun_out/attn_decode_r1 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile > gpurun_out/ncu_attn.log 2>&1 ) 2>&1 | tail -3
tail -3 gpurun_out/ncu_attn.log
ls -la gpurun_out/
for i in 1 2; do python - <<'PY'
import importlib, torch, hashlib, sys
sys.path.insert(0,'.')
syn = importlib.import_module("mr-mt3_b200.synthetic"); t5 = importlib.import_module("mr-mt3_b200.t5")
m = t5.T5ForConditionalGeneration(t5.T5Config()); m.load_state_dict(syn.synthetic_state_dict(1234)); m = m.eval().cuda()
x = syn.synthetic_features(7, 4).cuda()
e = m.encode(x); print("enc sha", hashlib.sha1(e.cpu().numpy().tobytes()).hexdigest()[:16])
ids = m.generate(x, max_length=64); print("ids sha", hashlib.sha1(ids.cpu().numpy().tobytes()).hexdigest()[:16])
PY
done
