python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python bench.py 2> gpurun_out/bench_r1c.err | tail -1 > gpurun_out/bench_r1c.json
python -c "import json; d=json.load(open('gpurun_out/bench_r1c.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'], d['clocks'], d.get('cpu_baseline',{}).get('value'))"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 2> gpurun_out/bench_2gpu.err | tail -1 > gpurun_out/bench_2gpu.json
tail -3 gpurun_out/bench_2gpu.err; cut -c1-700 gpurun_out/bench_2gpu.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-300
MRMT3_GROUP_LANES=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_decode_mma -s 8000 -c 2 -o gpurun_out/attn_mma_r1b -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-profile > gpurun_out/ncu_attn_mma.log 2>&1
tail -1 gpurun_out/ncu_attn_mma.log | cut -c1-200
