# Round 2, GPU call 22 (2 GPUs): the N > 1 bench paths the driver will run, on the final build.
set -x
O=gpurun_out/r2z; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 1 2> $O/bench_slakh_n2.err | tail -1 > $O/bench_slakh_n2.json; cut -c1-300 $O/bench_slakh_n2.json; tail -2 $O/bench_slakh_n2.err
timeout 600 $TR bench.py --gpus 2 --workload finetune --steps 10 --warmup 3 2> $O/bench_ft_n2.err | tail -1 > $O/bench_finetune_n2.json
python -c "import json; d=json.load(open('$O/bench_finetune_n2.json')); print('finetune n2', d['ms_per_step'], d['training'], d['clocks'])"; tail -2 $O/bench_ft_n2.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 > $O/bench_ref_n2.json; cut -c1-200 $O/bench_ref_n2.json
ls $O
