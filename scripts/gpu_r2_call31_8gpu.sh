# Round 2, GPU call 31 (8 GPUs): the fine-tune step on the final kernels, gradient all-reduce overlapped.
set -x
O=gpurun_out/r3i; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517"
timeout 500 $TR bench.py --gpus 8 --workload finetune --steps 20 --warmup 5 2> $O/bench_ft_n8.err | tail -1 > $O/bench_finetune_n8.json
python -c "import json; d=json.load(open('$O/bench_finetune_n8.json')); print('finetune n8', d['ms_per_step'], d['training'], d['clocks'])"; tail -2 $O/bench_ft_n8.err
