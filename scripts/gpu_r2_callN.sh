# Round 2: the N-GPU bench paths exactly as the driver launches them (N = $1).
set -x
N=$1
O=gpurun_out/r2n$N; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 1 2> $O/bench_slakh.err | tail -1 > $O/bench_slakh_n$N.json; cut -c1-300 $O/bench_slakh_n$N.json; tail -2 $O/bench_slakh.err
timeout 400 $TR bench.py --gpus $N --workload finetune --steps 8 --warmup 3 2> $O/bench_ft.err | tail -1 > $O/bench_finetune_n$N.json; cut -c1-300 $O/bench_finetune_n$N.json; tail -2 $O/bench_ft.err
ls $O
