# Round 2, GPU call 5 (2 GPUs): the N > 1 bench paths the driver will run, and the overlapped gradient all-reduce.
set -x
O=gpurun_out/r2e; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 2 --warmup 1 2> $O/bench_slakh_n2.err | tail -1 > $O/bench_slakh_n2.json; cut -c1-400 $O/bench_slakh_n2.json; tail -3 $O/bench_slakh_n2.err
timeout 600 $TR bench.py --gpus 2 --workload finetune --steps 5 --warmup 2 2> $O/bench_ft_n2.err | tail -1 > $O/bench_finetune_n2.json; cut -c1-400 $O/bench_finetune_n2.json; tail -3 $O/bench_ft_n2.err
timeout 600 $TR bench.py --gpus 2 --workload mt3_256 --steps 3 --warmup 2 --no-profile 2> $O/bench_mt3_n2.err | tail -1 > $O/bench_mt3_n2.json; cut -c1-300 $O/bench_mt3_n2.json
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>/dev/null | tail -1 > $O/bench_ref_n2.json; cut -c1-300 $O/bench_ref_n2.json
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/ -q -m gpu -s 2>&1 | tail -100 > $O/pytest.log; tail -4 $O/pytest.log
for v in 1 2; do CUDA_VISIBLE_DEVICES=0 MRMT3_FRONTEND_VARIANT=$v timeout 120 python scripts/gpu_frontend_bench.py 256 2>&1 | tail -1 >> $O/frontend_bench.jsonl; done; cat $O/frontend_bench.jsonl
for cfg in "S4:MRMT3_ATTN_STAGES=4" "S6:MRMT3_ATTN_STAGES=6" "S8:MRMT3_ATTN_STAGES=8"; do
  tag=${cfg%%:*}; envs=${cfg#*:}
  for lanes in 16 64; do
    r=$(CUDA_VISIBLE_DEVICES=0 env $envs timeout 120 python scripts/gpu_config3.py $lanes 2 1024 2>&1 | tail -1)
    echo "{\"lanes\": $lanes, \"cfg\": \"$tag\", \"r\": $r}" >> $O/ab_stages.jsonl
  done
done
cat $O/ab_stages.jsonl | cut -c1-200
CUDA_VISIBLE_DEVICES=0 MRMT3_GROUP_LANES=0 timeout 120 python scripts/gpu_trace_segmem.py 16 512 2>&1 | tail -1 > $O/trace_segmem_16.json
ls $O
