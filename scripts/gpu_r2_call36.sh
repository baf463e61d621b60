# Round 2, GPU call 36: ncu --set full of the three whole-sequence attention kernels inside a fine-tune step
# (decoder self-attention launches), for the issue-slot / tensor-pipe record of the final kernels.
set -x
O=gpurun_out/r3n; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 60 -c 9 \
    -o $O/train_attn_final -f python scripts/gpu_train_bench.py 32 1024 1 0.1 > $O/ncu_train_attn.log 2>&1
ls -la $O; tail -2 $O/ncu_train_attn.log
