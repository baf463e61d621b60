# Round 2, GPU call 34: T5SegMem (V1) teacher-forced forward against the reference's golden logits.
set -x
O=gpurun_out/r3l; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -s -k "v1 or teacher_forced or segmem" 2>&1 | tail -25 > $O/pytest.txt; cat $O/pytest.txt
