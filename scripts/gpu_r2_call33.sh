# Round 2, GPU call 33: the dataset's batched GPU frontend against the per-row oracle.
set -x
O=gpurun_out/r3k; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_dataset_gpu.py -x -q -m gpu 2>&1 | tail -15 > $O/pytest.txt; cat $O/pytest.txt
