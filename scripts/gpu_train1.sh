python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_train_gpu.py -x -q -s 2>&1 | tail -30
