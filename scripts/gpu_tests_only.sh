python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -12
