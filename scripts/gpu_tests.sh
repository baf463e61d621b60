set -x
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests/ -x -q -s -m gpu 2>&1 | tail -40
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
