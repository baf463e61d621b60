# Round 2, GPU call 32: last check of the shipped build -- whole GPU suite and smoke.
set -x
O=gpurun_out/r3j; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
( time timeout 1200 python -m pytest tests/ -q -m gpu 2>&1 | tail -3 ) > $O/pytest.log 2>&1; tail -6 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke.log
