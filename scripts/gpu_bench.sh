set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench_ref.json
# launch list of a shortened run (128 tokens) to learn ncu's per-launch overhead
( time timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_t128.csv python bench.py --steps 1 --warmup 0 --max-length 128 --no-cpu-baseline --no-profile > gpurun_out/ncu_t128.log 2>&1 ) 2>&1 | tail -3
wc -l gpurun_out/launches_t128.csv
