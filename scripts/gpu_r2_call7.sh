# Round 2, GPU call 7: software-pipelined attention chunk loop (bit-identical arithmetic).
set -x
O=gpurun_out/r2g; mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_parity_size_gpu.py -q -m gpu 2>&1 | tail -5 > $O/pytest.log; tail -3 $O/pytest.log
for lanes in 8 16 64; do
  r=$(timeout 120 python scripts/gpu_config3.py $lanes 2 1024 2>&1 | tail -1)
  echo "{\"lanes\": $lanes, \"cfg\": \"pipelined\", \"r\": $r}" >> $O/ab_small.jsonl
done
cut -c1-200 $O/ab_small.jsonl
MRMT3_GROUP_LANES=0 timeout 120 python scripts/gpu_trace_segmem.py 16 512 2>&1 | tail -1 > $O/trace_segmem_16.json
python -c "
import json; d=json.load(open('$O/trace_segmem_16.json')); print(d['step_us'], d['per_kernel_avg_us']); [print(r['name'], {k:round(v,2) for k,v in r.items() if k not in ('name','begin_us','end_us')}) for r in d['layer3'] if 'self' in r['name'] or 'cross' in r['name']]"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 > $O/bench_mt3.json
python -c "import json; d=json.load(open('$O/bench_mt3.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['other'], d['roofline']['decode_loop']['frac_of_peak_timed_region'])"
ls $O
