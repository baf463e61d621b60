set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
MRMT3_NO_GRAPH=1 timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -s -k "logmel or compute_spectrogram or encoder_states or teacher_forced" 2>&1 | tail -40
