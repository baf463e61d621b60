"""BASELINE configs[2] shape at reduced length: MR-MT3 (V2WithPrev) segment-memory transcription of
N tracks x S segments, max_length 1024 (synthetic weights never emit EOS => 1024 tokens/segment).
Reports audio-s/s and ms per segment round (memory block + cross-K/V + 1024 decode steps)."""
import importlib, sys, time, json, torch
sys.path.insert(0, '.')
syn = importlib.import_module("mr-mt3_b200.synthetic"); t5 = importlib.import_module("mr-mt3_b200.t5")
v2 = importlib.import_module("mr-mt3_b200.t5_segmem_v2_with_prev")
n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_seg = int(sys.argv[2]) if len(sys.argv) > 2 else 6
L = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
m = v2.T5SegMemV2WithPrev(t5.T5Config(), 1, 64); m.load_state_dict(syn.synthetic_state_dict(4322, segmem=True)); m = m.eval().cuda()
eng = m.engine()
x = syn.synthetic_features(3, n_tracks * n_seg).cuda()
counts = [n_seg] * n_tracks
out = None
for it in range(3):
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count
    e0.record(); out = eng.generate_segmem(x, seg_counts=counts, max_length=L); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"tracks": n_tracks, "segments_per_track": n_seg, "max_length": L, "ms": round(ms, 1),
                      "ms_per_round": round(ms / n_seg, 2), "us_per_decode_step": round(ms * 1e3 / n_seg / L, 1),
                      "audio_s_per_s": round(n_tracks * n_seg * 2.048 / (ms / 1e3), 1), "launches": eng.launch_count - l0}), flush=True)
