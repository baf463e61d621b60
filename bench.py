#!/usr/bin/env python
"""bench.py -- transcribed audio-seconds per second of the MR-MT3 hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload W]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads (BASELINE.json configs; `--workload`, default by N):
  mt3_256          configs[1], the N = 1 default and the headline: MT3, 256 synthetic 2.048 s segments
                   per GPU, log-mel + encoder + KV-cached greedy decode, max 1024 tokens.  N > 1 with
                   this workload = independent replicas of it (weak scaling).
  mrmt3_64x4min    configs[2]: MR-MT3 (T5SegMemV2WithPrev) on 64 four-minute tracks (118 segments each,
                   memory chained through every track, tracks batched across lanes).
  mrmt3_512_slakh  configs[3], the default for N > 1: 512 ragged Slakh-shaped tracks, LPT-sharded by
                   segment count over the N ranks (sharding.shard_tracks), every rank transcribes its
                   tracks with the cross-track batched MR-MT3 loop, token rows gathered on rank 0
                   (sharding.gather_token_rows) inside the timed region.  STRONG scaling: the total
                   work is fixed.  Track durations are multiplied by --duration-scale (default 1/16)
                   so that a pass takes seconds, not minutes: the lane count per GPU, the raggedness
                   (relative length distribution) and the decode length -- what determines the rate
                   -- are unchanged; --duration-scale 1 runs the full set.
  finetune         configs[4]: MR-MT3 fine-tune step, batch 32 per GPU, L = L_p = 1024, dropout 0.1,
                   forward + backward + gradient all-reduce (bucketed, overlapped with the backward)
                   + AdamW; audio-s/s = samples/s x 2.048.

The seeded synthetic weights never emit EOS (SURVEY 8d "decode length convention"): every segment
decodes T_dec = max_length tokens, the worst case.  One "step" = one pass of the whole path over the
workload.

  value : inputs (fp32 audio) already in HBM; log-mel -> generate, device-timed, max over ranks.
  e2e   : the same through the host-buffer C-ABI call (`mrmt3_transcribe_host`): pinned host audio
          in, host token rows out, H2D + D2H inside the timed region.
  roofline / cpu_baseline : DESIGN.md section 6.

`--impl reference` times the reference's CPU ALGORITHM (no KV cache, fp32, all host threads) through
the oracle port -- the reference's own modules cannot be imported on the GPU box.  Each step is one
bounded, really-executed sample; the quoted rate extrapolates it (see `cpu_reference_sample`).
"""
import argparse
import importlib
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT,):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

SEG_SAMPLES = 32768
SEG_SECONDS = SEG_SAMPLES / 16000.0
METRIC = "transcribed audio-sec/sec (log-mel + T5 encode + greedy decode)"
UNIT = "audio-s/s"
N_LAYERS, D_KV_BYTES = 8, 12288          # decoder layers; K+V bytes of one position over all layers (8*2*384*2)
WEIGHT_BYTES = 45.64e6                   # decoder + lm_head weights read once per decode step


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None,
                    choices=["mt3_256", "mrmt3_64x4min", "mrmt3_512_slakh", "finetune"])
    ap.add_argument("--segments", type=int, default=256, help="mt3_256: segments per GPU")
    ap.add_argument("--tracks", type=int, default=None, help="mrmt3_*: number of tracks (default 64 / 512)")
    ap.add_argument("--duration-scale", type=float, default=None,
                    help="mrmt3_*: multiply every track duration (default 1/16; 1 = the full set)")
    ap.add_argument("--batch", type=int, default=32, help="finetune: samples per GPU")
    ap.add_argument("--dropout", type=float, default=0.1, help="finetune: dropout rate (reference 0.1)")
    ap.add_argument("--max-length", type=int, default=1024)
    ap.add_argument("--eos-scale", type=float, default=1.0, help="scale of the EOS row of lm_head (early-exit runs)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    a = ap.parse_args()
    if a.workload is None:
        a.workload = "mt3_256" if a.gpus == 1 else "mrmt3_512_slakh"
    if a.duration_scale is None:
        a.duration_scale = 1.0 / 16.0
    return a


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def workload_config(args, world):
    """The `config` object: identical in the `ours` and `reference` arms."""
    T = args.max_length
    if args.workload == "mt3_256":
        return {"workload": f"mt3_256: MT3 batched inference, {args.segments} synthetic 2.048 s segments per GPU, log-mel + "
                            f"encoder + KV-cached greedy decode, T_dec = max_length = {T} (BASELINE.json configs[1])",
                "segments_per_gpu": args.segments, "max_length": T, "weights": "seeded synthetic (never emit EOS)"
                if args.eos_scale == 1.0 else f"seeded synthetic, EOS row of lm_head x{args.eos_scale}",
                "parallelism": f"independent replicas x{world}" if world > 1 else "1 GPU",
                "l2": "per-step working set (KV caches, ~4 GB) exceeds L2; no explicit flush"}
    if args.workload in ("mrmt3_64x4min", "mrmt3_512_slakh"):
        n_tracks = args.tracks or (64 if args.workload == "mrmt3_64x4min" else 512)
        what = ("64 four-minute tracks (BASELINE.json configs[2])" if args.workload == "mrmt3_64x4min" else
                "512 Slakh-shaped ragged tracks, durations clip(lognormal(mean 249 s, sigma 0.35), 60, 600) "
                "(BASELINE.json configs[3], SURVEY 8d)")
        return {"workload": f"{args.workload}: MR-MT3 (T5SegMemV2WithPrev, L_agg 64) segment-memory transcription of {what}, "
                            f"durations x{args.duration_scale:g}, T_dec = max_length = {T}",
                "tracks": n_tracks, "duration_scale": args.duration_scale, "max_length": T,
                "weights": "seeded synthetic (never emit EOS)",
                "parallelism": f"tracks LPT-sharded by segment count over {world} GPU(s), token rows gathered on rank 0",
                "l2": "per-step working set (KV caches) exceeds L2; no explicit flush"}
    return {"workload": f"finetune: MR-MT3 V2WithPrev fine-tune step, batch {args.batch}/GPU, L = L_p = 1024, dropout "
                        f"{args.dropout}, AdamW lr 1e-5 (BASELINE.json configs[4])",
            "batch_per_gpu": args.batch, "max_length": 1024, "weights": "seeded synthetic",
            "parallelism": f"data parallel x{world}, flat fp32 gradient all-reduced in buckets overlapped with the backward",
            "l2": "activations (7 GB) exceed L2; no explicit flush"}


# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (pynvml, 100 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown")
            else nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), \
            "measured (MEASURED_PEAKS.json hbm_gbs / bf16_tflops_sustained)"
    except Exception:
        return 6650.0, 1600.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kind):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/*_ncu_traffic.json, written from the .ncu-rep by scripts/ncu_summary.py)."""
    for name in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                d = json.load(f)[kind]
            return int(d["dram_bytes_per_launch"]), d["note"]
        except Exception:
            continue
    return None, "no ncu capture committed"


def synth_segments(n_seg, seed0):
    syn = importlib.import_module("mr-mt3_b200.synthetic")
    audio = np.stack([syn.synthetic_audio(seed=seed0 + i, n_samples=SEG_SAMPLES, n_tones=4) for i in range(n_seg)])
    return audio.astype(np.float32)


# ---------------------------------------------------------------------------------------------
# track tables of the MR-MT3 workloads (host-side framing exactly as inference.py:64-95)
def track_durations(args):
    syn = importlib.import_module("mr-mt3_b200.synthetic")
    if args.workload == "mrmt3_64x4min":
        n = args.tracks or 64
        dur = np.full(n, 240.0)
    else:
        n = args.tracks or 512
        dur = syn.slakh_shaped_durations(n, seed=0)
    return dur * args.duration_scale


def frame_track(n_samples):
    """-> (segments, valid frames of the last segment) for a track of n_samples (inference.py:64-95:
    pads at least one sample, a whole hop when already aligned; 256-frame segments)."""
    n_pad = n_samples + (128 - n_samples % 128)
    n_frames = n_pad // 128
    return math.ceil(n_frames / 256), n_frames


class TrackSet:
    """Audio + segment tables of a list of tracks, concatenated (what `mrmt3_transcribe_host` takes)."""

    def __init__(self, samples_per_track, pool, seed):
        rng = np.random.default_rng(seed)
        self.n_samples = [int(s) for s in samples_per_track]
        self.seg_counts, starts, lens, valid = [], [], [], []
        chunks, base = [], 0
        for n in self.n_samples:
            S, n_frames = frame_track(n)
            self.seg_counts.append(S)
            picks = rng.integers(0, len(pool), S)
            track = np.concatenate([pool[i] for i in picks])[:n] if n > 0 else np.zeros(0, np.float32)
            chunks.append(track)
            st = base + np.arange(S, dtype=np.int64) * SEG_SAMPLES
            starts.append(st)
            lens.append(np.clip(n - np.arange(S) * SEG_SAMPLES, 0, SEG_SAMPLES))     # per-segment STFT
            valid.append(np.clip(n_frames - np.arange(S) * 256, 0, 256))
            base += n
        self.audio = np.concatenate(chunks).astype(np.float32) if chunks else np.zeros(0, np.float32)
        self.seg_start = np.concatenate(starts).astype(np.int64)
        self.seg_len = np.concatenate(lens).astype(np.int32)
        self.valid = np.concatenate(valid).astype(np.int32)
        self.n_seg = int(sum(self.seg_counts))
        self.audio_seconds = float(sum(self.n_samples)) / 16000.0


def decode_bytes(lanes_per_round, T, Tk):
    """SURVEY 8(d) bytes of the decode loops: per round of T steps over `lanes` active lanes."""
    total = 0.0
    for lanes in lanes_per_round:
        total += T * WEIGHT_BYTES + lanes * (D_KV_BYTES * Tk * T + D_KV_BYTES * T * (T + 1) / 2 + (D_KV_BYTES + 4) * T)
    return total


def attn_algo_bytes(lanes, T, Tk):
    """Algorithmic bytes of all launches of the two decode attention kernels over one round."""
    return {
        # per launch (one layer, all lanes): K/V pages read + the step's K/V row written + the fused
        # q|k|v row read + the context row written
        "attn_self": lanes * N_LAYERS * (1536 * T * (T + 1) // 2 + T * (1536 + 2304 + 768)),
        # cross cache read + q row read + context row written
        "attn_cross": lanes * N_LAYERS * T * (1536 * Tk + 768 + 768),
    }


# ---------------------------------------------------------------------------------------------
# the reference's CPU algorithm on a bounded sample
def cpu_reference_sample(args, seed=0, anchor_steps=64):
    """One bounded, really-executed sample of the reference ALGORITHM (oracle port, fp32, all host
    threads) for ONE segment of the workload:
      * frontend + encoder (+ MR-MT3 memory block as written: all max_length query rows) in full;
      * the no-KV-cache greedy loop (models/t5.py:267-295 / t5_segmem_v2_with_prev.py:268-286 re-run the
        decoder over the whole prefix every step) REALLY RUN for the first `anchor_steps` steps;
      * the remaining steps costed from single full-prefix decoder passes at 8 prefix lengths up to
        max_length, integrated over the step index (a full run is ~40 s per segment on 16 cores,
        hours for the workload).
    -> dict(sec_per_segment extrapolated, wall_s of this sample, anchor, description)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mt3_oracle as O
    syn = importlib.import_module("mr-mt3_b200.synthetic")
    torch.set_num_threads(os.cpu_count())
    segmem = args.workload.startswith("mrmt3") or args.workload == "finetune"
    T = args.max_length
    t_wall0 = time.perf_counter()
    sd = O.cast_state_dict(syn.synthetic_state_dict(4322 if segmem else 1234, segmem=segmem), torch.float32)
    audio = syn.synthetic_audio(seed=seed, n_samples=32767, n_tones=4)
    with torch.no_grad():
        t0 = time.perf_counter()
        mel, _ = O.preprocess(audio, mel_norm=True, dtype=np.float32)
        t_front = time.perf_counter() - t0
        x = torch.from_numpy(mel.astype(np.float32))
        O.encode(x, sd)                                          # warm-up
        t0 = time.perf_counter()
        enc = O.encode(x, sd)
        t_enc = time.perf_counter() - t0
        t_mem = 0.0
        if segmem:
            ids0 = torch.zeros((1, T), dtype=torch.long)
            ids0[0, 0], ids0[0, 1] = O.TIE_ID, 1
            t0 = time.perf_counter()
            mem = O.memory_block(ids0, sd)
            t_mem = time.perf_counter() - t0
            enc = torch.cat([enc, mem], dim=1)
        # the real loop, first anchor_steps steps
        n_anchor = min(anchor_steps, T)
        ids = torch.zeros((1, 1), dtype=torch.long)
        t0 = time.perf_counter()
        for _ in range(n_anchor):
            logits = O.decoder_logits(ids, enc, sd)[:, -1, :]
            ids = torch.cat([ids, torch.argmax(logits, dim=-1)[:, None]], dim=-1)
        t_anchor = time.perf_counter() - t0
        # cost model of the remaining steps
        lens = sorted(set([n_anchor] + [max(n_anchor, round(T * i / 8)) for i in range(1, 9)]))
        gen = torch.Generator().manual_seed(seed)
        costs = []
        for L in lens:
            pid = torch.randint(3, 1391, (1, L), generator=gen)
            pid[:, 0] = 0
            t0 = time.perf_counter()
            O.decoder_logits(pid, enc, sd)
            costs.append(time.perf_counter() - t0)
        trapz = getattr(np, "trapezoid", None) or np.trapz
        t_rest = float(trapz(costs, lens)) if len(lens) > 1 else 0.0
    total = t_front + t_enc + t_mem + t_anchor + t_rest
    wall = time.perf_counter() - t_wall0
    desc = (f"1 segment (2.048 s) of the workload, reference ALGORITHM via the CPU oracle port (not the reference's own "
            f"modules), fp32, {os.cpu_count()} threads: frontend {t_front * 1e3:.1f} ms + encoder {t_enc * 1e3:.1f} ms"
            + (f" + memory block {t_mem * 1e3:.1f} ms" if segmem else "") +
            f" timed in full; no-KV-cache greedy loop REALLY RUN for steps 1..{n_anchor} = {t_anchor:.2f} s; steps "
            f"{n_anchor + 1}..{T} EXTRAPOLATED from single full-prefix decoder passes at prefix lengths {lens} = {t_rest:.1f} s")
    return {"sec_per_segment": total, "wall_s": wall, "anchor_steps": n_anchor, "anchor_s": t_anchor,
            "anchor_audio_s_per_s_at_T_dec_eq_anchor": SEG_SECONDS / (t_front + t_enc + t_mem + t_anchor), "desc": desc}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    if args.workload == "finetune":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU oracle port of the fine-tune step (fp64 autograd) "
                          "is a parity tool, minutes per sample; only the transcription workloads have a CPU arm"}))
        return
    samples = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_sample(args, seed=i)
        if i >= args.warmup:
            samples.append(r)
    if not samples:
        samples.append(cpu_reference_sample(args, seed=0))
    sec_per_seg = float(np.median([s["sec_per_segment"] for s in samples]))
    wall = float(np.median([s["wall_s"] for s in samples]))
    value = SEG_SECONDS / sec_per_seg
    last = samples[-1]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        # one step of this arm = one bounded sample that really ran; its wall time, not the extrapolation
        "ms_per_step": wall * 1e3, "extrapolated_s_per_segment": sec_per_seg,
        "higher_is_better": True, "scaling": "strong" if args.workload == "mrmt3_512_slakh" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "label": "CPU oracle port, partly extrapolated (see sample)", "sample": last["desc"],
                         "anchor": {"what": f"the same segment with T_dec = {last['anchor_steps']} run in full, no extrapolation",
                                    "value": last["anchor_audio_s_per_s_at_T_dec_eq_anchor"], "unit": UNIT}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
class Bench:
    """Timing plumbing shared by the workloads: barrier + synchronize on both sides, CUDA events on
    the current stream, max over ranks."""

    def __init__(self, args):
        self.args = args
        self.rank, self.world, self.local = dist_env()
        if self.world != args.gpus and not (self.world == 1 and args.gpus == 1):
            if self.world == 1 and args.gpus > 1:
                raise SystemExit("launch with torch.distributed.run for --gpus > 1")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        import torch.distributed as dist
        self.dist = dist
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.sampler = ClockSampler(self.local)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, k, counter=None):
        """-> (max-over-ranks ms of k calls, this rank's ms, launches, last result)"""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = counter() if counter else 0
        e0.record()
        out = None
        for _ in range(k):
            out = fn()
        e1.record()
        self.barrier()
        mine = e0.elapsed_time(e1)
        ms = mine
        if self.world > 1:
            t = torch.tensor([mine], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, mine, (counter() - l0) if counter else 0, out

    def all_ranks(self, value):
        if self.world == 1:
            return [float(value)]
        t = torch.zeros(self.world, device=self.dev)
        t[self.rank] = float(value)
        self.dist.all_reduce(t)
        return [float(v) for v in t.cpu()]

    def finish(self, line):
        if self.rank == 0:
            print(json.dumps(line))
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def base_line(args, b, value, ms_per_step, scaling, launches, e2e):
    return {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": b.world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 2), "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, b.world), "e2e": e2e, "gpu_launches": int(launches),
    }


def roofline_block(prof, algo, lanes, T, peak, how_peak, note):
    total_ms = sum(v[0] for v in prof.values())
    breakdown = {k: {"ms": round(v[0], 3), "launches": v[1], "share": round(v[0] / total_ms, 4)}
                 for k, v in prof.items() if v[1]}
    dom = max(("attn_self", "attn_cross"), key=lambda k: prof[k][0])
    ms_dom, n_dom = prof[dom]
    achieved = algo[dom] / (ms_dom / 1e3) / 1e9
    traffic, traffic_note = ncu_traffic(dom)
    roofline = {
        "bound": "hbm", "kernel": f"attn_decode_mma_kernel<{'paged self' if dom == 'attn_self' else 'cross'}>",
        "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
        "traffic": traffic, "traffic_note": traffic_note,
        "peak_source": how_peak, "launches": n_dom, "avg_launch_us": round(ms_dom * 1e3 / n_dom, 2),
        "lanes": lanes,
        "how": "eager pass outside the timed region, one lane group, CUDA events on the launching stream around "
               "every launch; achieved = sum of algorithmic bytes / sum of launch durations over all decode steps" + note,
        "algorithmic_bytes_per_launch_avg": int(algo[dom] / n_dom), "share_of_decode_step": breakdown[dom]["share"],
        "other": {k: round(algo[k] / (prof[k][0] / 1e3) / 1e9, 1) for k in algo if k != dom},
    }
    return roofline, breakdown, total_ms


# ---------------------------------------------------------------------------------------------
def run_mt3(args):
    b = Bench(args)
    dev, rank, world = b.dev, b.rank, b.world
    syn = importlib.import_module("mr-mt3_b200.synthetic")
    t5 = importlib.import_module("mr-mt3_b200.t5")
    model = t5.T5ForConditionalGeneration(t5.T5Config())
    model.load_state_dict(syn.synthetic_state_dict(1234, eos_scale=args.eos_scale), strict=True)
    model = model.eval().to(dev)
    eng = model.engine()

    S, T = args.segments, args.max_length
    audio_np = synth_segments(S, seed0=1000 * rank)
    host_audio = torch.from_numpy(audio_np.reshape(-1)).pin_memory()
    dev_audio = host_audio.to(dev)
    start_np = np.arange(S, dtype=np.int64) * SEG_SAMPLES
    len_np = np.full(S, SEG_SAMPLES, dtype=np.int32)
    valid_np = np.full(S, 256, dtype=np.int32)
    d_start, d_len, d_valid = (torch.from_numpy(a).to(dev) for a in (start_np, len_np, valid_np))
    host_out = torch.empty((S, T + 1), dtype=torch.int64).pin_memory()
    gather_bufs = [torch.empty((S, T + 1), dtype=torch.int64, device=dev) for _ in range(world)] \
        if (world > 1 and rank == 0) else None

    def gather(ids):
        if world > 1:
            full = ids if ids.shape[1] == T + 1 else torch.nn.functional.pad(ids, (0, T + 1 - ids.shape[1]))
            b.dist.gather(full.contiguous(), gather_bufs, dst=0)

    def step_resident(m=model, e=eng, t=T):
        mel = e.logmel(dev_audio, d_start, d_len, d_valid, mel_norm=True)
        ids = m.generate(mel, max_length=t)
        gather(ids)
        return ids

    def step_e2e():
        ids = eng.transcribe_host(host_audio, start_np, len_np, valid_np, mel_norm=True, max_length=T,
                                  out=host_out)
        if world > 1:
            gather(ids.to(dev, non_blocking=True))
        return ids

    for _ in range(args.warmup):
        ids = step_resident()
    b.sampler.start()
    ms_res, _, launches, ids = b.timed(step_resident, args.steps, lambda: eng.launch_count)
    for _ in range(max(1, args.warmup // 3)):
        step_e2e()
    ms_e2e, _, _, ids_e2e = b.timed(step_e2e, args.steps)
    b.sampler.stop_flag = True
    b.sampler.join(timeout=2)

    ids = ids.cpu()
    decode_steps = int(ids.shape[1] - 1)
    same = bool(torch.equal(ids, ids_e2e[:, :ids.shape[1]].cpu()))
    audio_s = S * SEG_SECONDS * world
    value = audio_s / (ms_res / 1e3 / args.steps)
    e2e_value = audio_s / (ms_e2e / 1e3 / args.steps)
    e2e = {"value": round(e2e_value, 2), "unit": UNIT, "ms_per_step": round(ms_e2e / args.steps, 2),
           "h2d_bytes_per_step": int(host_audio.numel() * 4 + start_np.nbytes + len_np.nbytes + valid_np.nbytes),
           "d2h_bytes_per_step": int(host_out.numel() * 8), "tokens_equal_resident_path": same}
    line = base_line(args, b, value, ms_res / args.steps, "weak", launches, e2e)
    line["config"]["decode_steps"] = decode_steps
    line["clocks"] = b.sampler.summary()

    # ---- roofline of the dominant kernel class, from an event-bracketed eager pass ----------
    peak, _, how = measured_peaks()
    if not args.no_profile:
        eng.profile_enable(True)
        mel = eng.logmel(dev_audio, d_start, d_len, d_valid, mel_norm=True)
        model.generate(mel, max_length=T)
        prof = eng.profile_read()
        eng.profile_enable(False)
        algo = attn_algo_bytes(S, decode_steps, 256)
        roofline, breakdown, total_ms = roofline_block(prof, algo, S, decode_steps, peak, how, "")
        # whole decode loop against SURVEY 8(d)'s per-step bytes; the timed step also holds the frontend,
        # the encoder and the cross-K/V projection, so this is a lower bound of what the loop sustains
        step_bytes = decode_bytes([S], decode_steps, 256)
        step_s = ms_res / 1e3 / args.steps
        roofline["decode_loop"] = {
            "algorithmic_GB": round(step_bytes / 1e9, 2), "eager_event_ms": round(total_ms, 1),
            "achieved_GBs_eager": round(step_bytes / (total_ms / 1e3) / 1e9, 1),
            "achieved_GBs_timed_region": round(step_bytes / step_s / 1e9, 1),
            "frac_of_peak_timed_region": round(step_bytes / step_s / 1e9 / peak, 4),
        }
        # the same two kernels INSIDE the replayed step graph (one lane group, programmatic dependent launch
        # active, no events between launches): %globaltimer stamps of the last step of runs of growing
        # length; a kernel's time on the chain = its last CTA's end - the previous kernel's end
        try:
            eng.set_option("group_lanes", 0)
            eng.trace_enable(True)
            in_graph = {}
            for t_len in (256, 512, 768, 1024):
                if t_len > T:
                    continue
                model.generate(mel, max_length=t_len)
                tr = eng.trace_read(80)
                spans = {"attn_self": [], "attn_cross": []}
                for layer in range(N_LAYERS):
                    for name, off in (("attn_self", 2), ("attn_cross", 5)):
                        i = 8 * layer + off
                        if tr[i][1] and tr[i - 1][1]:
                            spans[name].append((tr[i][1] - tr[i - 1][1]) / 1e3)
                if spans["attn_self"] and spans["attn_cross"]:
                    us_s, us_c = float(np.median(spans["attn_self"])), float(np.median(spans["attn_cross"]))
                    b_s = S * (1536 * t_len + 1536 + 2304 + 768)
                    b_c = S * (1536 * 256 + 768 + 768)
                    in_graph[f"position_{t_len - 1}"] = {
                        "attn_self_us": round(us_s, 2), "attn_self_GBs": round(b_s / us_s / 1e3, 1),
                        "attn_self_frac": round(b_s / us_s / 1e3 / peak, 4),
                        "attn_cross_us": round(us_c, 2), "attn_cross_GBs": round(b_c / us_c / 1e3, 1),
                        "attn_cross_frac": round(b_c / us_c / 1e3 / peak, 4)}
            roofline["in_graph"] = dict(in_graph, how="median over the 8 layers of (end of the kernel's last CTA - end of "
                                        "the previous kernel) inside the replayed graph, one lane group")
        finally:
            eng.trace_enable(False)
            eng.set_option("group_lanes", -1)
        line["roofline"] = roofline
        line["decode_step_breakdown"] = breakdown

    # ---- secondary decode-length numbers: T_dec = 256, and EOS early exit -------
    if not args.no_secondary and world == 1 and T == 1024 and args.eos_scale == 1.0:
        sec = {}
        for _ in range(2):
            step_resident(t=256)
        ms256, _, _, ids256 = b.timed(lambda: step_resident(t=256), 3)
        sec["t_dec_256"] = {"value": round(audio_s / (ms256 / 3e3), 1), "unit": UNIT, "ms_per_step": round(ms256 / 3, 2),
                            "decode_steps": int(ids256.shape[1] - 1),
                            "what": "same workload with max_length 256 (typical transcription length; not a reference figure)"}
        m2 = t5.T5ForConditionalGeneration(t5.T5Config())
        m2.load_state_dict(syn.synthetic_state_dict(1234, eos_scale=6.0), strict=True)
        m2 = m2.eval().to(dev)
        e2 = m2.engine()
        for _ in range(2):
            step_resident(m2, e2)
        ms_eos, _, _, ids_eos = b.timed(lambda: step_resident(m2, e2), 3)
        n_tok = ((ids_eos[:, 1:] != 0).sum(1)).float()
        sec["eos_early_exit"] = {
            "value": round(audio_s / (ms_eos / 3e3), 1), "unit": UNIT, "ms_per_step": round(ms_eos / 3, 2),
            "decode_steps_run": int(ids_eos.shape[1] - 1), "tokens_per_segment_mean": round(float(n_tok.mean()), 1),
            "tokens_per_segment_max": int(n_tok.max()),
            "what": "EOS row of lm_head x6: rows finish at different steps (masked), the loop exits when all have "
                    "(models/t5.py:288-294)"}
        line["secondary"] = sec

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_sample(args)
        line["cpu_baseline"] = {"value": round(SEG_SECONDS / r["sec_per_segment"], 5), "unit": UNIT, "cores": os.cpu_count(),
                                "kind": "port", "label": "CPU oracle port, partly extrapolated (see sample)",
                                "sample": r["desc"],
                                "anchor": {"what": f"the same segment with T_dec = {r['anchor_steps']} run in full",
                                           "value": round(r["anchor_audio_s_per_s_at_T_dec_eq_anchor"], 4), "unit": UNIT}}
    b.finish(line)


# ---------------------------------------------------------------------------------------------
def run_mrmt3(args):
    b = Bench(args)
    dev, rank, world = b.dev, b.rank, b.world
    syn = importlib.import_module("mr-mt3_b200.synthetic")
    t5 = importlib.import_module("mr-mt3_b200.t5")
    v2 = importlib.import_module("mr-mt3_b200.t5_segmem_v2_with_prev")
    sharding = importlib.import_module("mr-mt3_b200.sharding")
    model = v2.T5SegMemV2WithPrev(t5.T5Config(), 1, 64)
    model.load_state_dict(syn.synthetic_state_dict(4322, segmem=True, eos_scale=args.eos_scale), strict=True)
    model = model.eval().to(dev)
    eng = model.engine()
    T = args.max_length

    # every rank derives the same track table; LPT by segment count; audio only for the local tracks
    dur = track_durations(args)
    samples = np.maximum(1, np.round(dur * 16000.0)).astype(np.int64)
    seg_counts_all = np.array([frame_track(int(n))[0] for n in samples], dtype=np.int64)
    shards = sharding.shard_tracks(seg_counts_all, world)
    local = shards[rank]
    pool = [syn.synthetic_audio(seed=7000 + i, n_samples=SEG_SAMPLES, n_tones=4) for i in range(32)]
    ts = TrackSet(samples[local], pool, seed=100 + rank)
    assert ts.seg_counts == [int(seg_counts_all[t]) for t in local]
    host_audio = torch.from_numpy(ts.audio).pin_memory()
    dev_audio = host_audio.to(dev)
    d_start, d_len, d_valid = (torch.from_numpy(a).to(dev) for a in (ts.seg_start, ts.seg_len, ts.valid))
    host_out = torch.empty((ts.n_seg, T), dtype=torch.int64).pin_memory()
    counts = np.asarray(ts.seg_counts, dtype=np.int32)

    def gather(ids_dev):
        if world > 1:
            return sharding.gather_token_rows(ids_dev, local, seg_counts_all, T)
        return ids_dev

    def step_resident():
        mel = eng.logmel(dev_audio, d_start, d_len, d_valid, mel_norm=True)
        ids = eng.generate_segmem(mel, counts, max_length=T)
        gather(ids)
        return ids

    def step_e2e():
        ids = eng.transcribe_host(host_audio, ts.seg_start, ts.seg_len, ts.valid, seg_counts=counts, mel_norm=True,
                                  max_length=T, out=host_out)
        if world > 1:
            gather(ids.to(dev, non_blocking=True))
        return ids

    for _ in range(args.warmup):
        step_resident()
    b.sampler.start()
    ms_res, mine_res, launches, ids = b.timed(step_resident, args.steps, lambda: eng.launch_count)
    step_e2e()
    e2e_steps = min(args.steps, 5)       # a pass takes seconds: the end-to-end leg times at most 5 of them
    ms_e2e, _, _, ids_e2e = b.timed(step_e2e, e2e_steps)
    b.sampler.stop_flag = True
    b.sampler.join(timeout=2)

    same = bool(torch.equal(ids.cpu(), ids_e2e.cpu()))
    audio_s = float(samples.sum()) / 16000.0                         # the whole job, all ranks
    value = audio_s / (ms_res / 1e3 / args.steps)
    e2e_value = audio_s / (ms_e2e / 1e3 / e2e_steps)
    h2d = int(host_audio.numel() * 4 + ts.seg_start.nbytes + ts.seg_len.nbytes + ts.valid.nbytes)
    h2d_all = b.all_ranks(h2d)
    d2h_all = b.all_ranks(host_out.numel() * 8)
    per_rank_ms = [round(v / args.steps, 2) for v in b.all_ranks(mine_res)]
    seg_load = [int(sum(seg_counts_all[t] for t in s)) for s in shards]
    rounds = [int(max((seg_counts_all[t] for t in s), default=0)) for s in shards]
    e2e = {"value": round(e2e_value, 2), "unit": UNIT, "ms_per_step": round(ms_e2e / e2e_steps, 2), "steps": e2e_steps,
           "h2d_bytes_per_step": int(sum(h2d_all)), "d2h_bytes_per_step": int(sum(d2h_all)),
           "tokens_equal_resident_path": same}
    scaling = "strong" if args.workload == "mrmt3_512_slakh" else ("weak" if world == 1 else "strong")
    line = base_line(args, b, value, ms_res / args.steps, scaling, launches, e2e)
    line["config"].update({"segments_total": int(seg_counts_all.sum()), "audio_seconds_total": round(audio_s, 1),
                           "lanes_per_gpu": [len(s) for s in shards]})
    line["clocks"] = b.sampler.summary()
    if args.workload == "mrmt3_512_slakh" and args.duration_scale == 1.0 / 16.0 and (args.tracks or 512) == 512:
        # the N = 1 point of this strong-scaling curve (the N = 1 default of bench.py is configs[1]): measured in
        # round 2 with this same command at --gpus 1 on one B200 of this pool
        line["strong_scaling_context"] = {"same_workload_n1_value": 785.72, "unit": UNIT,
                                          "from": "profiles/r2f_bench_slakh_n1.json (python bench.py --workload "
                                                  "mrmt3_512_slakh, 1 GPU, round 2)"}
    line["sharding"] = {
        "per_rank_ms": per_rank_ms, "segments_per_rank": seg_load, "rounds_per_rank": rounds,
        "time_imbalance_max_over_mean": round(max(per_rank_ms) / (sum(per_rank_ms) / len(per_rank_ms)), 4),
        "segment_imbalance_max_over_mean": round(max(seg_load) / (sum(seg_load) / len(seg_load)), 4),
        # a rank runs max-segments rounds of max_length steps whatever its lane count: lanes whose track
        # has ended idle, so the mean active fraction bounds the efficiency of the latency-bound regime
        "mean_active_lane_fraction": round(float(np.mean([seg_load[r] / max(1, rounds[r] * len(shards[r]))
                                                          for r in range(world)])), 4),
    }

    peak, _, how = measured_peaks()
    lanes_per_round = [int(sum(1 for t in local if seg_counts_all[t] > r)) for r in range(rounds[rank])]
    step_bytes = decode_bytes(lanes_per_round, T, 320)
    step_s = mine_res / 1e3 / args.steps
    if not args.no_profile and len(local):
        # one round (every local track's first segment), eager, event-bracketed: the lane regime of this rank
        first = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
        mel = eng.logmel(dev_audio, d_start[first], d_len[first], d_valid[first], mel_norm=True)
        eng.profile_enable(True)
        eng.generate_segmem(mel, np.ones(len(local), dtype=np.int32), max_length=T)
        prof = eng.profile_read()
        eng.profile_enable(False)
        algo = attn_algo_bytes(len(local), T, 320)
        roofline, breakdown, total_ms = roofline_block(
            prof, algo, len(local), T, peak, how,
            f"; rank 0's lane regime ({len(local)} lanes, T_k 320), one round of {T} steps")
        roofline["decode_loop"] = {
            "algorithmic_GB_rank0": round(step_bytes / 1e9, 2),
            "achieved_GBs_timed_region_rank0": round(step_bytes / step_s / 1e9, 1),
            "frac_of_peak_timed_region_rank0": round(step_bytes / step_s / 1e9 / peak, 4),
            "us_per_decode_step_timed_region_rank0": round(step_s * 1e6 / max(1, rounds[rank] * T), 1),
            "hbm_floor_us_per_decode_step_full_lanes": round(decode_bytes([len(local)], T, 320) / T / peak / 1e3, 1),
        }
        line["roofline"] = roofline
        line["decode_step_breakdown"] = breakdown
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_sample(args)
        line["cpu_baseline"] = {"value": round(SEG_SECONDS / r["sec_per_segment"], 5), "unit": UNIT, "cores": os.cpu_count(),
                                "kind": "port", "label": "CPU oracle port, partly extrapolated (see sample)",
                                "sample": r["desc"],
                                "anchor": {"what": f"the same segment with T_dec = {r['anchor_steps']} run in full",
                                           "value": round(r["anchor_audio_s_per_s_at_T_dec_eq_anchor"], 4), "unit": UNIT}}
    b.finish(line)


# ---------------------------------------------------------------------------------------------
def run_finetune(args):
    b = Bench(args)
    dev, rank, world = b.dev, b.rank, b.world
    syn = importlib.import_module("mr-mt3_b200.synthetic")
    t5 = importlib.import_module("mr-mt3_b200.t5")
    v2 = importlib.import_module("mr-mt3_b200.t5_segmem_v2_with_prev")
    model = v2.T5SegMemV2WithPrev(t5.T5Config(), 1, 64)
    model.load_state_dict(syn.synthetic_state_dict(4322, segmem=True), strict=True)
    model = model.eval().to(dev)
    B, L = args.batch, 1024
    g = torch.Generator().manual_seed(100 + rank)
    x_host = torch.rand((B, 256, 512), generator=g).pin_memory()

    def rows():
        t = torch.randint(3, 1391, (B, L), generator=g)
        for i in range(B):
            n = int(torch.randint(64, 901, (1,), generator=g))
            t[i, n] = 1
            t[i, n + 1:] = -100
        return t
    labels_host, prev_host = rows().pin_memory(), rows().pin_memory()
    x, labels, prev = x_host.to(dev), labels_host.to(dev), prev_host.to(dev)
    trainer = model.trainer(lr=1e-5, dropout=args.dropout, seed=1234 + rank)

    def step_resident():
        return trainer.step(x, labels, prev, want_loss=False)

    def step_e2e():
        xs = x_host.to(dev, non_blocking=True)
        ls = labels_host.to(dev, non_blocking=True)
        ps = prev_host.to(dev, non_blocking=True)
        return trainer.step(xs, ls, ps, want_loss=True)              # the loss scalar comes back to the host

    for _ in range(args.warmup):
        step_resident()
    b.sampler.start()
    eng = model.engine()
    ms_res, _, launches, _ = b.timed(step_resident, args.steps, lambda: eng.launch_count)
    step_e2e()
    ms_e2e, _, _, loss = b.timed(step_e2e, args.steps)
    b.sampler.stop_flag = True
    b.sampler.join(timeout=2)
    phases = trainer.phase_times(x, labels, prev)                     # event-bracketed, outside the timed region
    samples = B * world
    step_s = ms_res / 1e3 / args.steps
    value = samples * SEG_SECONDS / step_s
    e2e = {"value": round(samples * SEG_SECONDS / (ms_e2e / 1e3 / args.steps), 2), "unit": UNIT,
           "ms_per_step": round(ms_e2e / args.steps, 2),
           "h2d_bytes_per_step": int((x_host.numel() * 4 + labels_host.numel() * 8 + prev_host.numel() * 8) * world),
           "d2h_bytes_per_step": 4 * world, "loss": loss}
    line = base_line(args, b, value, ms_res / args.steps, "weak", launches, e2e)
    line["clocks"] = b.sampler.summary()
    _, peak_tf, how = measured_peaks()
    flop = B * 249.6e9                                                 # 83.2 GFLOP forward per sample, x3
    tf = flop / step_s / 1e12
    identical = True
    if world > 1:                                                      # data-parallel replicas must stay bit-identical
        flat = eng.train_read_master()
        ref = flat.clone()
        b.dist.broadcast(ref, 0)
        ok = torch.tensor([1.0 if torch.equal(flat, ref) else 0.0], device=dev)
        b.dist.all_reduce(ok, op=b.dist.ReduceOp.MIN)
        identical = bool(ok.item() == 1.0)
    line["training"] = {"samples_per_s": round(samples / step_s, 1), "ms_per_step": round(ms_res / args.steps, 2),
                        "phases_ms": phases, "grad_allreduce": trainer.comm_description(),
                        "replicas_bit_identical_after_run": identical}
    line["roofline"] = {"bound": "tensor", "kernel": "fine-tune step (all kernels; SURVEY 8d 249.6 GFLOP per sample)",
                        "achieved": round(tf, 1), "peak": peak_tf, "unit": "TFLOP/s", "frac": round(tf / peak_tf, 4),
                        "traffic": None, "peak_source": how,
                        "how": "7.99 TFLOP per GPU per step / device-timed step (CUDA events, max over ranks)"}
    b.finish(line)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "mt3_256":
        run_mt3(a)
    elif a.workload == "finetune":
        run_finetune(a)
    else:
        run_mrmt3(a)
