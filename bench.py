#!/usr/bin/env python
"""bench.py -- transcribed audio-seconds per second of the MR-MT3 hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): MT3 batched inference, 256 synthetic 2.048 s segments per
GPU: log-mel + T5 encoder + KV-cached greedy decode, max 1024 tokens.  The seeded synthetic
weights never emit EOS (SURVEY 8d "decode length convention"), so every segment decodes the full
T_dec = 1024 tokens -- the worst case of the reference's `max_length=1024`.  One "step" = one pass
of the whole path over the batch.  N GPUs: every rank transcribes its own 256 segments (tracks are
independent, SURVEY 8e: no collective on the data path) and rank 0 gathers the token rows.

  value : inputs (fp32 audio) already in HBM; log-mel -> generate, device-timed.
  e2e   : the same through the host-buffer C-ABI call (`mrmt3_transcribe_host`): pinned host audio
          in, host token rows out, H2D + D2H inside the timed region.
  roofline / cpu_baseline : see DESIGN.md section "Measurement".

`--impl reference` times the reference's CPU algorithm (no KV cache, fp32, all host threads)
through the oracle port (the reference's own modules cannot be imported on the GPU box).
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

SEG_SECONDS = 32768 / 16000.0
METRIC = "transcribed audio-sec/sec (log-mel + T5 encode + greedy decode)"
UNIT = "audio-s/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--segments", type=int, default=256, help="segments per GPU")
    ap.add_argument("--max-length", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-profile", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


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


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kind):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r1_ncu_traffic.json, written from the .ncu-rep by scripts/ncu_summary.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")) as f:
            d = json.load(f)[kind]
        return int(d["dram_bytes_per_launch"]), d["note"]
    except Exception:
        return None, "no ncu capture committed"


def synth_segments(n_seg, seed0):
    syn = importlib.import_module("mr-mt3_b200.synthetic")
    audio = np.stack([syn.synthetic_audio(seed=seed0 + i, n_samples=32768, n_tones=4) for i in range(n_seg)])
    return audio.astype(np.float32)


# ---------------------------------------------------------------------------------------------
# the reference's CPU algorithm, costed on a bounded sample
def cpu_reference_sample(max_length, seed=0):
    """Time the reference algorithm for ONE segment on this host: frontend and encoder in full; the
    no-KV-cache greedy loop (models/t5.py:267-295 re-runs the decoder over the whole prefix every
    step) is costed by timing full-prefix decoder passes at 9 prefix lengths and integrating over
    the max_length steps.  -> (seconds per segment, description)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import mt3_oracle as O
    syn = importlib.import_module("mr-mt3_b200.synthetic")
    torch.set_num_threads(os.cpu_count())
    sd = O.cast_state_dict(syn.synthetic_state_dict(1234), torch.float32)
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
        lens = sorted(set([1] + [max(1, round(max_length * i / 8)) for i in range(1, 9)]))
        gen = torch.Generator().manual_seed(seed)
        costs = []
        for L in lens:
            ids = torch.randint(3, 1391, (1, L), generator=gen)
            ids[:, 0] = 0
            t0 = time.perf_counter()
            O.decoder_logits(ids, enc, sd)
            costs.append(time.perf_counter() - t0)
        trapz = getattr(np, "trapezoid", None) or np.trapz
        t_dec = float(trapz(costs, lens)) + costs[0]          # sum_{t=1..L} cost(t)
    total = t_front + t_enc + t_dec
    desc = (f"1 segment (2.048 s) of the workload, reference algorithm via the oracle port, fp32, "
            f"{os.cpu_count()} threads: frontend {t_front * 1e3:.1f} ms + encoder {t_enc * 1e3:.1f} ms timed in full; "
            f"no-KV-cache greedy loop costed from full-prefix decoder passes at prefix lengths {lens} "
            f"integrated over {max_length} steps = {t_dec:.1f} s")
    return total, desc


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    times = []
    desc = ""
    for i in range(args.warmup + args.steps):
        # a "step" of this arm is the bounded one-segment sample; warm-ups are not skipped because
        # each sample already contains its own warm-up pass
        if i >= args.warmup or i == 0:
            t, desc = cpu_reference_sample(args.max_length, seed=i)
            if i >= args.warmup:
                times.append(t)
    if not times:
        t, desc = cpu_reference_sample(args.max_length, seed=0)
        times.append(t)
    sec_per_seg = float(np.median(times))
    value = SEG_SECONDS / sec_per_seg
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_seg * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"MT3 batched inference: {args.segments} synthetic 2.048 s segments, "
                               f"greedy decode T_dec={args.max_length} (BASELINE.json configs[1])",
                   "segments_per_gpu": args.segments, "max_length": args.max_length,
                   "step": "one bounded CPU sample (1 segment), scaled linearly: segments are independent"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    rank, world, local = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    syn = importlib.import_module("mr-mt3_b200.synthetic")
    t5 = importlib.import_module("mr-mt3_b200.t5")
    lib = importlib.import_module("mr-mt3_b200._lib")
    model = t5.T5ForConditionalGeneration(t5.T5Config())
    model.load_state_dict(syn.synthetic_state_dict(1234), strict=True)
    model = model.eval().to(dev)
    eng = model.engine()

    S, T = args.segments, args.max_length
    audio_np = synth_segments(S, seed0=1000 * rank)
    host_audio = torch.from_numpy(audio_np.reshape(-1)).pin_memory()
    dev_audio = host_audio.to(dev)
    start_np = np.arange(S, dtype=np.int64) * 32768
    len_np = np.full(S, 32768, dtype=np.int32)
    valid_np = np.full(S, 256, dtype=np.int32)
    d_start, d_len, d_valid = (torch.from_numpy(a).to(dev) for a in (start_np, len_np, valid_np))
    host_out = torch.empty((S, T + 1), dtype=torch.int64).pin_memory()
    gather_bufs = [torch.empty((S, T + 1), dtype=torch.int64, device=dev) for _ in range(world)] \
        if (world > 1 and rank == 0) else None

    def gather(ids):
        if world > 1:
            full = ids if ids.shape[1] == T + 1 else torch.nn.functional.pad(ids, (0, T + 1 - ids.shape[1]))
            dist.gather(full.contiguous(), gather_bufs, dst=0)

    def step_resident():
        mel = eng.logmel(dev_audio, d_start, d_len, d_valid, mel_norm=True)
        ids = model.generate(mel, max_length=T)
        gather(ids)
        return ids

    def step_e2e():
        ids = eng.transcribe_host(host_audio, start_np, len_np, valid_np, mel_norm=True, max_length=T,
                                  out=host_out)
        if world > 1:
            gather(ids.to(dev, non_blocking=True))
        return ids

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count
        e0.record()
        for _ in range(k):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, eng.launch_count - l0, out

    for _ in range(args.warmup):
        ids = step_resident()
    sampler = ClockSampler(local)
    sampler.start()
    ms_res, launches, ids = timed(step_resident, args.steps)
    for _ in range(max(1, args.warmup // 3)):
        step_e2e()
    ms_e2e, _, ids_e2e = timed(step_e2e, args.steps)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    ids = ids.cpu()
    decode_steps = int(ids.shape[1] - 1)
    same = bool(torch.equal(ids, ids_e2e[:, :ids.shape[1]].cpu()))
    audio_s = S * SEG_SECONDS * world
    value = audio_s / (ms_res / 1e3 / args.steps)
    e2e_value = audio_s / (ms_e2e / 1e3 / args.steps)

    # ---- roofline of the dominant kernel class, from an event-bracketed eager pass ----------
    roofline, breakdown = None, None
    if not args.no_profile:
        eng.profile_enable(True)
        mel = eng.logmel(dev_audio, d_start, d_len, d_valid, mel_norm=True)
        model.generate(mel, max_length=T)
        prof = eng.profile_read()
        eng.profile_enable(False)
        total_ms = sum(v[0] for v in prof.values())
        breakdown = {k: {"ms": round(v[0], 3), "launches": v[1], "share": round(v[0] / total_ms, 4)}
                     for k, v in prof.items()}
        n_layers, Tk = 8, 256
        n_tok = decode_steps
        algo = {
            # per launch (one layer, all lanes): K/V pages read + the step's K/V row written + the
            # fused q|k|v row read + the context row written
            "attn_self": S * n_layers * (1536 * n_tok * (n_tok + 1) // 2 + n_tok * (1536 + 2304 + 768)),
            # cross cache read + q row read + context row written
            "attn_cross": S * n_layers * n_tok * (1536 * Tk + 768 + 768),
        }
        dom = max(("attn_self", "attn_cross"), key=lambda k: prof[k][0])
        peak, how = measured_peak_hbm()
        ms_dom, n_dom = prof[dom]
        achieved = algo[dom] / (ms_dom / 1e3) / 1e9
        traffic, traffic_note = ncu_traffic(dom)
        roofline = {
            "bound": "hbm", "kernel": f"attn_decode_mma_kernel<{'paged self' if dom == 'attn_self' else 'cross'}>",
            "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
            "traffic": traffic, "traffic_note": traffic_note,
            "peak_source": how, "launches": n_dom, "avg_launch_us": round(ms_dom * 1e3 / n_dom, 2),
            "how": "eager pass outside the timed region, one lane group, CUDA events on the launching stream around "
                   "every launch; achieved = sum of algorithmic bytes / sum of launch durations over all decode steps",
            "algorithmic_bytes_per_launch_avg": int(algo[dom] / n_dom), "share_of_decode_step": breakdown[dom]["share"],
            "other": {k: round(algo[k] / (prof[k][0] / 1e3) / 1e9, 1) for k in algo if k != dom},
        }
        # whole decode step against SURVEY 8(d)'s per-step bytes
        step_bytes = n_tok * 45.64e6 + S * (12288 * Tk * n_tok + 12288 * n_tok * (n_tok + 1) / 2 + 12292 * n_tok)
        # the same bytes against the TIMED region (graph replay, concurrent lane groups); the timed
        # step also holds the frontend, the encoder and the cross-K/V projection, so this is a lower
        # bound of what the decode loop itself sustains
        step_s = ms_res / 1e3 / args.steps
        roofline["decode_loop"] = {
            "algorithmic_GB": round(step_bytes / 1e9, 2), "eager_event_ms": round(total_ms, 1),
            "achieved_GBs_eager": round(step_bytes / (total_ms / 1e3) / 1e9, 1),
            "achieved_GBs_timed_region": round(step_bytes / step_s / 1e9, 1),
            "frac_of_peak_timed_region": round(step_bytes / step_s / 1e9 / peak, 4),
        }

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_res / args.steps, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"MT3 batched inference: {S} synthetic 2.048 s segments per GPU, log-mel + encoder + "
                               f"KV-cached greedy decode, T_dec={decode_steps} of max_length {T} "
                               f"(BASELINE.json configs[1])",
                   "segments_per_gpu": S, "max_length": T, "decode_steps": decode_steps,
                   "weights": "seeded synthetic (never emit EOS)", "parallelism": f"track-sharded x{world}",
                   "l2": "per-step working set (KV caches, ~4 GB) exceeds L2; no explicit flush"},
        "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "ms_per_step": round(ms_e2e / args.steps, 2),
                "h2d_bytes_per_step": int(host_audio.numel() * 4 + start_np.nbytes + len_np.nbytes + valid_np.nbytes),
                "d2h_bytes_per_step": int(host_out.numel() * 8), "tokens_equal_resident_path": same},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if roofline:
        line["roofline"] = roofline
        line["decode_step_breakdown"] = breakdown
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, desc = cpu_reference_sample(T)
        line["cpu_baseline"] = {"value": round(SEG_SECONDS / sec, 5), "unit": UNIT, "cores": os.cpu_count(),
                                "kind": "port", "sample": desc}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
