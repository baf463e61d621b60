"""BASELINE configs[4] on one GPU (or under torchrun on N): MR-MT3 V2WithPrev fine-tune step,
batch 32 per GPU, labels and targets_prev of length 1024 (reference batch recipe
dataset_2_random_segmem_prev.py:98-134), forward + backward + gradient all-reduce + AdamW.
Reports ms per phase and TFLOP/s against SURVEY 8d's 249.6 GFLOP per sample (forward 83.2, x3)."""
import importlib, json, os, sys, torch
sys.path.insert(0, '.')
import torch.distributed as dist
syn = importlib.import_module("mr-mt3_b200.synthetic"); t5 = importlib.import_module("mr-mt3_b200.t5")
v2 = importlib.import_module("mr-mt3_b200.t5_segmem_v2_with_prev")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dropout = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
m = v2.T5SegMemV2WithPrev(t5.T5Config(), 1, 64); m.load_state_dict(syn.synthetic_state_dict(4322, segmem=True)); m = m.eval().cuda()
eng = m.engine(); n_params = eng.train_init()
eng.train_set_dropout(dropout, 1234 + rank)
g = torch.Generator().manual_seed(100 + rank)
x = torch.rand((B, 256, 512), generator=g).cuda()
def toks():
    t = torch.randint(3, 1391, (B, L), generator=g)
    for b in range(B):
        n = int(torch.randint(64, min(900, L - 1), (1,), generator=g)) if L > 128 else L // 2
        t[b, n] = 1; t[b, n + 1:] = -100
    return t
labels, prev = toks().cuda(), toks().cuda()
prev[prev == -100] = 0
dec_in = m._shift_right(labels)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
res = []
for it in range(steps + 1):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    ev[0].record(); logits, loss = eng.train_forward(x, dec_in, labels, prev)
    ev[1].record(); grad = eng.train_backward()
    ev[2].record()
    if world > 1:
        dist.all_reduce(grad); grad /= world
    ev[3].record(); eng.train_apply(grad, 1e-5)
    ev[4].record(); torch.cuda.synchronize()
    if it:  # first iteration warms up allocations
        res.append([ev[i].elapsed_time(ev[i + 1]) for i in range(4)] + [loss])
import numpy as np
r = np.array(res)
tot = float(np.median(r[:, :4].sum(1)))  # median: a descheduled host thread shows up as one slow forward
if world > 1:
    t = torch.tensor([tot], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); tot = float(t)
if rank == 0:
    print(json.dumps({"config": f"MR-MT3 V2WithPrev fine-tune step, batch {B}/GPU, L = Lp = {L}, {world} GPU(s), dropout {dropout}", "params": n_params,
                      "ms_forward": round(float(np.median(r[:, 0])), 2), "ms_backward": round(float(np.median(r[:, 1])), 2),
                      "ms_allreduce": round(float(np.median(r[:, 2])), 2), "ms_adamw": round(float(np.median(r[:, 3])), 2), "ms_step": round(tot, 2),
                      "ms_step_each": [round(v, 2) for v in r[:, :4].sum(1)],
                      "samples_per_s": round(B * world / (tot / 1e3), 1),
                      "tflops_per_gpu_vs_249.6_gflop_per_sample": round(B * 249.6e9 / (tot / 1e3) / 1e12, 1),
                      "losses": [round(v, 4) for v in r[:, 4]], "launches": eng.launch_count}))
if world > 1:
    # the replicas must stay identical
    flat = eng.train_read_master(); ref = flat.clone(); dist.broadcast(ref, 0)
    assert torch.equal(flat, ref), "replicas diverged"
    dist.destroy_process_group()
