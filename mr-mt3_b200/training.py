"""Data-parallel fine-tune loop over the C-ABI fine-tune step (BASELINE configs[4]).

Reference: `training_step` of tasks/mt3_net.py / tasks/mt3_net_segmem_v2_with_prev.py:25-39 under
Lightning's DDP trainer (train.py, config/config.yaml:45-46; `train.sh:65-84`: AdamW lr 1e-5).  DDP
all-reduces gradient buckets while the backward is still running; here the flat fp32 gradient is laid
out in the order the hand-written backward FINISHES it (`mrmt3_train_bucket`), the library records
one CUDA event per bucket, and this class all-reduces bucket i (NCCL, ReduceOp.AVG) on a side stream
as soon as its event has fired -- the collectives hide behind the rest of the backward pass, and only
the last bucket (proj, embedding, norm weights: 1.1 M of 48.5 M elements) is exposed.
"""
import contextlib

import torch
import torch.distributed as dist


class Trainer:
    def __init__(self, model, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, dropout=None, seed=0,
                 process_group=None, overlap=True, engine=None):
        self.model = model
        self.eng = engine if engine is not None else model._train_engine()   # `engine`: test stub
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.group = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.overlap = overlap
        self.buckets = self.eng.train_buckets()
        self.grad = torch.empty(self.eng._n_params, dtype=torch.float32, device=self.eng.device)
        self.side = torch.cuda.Stream(device=self.eng.device) if (self.world > 1 and self.grad.is_cuda) else None
        p = float(getattr(model.config, "dropout_rate", 0.0) or 0.0) if dropout is None else float(dropout)
        self.eng.train_set_dropout(p, seed)          # the library derives a new mask seed after every forward
        self._avg = dist.ReduceOp.AVG if (self.world > 1 and dist.get_backend(process_group) == "nccl") else None

    # ---- gradient exchange --------------------------------------------------------------------------
    def _allreduce(self):
        """All-reduce (mean) of the flat gradient; with `overlap` bucket by bucket behind the backward."""
        if self.world == 1:
            return
        if not self.overlap:
            dist.all_reduce(self.grad, op=self._avg or dist.ReduceOp.SUM, group=self.group)
            if self._avg is None:
                self.grad /= self.world
            return
        works = []
        for i, (off, cnt) in enumerate(self.buckets):
            with (torch.cuda.stream(self.side) if self.side is not None else contextlib.nullcontext()):
                self.eng.train_wait_bucket(i, self.side)               # side stream waits for the bucket's event
                view = self.grad[off:off + cnt]
                works.append(dist.all_reduce(view, op=self._avg or dist.ReduceOp.SUM, group=self.group, async_op=True))
        for w in works:
            w.wait()                                                   # the launching stream waits for NCCL
        if self._avg is None:                                          # backends without AVG (gloo)
            self.grad /= self.world

    def step(self, inputs, labels, targets_prev=None, want_loss=True):
        """forward -> loss -> backward -> all-reduce -> AdamW.  Device tensors.  Returns the mean loss
        of this rank's batch (None with want_loss=False: the step then has no host synchronisation)."""
        m, eng = self.model, self.eng
        if self.side is not None:
            self.side.wait_stream(torch.cuda.current_stream())         # last step's AdamW read the gradient
        eng.train_forward(inputs, m._shift_right(labels), labels, targets_prev, want_loss=False)
        eng.train_backward(self.grad)
        self._allreduce()
        eng.train_apply(self.grad, self.lr, self.betas, self.eps, self.weight_decay)
        return eng.train_loss() if want_loss else None

    # ---- measurement helpers (bench.py; never inside a timed region) --------------------------------
    def phase_times(self, inputs, labels, targets_prev=None, repeats=3):
        """Median device ms of forward / backward (+ overlapped all-reduce) / exposed all-reduce tail / AdamW."""
        m, eng = self.model, self.eng
        rows = []
        for _ in range(repeats):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier(group=self.group)
            if self.side is not None:
                self.side.wait_stream(torch.cuda.current_stream())
            ev[0].record()
            eng.train_forward(inputs, m._shift_right(labels), labels, targets_prev, want_loss=False)
            ev[1].record()
            eng.train_backward(self.grad)
            ev[2].record()
            self._allreduce()
            ev[3].record()
            eng.train_apply(self.grad, self.lr, self.betas, self.eps, self.weight_decay)
            ev[4].record()
            torch.cuda.synchronize()
            rows.append([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
        t = torch.tensor(rows).median(0).values
        return {"forward": round(float(t[0]), 3), "backward": round(float(t[1]), 3),
                "allreduce_exposed_after_backward": round(float(t[2]), 3), "adamw": round(float(t[3]), 3)}

    def comm_description(self):
        sizes = [c * 4 for _, c in self.buckets]
        return {"world": self.world, "overlapped": bool(self.overlap and self.world > 1), "buckets": len(self.buckets),
                "bucket_MB_min_max": [round(min(sizes) / 1e6, 2), round(max(sizes) / 1e6, 2)],
                "total_MB": round(sum(sizes) / 1e6, 1), "op": "nccl all-reduce AVG per bucket on a side stream"
                if self.world > 1 else "none (1 GPU)"}
