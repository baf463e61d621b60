"""TEST INFRASTRUCTURE ONLY -- mint tests/golden/segmem_v1_forward.npz from the REFERENCE ITSELF.

    python oracle/make_golden_v1_forward.py        (in the build container, /root/reference mounted)

The teacher-forced logits of the reference's unmodified `T5SegMem` (models/t5_segmem.py:68-170, reached
through `forward` of models/t5.py:182-249), imported through oracle/ref_shim.py, fp32, CPU, eval mode, with
the seeded synthetic state dict of `mr-mt3_b200/synthetic.py`.  The rows of the batch are consecutive
segments: row i's memory block is built from row i - 1's shifted labels.  Two cases: L = 64 (= segmem_length,
the shortest the reference handles: it drops `segmem_length` output rows whatever the number of memory rows,
so shorter labels lose token rows) and L = 80 with -100 padding behind an EOS.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

spec = importlib.util.spec_from_file_location("mrmt3_synthetic", os.path.join(ROOT, "mr-mt3_b200", "synthetic.py"))
syn = importlib.util.module_from_spec(spec)
spec.loader.exec_module(syn)


@torch.no_grad()
def main():
    torch.set_num_threads(os.cpu_count())
    sd = syn.synthetic_state_dict(4322, segmem=True)
    model = ref_shim.build_segmem_v1(sd)
    x = syn.synthetic_features(11, 3)
    out = {}
    gen = torch.Generator().manual_seed(9)
    for tag, L in (("short", 64), ("long", 80)):
        labels = torch.randint(3, 1391, (3, L), generator=gen)
        if tag == "long":
            labels[1, 50] = 1
            labels[1, 51:] = -100
            labels[2, 70] = 1
            labels[2, 71:] = -100
        logits = model(inputs=x, labels=labels.clone())
        out[f"{tag}_labels"] = labels.numpy()
        out[f"{tag}_logits_sub"] = logits.numpy().astype(np.float32)[:, :, ::4]     # every 4th vocabulary column
        out[f"{tag}_argmax"] = logits.argmax(-1).numpy()
        print(tag, logits.shape, float(logits.abs().max()))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "segmem_v1_forward.npz"), sd_seed=4322, feat_seed=11, **out)


if __name__ == "__main__":
    main()
