#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into the handful of counters DESIGN.md / bench.py cite.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/r1_x_summary.json

Per captured launch: duration, DRAM bytes read/written (the `traffic` of bench.py's roofline),
DRAM and L2 throughput, tensor-pipe activity, issue-slot use, occupancy, registers, shared memory
and the top warp-stall reasons.  Runs where ncu is installed (no GPU needed)."""
import csv
import io
import json
import subprocess
import sys

RAW = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "dram__bytes_read.sum.per_second": "dram_read_rate",
    "lts__t_bytes.sum": "l2_bytes_MB",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct_of_nominal",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active": "hmma_pipe_pct",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active": "tmem_pipe_pct",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active": "tma_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_per_block",
    "launch__waves_per_multiprocessor": "waves_per_sm",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
}


def page(rep, which):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], capture_output=True, text=True).stdout
    start = out.find('"ID"')
    return list(csv.reader(io.StringIO(out[start:])))


def main(rep, dst):
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:100]}
        stalls = {}
        for i, h in enumerate(hdr):
            if h in RAW and r[i] != "":
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[RAW[h]] = v if not units[i] else {"v": v, "unit": units[i]}
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(r[i])
                except ValueError:
                    pass
        d["top_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:5])
        launches.append(d)
    json.dump({"report": rep, "launches": launches}, open(dst, "w"), indent=1)
    for d in launches:
        print({k: d[k] for k in d if k in ("kernel", "duration_us", "dram_read_MB", "dram_write_MB",
                                           "achieved_occupancy_pct", "issue_slots_busy_pct")})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
