"""Builds profiles/r2_ncu_summary.json (read by bench.py for `roofline.traffic` and `roofline.ncu`) from `ncu --set full`
captures of the score/divergence kernels.  Run here, no GPU needed:
    python profiles/ncu_to_json.py n55:592:gpurun_out/r2v_tri55.ncu-rep n13:2664:gpurun_out/r2v_tri13.ncu-rep
(label : particles per launch : report).  Per-launch numbers; ncu serialises and replays kernels, so only ratios and byte
counts are quoted from it, never times as benchmark values."""
import csv
import json
import os
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
         "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def summarise(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]
        k = {}
        for h, u, v in zip(hdr, units, r):
            if h in WANT:
                val = float(v.replace(",", "")) * SCALE.get(u, 1.0)
                k[WANT[h]] = val
        if "time" in k:
            k["time_us"] = k.pop("time")
        kernels.setdefault(name, k)  # first captured launch of each kernel
    return kernels


def main():
    res = {}
    for spec in sys.argv[1:]:
        label, per_launch, rep = spec.split(":")
        ks = summarise(rep)
        per_launch = int(per_launch)
        dram = sum(k.get("dram_read", 0.0) + k.get("dram_write", 0.0) for n, k in ks.items() if n.startswith("tri_phase"))
        res[label] = {"source": os.path.basename(rep), "particles_per_launch": per_launch, "kernels": ks,
                      "dram_bytes_per_particle": dram / per_launch,
                      "note": "dram bytes = dram__bytes_read.sum + dram__bytes_write.sum of tri_phase_a + tri_phase_b, one launch each"}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r2_ncu_summary.json")
    json.dump(res, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(res, indent=1, sort_keys=True)[:3000])


if __name__ == "__main__":
    main()
