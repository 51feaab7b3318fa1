"""Turns one `ncu --set full` capture of ONE bench step (bench.py --profile-step) into the committed evidence:

    python scripts/ncu_summary.py gpurun_out/r02_full_cfg2.ncu-rep gpurun_out/profile_step_cfg2.json cfg2

  profiles/r02_ncu_full_<cfg>.csv     one row per launch: schedule entry, kernel, duration, DRAM bytes read/written, DRAM / tensor-pipe /
                                      SM throughput %, active warps %, registers, shared memory
  profiles/r02_traffic_<cfg>.json     {schedule entry: dram bytes read + written per launch} — what bench.py reports as roofline.traffic
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def main():
    rep, sched_json, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    units = rows[1]
    data = rows[2:]
    ki = hdr.index("Kernel Name")
    cols = {m: hdr.index(m) for m in METRICS if m in hdr}
    with open(sched_json) as fh:
        sched = json.loads([l for l in fh.read().splitlines() if l.startswith("{")][-1])["schedule"]

    def num(v):
        try:
            return float(v.replace(",", ""))
        except ValueError:
            return None

    def to_bytes(v, unit):
        f = num(v)
        if f is None:
            return None
        return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

    out_rows, traffic = [], {}
    si = 0   # index into the schedule; split-K reducers and the decode/NMS pair are attached by kernel name
    last = None
    for r in data:
        k = r[ki]
        short = k.split("(")[0].replace("void ", "").replace("k2y::<unnamed>::", "").replace("k2y::", "")
        if "splitk_reduce" in k:
            entry = f"{last} (split-K reduce)"
        elif "detect_scan" in k:
            entry = "detect: decode scan"
        elif "detect_nms" in k:
            entry = "detect: per-class NMS"
        elif k.startswith("void at::") or "at::native" in k:
            entry = "(torch fill, outside the step)"
        else:
            while si < len(sched) and sched[si].startswith("conv_pw") and si > 0 and "+" in sched[si - 1] and sched[si] in sched[si - 1]:
                si += 1   # a pointwise layer that ran inside the previous fused launch has no kernel of its own
            entry = sched[si] if si < len(sched) else "?"
            last = entry
            si += 1
        rec = {"entry": entry, "kernel": short}
        for m, ci in cols.items():
            v = r[ci]
            rec[m] = to_bytes(v, units[ci]) if m.startswith("dram__bytes") else num(v)
            if m == "gpu__time_duration.sum" and rec[m] is not None:
                rec[m] = rec[m] * {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(units[ci], 1.0)   # -> us
        out_rows.append(rec)
        if rec.get("dram__bytes_read.sum") is not None:
            traffic[entry] = traffic.get(entry, 0.0) + rec["dram__bytes_read.sum"] + rec["dram__bytes_write.sum"]
    # the decode/NMS pair is one roofline entry in bench.py
    traffic["detect (decode scan + per-class NMS)"] = traffic.get("detect: decode scan", 0.0) + traffic.get("detect: per-class NMS", 0.0)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"r02_ncu_full_{tag}.csv"), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["entry", "kernel"] + [m + (" [us]" if m.startswith("gpu__time") else " [bytes]" if m.startswith("dram__bytes") else "") for m in cols])
        for rec in out_rows:
            w.writerow([rec["entry"], rec["kernel"]] + [("" if rec[m] is None else (f"{rec[m]:.0f}" if abs(rec[m]) >= 100 else f"{rec[m]:.3f}")) for m in cols])
    with open(os.path.join(ROOT, "profiles", f"r02_traffic_{tag}.json"), "w") as fh:
        json.dump({k: round(v) for k, v in traffic.items()}, fh, indent=1)
    tot = sum(r["gpu__time_duration.sum"] or 0 for r in out_rows)
    for rec in out_rows:
        print(f"{rec['entry'][:34]:34s} {rec['kernel'][:28]:28s} {rec['gpu__time_duration.sum'] or 0:8.2f} us  dram {100 * (rec.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed') or 0) / 100:5.1f}%  "
              f"tensor {rec.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active') or 0:5.1f}%  warps {rec.get('sm__warps_active.avg.pct_of_peak_sustained_active') or 0:5.1f}%")
    print(f"total {tot:.1f} us over {len(out_rows)} launches")


if __name__ == "__main__":
    main()
