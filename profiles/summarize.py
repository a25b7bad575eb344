"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under
profiles/:   python profiles/summarize.py <round tag> <launches.csv> <full.ncu-rep> [workload]
  <tag>_launches.md   per-kernel launch count / avg device time / share of the step (gpu__time_duration pass)
  <tag>_kernels.md    per-kernel metrics of the --set full capture (DRAM bytes, throughput %, tensor pipe %, ...)
  <tag>_traffic.json  dram__bytes_read+write summed over the kernels of ONE step (bench.py's roofline.traffic)"""
import csv
import json
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = {}
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg.setdefault(r[ki], []).append(v)
    return agg


def short(name):
    name = name.replace("kge::<unnamed>::", "").replace("kge::", "").replace("void ", "")
    return name.split("(")[0][:60]


def main():
    tag, lcsv, rep = sys.argv[1], sys.argv[2], sys.argv[3]
    workload = sys.argv[4] if len(sys.argv) > 4 else "fb15k_transe_l2"
    agg = launches(lcsv)
    tot = sum(sum(v) for v in agg.values())
    with open("profiles/%s_launches.md" % tag, "w") as f:
        f.write("# %s: every launch with its device time (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n" % tag)
        f.write("Cold-cache, serialised launches: compare SHARES, not absolutes.\n\n| kernel | launches | avg us | share |\n|---|---|---|---|\n")
        for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.2f | %.3f |\n" % (short(n), len(v), sum(v) / len(v), sum(v) / tot))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]
    idx = [(h.index(m), lbl) for m, lbl in want if m in h]
    units = rows[1]
    ki = h.index("Kernel Name")
    per_kernel = {}
    with open("profiles/%s_kernels.md" % tag, "w") as f:
        f.write("# %s: ncu --set full --clock-control none (one row per captured launch)\n\n" % tag)
        f.write("| kernel | " + " | ".join("%s [%s]" % (lbl, units[i]) for i, lbl in idx) + " |\n")
        f.write("|---|" + "---|" * len(idx) + "\n")
        for r in rows[2:]:
            f.write("| `%s` | " % short(r[ki]) + " | ".join(r[i] for i, _ in idx) + " |\n")
            try:
                rd, wr = float(r[h.index("dram__bytes_read.sum")]), float(r[h.index("dram__bytes_write.sum")])
                u = units[h.index("dram__bytes_read.sum")]
                mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                per_kernel.setdefault(short(r[ki]), []).append((rd + wr) * mul)
            except Exception:
                pass
    # traffic of one step = sum over kernels of (avg bytes per launch) x (launches per step in the launch list)
    steps = max(1, min(len(v) for n, v in agg.items() if "k_chain" in n or "k_rescal_bwd" in n)) if agg else 1
    per_step = 0.0
    detail = {}
    for n, v in agg.items():
        s = short(n)
        if s in per_kernel:
            b = sum(per_kernel[s]) / len(per_kernel[s]) * (len(v) / steps)
            per_step += b
            detail[s] = b
    try:
        cur = json.load(open("profiles/%s_traffic.json" % tag))
    except Exception:
        cur = {}
    cur[workload] = per_step
    cur[workload + "_detail"] = detail
    json.dump(cur, open("profiles/%s_traffic.json" % tag, "w"), indent=1)
    print("wrote profiles/%s_{launches,kernels}.md, traffic %.1f MB/step" % (tag, per_step / 1e6))


if __name__ == "__main__":
    main()
