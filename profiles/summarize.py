"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/:

    python profiles/summarize.py <tag> <launches.csv> <full.ncu-rep> <traffic key> [launches per step]

  <tag>_launches.md   per-kernel launch count / avg device time / share of the step (gpu__time_duration pass)
  <tag>_kernels.md    per-kernel metrics of the --set full capture (DRAM bytes, throughput %, tensor pipe %, ...)
  <tag>_traffic.json  {<traffic key>: dram__bytes_read+write of ONE step, <key>_detail: per kernel, <key>_missing: [...]}

bench.py reads <tag>_traffic.json and reports roofline.traffic only for the exact key (workload | batch | schedule) the
capture was taken on -- null otherwise.  The traffic of a step is the sum over the kernels of the launch list of
(average bytes per captured launch) x (launches of that kernel per step); a kernel of the launch list without a full
capture is listed under <key>_missing and makes the total null (a partial sum would under-report)."""
import csv
import json
import re
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
        agg.setdefault(short(r[ki]), []).append(v)
    return agg


def short(name):
    """Canonical kernel label: namespaces (kge::, the anonymous namespace in either spelling) and the argument list
    dropped, template arguments kept (k_fused<0> and k_fused<1> are different kernels)."""
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|unnamed>::|kge::|^void ", "", name)
    depth, out = 0, []
    for ch in name:                      # cut at the first '(' outside template brackets
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            break
        out.append(ch)
    return "".join(out).strip()[:80]


def main():
    tag, lcsv, rep, key = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
    agg = launches(lcsv)
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    if not steps:   # every step runs k_chain (or the RESCAL backward) exactly once
        steps = max(1, min([len(v) for n, v in agg.items() if n.startswith("k_chain") or n.startswith("k_rescal_bwd")] or [1]))
    tot = sum(sum(v) for v in agg.values())
    with open("profiles/%s_launches.md" % tag, "w") as f:
        f.write("# %s: every launch with its device time (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n" % tag)
        f.write("Cold-cache, serialised launches: compare SHARES, not absolutes.  %d steps in the list.\n\n" % steps)
        own = {n: v for n, v in agg.items() if n.startswith("k_") and len(v) >= steps}
        tot_own = sum(sum(v) for v in own.values())
        f.write("Share = of the step's own kernels (k_*, launched every step); the other rows are the bench harness (table "
                "initialisation, L2 flush, profiling spin) and one-time set-up.\n\n")
        f.write("| kernel | launches | per step | avg us | share of the step |\n|---|---|---|---|---|\n")
        for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.2f | %.2f | %s |\n" % (n, len(v), len(v) / steps, sum(v) / len(v),
                                                           "%.3f" % (sum(v) / tot_own) if n in own else "-"))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h = rows[0]
    want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("lts__t_sector_hit_rate.pct", "L2 hit %"),
            ("lts__t_bytes.sum", "L2 bytes"),
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
                rd, wr = float(r[h.index("dram__bytes_read.sum")].replace(",", "")), float(r[h.index("dram__bytes_write.sum")].replace(",", ""))
                mul = lambda col: {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[h.index(col)], 1)
                per_kernel.setdefault(short(r[ki]), []).append(rd * mul("dram__bytes_read.sum") + wr * mul("dram__bytes_write.sum"))
            except Exception:
                pass
    # the kernels of a STEP: the library's own (k_*) that run every step.  The bench harness's kernels (torch's table
    # initialisation, the 256 MiB L2-flush fill, the profiling spin) and one-time set-up kernels (fewer launches than steps)
    # are listed in the launch table but are not part of the step's traffic.
    detail, missing = {}, []
    for n, v in agg.items():
        if not n.startswith("k_") or len(v) < steps:
            continue
        if n in per_kernel:
            detail[n] = sum(per_kernel[n]) / len(per_kernel[n]) * (len(v) / steps)
        else:
            missing.append(n)
    total = None if missing else sum(detail.values())
    assert total is None or abs(total - sum(detail.values())) < 1e-6
    try:
        cur = json.load(open("profiles/%s_traffic.json" % tag))
    except Exception:
        cur = {}
    cur[key] = total
    cur[key + "_detail"] = detail
    cur[key + "_missing"] = missing
    json.dump(cur, open("profiles/%s_traffic.json" % tag, "w"), indent=1)
    print("wrote profiles/%s_{launches,kernels}.md; traffic[%s] = %s MB/step (%d kernels, missing from the full capture: %s)"
          % (tag, key, "%.1f" % (total / 1e6) if total is not None else "null", len(detail), missing))


if __name__ == "__main__":
    main()
