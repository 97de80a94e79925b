#!/usr/bin/env python3
"""Summarises an ncu report (full set) and a launch list (gpu__time_duration) into profiles/*.md / *.csv.

    python tools/ncu_summary.py gpurun_out/prof_r1.ncu-rep gpurun_out/launches_r1.csv r1
"""
import collections
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in WANT if c in idx]
    per = collections.OrderedDict()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("unnamed>::", "").strip()
        per.setdefault(name, []).append(r)
    out = [f"# ncu --set full --clock-control none, round {tag}", "",
           "Command: `python tools/ncu_step.py 3` (eager launches of the bench step: configs[1], M ~ 721k samples/step, 4096 rays).",
           "One row per kernel = mean over the captured launches. Times are ncu's serialised cold-clock times: compare",
           "shares, not absolutes (bench.py times with CUDA events).", ""]
    out.append("| kernel | launches | " + " | ".join(f"{c} [{units[idx[c]]}]" for c in cols) + " |")
    out.append("|---|---|" + "---|" * len(cols))
    for name, rs in per.items():
        vals = []
        for c in cols:
            xs = []
            for r in rs:
                try:
                    xs.append(float(r[idx[c]].replace(",", "")))
                except ValueError:
                    pass
            vals.append(f"{sum(xs) / len(xs):.4g}" if xs else "-")
        out.append(f"| {name} | {len(rs)} | " + " | ".join(vals) + " |")
    # launch list shares
    lrows = list(csv.reader(open(launches)))
    h = [i for i, r in enumerate(lrows) if r and r[0] == "ID"][0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in lrows[h + 1:]:
        if len(r) < 15:
            continue
        v = float(r[14].replace(",", ""))
        v = v / 1000 if r[13] == "ns" else (v * 1000 if r[13] == "ms" else v)
        name = r[4].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:80]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    out += ["", f"## Launch list ({launches.split('/')[-1]}): {sum(v[0] for v in agg.values())} launches, {tot / 1000:.2f} ms total", "",
            "Every launch of the same command (`python tools/ncu_step.py 4`: scene set-up + 4 eager training steps).", "",
            "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        out.append(f"| {k} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
    # the kernels of ONE training step (FusedTrainStep): average launch time and share of the step
    step_names = ["march_train_count_coop_kernel", "march_train_write_coop_kernel", "field_forward_ws_kernel", "composite_train_mse_kernel",
                  "field_backward_ws_kernel", "check_finite_kernel", "fused_adam_kernel"]
    per_step = []
    for sn in step_names:
        hits = [(k, v) for k, v in agg.items() if sn in k]
        if hits:
            n = sum(v[0] for _, v in hits)
            t = sum(v[1] for _, v in hits)
            calls = 2 if sn in ("check_finite_kernel", "fused_adam_kernel") else 1     # table + MLP weights
            per_step.append((sn, t / n * calls))
    if per_step:
        tot_step = sum(t for _, t in per_step)
        out += ["", f"## One training step from the launch list: {tot_step:.1f} us of kernel time", "",
                "| kernel | us per step | share |", "|---|---|---|"]
        for sn, t in per_step:
            out.append(f"| {sn} | {t:.1f} | {100 * t / tot_step:.1f}% |")
    open(f"profiles/ncu_summary_{tag}.md", "w").write("\n".join(out) + "\n")
    print("\n".join(out[:14]))


if __name__ == "__main__":
    main()
