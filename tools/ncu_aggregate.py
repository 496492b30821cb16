#!/usr/bin/env python
"""Aggregate per-launch ncu summaries (tools/ncu_summary.py output) by kernel flavour (name, grid) into the table the
roofline discussion cites, and write the per-kernel DRAM-traffic files bench.py attaches to its `roofline` object.

    python tools/ncu_aggregate.py profiles/r02 gpurun_out/r02/ncu_x3_step.json gpurun_out/r02/ncu_misc_step.json ...
"""
import collections
import json
import os
import sys

PEAK_HBM_GBS = 6575.4      # MEASURED_PEAKS.json hbm_gbs on this pool (copy bandwidth)
WORKLOAD = {"workload": "egolane_2lane_b32_256x512_fp32", "batch": 32}
CAPI = {"conv1d_tc_x3_kernel": "lf_conv1d_tc_x3", "wgrad_tc_x3_kernel": "lf_wgrad3_tc_x3"}


def main(outdir, files):
    rows = []
    per_capi = collections.defaultdict(list)
    for f in files:
        d = json.load(open(f))
        groups = collections.OrderedDict()
        for l in d["launches"]:
            name = l["kernel"].replace("void ", "").replace("lf::", "")
            groups.setdefault((name, int(l.get("grid", 0))), []).append(l)
            for k, capi in CAPI.items():
                if k in name:
                    per_capi[capi].append(l)
        for (name, grid), ls in groups.items():
            n = len(ls)
            avg = lambda key: sum(x.get(key, 0.0) for x in ls) / n
            dur = avg("duration_us")
            traffic = avg("dram_read") + avg("dram_write")
            rows.append({"source": os.path.basename(f), "kernel": name, "grid": grid, "launches": n, "duration_us": round(dur, 2),
                         "dram_read_MB": round(avg("dram_read") / 1e6, 2), "dram_write_MB": round(avg("dram_write") / 1e6, 2),
                         "dram_GBps": round(traffic / dur / 1e3, 1) if dur else None,
                         "dram_frac_of_measured_peak": round(traffic / dur / 1e3 / PEAK_HBM_GBS, 3) if dur else None,
                         "ncu_dram_throughput_pct": round(avg("dram_throughput_pct"), 1),
                         "tensor_pipe_active_pct": round(avg("tensor_pipe_active_pct"), 1),
                         "sm_throughput_pct": round(avg("sm_throughput_pct"), 1), "l2_throughput_pct": round(avg("l2_throughput_pct"), 1),
                         "registers": int(avg("registers")), "block": int(avg("block"))})
    how = ("ncu --set full --clock-control none, one eager step of `bench.py --steps 1 --warmup 1 --no-graph` (config 2, batch 32) / "
           "tools/bench_lsq.py; per-launch numbers are cold-cache and serialised (shares, not absolutes, carry over to the graph "
           "replay); averaged over the launches of each (kernel, grid) flavour")
    json.dump({"how": how, "hbm_peak_GBps": PEAK_HBM_GBS, "flavours": rows}, open(os.path.join(outdir, "ncu_kernel_flavours.json"), "w"), indent=1)
    with open(os.path.join(outdir, "ncu_kernel_flavours.md"), "w") as md:
        md.write("# ncu --set full summaries by kernel flavour (round 2)\n\n%s.\n\n" % how)
        md.write("| kernel | grid | launches | us | DRAM read MB | write MB | DRAM GB/s | frac of %.0f | tensor pipe %% | SM %% | L2 %% | regs |\n" % PEAK_HBM_GBS)
        md.write("|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            md.write("| %s | %d | %d | %.1f | %.1f | %.1f | %s | %s | %.1f | %.1f | %.1f | %d |\n" % (
                r["kernel"][:60], r["grid"], r["launches"], r["duration_us"], r["dram_read_MB"], r["dram_write_MB"], r["dram_GBps"],
                r["dram_frac_of_measured_peak"], r["tensor_pipe_active_pct"], r["sm_throughput_pct"], r["l2_throughput_pct"], r["registers"]))
    for capi, ls in per_capi.items():
        n = len(ls)
        json.dump(dict(WORKLOAD, kernel=capi, launches_captured=n,
                       mean_traffic_bytes_per_launch=sum(x["dram_read"] + x["dram_write"] for x in ls) / n,
                       mean_duration_us=sum(x["duration_us"] for x in ls) / n,
                       mean_tensor_pipe_active_pct=sum(x.get("tensor_pipe_active_pct", 0) for x in ls) / n,
                       how=how), open(os.path.join(outdir, "ncu_%s.json" % capi), "w"), indent=1)
    print("wrote", len(rows), "flavours")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
