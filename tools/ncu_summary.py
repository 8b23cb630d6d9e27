"""Summarise an Nsight Compute report (read on the CPU box with `ncu -i`) into the small files kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/round1_cfg2_tiled [--traffic profiles/traffic_cfg2.json]

Writes <out>.csv (selected raw metrics of every captured launch) and <out>.md (a readable digest); with --traffic also
the dram bytes per launch that bench.py reports as roofline.traffic.
"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    traffic = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    cols = [hdr.index(k) for k in KEEP if k in hdr]
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + [hdr[c] for c in cols])
        w.writerow([""] + [units[c] for c in cols])
        for r in data:
            w.writerow([r[name_i]] + [r[c] for c in cols])
    with open(out + ".md", "w") as f:
        f.write("# ncu --set full digest of `%s`\n\n" % rep)
        for r in data:
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % r[name_i][:90])
            for c in cols:
                f.write("| %s | %s | %s |\n" % (hdr[c], r[c], units[c]))
            f.write("\n")
    if traffic and data:
        def val(k, r):
            v = float(r[hdr.index(k)])
            u = units[hdr.index(k)].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        per = [val("dram__bytes_read.sum", r) + val("dram__bytes_write.sum", r) for r in data]
        if "--per-step" in sys.argv:      # the captured launches are the kernels of ONE step (e.g. masked + dense): sum them
            json.dump({"dram_bytes_per_launch": sum(per), "launches": len(per), "source": rep, "per_kernel": per,
                       "kernel": " + ".join(r[name_i][:40] for r in data)}, open(traffic, "w"))
        else:
            json.dump({"dram_bytes_per_launch": sum(per) / len(per), "launches": len(per), "source": rep,
                       "kernel": data[0][name_i][:60]}, open(traffic, "w"))
    print("wrote", out + ".csv", out + ".md", traffic or "")


if __name__ == "__main__":
    main()
