"""Summarise an `ncu --set full` report for profiles/: the details page as a flat CSV (id,kernel,section,metric,unit,value)
plus a compact key-metric table on stdout.   usage: python tools/ncu_summary.py report.ncu-rep out_summary.csv"""
import csv
import subprocess
import sys

KEY = [("gpu__time_duration.sum", "duration"),
       ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
       ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_active_pct"),
       ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_smem_feed_pct"),
       ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
       ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
       ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
       ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "smem_dyn"),
       ("launch__grid_size", "grid"), ("launch__occupancy_limit_shared_mem", "occ_limit_smem"),
       ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
       ("smsp__pcsamp_warps_issue_stalled_long_scoreboard", "stall_long_scoreboard"),
       ("smsp__pcsamp_warps_issue_stalled_barrier", "stall_barrier"),
       ("smsp__pcsamp_warps_issue_stalled_branch_resolving", "stall_branch"),
       ("smsp__pcsamp_warps_issue_stalled_wait", "stall_wait"),
       ("smsp__pcsamp_warps_issue_stalled_lg_throttle", "stall_lg_throttle"),
       ("smsp__pcsamp_warps_issue_stalled_short_scoreboard", "stall_short_scoreboard"),
       ("smsp__pcsamp_sample_buffers", "samples")]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    det = page(rep, "details")
    h = det[0]
    ix = {k: h.index(k) for k in ("ID", "Kernel Name", "Section Name", "Metric Name", "Metric Unit", "Metric Value")}
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "section", "metric", "unit", "value"])
        for r in det[1:]:
            if len(r) <= ix["Metric Value"] or not r[ix["Metric Name"]]:
                continue
            w.writerow([r[ix["ID"]], r[ix["Kernel Name"]], r[ix["Section Name"]], r[ix["Metric Name"]], r[ix["Metric Unit"]],
                        r[ix["Metric Value"]]])
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]

    def col(name):
        if name in hdr:
            return hdr.index(name)
        for i, x in enumerate(hdr):
            if x.endswith(name):
                return i
        return -1
    for r in raw[2:]:
        print(r[hdr.index("Kernel Name")])
        for m, short in KEY:
            i = col(m)
            if i >= 0 and r[i] != "":
                print("   %-26s %s %s" % (short, r[i], units[i]))


if __name__ == "__main__":
    main()
