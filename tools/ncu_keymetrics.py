"""Print the handful of `ncu --set full` metrics the roofline discussion uses, per profiled launch.

    python tools/ncu_keymetrics.py report.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.max.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "smsp__cycles_active.avg", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_lsu.sum"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    for n, r in enumerate(rows[2:]):
        print("== launch %d: %s" % (n, r[ik][:150]))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("   %-70s %s %s" % (k, r[i], units[i]))


if __name__ == "__main__":
    main()
