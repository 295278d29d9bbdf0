#!/usr/bin/env python
"""Print the judged metrics of every kernel in an .ncu-rep (duration, DRAM bytes, tensor / XU / L2 activity, registers)."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        # tcgen05 on B200: the legacy sm__inst_executed_pipe_tensor* / sm__ops_path_tensor_op_hmma* counters stay at 0; the tensor
        # pipe's busy cycles show up in the TriageCompute group (realtime counters) and the TMEM pipe has its own instruction counter
        "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.avg.per_second"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        print(f"== {r[ki][:100]}")
        vals = {}
        for h, u, v in zip(hdr, units, r):
            if h in WANT:
                print(f"   {h:75s} {v:>16s} {u}")
                vals[h] = v
        try:  # tensor-pipe utilisation = busy cycles of the (h)mma sub-pipe, which executes tcgen05.mma kind::f16, over elapsed cycles
            busy = float(vals["TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg"].replace(",", ""))
            el = float(vals["sm__cycles_elapsed.avg"].replace(",", ""))
            # the TriageCompute counter is per TPC (two SMs): halve it for the per-SM busy fraction
            print(f"   {'tensor pipe busy per SM (TPC hmma sub-pipe busy cycles / 2 / elapsed cycles)':75s} {50.0 * busy / el:16.2f} %")
        except (KeyError, ValueError, ZeroDivisionError):
            pass


if __name__ == "__main__":
    main(sys.argv[1])
