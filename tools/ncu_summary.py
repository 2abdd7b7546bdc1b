"""Extract the handful of ncu metrics the roofline discussion needs from a .ncu-rep (run where ncu is installed).

    python tools/ncu_summary.py gpurun_out/prof_tapgemm_layer4.ncu-rep > profiles/r1_tapgemm_layer4.txt
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__cycles_active.avg", "sm__cycles_active.avg",
    "sm__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("kernel:", r[col["Kernel Name"]][:110])
        for k in KEYS:
            if k in col:
                print(f"  {k:75s} {r[col[k]]:>18s} {units[col[k]]}")
        extra = [h for h in hdr if ("tensor" in h and "pct_of_peak_sustained_active" in h and r[col[h]] not in ("0", ""))]
        for h in extra:
            if h not in KEYS:
                print(f"  {h:75s} {r[col[h]]:>18s} {units[col[h]]}")
        print()


if __name__ == "__main__":
    main()
