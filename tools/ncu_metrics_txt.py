"""One line of key metrics per kernel launch of an `ncu --set full` report (read here, without a GPU, through
`ncu -i ... --page raw --csv`): the format of profiles/r0N_ncu_metrics.txt.

    python tools/ncu_metrics_txt.py gpurun_out/r02i_prof.ncu-rep "title" > profiles/r02_ncu_metrics.txt
"""
import csv
import io
import subprocess
import sys

COLS = [  # (label, ncu metric)
    ("time_us", "gpu__time_duration.sum"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("tensor_inst_pct_while_active", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"),
    ("tensor_cycles_active_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("sm_busy_pct", "sm__instruction_throughput.avg.pct_of_peak_sustained_active"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("l2_to_sm_sectors", "lts__t_sectors_srcunit_tex.sum"),
    ("regs", "launch__registers_per_thread"),
    ("occupancy_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("sm_cycles_active", "sm__cycles_active.avg"),
    ("sm_cycles_elapsed", "sm__cycles_elapsed.max"),
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    have = [(lab, m) for lab, m in COLS if m in col]
    print(f"# {title}")
    print("# units: " + ", ".join(f"{lab}[{units[col[m]]}]" for lab, m in have))
    for r in rows[2:]:
        name = r[col["Kernel Name"]].replace("b200::", "")[:78]
        parts = [f"kernel={name}", f"grid={r[col['Grid Size']]}"]
        for lab, m in have:
            parts.append(f"{lab}={r[col[m]].replace(',', '')}")
        print(" | ".join(parts))


if __name__ == "__main__":
    main()
