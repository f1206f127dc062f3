"""Extracts the dominant conv kernel's DRAM traffic / tensor-pipe activity from an `ncu --set full` report
(read here with `ncu -i ... --page raw --csv`) into profiles/dominant_kernel_ncu.json, which bench.py quotes as
`roofline.traffic` / `tensor_pipe_util.ncu_*`.

    python tools/ncu_to_json.py gpurun_out/r02b_prof.ncu-rep "conv_tc2h_kernel<(int)64, (int)1, (bool)0, (bool)1>" \
        profiles/dominant_kernel_ncu.json
"""
import csv
import io
import json
import subprocess
import sys


def main():
    rep, pattern, out = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        v = r[col[name]].replace(",", "")
        return float(v) if v else None

    def scale(name):   # bytes per unit
        u = units[col[name]].lower()
        return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)

    picked = [r for r in rows[2:] if pattern in r[col["Kernel Name"]]]
    if not picked:
        raise SystemExit(f"no kernel matching {pattern!r} in {rep}")
    rd = [val(r, "dram__bytes_read.sum") * scale("dram__bytes_read.sum") for r in picked]
    wr = [val(r, "dram__bytes_write.sum") * scale("dram__bytes_write.sum") for r in picked]
    rec = {
        "kernel": picked[0][col["Kernel Name"]][:120],
        "launches": len(picked),
        "dram_bytes_read_per_launch": sum(rd) / len(rd),
        "dram_bytes_write_per_launch": sum(wr) / len(wr),
        "dram_bytes_per_launch": (sum(rd) + sum(wr)) / len(rd),
        "duration_us": sum(val(r, "gpu__time_duration.sum") for r in picked) / len(picked),
        "pipe_tensor_cycles_active_pct": sum(
            val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") for r in picked) / len(picked),
        "tensor_inst_pct_of_peak_while_active": 100.0 * sum(
            val(r, "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active")
            for r in picked) / len(picked),
        "source": f"ncu --set full --clock-control none ({rep.split('/')[-1]}), "
                  "dram__bytes_read.sum + dram__bytes_write.sum per launch, cold cache, one kernel at a time",
    }
    with open(out, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
