"""Condenses an `ncu --set full` report into the table profiles/ keeps:  python tools/summarize_ncu.py gpurun_out/r2_full.ncu-rep > profiles/r2_full_summary.txt
(reads the report with `ncu -i ... --page raw --csv`, or that export itself when given a .csv; no GPU needed)."""
import csv
import io
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time_us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", 1e-6),
    ("dram__bytes_write.sum", "dram_wr_MB", 1e-6),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%", 1),
    ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "tensor_hmma_%", 1),
    ("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_act_%", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%", 1),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wavefront_%", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%", 1),
    ("l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "lsu_wb_%", 1),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "lsu_wavefronts", 1),
    ("smsp__inst_executed_pipe_lsu.sum", "lsu_inst", 1),
    ("lts__t_sectors_op_read.sum", "l2_rd_sectors", 1),
    ("lts__t_sector_hit_rate.pct", "l2_hit_%", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_%", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
    ("launch__block_size", "block", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%", 1),
]


def main(path):
    if path.endswith(".csv"):          # already exported on the GPU box (`ncu -i report --page raw --csv`)
        raw = open(path).read()
    else:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units, data = rows[0], rows[1], rows[2:]
    col = {name: i for i, name in enumerate(header)}
    name_i = col.get("Kernel Name")
    have = [(m, label, scale) for m, label, scale in WANT if m in col]
    print("# " + path)
    print("# metrics: " + ", ".join(f"{label}={m} [{units[col[m]]}]" for m, label, _ in have))
    for r in data:
        if len(r) <= name_i:
            continue
        kernel = re.sub(r"\(.*", "", r[name_i]).replace("gcd::<unnamed>::", "").replace("void ", "")[:70]
        vals = []
        for m, label, scale in have:
            v = r[col[m]].replace(",", "")
            try:
                unit = units[col[m]]        # the export picks its own unit per column
                scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, scale if unit in ("", "%") else 1)
                x = float(v) * scale
                vals.append(f"{label}={x:.4g}")
            except ValueError:
                vals.append(f"{label}={v}")
        print(f"{kernel:70s} " + " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
