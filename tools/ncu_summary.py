#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into one row per launch: duration, DRAM bytes, throughput %,
occupancy, registers, pipe utilisation (xu% = the pipe POPC / F2I / MUFU issue on: the "integer-pipe utilisation for matching"
BASELINE.json's metric asks for), warp- and thread-level instruction counts, shared-memory wavefronts and bank conflicts.
Usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [out.md [frames_per_launch]]
The first line of out.md records the frames one launch processed (bench.py scales DRAM bytes per frame from it)."""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("smsp__inst_executed.sum", "inst"),
    ("sass__thread_inst_executed_true_per_opcode", "tinst"),      # thread-level instructions executed (predicated-off lanes not counted)
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
]


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    head = ["kernel"] + [s for _, s in WANT]
    out.append("| " + " | ".join(head) + " |")
    out.append("|" + "---|" * len(head))
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("nav24::", "")[:28]
        cells = [name]
        for m, s in WANT:
            if m not in col:
                cells.append("-"); continue
            v, u = r[col[m]], units[col[m]]
            try:
                if s == "dur":
                    cells.append(f"{to_us(v, u):.1f}us")
                elif s.startswith("dram_"):
                    cells.append(f"{to_bytes(v, u) / 1e6:.1f}MB")
                elif s in ("inst", "tinst", "smem_conflicts", "smem_wavefronts"):
                    cells.append(f"{float(v.replace(',', '')) / 1e6:.1f}M")
                else:
                    cells.append(f"{float(v.replace(',', '')):.1f}" if "." in v else v)
            except ValueError:
                cells.append(v)
        out.append("| " + " | ".join(cells) + " |")
    s = "\n".join(out)
    if len(sys.argv) > 3:
        s = f"<!-- frames_per_launch: {int(sys.argv[3])} -->\n" + s
    print(s)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(s + "\n")


if __name__ == "__main__":
    main()
