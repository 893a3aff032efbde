"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small text file for profiles/."""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    lines = []
    for v in rows[2:]:
        name = v[h.index("Kernel Name")] if "Kernel Name" in h else "?"
        lines.append(f"kernel: {name}")
        for k in KEYS:
            if k in h:
                lines.append(f"  {k} = {v[h.index(k)]} {units[h.index(k)]}")
        for i, k in enumerate(h):
            if "issue_stalled" in k and "per_issue_active" in k:
                try:
                    if float(v[i]) > 0.05:
                        lines.append(f"  stall {k.split('stalled_')[1].split('_per')[0]} = {float(v[i]):.3f} warps/issue")
                except ValueError:
                    pass
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(sass)))
    hi = [i for i, r in enumerate(rows) if "# Samples" in r]
    if hi:
        hdr = rows[hi[0]]
        ix = {x: i for i, x in enumerate(hdr)}
        stop = hi[1] if len(hi) > 1 else len(rows)   # several kernels in one report: the first kernel's listing only
        data = [r for r in rows[hi[0] + 1:stop] if len(r) == len(hdr) and r != hdr]
        cls = {}
        for r in data:
            src = r[ix["Source"]].strip()
            op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
            c = cls.setdefault(op, [0, 0])
            c[0] += int(r[ix["Instructions Executed"]] or 0)
            c[1] += 1
        tot = sum(c[0] for c in cls.values())
        lines.append("  SASS mix (executed warp instructions, first kernel): " + ", ".join(
            f"{op} {100 * c[0] / tot:.1f}%" for op, c in sorted(cls.items(), key=lambda x: -x[1][0])[:10]))
        lines.append("  SASS evidence: " + ", ".join(f"{op} x{cls[op][1]}" for op in ("DMMA", "UBLKCP", "LDGSTS", "SYNCS",
                                                                                      "UTMALDG") if op in cls))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
