"""Summarise an ncu report (--set full) into one line per profiled launch with the metrics the
roofline uses:  python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/ncu_<name>.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "dur_us"),
    ("sm__cycles_elapsed.avg.per_second", "sm_ghz"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "l2_read_sectors"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for d in data:
        name = d[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("msig::", "")
        parts = [f"{name:28s}"]
        dur_s = None
        for key, label in WANT:
            if key not in idx:
                continue
            v, u = d[idx[key]].replace(",", ""), units[idx[key]]
            try:
                f = float(v)
            except ValueError:
                continue
            if label == "dur_us":
                f = f / 1e3 if u == "ns" else (f * 1e3 if u == "ms" else f)
                dur_s = f * 1e-6
                parts.append(f"{label}={f:.1f}")
            elif label in ("dram_read", "dram_write"):
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                parts.append(f"{label}_MB={f * scale / 1e6:.1f}")
                if dur_s:
                    parts.append(f"{label}_GBs={f * scale / dur_s / 1e9:.0f}")
            elif label == "l2_read_sectors":
                parts.append(f"l2_to_sm_MB={f * 32 / 1e6:.0f}")
                if dur_s:
                    parts.append(f"l2_to_sm_TBs={f * 32 / dur_s / 1e12:.2f}")
            else:
                parts.append(f"{label}={f:.4g}")
        print("  ".join(parts))


if __name__ == "__main__":
    main()
