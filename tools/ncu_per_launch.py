"""One row per captured launch of an ncu --set full report: python tools/ncu_per_launch.py <rep> <out.csv>
(the format of profiles/r0N_wide_chain_ncu_full_per_launch.csv, which bench.py reads roofline.traffic from)."""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic"]
scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def value(r, m):
    v, u = float(r[col[m]].replace(",", "")), units[col[m]]
    if m.startswith("dram__bytes") or m == "gpu__time_duration.sum":
        return v * scale.get(u, 1.0)
    return v


def role(name, n_chain):
    if "chain_tc" in name:
        return "forward chain" if n_chain % 2 == 0 else "adjoint chain"
    if "dw_tc" in name:
        return "weight gradient + column sums of one layer"
    return ""


with open(out, "w") as f:
    f.write("launch,role,kernel,duration [us],dram bytes read [byte],dram bytes written [byte]," + ",".join("%s [%s]" % (m, units[col[m]]) for m in want[3:]) + "\n")
    n_chain = 0
    for i, r in enumerate(rows[2:]):
        name = r[col["Kernel Name"]]
        short = name.split("(")[0].split("::")[-1].split("<")[0]
        f.write("%d,%s,%s,%s\n" % (i, role(name, n_chain), short, ",".join("%.6f" % value(r, m) if m in want[:3] or "." in r[col[m]] else r[col[m]] for m in want)))
        n_chain += "chain_tc" in name
print("wrote", out, len(rows) - 2, "launches")
