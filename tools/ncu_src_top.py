"""Read an .ncu-rep (here, no GPU needed): key raw metrics + the instructions with most stall samples, grouped by
mnemonic class, to see which warp role waits on what.   python tools/ncu_src_top.py gpurun_out/x.ncu-rep [n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
h, u = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__inst_executed_pipe_tma.sum", "smsp__inst_executed_pipe_tma.sum",
        "launch__grid_size", "launch__block_size", "Context", "ID"]
for k, r in enumerate(rows[2:]):
    if len(r) != len(h):
        continue
    print(f"==== launch {k} (raw metrics)")
    for i, name in enumerate(h):
        if name in want:
            print(f"{name} = {r[i]} {u[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
heads = [i for i, x in enumerate(rows) if "Source" in x and "# Samples" in x]
for n, hi in enumerate(heads):
    h = rows[hi]
    end = heads[n + 1] - 1 if n + 1 < len(heads) else len(rows)     # the row before the next header names the kernel
    name = rows[hi - 1][1] if hi > 0 and len(rows[hi - 1]) > 1 else "?"
    si, so, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    data = []
    for k, x in enumerate(rows[hi + 1:end]):
        if len(x) <= si or not x[si].isdigit():
            continue
        st = {h[i]: int(x[i]) for i in stall_cols if i < len(x) and x[i].isdigit() and x[i] != "0"}
        data.append((int(x[si]), x[so].strip(), int(x[ie] or 0), k, st))
    tot = sum(d[0] for d in data) or 1
    print(f"==== launch {n} (source page): {name}\ntotal samples {tot}")
    for d in sorted(data, key=lambda d: -d[0])[:topn]:
        top = sorted(d[4].items(), key=lambda kv: -kv[1])[:3]
        print(f"{d[0]:6d} {100 * d[0] / tot:5.1f}% exec={d[2]:9d} #{d[3]:5d} {d[1][:70]:70s} {top}")
