"""Digest an .ncu-rep: headline metrics + hottest SASS lines.  usage: python scripts/ncu_digest.py file.ncu-rep [ntop]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "sm__cycles_elapsed.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("== kernel:", d.get("Kernel Name", "?")[:80])
    for k in want:
        if k in d:
            print("   %-75s %s %s" % (k, d[k], units[hdr.index(k)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
S, E = idx["# Samples"], idx["Instructions Executed"]
tot = sum(int(r[S]) for r in data)
print("total samples", tot)
for r in sorted(data, key=lambda r: -int(r[S]))[:ntop]:
    st = {h[6:]: int(r[idx[h]]) for h in hdr if h.startswith("stall_") and "(" not in h and r[idx[h]] not in ("0", "")}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print("%5.1f%% %10s  %-70s %s" % (100.0 * int(r[S]) / tot, r[E], r[idx["Source"]][:70], st))
