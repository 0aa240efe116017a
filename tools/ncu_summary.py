"""Developer tool: print the key metrics, stall reasons and SASS opcode mix of every kernel in an
.ncu-rep (read here on the CPU box with `ncu -i`).  usage: python tools/ncu_summary.py rep [regex]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H = rows[0]
idx = {h: i for i, h in enumerate(H)}
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_shared_atom.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warp_latency_per_inst_issued.ratio"]
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    if pat and not re.search(pat, name):
        continue
    print("=====", name[:100])
    for w in WANT:
        if w in idx:
            print(f"  {w:70s} {r[idx[w]]}")
    for h in H:
        if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                v = float(r[idx[h]].replace(",", ""))
            except ValueError:
                continue
            if v > 0.08:
                print(f"  stall {h.split('stalled_')[1].split('_per_issue')[0]:30s} {v:.3f}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name",
                          "regex:" + re.escape(name.split("(")[0].split("<")[0].split()[-1].split("::")[-1])],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    hdr = next((i for i, x in enumerate(srows) if x and x[0] == "Address"), None)
    if hdr is None:
        continue
    SH = srows[hdr]
    si = {h: i for i, h in enumerate(SH)}
    cnt, st = collections.Counter(), collections.Counter()
    for x in srows[hdr + 1:]:
        if len(x) < len(SH) or not x[0].startswith("0x"):
            if x and x[0] == "Kernel Name":
                break
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", x[si["Source"]].strip())
        op = m.group(2).split(".")[0] if m else "?"
        cnt[op] += int(x[si["Instructions Executed"]])
        st[op] += int(x[si["Warp Stall Sampling (All Samples)"]])
    tot, tots = sum(cnt.values()) or 1, sum(st.values()) or 1
    print("  opcode mix:", ", ".join(f"{o} {100 * n / tot:.1f}%/{100 * st[o] / tots:.0f}%" for o, n in cnt.most_common(14)))
