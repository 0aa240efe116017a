"""Developer tool: per-source-line dynamic instruction counts of one kernel, joining the SASS page of
an .ncu-rep with nvdisasm line info of the in-tree library (same build).
usage: python tools/ncu_lines.py rep kernel_regex mangled_substring [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "qcpinn-convection-diffusion-qiskit_b200", "libqcpinn_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
lines_of = []   # per instruction (in order): (file, line)
for cub in sorted(os.listdir(tmp)):
    out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    cur, line, found = None, None, False
    for l in out.split("\n"):
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            cur = m.group(1); continue
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', l)
        if m:
            line = (m.group(1), int(m.group(2))); continue
        if cur and mangled in cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            lines_of.append(line); found = True
    if found:
        break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(i for i, x in enumerate(rows) if x and x[0] == "Address")
H = rows[hdr]; si = {h: i for i, h in enumerate(H)}
ins = [x for x in rows[hdr + 1:] if len(x) >= len(H) and x[0].startswith("0x")]
if len(ins) != len(lines_of):
    print(f"warning: {len(ins)} profiled instructions vs {len(lines_of)} disassembled (different build?)")
cnt, st, ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for x, ln in zip(ins, lines_of):
    n = int(x[si["Instructions Executed"]])
    cnt[ln] += n
    st[ln] += int(x[si["Warp Stall Sampling (All Samples)"]])
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", x[si["Source"]].strip())
    ops[ln][m.group(2).split(".")[0] if m else "?"] += n
tot, tots = sum(cnt.values()) or 1, sum(st.values()) or 1
for ln, n in cnt.most_common(top):
    txt = ""
    if ln and ln[0].startswith("qcp_"):
        try:
            txt = open(os.path.join(ROOT, "qcpinn-convection-diffusion-qiskit_b200", "csrc", ln[0])).read().split("\n")[ln[1] - 1].strip()[:60]
        except Exception:
            pass
    mix = " ".join(f"{o}:{100 * v // n}" for o, v in ops[ln].most_common(3))
    print(f"{str(ln):34s} {100 * n / tot:5.1f}% stall {100 * st[ln] / tots:5.1f}%  [{mix}]  {txt}")
