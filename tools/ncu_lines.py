"""Attribute executed warp instructions of one kernel in an .ncu-rep to CUDA source lines: zips the SASS rows of ncu's
source page (per-instruction 'Instructions Executed') with `nvdisasm -g` line annotations of the same kernel.
usage: python tools/ncu_lines.py report.ncu-rep <kernel regex> <cubin> <mangled-name substring> [top]"""
import csv, re, subprocess, sys, collections
rep, kre, cubin, sub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
ii, si, sm = rows[h].index("Instructions Executed"), rows[h].index("Source"), rows[h].index("# Samples")
sass = []
for r in rows[h + 1:]:
    if len(r) <= ii: continue
    if r[0].startswith("0x"): sass.append((r[si].strip(), float(r[ii] or 0), float(r[sm] or 0)))
    elif sass and r[0] == "Kernel Name": break         # next kernel instance
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
cur = None; infn = False; lines = []
for l in dis:
    if l.startswith("//---") and ".text." in l: infn = sub in l; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4}\*/", l): lines.append(cur)
n = min(len(lines), len(sass))
print(f"sass rows {len(sass)}, disasm instrs {len(lines)}")
agg = collections.Counter(); samp = collections.Counter()
for k in range(n):
    agg[lines[k]] += sass[k][1]; samp[lines[k]] += sass[k][2]
tot = sum(agg.values()); ts = sum(samp.values())
print(f"total warp instrs {tot:.0f}")
for key, v in agg.most_common(top):
    print(f"{100*v/tot:5.1f}% instr  {100*samp[key]/max(ts,1):5.1f}% samples  {key}")
