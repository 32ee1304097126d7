"""Per-kernel counts of the Blackwell-specific SASS instructions in libscd_b200.so (runs on the CPU box: cuobjdump).
usage: python tools/sass_counts.py [out.txt]"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "scd_resnet_b200", "libscd_b200.so")
PAT = {"UTCHMMA": r"\bUTCHMMA", "UTCHMMA.2CTA": r"UTCHMMA\.2CTA", "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG",
       "LDTM": r"\bLDTM", "UTCBAR": r"\bUTCBAR", "UTCATOMSWS (tmem alloc)": r"\bUTCATOMSWS", "SYNCS (mbarrier)": r"\bSYNCS",
       "UTMAPF / UTMACCTL": r"\bUTMA(PF|CCTL)", "ACQBULK / griddepcontrol": r"\bACQBULK|\bPREEXIT", "REDG.E.ADD.F32 (split-K)": r"\bREDG\.E\.ADD\.F32",
       "F2FP.SATFINITE": r"F2FP\.SATFINITE", "FMNMX3": r"\bFMNMX3"}
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern).replace("void ", "").replace("scd::", "")
        counts[kern] = collections.Counter()
        continue
    if kern is None or "/*" not in line:
        continue
    counts[kern]["instructions"] += 1
    for name, pat in PAT.items():
        if re.search(pat, line):
            counts[kern][name] += 1
            total[name] += 1
lines = ["SASS instruction counts per kernel, libscd_b200.so (nvcc sm_100a; `cuobjdump -sass`), kernels that use tensor / TMA / TMEM paths first", ""]
hdr = ["kernel", "instr"] + list(PAT)
rows = []
for k, c in counts.items():
    rows.append([k, c["instructions"]] + [c.get(n, 0) for n in PAT])
rows.sort(key=lambda r: (-(r[2] + r[4] + r[6]), r[0]))
w = max(len(r[0]) for r in rows)
lines.append("  ".join([hdr[0].ljust(w)] + [h[:12].rjust(12) for h in hdr[1:]]))
for r in rows:
    lines.append("  ".join([r[0].ljust(w)] + [str(v).rjust(12) for v in r[1:]]))
lines.append("")
lines.append("totals: " + ", ".join("%s %d" % (n, total[n]) for n in PAT))
text = "\n".join(lines)
print(text)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text + "\n")
