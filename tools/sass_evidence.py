"""profiles/r02_sass_evidence.txt: which Blackwell-specific SASS each kernel of libokb200.so contains
(cuobjdump -sass; runs without a GPU).  tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA / bulk copies -> UBLKCP / UTMALDG,
packed fp32 -> FADD2, mbarrier -> SYNCS, cluster barrier -> UCGABAR, griddepcontrol -> ACQBULK / PREEXIT-like control."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "openkeonspark_b200", "libokb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UBLKCP|UBLKPF|UTMALDG|UTMASTG|FADD2|FMUL2|FFMA2|SYNCS|UCGABAR_ARV|UCGABAR_WAIT|MATCH|REDUX|HMMA|LDGSTS|ATOMG|ATOMS|REDG)\b")
kern, cur = collections.OrderedDict(), None
samples = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*\)$", "", cur)
        kern[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = pat.search(line)
    if m:
        kern[cur][m.group(1)] += 1
        key = (cur, m.group(1))
        if key not in samples and m.group(1) in ("UTCHMMA", "LDTM", "UBLKCP", "FADD2", "UTCBAR", "UBLKPF"):
            samples[key] = re.sub(r"/\*[0-9a-fx]+\*/", "", line).strip()
lines = ["# SASS evidence (cuobjdump -sass openkeonspark_b200/libokb200.so, sm_100a) — instruction counts per kernel", ""]
for k, c in kern.items():
    if not c:
        continue
    lines.append("%-90s %s" % (k[:90], "  ".join("%s x%d" % kv for kv in sorted(c.items()))))
lines += ["", "# one instance of each Blackwell-specific mnemonic, as disassembled", ""]
for (k, mn), l in samples.items():
    lines.append("%-60s %s" % (k[:60], l))
txt = "\n".join(lines) + "\n"
open(os.path.join(ROOT, "profiles", "r02_sass_evidence.txt"), "w").write(txt)
sys.stdout.write(txt)
