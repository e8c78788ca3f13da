#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-native SASS mnemonics in libnavc.so (cuobjdump -sass): UTC*MMA (tcgen05.mma, with
the .2CTA forms of cta_group::2), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA loads / stores), UTCBAR
(tcgen05.commit), HMMA (legacy mma.sync: none expected).   python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "non-autoregressive-video-captioning_b200", "csrc", "libnavc.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pats = [("UTCHMMA.2CTA", r"UTC\w*MMA\.2CTA"), ("UTCHMMA", r"UTC\w*MMA(?!\.2CTA)"), ("UTCHMMA(A in TMEM)", r"UTC\w*MMA\S* tmem\["),
        ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"UTMALDG(?!\S*2CTA)"), ("UTMALDG.2CTA", r"UTMALDG\S*2CTA"), ("UTMASTG", r"UTMASTG"),
        ("UTCBAR", r"UTCBAR"), ("HMMA", r"\bHMMA"), ("STG.256", r"STG\S*\.256"), ("LDG.256", r"LDG\S*\.256")]
cur, counts, order = None, collections.defaultdict(lambda: collections.Counter()), []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("navc::", "")
        order.append(cur)
        continue
    if cur is None or "/*" not in line:
        continue
    for name, pat in pats:
        if re.search(pat, line):
            counts[cur][name] += 1
    counts[cur]["instructions"] += 1
names = [n for n, _ in pats]
print("%-46s %7s " % ("kernel (libnavc.so, sm_100a)", "instr") + " ".join("%12s" % n[:12] for n in names))
tot = collections.Counter()
for k in order:
    c = counts[k]
    if not any(c[n] for n in names):
        continue
    print("%-46s %7d " % (k[:46], c["instructions"]) + " ".join("%12d" % c[n] for n in names))
    tot.update(c)
print("%-46s %7d " % ("TOTAL (kernels with any of the above)", tot["instructions"]) + " ".join("%12d" % tot[n] for n in names))
