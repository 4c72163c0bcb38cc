"""cuobjdump -sass of the shipped library -> per-kernel counts of the Blackwell-native mnemonics (profiles/r02_sass_summary.txt)."""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "siammask_b200/libsiammask_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "UTMALDG.2D", "UTMALDG.4D", "UTMALDG.4D.IM2COL", "UTMASTG", "UBLKCP",
        "SYNCS", "BRA.U.ANY", "R2UR.BROADCAST", "HMMA", "FFMA"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"smk::\(anonymous namespace\)::", "", name)
        cur = per.setdefault(name[:110], collections.Counter())
        continue
    if cur is None:
        continue
    for p in pats:
        if re.search(r"\b" + re.escape(p) + r"(\b|\.)", line) and (p != "UTCHMMA" or ".2CTA" not in line or True):
            if p == "UTMALDG.4D" and "IM2COL" in line:
                continue
            cur[p] += 1
print(f"# SASS mnemonics per kernel in {lib} (cuobjdump -sass, nvcc 12.9, sm_100a)")
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = cp.async.bulk,")
print("# BRA.U.ANY / R2UR.BROADCAST = per-instruction divergence ('waterfall') loops around uniform-datapath instructions")
tot = collections.Counter()
for k, c in per.items():
    if not any(c[p] for p in pats[:9]):
        continue
    print(k)
    print("    " + "  ".join(f"{p}={c[p]}" for p in pats if c[p]))
    tot.update(c)
print("TOTAL  " + "  ".join(f"{p}={tot[p]}" for p in pats if tot[p]))
