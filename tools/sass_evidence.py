"""tcgen05 / TMEM / TMA instruction counts per kernel from `cuobjdump -sass` of the built library (no GPU needed).
    python tools/sass_evidence.py > profiles/r02_sass_evidence.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sp_orb_slam_b200", "lib", "libspfe.so")
OPS = ("UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "UTMALDG")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
cur, cnt = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        for op in OPS:
            if re.search(r"\s" + re.escape(op) + r"[\s.]", line):
                cnt[cur][op] += 1
                break
print("# SASS evidence, round 2 (`cuobjdump -sass sp_orb_slam_b200/lib/libspfe.so`, sm_100a): tcgen05 / TMEM / TMA instructions per kernel\n")
print("`UTCHMMA` = tcgen05.mma (`.2CTA`: cta_group::2), `UTCBAR` = tcgen05.commit, `LDTM` = tcgen05.ld, `UTMALDG` = cp.async.bulk.tensor (TMA).")
print("ConvCfg<TAPS, CB, N, EPI, WRES, SA, SB, HALVES, EG, PAIR, XP>: EPI 4 / 5 = descriptor matching (top-2 pipeline / top-3 set path,")
print("what `spfe_match_mutual_nn` / `spfe_match_knn2` launch), XP = exact mode.\n")
print("| kernel | UTCHMMA | UTCHMMA.2CTA | UTCBAR | LDTM | UTMALDG |\n|---|---|---|---|---|---|")
tot = collections.Counter()
for f, c in sorted(cnt.items(), key=lambda kv: kv[0]):
    if not (c["UTCHMMA"] + c["UTCHMMA.2CTA"]):
        continue
    name = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"\(CUtensorMap.*", "", name).replace("void spfe::", "").replace("spfe::", "")
    print(f"| `{name}` | {c['UTCHMMA']} | {c['UTCHMMA.2CTA']} | {c['UTCBAR']} | {c['LDTM']} | {c['UTMALDG']} |")
    tot.update(c)
print(f"| **total** | {tot['UTCHMMA']} | {tot['UTCHMMA.2CTA']} | {tot['UTCBAR']} | {tot['LDTM']} | {tot['UTMALDG']} |")
