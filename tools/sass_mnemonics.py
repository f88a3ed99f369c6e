"""cuobjdump -sass lance_b200/liblance_b200.so | python tools/sass_mnemonics.py > profiles/sass_rNN_mnemonics.txt
Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (tcgen05 / TMEM / TMA / clusters)."""
import collections, re, subprocess, sys

pat = re.compile(r'\b(UTCHMMA|UTCQMMA|UTCIMMA|UTCOMMA|LDTM|STTM|UTMALDG|UTMASTG|UTCBAR|UTCCP|SYNCS|FMNMX3|UBLKCP|UCGABAR_ARV|UCGABAR_WAIT)\b')
cur, counts = None, collections.defaultdict(collections.Counter)
for line in sys.stdin:
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        for k in pat.findall(line):
            counts[cur][k] += 1


def dem(n):
    try:
        return subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()[:120]
    except Exception:
        return n


print("# SASS mnemonics per kernel in the shipped lance_b200/liblance_b200.so (cuobjdump -sass)")
print("# UTCHMMA = tcgen05.mma (kind::tf32 / kind::f16), LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = TMA tensor load,")
print("# UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier operations, UCGABAR_* = cluster barrier")
rows = sorted((dem(f), c) for f, c in counts.items()
              if any(k in c for k in ('UTCHMMA', 'LDTM', 'UTMALDG', 'UBLKCP', 'UCGABAR_ARV')))
for n, c in rows:
    print(n)
    print("    " + ", ".join(f"{k} x{v}" for k, v in sorted(c.items())))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("library-wide: " + ", ".join(f"{k} x{v}" for k, v in sorted(tot.items())))
