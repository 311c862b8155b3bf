"""Per-kernel counts of the Blackwell-specific SASS opcodes in the built library (cuobjdump -sass): the evidence that the
tensor-core kernels are tcgen05 / TMA code.  `python tools/sass_opcodes.py > profiles/rNN/sass_opcodes.txt`"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rnnt_speech_recognition_b200", "librnnt_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
WANT = re.compile(r"\b(UTCHMMA[.\w]*|UTCBAR[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UBLKCP[.\w]*|UCGABAR[.\w]*|MUFU\.(?:TANH|EX2|RCP|LG2)|STG\.E\.ENL2\.256|ATOMG[.\w]*|REDG[.\w]*|RED[.\w]*)")
cnt, fn = collections.Counter(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", "-p", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        continue
    m = WANT.search(line)
    if m and fn:
        cnt[(fn, m.group(1))] += 1
print("# per-kernel counts of the Blackwell-specific SASS opcodes in librnnt_b200.so (cuobjdump -sass)")
print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor load (.2CTA = cta_group::2 form),")
print("# UBLKCP = cp.async.bulk (1-D), UTCBAR = tcgen05.commit (.2CTA.MULTICAST = multicast to the pair), UCGABAR = cluster barrier")
for (f, op), n in sorted(cnt.items()):
    print("%-70s %-28s %d" % (f[:70], op, n))
