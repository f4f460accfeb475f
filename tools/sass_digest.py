"""Per-kernel opcode histogram of the Blackwell-specific SASS in the built objects (fast-cwdm_b200/fcwdm/_obj/*.o):
UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor loads), UBLKCP (bulk copies), LDTM/STTM (TMEM loads/stores), UTCBAR
(tcgen05.commit), SYNCS (mbarrier), 256-bit global loads/stores.  Written to profiles/ so the tree carries its own proof.

    python tools/sass_digest.py > profiles/r02_sass_digest.txt"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "fast-cwdm_b200", "fcwdm", "_obj")
PAT = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTMALDG\.\dD(?:\.2CTA)?(?:\.MULTICAST)?|UBLKCP[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTCBAR[.\w]*|"
                 r"SYNCS[.\w]*|LDG\.E[.\w]*256[.\w]*|STG\.E[.\w]*256[.\w]*|STAS[.\w]*|RED\.E[.\w]*|ATOMG[.\w]*|MUFU\.TANH|CCTL[.\w]*)")
print("# SASS opcode digest of libfcwdm.so's objects (cuobjdump -sass, sm_100a); digest.txt =",
      open(os.path.join(OBJ, "digest.txt")).read().strip()[:16] if os.path.exists(os.path.join(OBJ, "digest.txt")) else "?")
for obj in sorted(f for f in os.listdir(OBJ) if f.endswith(".o")):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    rows = []
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem).replace("fcwdm::", "").replace("void ", "")
        hist = collections.Counter(m.group(1) for m in PAT.finditer(f))
        key = {k: v for k, v in hist.items() if k.startswith(("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "LDG", "STG", "STAS", "MUFU"))}
        if key:
            rows.append((dem, key))
    if rows:
        print(f"\n== {obj}")
        for dem, key in rows:
            print(f"  {dem[:70]:70s} " + "  ".join(f"{k}:{v}" for k, v in sorted(key.items())))
