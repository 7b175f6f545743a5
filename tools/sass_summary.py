"""profiles/sass_summary.md: per-kernel counts of the Blackwell-specific SASS mnemonics in the built library
(cuobjdump -sass of the object files): UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / .st,
UTMALDG = TMA tensor load, UBLKCP = bulk async copy, LDGMC = multimem.ld_reduce, STG .MC / multimem stores, SYNCS = mbarrier.
usage: python tools/sass_summary.py > profiles/sass_summary.md"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "LDGMC", "REDG", "SYNCS", "DFMA"]
print("# SASS evidence (cuobjdump -sass of exemplar_vae_b200/csrc/*.o, sm_100a)\n")
print("Counts of Blackwell-specific instructions per kernel: `UTCHMMA` = tcgen05.mma (`.2CTA` = cta_group::2), `LDTM`/`STTM` = "
      "tcgen05.ld/st (tensor memory), `UTMALDG` = TMA tensor load, `UBLKCP` = cp.async.bulk, `UTCBAR` = tcgen05.commit, "
      "`LDGMC` = multimem.ld_reduce (NVSwitch in-fabric reduction), `SYNCS` = mbarrier ops, `DFMA` = fp64 FMA (bit-exact "
      "kNN distances).\n")
print("| object | kernel | " + " | ".join(PAT) + " |\n|---|---|" + "---|" * len(PAT))
for obj in sorted(glob.glob(os.path.join(ROOT, "exemplar_vae_b200", "csrc", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("exvae::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void ", "")
            name = re.sub(r"\(.*", "", name)
            cur = counts.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        for p_ in PAT:
            if re.search(r"\b" + re.escape(p_) + r"\b", line) or (p_ != "UTCHMMA" and p_ in line):
                if p_ == "UTCHMMA" and "UTCHMMA.2CTA" in line:
                    continue
                cur[p_] += 1
    for name, c in counts.items():
        if sum(c.values()) == 0:
            continue
        print(f"| {os.path.basename(obj)} | `{name[:90]}` | " + " | ".join(str(c.get(p_, 0)) for p_ in PAT) + " |")
