"""Per-kernel SASS mnemonic counts of libmargin_head.so (cuobjdump -sass; no GPU needed): the evidence that the hot kernels
are tcgen05 / TMEM / TMA code (UTCHMMA.2CTA, LDTM, UTMALDG, UTCBAR) and carry no legacy mma.sync (HMMA).

    python scripts/sass_summary.py > profiles/r2_sass_per_kernel.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "face_recognition_models_b200", "libmargin_head.so")
KEYS = ["UTCHMMA", "HMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "MUFU.EX2", "MUFU.LG2",
        "STG.E.128", "LDG.E.128", "LDS.128", "STS.128", "BAR.SYNC", "FFMA", "FADD", "FMNMX", "F2FP", "MEMBAR", "RED",
        "NANOSLEEP", "ACQBULK", "PREEXIT"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            op = m.group(1)
            kernels[cur]["inst"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k in ("MUFU.EX2", "MUFU.LG2", "STG.E.128", "LDG.E.128", "LDS.128", "STS.128", "BAR.SYNC") and op.startswith(k)):
                    kernels[cur][k] += 1
                    break
    names = demangle(list(kernels))
    print("# cuobjdump -sass face_recognition_models_b200/libmargin_head.so (sm_100a), mnemonic counts per kernel (scripts/sass_summary.py).")
    print("# tcgen05.mma.cta_group::2 -> UTCHMMA.2CTA, tcgen05.ld -> LDTM, TMA -> UTMALDG, bulk async copy -> UBLKCP, tcgen05.commit -> UTCBAR,")
    print("# mbarrier -> SYNCS, cp.async -> LDGSTS, griddepcontrol.wait / launch_dependents -> ACQBULK / PREEXIT; HMMA (legacy mma.sync) must be absent.")
    print("# tc_kernel<MODE, V>: MODE 0 FWD, 1 FWDS (forward + stash), 2 BWD_G, 3 DX, 4 DW; V 0 plain, 1 clamp, 2 sphere, 3 MV, 4 curricular, 5 none")
    print("# tc_kernel_dxdw: the merged backward (DX role + DW role); tc_kernel_pwfwd<MODE, V>: the optional merged W prologue + forward\n")
    rows = []
    for mangled, cnt in kernels.items():
        nm = names.get(mangled, mangled)
        nm = re.sub(r"\(anonymous namespace\)::", "", nm)
        nm = re.sub(r"^void ", "", nm)
        nm = re.sub(r"\(.*$", "", nm)
        rows.append((nm, cnt))
    tot_hmma = 0
    for nm, cnt in sorted(rows):
        tot_hmma += cnt["HMMA"]
        body = " ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k])
        print(f"{nm:52s} inst {cnt['inst']:6d}  {body}")
    print(f"\n# HMMA (mma.sync) instructions in the whole library: {tot_hmma}")


if __name__ == "__main__":
    sys.exit(main())
