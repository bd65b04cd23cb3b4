"""Per-kernel SASS mnemonic counts (tensor core / TMA / TMEM / async copy) and ptxas resources of the built library.

    python tools/sass_summary.py > profiles/rNN_sass_mnemonics.txt        (needs cuobjdump; no GPU)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fer_vit_b200", "libfervit_b200.so")
PAT = re.compile(r"\b(UTCHMMA[.\w]*|UTCQMMA[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTMAPF[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTCBAR[.\w]*|"
                 r"UTCATOMSWS[.\w]*|SYNCS[.\w]*|HMMA[.\w]*|LDSM[.\w]*|LDGSTS[.\w]*|FFMA2|FMUL2|FADD2|ELECT|UBLKPF[.\w]*)")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return [re.sub(r"\(.*", "", o).replace("void ", "").replace("fervit::", "") for o in out]


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
        if cur:
            for mm in PAT.findall(line):
                kernels[cur][mm] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for line in res.split("\n"):
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    names = list(kernels)
    pretty = demangle(names)
    print("# cuobjdump -sass / -res-usage of fer_vit_b200/libfervit_b200.so (sm_100a): per kernel, registers / static shared /")
    print("# local (spill) bytes, then the tensor-core / TMA / TMEM / async-copy mnemonics it contains.")
    print("# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld,")
    print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier operations, HMMA / LDSM = mma.sync / ldmatrix (attention kernels),")
    print("# LDGSTS = cp.async, FFMA2 / FMUL2 = packed fp32 (GELU epilogues)")
    for n, p in sorted(zip(names, pretty), key=lambda t: t[1]):
        r = usage.get(n)
        head = p + (f"   [regs {r[0]}, static smem {r[1]} B, local {r[2]} B]" if r else "")
        print(head)
        c = kernels[n]
        print("    " + (", ".join(f"{k} x{v}" for k, v in sorted(c.items())) if c else "(none of the listed mnemonics)"))


if __name__ == "__main__":
    sys.exit(main())
