"""Per-kernel counts of the SASS mnemonics that show which hardware path a kernel takes, from the built library (no GPU needed):

    python profiles/sass_summary.py [trinerflet_b200/lib/libtrinerflet_b200.so] > profiles/sass_summary.txt

  UTCHMMA  tcgen05.mma (5th-gen tensor cores)      LDTM / STTM  tcgen05.ld / st (TMEM)        UTCBAR  tcgen05.commit
  HMMA     mma.sync (legacy tensor path)           LDSM         ldmatrix
  LDGSTS   cp.async (global -> shared)             UTMALDG / UTMASTG  TMA tensor copies
  RED      red.global (scatter atomics)            LDGMC        multimem.ld_reduce (NVSwitch multicast load + in-switch add)
  FFMA2    packed fp32 FMA                         SYNCS        mbarrier try_wait / arrive
"""
import collections
import re
import subprocess
import sys

SO = sys.argv[1] if len(sys.argv) > 1 else "trinerflet_b200/lib/libtrinerflet_b200.so"
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "HMMA", "LDSM", "LDGSTS", "UTMALDG", "UTMASTG", "RED", "LDGMC", "FFMA2", "SYNCS", "ATOMG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for k in KEYS:
                if op.startswith(k):
                    kernels[cur][k] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# {SO}: {len(kernels)} kernels, cubin architectures: {', '.join(arch)}")
    print(f"# {'kernel':78s} {'instr':>7s} " + " ".join(f"{k:>8s}" for k in KEYS))
    for (mangled, c), name in zip(kernels.items(), names):
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("tnl::", "")
        print(f"{short[:80]:80s} {c['_total']:7d} " + " ".join(f"{c[k]:8d}" if c[k] else f"{'.':>8s}" for k in KEYS))
    tot = collections.Counter()
    for c in kernels.values():
        tot.update(c)
    print(f"{'TOTAL':80s} {tot['_total']:7d} " + " ".join(f"{tot[k]:8d}" for k in KEYS))


if __name__ == "__main__":
    main()
