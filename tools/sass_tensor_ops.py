"""Developer tool: per-kernel counts of the tensor-core / TMA / TMEM instructions in the built library (runs anywhere,
needs only cuobjdump):   python tools/sass_tensor_ops.py > profiles/r2_sass_tensor_ops.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "youtube-vln_b200", "yvb200", "libyvb200.so")
PAT = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTMALDG\.\dD(?:\.2CTA)?|UTMASTG\.\dD|LDTM\.x\d+|STTM\.x\d+|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|HMMA\S*)")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per, sample, cur = collections.OrderedDict(), {}, None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = PAT.search(line)
        if m and cur:
            per[cur][m.group(1)] += 1
            sample.setdefault((cur, m.group(1)), line.strip().split("*/", 1)[-1].strip()[:110])
    names = demangle(list(per))
    print("cuobjdump -sass youtube-vln_b200/yvb200/libyvb200.so : tensor-core / TMA / TMEM instructions per kernel (sm_100a)")
    print("(tcgen05.mma -> UTCHMMA, cp.async.bulk.tensor -> UTMALDG, tcgen05.ld/st -> LDTM/STTM, tcgen05.commit -> UTCBAR; "
          "no HMMA = no legacy mma.sync)\n")
    hmma = 0
    for k, c in sorted(per.items(), key=lambda kv: names[kv[0]]):
        if not c:
            continue
        short = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", names[k]).split("(")[0]
        print(short)
        for op, n in sorted(c.items()):
            hmma += n if op.startswith("HMMA") else 0
            print(f"    {n:5d} x {op:24s} e.g. {sample[(k, op)]}")
        print()
    print(f"kernels without any of these instructions: {sum(1 for c in per.values() if not c)}; HMMA instructions: {hmma}")


if __name__ == "__main__":
    main()
