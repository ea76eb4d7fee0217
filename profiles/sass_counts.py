"""SASS mnemonic counts of the built library (the proof that the tcgen05 / TMA / tensor-memory paths are in the
binary): writes profiles/r02_sass_counts.txt.   python profiles/sass_counts.py   (needs cuobjdump and c++filt)"""
import collections
import datetime
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vivit_b200", "csrc", "libvivit_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "DMMA", "HMMA.1688.F32.TF32", "SYNCS", "FFMA2",
             "STG.E.ENL2.256", "UCGABAR", "MEMBAR.SC.SYS", ".STRONG.SYS"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    out = [f"# cuobjdump -sass vivit_b200/csrc/libvivit_b200.so | grep -c <mnemonic>   (built {datetime.date.today()} from "
           "this tree, nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)"]
    for m in MNEMONICS:
        out.append(f"{m}: {sum(1 for line in sass.splitlines() if m in line)}")
    per = collections.OrderedDict()
    fn = None
    for line in sass.splitlines():
        hit = re.search(r"Function : (\S+)", line)
        if hit:
            fn = hit.group(1)
            per[fn] = collections.Counter()
            continue
        if fn:
            for m in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", ".STRONG.SYS"):
                if m in line:
                    per[fn][m] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    out += ["", "# per kernel (function name: UTCHMMA / LDTM / STTM / UTMALDG counts; .STRONG.SYS = system-scope flag",
            "# accesses of the peer-memory hand-over)"]
    for (fn, c), name in zip(per.items(), names):
        if c:
            out.append(f"{name}: " + " ".join(f"{m} {c[m]}" for m in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", ".STRONG.SYS") if c[m]))
    open(os.path.join(ROOT, "profiles", "r02_sass_counts.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:16]))


if __name__ == "__main__":
    main()
