"""cuobjdump -sass histogram of the mnemonics that tell a Blackwell-native kernel from a recompiled one, per kernel of the
product library (and of the test-only check library).  Run here (no GPU needed):

    python tools/sass_histogram.py > profiles/r02/sass_mnemonics.txt
"""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "HMMA", "MUFU.EX2", "MUFU.TANH",
        "FFMA2", "FADD2", "FMUL2", "LDGSTS", "STG.E.ENL2.256", "ATOM", "RED"]


def histogram(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|sasvqa::", "", name).split("(")[0]
            cur = per.setdefault(name, collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["_instructions"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k in ("HMMA", "ATOM", "RED", "LDGSTS") and op.startswith(k)):
                if k == "UTCHMMA" and op.startswith("UTCHMMA.2CTA"):
                    continue
                cur[k] += 1
    return per


def main():
    for lib in ("libsasvqa_b200.so", "libsasvqa_b200_test.so"):
        path = os.path.join(ROOT, "sas-vqa_b200", lib)
        if not os.path.exists(path):
            continue
        digest = hashlib.sha256(open(path, "rb").read()).hexdigest()[:16]
        print(f"== {lib} (sha256 {digest}) ==")
        per = histogram(path)
        for name, c in sorted(per.items(), key=lambda kv: -kv[1]["_instructions"]):
            tags = "  ".join(f"{k}={c[k]}" for k in KEYS if c[k])
            print(f"{name[:72]:72s} {c['_instructions']:6d} instr  {tags}")
        print()
        tot = collections.Counter()
        for c in per.values():
            tot.update(c)
        print("TOTAL  " + "  ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))
        print()


if __name__ == "__main__":
    main()
