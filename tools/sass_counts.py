"""Static SASS instruction mix per kernel of one object file (cuobjdump -sass), as markdown.

    python tools/sass_counts.py hymd_b200/csrc/_obj/bonded.o > profiles/<round>_bonded_sass_counts.md

Static counts (loops and branches are not weighted): evidence for which pipe a kernel leans on, not a timing."""
import re
import subprocess
import sys
from collections import defaultdict


def main():
    obj = sys.argv[1]
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    demangled = {}
    try:
        names = re.findall(r"Function : (\S+)", out)
        dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangled = dict(zip(names, dm))
    except Exception:
        pass
    cls = [("fp64", r"\b(DFMA|DMUL|DADD|DSETP|DMNMX)"), ("fp32", r"\b(FFMA|FMUL|FADD|FSETP|FMNMX)"),
           ("mufu", r"\bMUFU"), ("int", r"\b(IMAD|IADD3|LEA|LOP3|SHF|ISETP)"), ("ldg", r"\bLDG"), ("stg", r"\bSTG"),
           ("lds/sts", r"\b(LDS|STS)"), ("shfl", r"\bSHFL"), ("bar", r"\bBAR"), ("branch", r"\b(BRA|BSSY|BSYNC|CALL)")]
    counts = defaultdict(lambda: defaultdict(int))
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        if name and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            counts[name]["total"] += 1
            for key, pat in cls:
                if re.search(pat, line):
                    counts[name][key] += 1
    print(f"# Static SASS instruction mix, `{obj}` (sm_100a)\n")
    print("| kernel | total | " + " | ".join(k for k, _ in cls) + " |")
    print("|---|---|" + "---|" * len(cls))
    for n in sorted(counts, key=lambda k: demangled.get(k, k)):
        short = re.sub(r"\(.*", "", demangled.get(n, n).replace("(int)", "")).replace("hymd::", "").replace("void ", "")
        c = counts[n]
        print(f"| `{short}` | {c['total']} | " + " | ".join(str(c[k]) for k, _ in cls) + " |")


if __name__ == "__main__":
    main()
