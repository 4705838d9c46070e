"""Registers / spills of every kernel in libpsbax_b200.so (ptxas -v of a forced verbose build)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    res = subprocess.run([sys.executable, "-m", "psbax_b200.build", "--force", "--verbose"], cwd=ROOT,
                         capture_output=True, text=True)
    txt = res.stderr
    rows, cur, spill = [], None, ""
    for line in txt.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur, spill = m.group(1), ""
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and int(m.group(1)):
            spill = f"stack {m.group(1)} spill {m.group(2)}/{m.group(3)}"
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            rows.append((cur, m.group(1), spill))
            cur = None
    names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
    pat = sys.argv[1] if len(sys.argv) > 1 else ""
    for r, n in zip(rows, names):
        n = re.sub(r"\(.*", "", n).replace("void psbax::", "")
        if pat in n:
            print(f"{n:60s} regs {r[1]:4s} {r[2]}")


if __name__ == "__main__":
    main()
