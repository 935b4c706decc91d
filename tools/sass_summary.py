"""SASS evidence per kernel (profiles/sass_r2.txt): instruction counts that show which hardware paths a kernel uses -- DMMA.8x8x4 (fp64 tensor path),
UBLKCP (cp.async.bulk), SYNCS (mbarrier), LDGSTS (cp.async), BAR -- plus the first lines around the first DMMA / UBLKCP of the Riccati kernels.
usage: python tools/sass_summary.py > profiles/sass_r2.txt   (needs cuobjdump; no GPU)"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
so = ROOT / "wb_humanoid_mpc_b200" / "libb200sqp.so"
txt = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]+\*/\s+\S", line):
        funcs[cur].append(line)
demangle = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {so.name} (sm_100a): instruction counts per kernel\n")
print("| kernel | instructions | DMMA.8x8x4 | DFMA | UBLKCP | SYNCS | LDGSTS | BAR | SHFL |")
print("|---|---|---|---|---|---|---|---|---|")
pat = ["DMMA.8x8x4", " DFMA", "UBLKCP", "SYNCS", "LDGSTS", " BAR", "SHFL"]
for (name, lines), dm in zip(funcs.items(), demangle):
    short = re.sub(r"\(.*", "", dm).replace("b200sqp::", "")
    c = [sum(p in l for l in lines) for p in pat]
    print(f"| `{short}` | {len(lines)} | " + " | ".join(str(v) for v in c) + " |")
for key in ("riccati_bwd_kernel", "riccati_fwd_kernel"):
    for name, lines in funcs.items():
        if key in name:
            print(f"\n## {key}: bulk copies and mbarrier\n```")
            for i, l in enumerate(lines):
                if "UBLKCP" in l or "SYNCS" in l:
                    print(l.strip()[:150])
            print("```")
            idx = next((i for i, l in enumerate(lines) if "DMMA" in l), None)
            if idx is not None:
                print(f"\n## {key}: the DMMA stream (first 24 instructions from the first DMMA)\n```")
                for l in lines[idx: idx + 24]:
                    print(l.strip()[:150])
                print("```")
