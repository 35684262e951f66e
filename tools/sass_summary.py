"""Static evidence for the hot kernels (no GPU needed): registers / spills / shared memory from the ptxas logs the build
keeps (satellite_approximation_b200/csrc/_build/*.ptxas.log) and SASS mnemonic counts from `cuobjdump -sass` of
lib/libsatfill.so -- 128-bit global loads / stores, warp shuffles, reductions / atomics, barriers, local-memory traffic.
    python tools/sass_summary.py [kernel-name-substring ...] > profiles/<round>_sass_summary.txt"""
from __future__ import annotations

import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "satellite_approximation_b200", "lib", "libsatfill.so")
BUILD = os.path.join(ROOT, "satellite_approximation_b200", "csrc", "_build")
DEFAULT = ["k_update2", "k_direction2", "k_rb_down", "k_rb_up", "k_rb_coarsest", "k_setup2", "k_fetch", "k_scatter",
           "k_row_counts", "k_scan_rows", "k_row_number", "k_ccl", "k_morph", "k_split_u8", "k_merge_f64"]  # fmt: skip
COUNT = [("LDG.E.128", r"\bLDG\.E\S*\.128"), ("LDG.E.64", r"\bLDG\.E\S*\.64"), ("LDG (other)", r"\bLDG\b"),
         ("STG.E.128", r"\bSTG\.E\S*\.128"), ("STG.E.64", r"\bSTG\.E\S*\.64"), ("STG (other)", r"\bSTG\b"),
         ("LDS", r"\bLDS\b"), ("STS", r"\bSTS\b"), ("SHFL", r"\bSHFL\b"), ("RED/ATOM", r"\b(RED|ATOM|ATOMG|ATOMS)\b"),
         ("BAR", r"\bBAR\b"), ("LDL/STL (local)", r"\b(LDL|STL)\b"), ("DFMA/DADD/DMUL", r"\b(DFMA|DADD|DMUL)\b"),
         ("FFMA/FADD/FMUL", r"\b(FFMA|FADD|FMUL|FFMA2|FADD2|FMUL2)\b")]  # fmt: skip


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return r.stdout.splitlines() if r.returncode == 0 else names


def ptxas_info():
    info = {}
    for f in sorted(glob.glob(os.path.join(BUILD, "*.ptxas.log"))):
        cur = None
        for line in open(f):
            m = re.search(r"Function properties for (\S+)", line)
            if m:
                cur = m.group(1)
                info[cur] = {"file": os.path.basename(f).replace(".ptxas.log", ".cu")}
                continue
            if cur is None:
                continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m:
                info[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
            m = re.search(r"Used (\d+) registers(?:, used (\d+) barriers)?", line)
            if m:
                info[cur]["regs"] = int(m.group(1))
                s = re.search(r"(\d+) bytes smem", line)
                info[cur]["smem"] = int(s.group(1)) if s else 0
                cur = None
    return info


def sass_counts():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = {k: 0 for k, _ in COUNT}
            counts[cur]["instructions"] = 0
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1)
        counts[cur]["instructions"] += 1
        seen_ldg = seen_stg = False
        for key, pat in COUNT:
            if key.startswith("LDG (") and seen_ldg or key.startswith("STG (") and seen_stg:
                continue
            if re.search(pat, ins):
                counts[cur][key] += 1
                seen_ldg |= key.startswith("LDG")
                seen_stg |= key.startswith("STG")
    return counts


def main():
    want = sys.argv[1:] or DEFAULT
    info, counts = ptxas_info(), sass_counts()
    names = sorted(n for n in counts if any(w in n for w in want))
    pretty = dict(zip(names, demangle(names)))
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(counts)} kernels in the cubin (sm_100a); {len(names)} shown")
    print("# per kernel: ptxas resource use, then SASS mnemonic counts (static, per kernel body)")
    for n in names:
        p = info.get(n, {})
        short = re.sub(r"\(.*", "", pretty[n]).replace("void satfill::", "")
        print(f"\n{short}   [{p.get('file', '?')}]")
        print(f"  regs {p.get('regs', '?')}  smem {p.get('smem', '?')} B  stack {p.get('stack', '?')} B  "
              f"spill st/ld {p.get('spill_st', '?')}/{p.get('spill_ld', '?')} B  instructions {counts[n]['instructions']}")  # fmt: skip
        print("  " + "  ".join(f"{k}={v}" for k, v in counts[n].items() if k != "instructions" and v))


if __name__ == "__main__":
    main()
