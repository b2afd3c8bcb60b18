"""SASS opcode histogram and resource usage of every kernel in the product library (no GPU needed):
    python tools/sass_histogram.py > profiles/r02_sass_histogram.txt
Per kernel: instruction count, the opcodes that matter for this path (integer / logic / shift, shared and global memory,
warp collectives, atomics, the bulk-copy + mbarrier opcodes of the staged variant), registers, shared memory, spills."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "biorseo_b200", "lib", "libbiorseo_b200.so")
GROUPS = [("LOP3/SHF/POPC/FLO/BREV (bit work)", r"^(LOP3|SHF|POPC|FLO|BREV|PLOP3)"), ("IMAD/IADD3/LEA/ISETP/SEL (integer, addresses)", r"^(IMAD|IADD3|IADD|LEA|ISETP|SEL|IMNMX|VIMNMX|IABS|MOV|PRMT)"),
          ("LDS/STS (shared)", r"^(LDS|STS|LDSM)"), ("LDG/STG/LD/ST (global, generic)", r"^(LDG|STG|LD|ST)\b|^(LDG|STG)\."), ("LDL/STL (spills)", r"^(LDL|STL)"),
          ("SHFL/VOTE/MATCH/WARPSYNC (warp collectives)", r"^(SHFL|VOTE|MATCH|WARPSYNC|BAR|NANOSLEEP)"), ("ATOM/ATOMG/RED/ATOMS", r"^(ATOM|RED)"),
          ("BRA/BSSY/BSYNC/CALL/RET/EXIT (control)", r"^(BRA|BRX|BSSY|BSYNC|CALL|RET|EXIT|JMP|BREAK)"),
          ("UBLKCP/SYNCS/UTMA* (bulk copy + mbarrier)", r"^(UBLKCP|SYNCS|UTMA|UTMACCTL|ARRIVES)")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage, name = {}, None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        elif name and "REG:" in line:
            usage[name] = line.strip()
            name = None
    kernels, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if name and m:
            kernels[name][m.group(1).split(".")[0] if not m.group(1).startswith(("LDG", "STG")) else m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    for (name, ops), pretty in zip(kernels.items(), demangle):
        total = sum(ops.values())
        short = re.sub(r"\(anonymous namespace\)::|bss::", "", pretty.split("(bss::")[0].split("(unsigned")[0].split("(float")[0])
        print(f"== {short}   {total} instructions")
        print(f"   {usage.get(name, '')}")
        for label, pattern in GROUPS:
            n = sum(c for op, c in ops.items() if re.match(pattern, op))
            if n:
                print(f"   {n:6d}  {label}")
        wide = {op: c for op, c in ops.items() if re.match(r"^(LDG|STG).*\.(64|128)", op)}
        if wide:
            print("           wide global accesses: " + ", ".join(f"{op} x{c}" for op, c in sorted(wide.items())))
        print("           top: " + ", ".join(f"{op} {c}" for op, c in ops.most_common(8)))


if __name__ == "__main__":
    main()
