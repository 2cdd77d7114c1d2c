#!/usr/bin/env python
"""Opcode evidence from the built library (no GPU needed): for the hot kernels, how many of the instructions that matter
— 256-bit global loads/stores, cache-hinted loads, bulk-async (TMA engine) copies, atomics/fences of the slot claim,
tcgen05 MMA / TMEM loads — the SASS of libhpsx.so contains.  usage: python scripts/sass_counts.py > profiles/sass_r02.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "hugectr_backend_b200", "lib", "libhpsx.so")
WANT = [("probe_gather_v8_kernel", "ILi16ELi4ELb0EE"), ("pull_binned_kernel", "I6float4Lb1ELi1E"),
        ("pull_binned_kernel", "I6float4Lb1ELi4E"), ("tier_gather_kernel", "I6float4E"), ("insert_binned_kernel", "I6float4E"),
        ("probe_gather_tma_kernel", ""), ("mlp_gemm_tcgen05_kernel", "")]
PAT = re.compile(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
KEEP = re.compile(r"^(LDG|STG|LD\.|ST\.|LDS|STS|ATOMG|ATOM|REDG|RED|MEMBAR|CCTL|UBLKCP|UBLKPF|UTMALDG|UTMASTG|UTCHMMA|UTCQMMA|UTCBAR|"
                  r"LDTM|STTM|SYNCS|NANOSLEEP|SHFL|VOTE|MATCH|ERRBAR|FENCE|UTCCP|UTMACCTL)")

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = PAT.match(line)
    if m:
        funcs[cur][m.group(1)] += 1
print(f"# SASS opcode counts, {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a); memory / sync / tensor instructions only")
for name, tag in WANT:
    for f, c in funcs.items():
        if name in f and tag in f:
            total = sum(c.values())
            print(f"\n## {f}\n   {total} instructions")
            for op, n in sorted(c.items(), key=lambda kv: (-kv[1], kv[0])):
                if KEEP.match(op):
                    print(f"   {n:5d}  {op}")
            break
