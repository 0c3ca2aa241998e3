#!/usr/bin/env python
"""Per-kernel comparison of the SASS of two builds of the same object (addresses, encodings and the path-dependent anonymous-namespace
hash stripped).  Host-only edits, comment edits and explicit re-statements of what the compiler already generated must leave every
kernel identical.   usage: sass_diff.py before.o after.o"""
import re, subprocess, sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_", "_ANON_", m.group(1))
            cur = re.sub(r"_cu_[0-9a-f]{8}", "_cu_X", cur)
            res[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur:
            res[cur].append(m.group(1).strip())
    return res


a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
diff = only = 0
for k in sorted(set(a) | set(b)):
    if k not in a or k not in b:
        print("ONLY IN", "before" if k in a else "after", k); only += 1
    elif a[k] != b[k]:
        n = sum(1 for x, y in zip(a[k], b[k]) if x != y) + abs(len(a[k]) - len(b[k]))
        print(f"DIFF  {k}: {len(a[k])} -> {len(b[k])} instructions, {n} positions differ"); diff += 1
common = len(set(a) & set(b))
print(f"{common - diff} of {common} common kernels identical, {only} present in one build only")
sys.exit(1 if diff else 0)
