#!/usr/bin/env python
"""Print selected keys of the last JSON line on stdin: bench_pick.py LABEL key[.sub] ..."""
import json
import sys
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])
out = [sys.argv[1]]
for k in sys.argv[2:]:
    v = d
    for part in k.split("."):
        v = v[part] if v is not None else None
    out.append(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={json.dumps(v)[:400]}")
print(" ".join(out))
