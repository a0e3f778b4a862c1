#!/usr/bin/env python
"""Compare the SASS of two builds of libbpgpu.so function by function (instruction text only, addresses and encodings
stripped).  Used to show that a change which adds gated code paths left the kernels of the default path untouched:
  python scripts/sass_diff.py /tmp/prev/.../libbpgpu.so dnn-for-speech-enhancement_b200/lib/libbpgpu.so"""
import re
import subprocess
import sys


def functions(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    funcs, name, body = {}, None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                funcs[name] = body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);\s*/\*", line)
        if m and name:
            body.append(m.group(1).strip())
    if name:
        funcs[name] = body
    return funcs


def main():
    a, b = functions(sys.argv[1]), functions(sys.argv[2])
    same = [n for n in a if n in b and a[n] == b[n]]
    diff = [n for n in a if n in b and a[n] != b[n]]
    print(f"{len(a)} / {len(b)} functions; identical: {len(same)}; changed: {len(diff)}; "
          f"only in first: {len(set(a) - set(b))}; only in second: {len(set(b) - set(a))}")
    for n in diff:
        print("  changed:", n, len(a[n]), "->", len(b[n]), "instructions")
    for n in sorted(set(b) - set(a)):
        print("  new:", n)
    return 1 if diff else 0


if __name__ == "__main__":
    sys.exit(main())
