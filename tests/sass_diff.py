"""Compare the SASS instruction streams of two builds of libb200sim.so kernel by kernel (cuobjdump, no GPU needed):
    python tests/sass_diff.py <old.so> <new.so>
Used to show that a change compiled out of the measured kernels really left them instruction-identical (addresses and encodings are
ignored, the instruction text is compared)."""
import re
import subprocess
import sys


def kernels(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    d, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
        if cur and m:
            d[cur].append(m.group(1))
    return d


if __name__ == "__main__":
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    same = [k for k in a if a[k] == b.get(k)]
    print(f"{len(a)} kernels in {sys.argv[1]}: {len(same)} instruction-identical in {sys.argv[2]}")
    for k in sorted(a):
        if k not in same:
            print(f"  changed: {k[:90]} ({len(a[k])} -> {len(b.get(k, []))} instructions)")
    for k in sorted(b):
        if k not in a:
            print(f"  new:     {k[:90]} ({len(b[k])} instructions)")
