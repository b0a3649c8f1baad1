"""Instruction mix of the hot loop of a kernel from `cuobjdump -sass` output.
usage: python tools/sass_loop.py <lib.so> <kernel-name-substring>"""
import collections
import re
import subprocess
import sys

txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
for f in txt.split("Function : ")[1:]:
    name = f.split("\n")[0]
    if sys.argv[2] not in name:
        continue
    ins = re.findall(r"/\*([0-9a-f]{4})\*/\s+(.*?);", f)
    addr = {int(a, 16): t for a, t in ins}
    back = []
    for a, t in ins:
        m = re.search(r"BRA (?:\S+ )?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < int(a, 16):
            back.append((int(a, 16), int(m.group(1), 16)))
    print(name, "total", len(ins), "loops", back)
    if not back:
        continue
    a, b = max(back, key=lambda x: x[0] - x[1])
    body = [t for ad, t in addr.items() if b <= ad <= a]
    c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0] for t in body)
    print("loop body", len(body), c.most_common(30))
