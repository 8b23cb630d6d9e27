"""Print the opcode stream of one kernel from `cuobjdump -sass` output: python tools/sass_ops.py file.sass <kernel substring> [start end]"""
import re, sys
txt = open(sys.argv[1]).read()
key = sys.argv[2]
fn = [f for f in txt.split("Function : ")[1:] if key in f.split("\n")[0]][0]
ops = []
for line in fn.split("\n"):
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ops.append((int(m.group(1), 16), m.group(2).strip()))
if len(sys.argv) > 3:
    a, b = int(sys.argv[3]), int(sys.argv[4])
    for i in range(a, min(b, len(ops))):
        print(i, hex(ops[i][0]), ops[i][1][:100])
else:
    print(len(ops), "instructions")
    # run-length summary
    prev, cnt, start = None, 0, 0
    for i, (_, o) in enumerate(ops):
        t = o.split()[1] if o.startswith("@") else o.split()[0]
        t = t.split(".")[0]
        if t == prev:
            cnt += 1
        else:
            if prev is not None and (cnt >= 4 or prev in ("BRA", "BSSY", "BSYNC", "EXIT")):
                print(start, prev, cnt)
            prev, cnt, start = t, 1, i
