#!/usr/bin/env python
"""patch_reuse.py -- rewrite the .reuse flags of the DFMAs of ubench4's kernels to calibrate the operand-fetch model.
usage: patch_reuse.py in_exe out_exe none|ideal|skip|swapalt
  none    : clear every reuse flag
  ideal   : set the flag when the next instruction is a DFMA reading the same register in the same slot
  skip    : as ideal, but look through intervening non-FP64 instructions
  swapalt : exchange Ra/Rb of every other DFMA (the shared operand alternates between slots), flags cleared
"""
import struct, subprocess, sys
sys.path.insert(0, __file__.rsplit("/", 3)[0] + "/tools")
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from sass_reuse import parse_function, set_ctrl

src, dst, mode = sys.argv[1:4]
out = subprocess.run(["cuobjdump", "-sass", src], capture_output=True, text=True).stdout
data = bytearray(open(src, "rb").read())
for body in out.split("Function : ")[1:]:
    name = body.split("\n")[0].strip()
    ins = parse_function(body)
    blob = b"".join(struct.pack("<QQ", x.lo, x.hi) for x in ins)
    off = data.find(blob)
    assert off >= 0, name
    words = []
    nd = 0
    for i, x in enumerate(ins):
        lo, hi = x.lo, x.hi
        three = x.fp64 and x.op.startswith("DFMA") and all(r is not None for _, r in x.srcs)
        if three:
            srcs = dict(x.srcs)
            reuse = 0
            if mode in ("ideal", "skip"):
                j = i + 1
                while mode == "skip" and j < len(ins) and not ins[j].fp64 and not ins[j].op.startswith("BRA"):
                    j += 1
                if j < len(ins) and ins[j].fp64:
                    ns = dict(ins[j].srcs)
                    for s, r in srcs.items():
                        if ns.get(s) == r and x.dst not in (r, r + 1, r - 1):
                            reuse |= 1 << s
            if mode in ("first", "alt", "last"):
                # group structure by the shared slot-0 register: first = flag only the first DFMA of a run,
                # alt = flag every other one, last = flag every one INCLUDING the last of the run (stale flag)
                j = i + 1
                while j < len(ins) and not ins[j].fp64 and not ins[j].op.startswith("BRA"):
                    j += 1
                nxt_same = j < len(ins) and ins[j].fp64 and dict(ins[j].srcs).get(0) == srcs.get(0)
                j = i - 1
                while j >= 0 and not ins[j].fp64 and not ins[j].op.startswith("BRA"):
                    j -= 1
                prv_same = j >= 0 and ins[j].fp64 and dict(ins[j].srcs).get(0) == srcs.get(0)
                if mode == "first":
                    reuse = 1 if (nxt_same and not prv_same) else 0
                elif mode == "alt":
                    reuse = 1 if (nxt_same and not (nd & 1)) else 0
                else:
                    reuse = 1
            if mode == "swapalt":
                if nd & 1:
                    ra, rb = (lo >> 24) & 0xFF, (lo >> 32) & 0xFF
                    lo = (lo & ~(0xFFFF << 24)) | (rb << 24) | (ra << 32)
            nd += 1
            hi = set_ctrl(hi, x.stall, x.yld, reuse)
        words.append((lo, hi))
    nb = b"".join(struct.pack("<QQ", lo, hi) for lo, hi in words)
    data[off:off + len(blob)] = nb
    print(f"{name}: {nd} DFMA rewritten ({mode})")
open(dst, "wb").write(bytes(data))
