"""Static size of the interpreter's dispatch frame in the V=16 sweep kernel: instructions between the fetch of the
instruction word and the head of the dispatch tree (cuobjdump -sass of ek_sweep.o)."""
import re, subprocess, sys, collections
obj = sys.argv[1] if len(sys.argv) > 1 else "enoki_b200/csrc/ek_sweep.o"
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = sass.split("Function : ")
for f in funcs:
    if "ek_sweep_kernelILi16" not in f.split("\n")[0]: continue
    ops = [re.sub(r'^\s*/\*[0-9a-f]+\*/\s+', '', l).split(';')[0].strip() for l in f.split("\n") if re.search(r'/\*[0-9a-f]{4,5}\*/\s+\S', l)]
    # the fetch: LDC.64 Rw, c[0x0][Rpc*16 + prog_inline] preceded by the LEA that scales the pc
    start = next(i for i, o in enumerate(ops) if o.startswith("LDC.64") and "+0x380]" in o and any("UMOV" in x and "0xa60" in x for x in ops[max(0, i - 10):i]))
    seg = ops[start - 3:start + 200]          # the frame is laid out contiguously after the fetch
    c = collections.Counter((o.split()[1] if o.startswith('@') else o.split()[0]).split('.')[0] for o in seg)
    movs = sum(1 for o in seg if re.match(r'(@\S+\s+)?(IMAD\.MOV\.U32 R\d+, RZ, RZ, R\d+|MOV R\d+, R\d+)', o))
    print(f"total kernel instrs {len(ops)}; {movs} register moves in the 200 instructions after the fetch;", c.most_common(8))
