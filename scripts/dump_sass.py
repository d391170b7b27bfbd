"""Offline (no GPU): NVRTC-compile a generated large system and dump the SASS of one kernel with an opcode histogram.
usage: python scripts/dump_sass.py [bistable|chain] [kernel-name] [halo]"""
import ctypes as C, collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
from odeintr_b200 import _abi, codegen
from systems import *

case = sys.argv[1] if len(sys.argv) > 1 else "bistable"
kernel = sys.argv[2] if len(sys.argv) > 2 else "odeintr_lattice_rk4_tiled_fast"
n = 1 << 28
if case == "bistable":
    g = codegen.generate_sys_openmp("b", BISTABLE_1D, BISTABLE_PARS, True, "rk4", sys_dim=n)
else:
    g = codegen.generate_sys_openmp("c", CHAIN, CHAIN_PARS, True, "rk4", sys_dim=n, globals=CHAIN_GLOBALS)
lib = _abi.load()
size = C.c_size_t(0)
rc = lib.odeintr_compile_cubin(g.code.encode(), 0, None, 0, C.byref(size))
assert rc == 0, _abi.last_error()
buf = C.create_string_buffer(size.value)
assert lib.odeintr_compile_cubin(g.code.encode(), 0, buf, size.value, C.byref(size)) == 0
path = "/tmp/sass/%s.cubin" % case
open(path, "wb").write(buf.raw[:size.value])
sass = subprocess.run(["cuobjdump", "-sass", "-fun", kernel, path], capture_output=True, text=True).stdout
open("/tmp/sass/%s_%s.sass" % (case, kernel), "w").write(sass)
ops = collections.Counter()
for line in sass.splitlines():
    m = re.search(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m:
        ops[m.group(1).split(".")[0]] += 1
print(sum(ops.values()), "instructions")
for k, v in ops.most_common(40):
    print("%6d %s" % (v, k))
