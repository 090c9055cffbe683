#!/usr/bin/env python
"""Build an experimental variant of the library: tools/build_variant.py NAME -DFOO=1 ...
-> radio-core_b200/build_NAME/libradiocore_b200.so (select it with RADIOCORE_B200_LIB)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
name, extra = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(g.PKG_DIR, "build_" + name)
os.makedirs(out_dir, exist_ok=True)
g.compile_units(g._units(), g.NVCC_FLAGS + extra, os.path.join(out_dir, "obj"), os.path.join(out_dir, "libradiocore_b200.so"))
print(os.path.join(out_dir, "libradiocore_b200.so"))
