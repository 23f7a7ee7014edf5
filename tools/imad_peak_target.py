"""ncu target: the IMAD.WIDE microbenchmarks that give bench.py its roofline denominator (cuzk_imad_peak), so that their own
sm__pipe_fmaheavy_cycles_active can be read next to the rate they report."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cuzk_b200 import api
from cuzk_b200.lib import get_lib

L = get_lib()
api.initialize(0)
out = {}
for name, variant in (("imad_wide_reg", 0), ("imad_wide_imm", 10), ("imad_wide_constbank", 11), ("imad_wide_x_chain_reg", 3)):
    v = C.c_double()
    L.check(L.cuzk_imad_peak(variant, 2000, C.byref(v)), "imad_peak")
    out[name] = v.value
print(json.dumps(out))
