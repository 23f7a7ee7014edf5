#!/usr/bin/env python
"""Integer-pipe microbenchmarks (cuzk_imad_peak variants) -> JSON, per second and per SM per clock."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cuzk_b200 import api, lib as cl
api.initialize(0)
L = cl.get_lib()
names = {0: "imad_wide_reg", 10: "imad_wide_imm", 11: "imad_wide_constbank", 3: "imad_wide_x_chain_reg", 12: "imad_wide_x_chain_imm",
         2: "imad_hi", 1: "imad_lo", 4: "iadd3_x_chain", 5: "wide+1add", 6: "wide+2add", 7: "wide+3add", 8: "sel", 9: "dfma",
         13: "wide+1dfma (wide counted)", 15: "wide+2dfma (wide counted)", 14: "wide+1imad_lo (wide counted)"}
out = {}
for v, name in names.items():
    x = C.c_double()
    L.check(L.cuzk_imad_peak(v, 2000, C.byref(x)), "imad_peak")
    out[name] = {"per_s": x.value, "per_clk_per_sm_at_1965": x.value / 148 / 1.965e9}
print(json.dumps(out, indent=1))
