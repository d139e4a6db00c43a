"""Regenerates tests/golden/part_layouts.json and reference_constants.json from
the oracle builds (oracle/_ref/libswiftref_<scheme>.so, i.e. offsetof() on the
reference's own struct part). Run in the container that has /root/reference:
    make -C oracle ref && python tests/golden/gen_layouts.py
"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from swift_b200.abi import PartLayout

CONST_NAMES = ["kernel_gamma", "kernel_gamma2", "kernel_root", "kernel_norm",
               "kernel_constant", "kernel_gamma_inv_dim",
               "kernel_gamma_inv_dim_plus_one", "hydro_gamma", "space_splitsize",
               "space_recurse_size_self_hydro", "space_recurse_size_pair_hydro",
               "space_maxreldx", "const_viscosity_beta", "sizeof_part",
               "sizeof_xpart", "sizeof_cell"]
layouts, consts = {}, {}
for s in ("minimal", "gadget2", "sphenix", "sphenix_chk"):
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", f"libswiftref_{s}.so"), mode=os.RTLD_NOW)
    L = PartLayout()
    lib.swiftref_layout(ctypes.byref(L))
    layouts[s] = L.as_dict()
    c = (ctypes.c_double * 16)()
    lib.swiftref_constants(c)
    consts[s] = dict(zip(CONST_NAMES, [float.hex(v) for v in c]))
    # kernel_deval known answers on a fixed u grid
    lib.swiftref_kernel_deval.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    if s == "minimal":
        kv = []
        for i in range(0, 41):
            u = i * 0.05
            w, dw = ctypes.c_float(), ctypes.c_float()
            lib.swiftref_kernel_deval(u, ctypes.byref(w), ctypes.byref(dw))
            kv.append([float.hex(ctypes.c_float(u).value), float.hex(w.value), float.hex(dw.value)])
        consts["kernel_deval"] = kv
json.dump(layouts, open(os.path.join(ROOT, "tests", "golden", "part_layouts.json"), "w"), indent=1)
json.dump(consts, open(os.path.join(ROOT, "tests", "golden", "reference_constants.json"), "w"), indent=1)
print("wrote layouts for", list(layouts))
