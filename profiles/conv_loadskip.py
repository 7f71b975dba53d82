"""Timing experiment: how the conv mainloop responds when A / B tile loads are skipped (gd_debug_set keys 4/5).
Separates "L2->SM bandwidth bound" from "MMA issue bound".  Results of these launches are garbage by design."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from guided_diffusion_clip_b200 import _lib as L  # noqa: E402
from profiles import conv_sweep as cs  # noqa: E402

if __name__ == "__main__":
    lib = L.load()
    names = ["256^2 256->256", "256^2 128->128 (clf)", "64^2 512->512"]
    for shape in [s for s in cs.SHAPES if s[0] in names]:
        for mode in (1, 0):
            lib.gd_debug_set(0, mode)
            for a_div, b_div in ((0, 0), (2, 0), (3, 0), (1000, 0), (0, 1000), (1000, 1000)):
                lib.gd_debug_set(4, a_div)
                lib.gd_debug_set(5, b_div)
                ms, tf = cs.run(shape)
                print(json.dumps({"shape": shape[0], "mode": mode, "a_div": a_div, "b_div": b_div, "ms": round(ms, 4),
                                  "tflops": round(tf, 1)}), flush=True)
    lib.gd_debug_set(4, 0)
    lib.gd_debug_set(5, 0)
    lib.gd_debug_set(0, 0)
