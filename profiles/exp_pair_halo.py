import sys, json
sys.path.insert(0, '.')
from guided_diffusion_clip_b200 import _lib as L
from profiles import conv_sweep as cs
lib = L.load()
names = ["256^2 128->128 (clf)", "256^2 256->256", "64^2 512->512"]
for pair in (1, 0):
    lib.gd_debug_set(3, pair)
    for halo in (1, 0):
        lib.gd_debug_set(4, halo)
        for shape in [s for s in cs.SHAPES if s[0] in names]:
            row = {"pair": pair, "halo": halo, "shape": shape[0]}
            for mode in (0, 1):
                lib.gd_debug_set(0, mode)
                ms, tf = cs.run(shape)
                row[f"mode{mode}"] = round(tf, 1)
            lib.gd_debug_set(0, 0)
            print(json.dumps(row), flush=True)
