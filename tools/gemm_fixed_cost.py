"""Fixed vs per-byte cost of the swap-AB weight-streaming GEMM: time(nout) at K=4096, batch 32, for the SwiGLU and the
plain transposed epilogue; a linear fit gives the launch's fixed cost (us) and the asymptotic bandwidth."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_bw import bench, copy_bw  # noqa: E402

if __name__ == "__main__":
    print(json.dumps(dict(copy_GBs=round(copy_bw(), 1))))
    for epi in (4, 3):
        xs, ys = [], []
        for tiles in (74, 148, 224, 296, 444, 592, 888, 1184):
            r = bench(tiles * 128, 4096, 32, epi, 32, 1, reps=40)
            xs.append(tiles * 128 * 4096 * 2 / 1e6); ys.append(r["us"])
            print(json.dumps(dict(tiles=tiles, **r)))
        b, a = np.polyfit(xs, ys, 1)
        print(json.dumps(dict(epi=epi, fixed_us=round(float(a), 2), asymptotic_GBs=round(1e3 / float(b), 1))))
