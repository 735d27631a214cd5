#!/usr/bin/env python3
"""The kernels added at the end of the round for ncu (run_r1q.sh): permutation / lookup grand products (31 columns x 2^20
rows), prefix product, batch inverse, fixed-base batch exponentiation (2^18 scalars, BLS12-381 G1 and BN254 G2)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crypto3_zk_b200 import Context  # noqa: E402
from crypto3_zk_b200.fields import CURVE_BY_NAME  # noqa: E402
from profiles.prof_run import rand  # noqa: E402


def main():
    ctx = Context(0)
    n = 1 << 20
    cols, sid, ssg = rand((31, n, 8), 1), rand((31, n, 8), 2), rand((31, n, 8), 3)
    inp, val, srt = rand((2, n, 8), 4), rand((2, n, 8), 5), rand((4, n, 8), 6)
    x = rand((n, 8), 7)
    sc = rand((1 << 18, 8), 8)
    g1, g2 = CURVE_BY_NAME["bls12_381_g1"], CURVE_BY_NAME["bn254_g2"]

    def everything():
        ctx.permutation_grand_product("pallas_fp", cols, sid, ssg, 12345, 67890)
        ctx.lookup_grand_product("pallas_fp", inp, val, srt, 12345, 67890, n - 5)
        ctx.prefix_product("pallas_fp", x)
        ctx.batch_inverse("pallas_fp", x)
        ctx.batch_exp("bls12_381_g1", (g1.gen_x, g1.gen_y), sc)
        ctx.batch_exp("bn254_g2", (g2.gen_x, g2.gen_y), sc)

    everything()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    everything()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ctx.close()


if __name__ == "__main__":
    main()
