#!/usr/bin/env python
"""The reference's BLER sweep (sim_script/journal_version/sim_all_bler.py:30-47) against this
package: same loop, same calls, same CSV rows `[g_it, it, Z_fin, *bler]` -- only the imports
change.  The "mmw" and "rand" arms run; the ladmm (cvxpy / SCS) and greedy (gm.py) arms of the
reference are outside this package.  Needs a CUDA device.

    python examples/sim_all_bler.py [--cells 5 6] [--repeat 2] [--out /tmp/all_bler]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sig_sdp_mmw_b200 import mmw, rand_sdp_solver                              # noqa: E402
from sig_sdp_mmw_b200.binary_search_relaxation import binary_search_relaxation  # noqa: E402
from sig_sdp_mmw_b200.topology import sparse_env as env                        # noqa: E402
from sig_sdp_mmw_b200.util import CSV_WRITER_OBJECT                            # noqa: E402


def run_point(CELL_SIZE, RHO, seed, log):
    e = env(cell_size=CELL_SIZE, sta_density_per_1m2=RHO, seed=seed)
    bs = binary_search_relaxation()
    alg = mmw(nit=150, eta=0.04)
    bs.feasibility_check_alg = alg
    z_vec, Z_fin, remainder = bs.run(e.generate_S_Q_hmax())
    bler = e.evaluate_bler(z_vec, Z_fin)
    tag = "%d-%d" % (CELL_SIZE, int(RHO * 10000))
    log.log_mul_scalar(data_name="mmw-" + tag, iteration=seed, values=[Z_fin] + bler.tolist())

    alg = rand_sdp_solver()
    _, gX = alg.run_with_state(0, Z_fin, e.generate_S_Q_hmax())
    z_rand, Z_tmp, _ = alg.rounding(Z_fin, gX, e.generate_S_Q_hmax())
    bler_rand = e.evaluate_bler(z_rand, Z_fin)
    log.log_mul_scalar(data_name="rand-" + tag, iteration=seed, values=[Z_fin] + bler_rand.tolist())
    return e.n_sta, Z_fin, remainder, bler, bler_rand


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, nargs="+", default=list(range(5, 16)))
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--rho", type=float, default=75e-4)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    log = CSV_WRITER_OBJECT(path=args.out)
    for CELL_SIZE in args.cells:
        for seed in range(args.repeat):
            K, Z_fin, rem, bler, bler_rand = run_point(CELL_SIZE, args.rho, seed, log)
            print("cell %2d seed %d: K=%d Z=%d rem=%d  BLER mean/max mmw %.2e/%.2e  rand %.2e/%.2e"
                  % (CELL_SIZE, seed, K, Z_fin, rem, bler.mean(), bler.max(), bler_rand.mean(), bler_rand.max()))
    log.close()


if __name__ == "__main__":
    main()
