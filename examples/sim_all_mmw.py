#!/usr/bin/env python
"""The reference's duality-gap experiment (sim_script/journal_version/sim_all_mmw.py:27-53)
against this package: for eta = 0.02 ... 0.10, nit = ceil(1 / eta^2) iterations with
LOG_GAP = True, then the two curves the reference logs: `gap[:, 3]` (max_c A_c . X_avgd, the upper
curve) and `gap[:, 4]` (K lambda_min(L(Y_avgd)), the lower one), one CSV row each.  The reference
takes Z from a cvxpy / SCS search (not installable here); Z comes from the MMW search instead, or
from --Z.  Needs a CUDA device.

    python examples/sim_all_mmw.py [--cells 5] [--etas 0.1] [--repeat 1] [--out /tmp/all_mmw]
"""
import argparse
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sig_sdp_mmw_b200 import mmw                                              # noqa: E402
from sig_sdp_mmw_b200.binary_search_relaxation import binary_search_relaxation  # noqa: E402
from sig_sdp_mmw_b200.topology import sparse_env as env                        # noqa: E402
from sig_sdp_mmw_b200.util import CSV_WRITER_OBJECT                            # noqa: E402


def gap_curves(state, Z, eta, nit, rank_radio=2):
    """One LOG_GAP solve; returns the reference's (ub, lb) = (gap[:, 3], gap[:, 4])."""
    alg = mmw(nit=nit, eta=eta, rank_radio=rank_radio)
    alg.LOG_GAP = True
    _, gX = alg.run_with_state(0, Z, state)
    return alg.LOGGED_NP_DATA["gap"][:, 3], alg.LOGGED_NP_DATA["gap"][:, 4]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, nargs="+", default=[5, 10, 15])
    ap.add_argument("--etas", type=float, nargs="+", default=[0.02 + 0.01 * i for i in range(9)])
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--rho", type=float, default=75e-4)
    ap.add_argument("--Z", type=int, default=0, help="slots (0: smallest Z the MMW search colours)")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    log = CSV_WRITER_OBJECT(path=args.out)
    for CELL_SIZE in args.cells:
        for ETA in args.etas:
            NIT = math.ceil(1. / ETA / ETA)
            for seed in range(args.repeat):
                e = env(cell_size=CELL_SIZE, sta_density_per_1m2=args.rho, seed=seed)
                Z_fin = args.Z
                if Z_fin <= 0:
                    bs = binary_search_relaxation()
                    bs.feasibility_check_alg = mmw(nit=150, eta=0.04)
                    _, Z_fin, _ = bs.run(e.generate_S_Q_hmax())
                ub, lb = gap_curves(e.generate_S_Q_hmax(), Z_fin, ETA, NIT)
                name = "mmw-dual-%d-%d-%d" % (CELL_SIZE, int(args.rho * 10000), int(round(ETA * 100)))
                log.log_mul_scalar(data_name=name, iteration=seed, values=ub.tolist())
                log.log_mul_scalar(data_name=name, iteration=seed, values=lb.tolist())
                print("cell %2d eta %.2f seed %d: Z=%d nit=%d  ub %.4f -> %.4f  lb %.4f -> %.4f"
                      % (CELL_SIZE, ETA, seed, Z_fin, NIT, ub[0], ub[-1], lb[0], lb[-1]))
    log.close()


if __name__ == "__main__":
    main()
