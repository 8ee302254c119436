#!/usr/bin/env python
"""The reference's per-phase timing experiment (sim_script/journal_version/sim_mmw_time.py)
against this package: same loop, same calls, same log rows -- only the imports change
(sparse_env for env, the CUDA mmw for mmw).  Needs a CUDA device.

    python examples/sim_mmw_time.py [--cells 5 8] [--repeat 2] [--out /tmp/mmw_time]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sig_sdp_mmw_b200 import mmw                                              # noqa: E402
from sig_sdp_mmw_b200.binary_search_relaxation import binary_search_relaxation  # noqa: E402
from sig_sdp_mmw_b200.topology import sparse_env as env                        # noqa: E402
from sig_sdp_mmw_b200.util import CSV_WRITER_OBJECT                            # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, nargs="+", default=[5, 8, 11])
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--rho", type=float, default=75e-4)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    log = CSV_WRITER_OBJECT(path=args.out)
    for CELL_SIZE in args.cells:
        for seed in range(args.repeat):
            e = env(cell_size=CELL_SIZE, sta_density_per_1m2=args.rho, seed=seed)
            bs = binary_search_relaxation()
            alg = mmw(nit=150, eta=0.04)
            bs.feasibility_check_alg = alg
            state = e.generate_S_Q_hmax()
            z_vec, Z_fin, remainder = bs.run(state)
            bler = e.evaluate_bler(z_vec, Z_fin)

            alg = mmw(nit=150, eta=0.04)
            _, X_half = alg.run_with_state(0, Z_fin, state)
            tic_rnd = alg._get_tic()
            alg.rounding(Z_fin, X_half, state)
            tim_rnd = alg._get_tim(tic_rnd)
            times = [np.mean(alg.LOGGED_NP_DATA[k][:, 5]) for k in
                     ("mmw_all_it", "mmw_dual", "mmw_loss", "mmw_expm", "mmw_xavg")] + [tim_rnd]
            print("cell %2d seed %d: K=%d Z=%d rem=%d mean BLER %.2e | us: all %.0f dual %.1f loss %.1f expm %.1f "
                  "xavg %.0f rounding %.0f" % (CELL_SIZE, seed, e.n_sta, Z_fin, remainder, bler.mean(), *times))
            log.log_mul_scalar(data_name="mmw150-time-%d-%d" % (CELL_SIZE, int(args.rho * 10000)), iteration=seed, values=times)
    log.close()


if __name__ == "__main__":
    main()
