#!/usr/bin/env python
"""The reference's search-cost experiment (sim_script/journal_version/sim_mmw_scs_iter_time.py:
27-70) against this package: number of probes and wall time of the binary search over Z with the
MMW solver, between set_bounds' bounds ("MMW") and between the trivial bounds 1..K
(`force_full_bound`, "MWM-NB").  The SCS arm (cvxpy) cannot run here; its two columns are logged
as nan so the row keeps the reference's layout `[g_it, it, d_mmw, t_mmw, d_scs, t_scs, d_nb, t_nb]`.
Needs a CUDA device.

    python examples/sim_mmw_scs_iter_time.py [--cells 5] [--repeat 1] [--out /tmp/iter_time]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sig_sdp_mmw_b200 import mmw                                              # noqa: E402
from sig_sdp_mmw_b200.binary_search_relaxation import binary_search_relaxation  # noqa: E402
from sig_sdp_mmw_b200.topology import sparse_env as env                        # noqa: E402
from sig_sdp_mmw_b200.util import CSV_WRITER_OBJECT                            # noqa: E402


def search_cost(state, full_bound):
    bs = binary_search_relaxation()
    tic = bs._get_tic()
    bs.feasibility_check_alg = mmw(nit=150, eta=0.04)
    bs.force_full_bound = full_bound
    z_vec, Z_fin, remainder = bs.run(state)
    d = bs.LOGGED_NP_DATA["bs_search_per_it"].shape[0]
    return d, bs._get_tim(tic), Z_fin, remainder


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, nargs="+", default=[13, 14, 15])
    ap.add_argument("--repeat", type=int, default=5)
    ap.add_argument("--rho", type=float, default=75e-4)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    log = CSV_WRITER_OBJECT(path=args.out)
    for CELL_SIZE in args.cells:
        for seed in range(args.repeat):
            e = env(cell_size=CELL_SIZE, sta_density_per_1m2=args.rho, seed=seed)
            d0, t0, Z0, r0 = search_cost(e.generate_S_Q_hmax(), False)
            d2, t2, Z2, r2 = search_cost(e.generate_S_Q_hmax(), True)
            res = [d0, t0, float("nan"), float("nan"), d2, t2]
            log.log_mul_scalar(data_name="time-%d-%d" % (CELL_SIZE, int(args.rho * 10000)), iteration=seed, values=res)
            print("cell %2d seed %d: bounded search %d probes %.0f ms (Z=%d rem=%d) | 1..K search %d probes %.0f ms (Z=%d rem=%d)"
                  % (CELL_SIZE, seed, d0, t0 / 1e3, Z0, r0, d2, t2 / 1e3, Z2, r2))
    log.close()


if __name__ == "__main__":
    main()
