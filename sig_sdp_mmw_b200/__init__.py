"""B200-native MMW SDP hot path of zhouyou-gu/sig-sdp-mmw (sim_src/alg/mmw.py and
sdp_solver.rounding) behind the reference's own solver interface."""
from .mmw import mmw  # noqa: F401
from .sdp_solver import rand_sdp_solver, sdp_solver  # noqa: F401
from .stats import STATS_OBJECT  # noqa: F401
