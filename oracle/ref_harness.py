"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference from /root/reference.

This module exists to (a) validate oracle/mmw_oracle.py against the live
reference and (b) generate the golden vectors under tests/golden/ (see
oracle/make_golden.py).  /root/reference only exists in the build container,
never on the GPU box, so nothing under tests/ -m gpu, bench.py or smoke() may
import this file.

The reference imports three modules that are absent from the image and that do
no arithmetic on the MMW path: matplotlib (util.py:7), line_profiler
(util.py:88) and cvxpy (sdp_solver.py:3).  They are replaced by empty stubs in
sys.modules before the import.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SIG_SDP_REFERENCE", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "sim_src"))


def _install_stubs():
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            m = types.ModuleType("matplotlib")
            p = types.ModuleType("matplotlib.pyplot")
            m.pyplot = p
            sys.modules["matplotlib"] = m
            sys.modules["matplotlib.pyplot"] = p
    if "line_profiler" not in sys.modules:
        try:
            import line_profiler  # noqa: F401
        except Exception:
            lp = types.ModuleType("line_profiler")
            lp.LineProfiler = object
            sys.modules["line_profiler"] = lp
    if "cvxpy" not in sys.modules:
        try:
            import cvxpy  # noqa: F401
        except Exception:
            cp = types.ModuleType("cvxpy")
            cp.SCS = "SCS"
            sys.modules["cvxpy"] = cp


def load_reference():
    """Return a namespace with the reference's mmw, env, binary_search_relaxation."""
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from sim_src.alg.mmw import mmw
    from sim_src.env.env import env
    from sim_src.alg.binary_search_relaxation import binary_search_relaxation
    from sim_src.alg.sdp_solver import sdp_solver
    from sim_src.alg.rounding import rand_rounding
    ns = types.SimpleNamespace(mmw=mmw, env=env, sdp_solver=sdp_solver, rand_rounding=rand_rounding,
                               binary_search_relaxation=binary_search_relaxation)
    return ns
