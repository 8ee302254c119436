"""Is the solver of a finished run_with_state freed when the mmw object goes away (no reference cycle)?"""
import gc, sys, weakref
sys.path.insert(0, ".")
import numpy as np, torch
from sig_sdp_mmw_b200 import mmw
from sig_sdp_mmw_b200.topology import sparse_env
state = sparse_env(cell_size=75, sta_density_per_1m2=75e-4, seed=1).generate_S_Q_hmax()
gc.disable()
alg = mmw(nit=3, rank_radio=2, eta=0.04, omega="device", seed=1)
alg.run_with_state(0, 8, state)
plan = alg._plan_cache["plan"]
w = weakref.ref(plan)
del plan, alg
print("plan alive after del (refcount only):", w() is not None)
print("gc.collect() found", gc.collect(), "objects; plan alive:", w() is not None)
