"""Log store and wall-clock timers with the reference's surface
(sim_src/util.py:114-217, STATS_OBJECT): drivers read
``alg.LOGGED_NP_DATA[key][:, 5]`` and call ``alg._get_tic() / alg._get_tim()``
(sim_script/journal_version/sim_mmw_time.py:43-52), so the row layout
``[g_step, step, time(), *data]`` and the microsecond timers are part of the
drop-in contract.  Written from the contract, not from the reference's code."""
import os
import time as _time

import numpy as np

LOGGED_NP_DATA_HEADER_SIZE = 3


class STATS_OBJECT:
    N_STEP = 0
    DISABLE_ALL_DEBUG = False
    DEBUG_STEP = 100
    DEBUG = False
    PRINT_DIM = 5
    LOGGED_CLASS_NAME = None

    # per-instance state is created lazily so subclasses need not call __init__
    def _stats(self):
        d = self.__dict__
        if "_stats_ready" not in d:
            d["LOGGED_NP_DATA"] = {}
            d["_timers"] = {}
            d["_ntimer"] = 0
            d["_stats_ready"] = True
        return d

    @property
    def timers(self):
        return [(k, v) for k, v in self._stats()["_timers"].items()]

    def __getattr__(self, name):
        if name == "LOGGED_NP_DATA":
            return self._stats()["LOGGED_NP_DATA"]
        raise AttributeError(name)

    def _add_np_log(self, key, step, float_row_data, g_step=0):
        logs = self._stats()["LOGGED_NP_DATA"]
        row = np.atleast_1d(np.squeeze(np.asarray(float_row_data, dtype=np.float64)))
        assert row.ndim == 1
        if key not in logs:
            logs[key] = np.zeros((0, row.size + LOGGED_NP_DATA_HEADER_SIZE))
        assert row.size + LOGGED_NP_DATA_HEADER_SIZE == logs[key].shape[1]
        logs[key] = np.vstack((logs[key], np.hstack(([g_step, step, _time.time()], row))))

    def _add_np_log_rows(self, key, steps, rows, g_step=0):
        """Many rows of one key at once (same layout as _add_np_log, one timestamp for the block): the per-iteration
        phase logs of a solve are written after it, and one vstack per row is 12 us x 4 keys x nit."""
        logs = self._stats()["LOGGED_NP_DATA"]
        rows = np.asarray(rows, dtype=np.float64)
        steps = np.asarray(steps, dtype=np.float64)
        assert rows.ndim == 2 and steps.shape == (rows.shape[0],)
        if key not in logs:
            logs[key] = np.zeros((0, rows.shape[1] + LOGGED_NP_DATA_HEADER_SIZE))
        assert rows.shape[1] + LOGGED_NP_DATA_HEADER_SIZE == logs[key].shape[1]
        head = np.empty((rows.shape[0], LOGGED_NP_DATA_HEADER_SIZE))
        head[:, 0] = g_step
        head[:, 1] = steps
        head[:, 2] = _time.time()
        logs[key] = np.vstack((logs[key], np.hstack((head, rows))))

    def save_np(self, path, postfix):
        os.makedirs(path, exist_ok=True)
        name = self.LOGGED_CLASS_NAME or self.__class__.__name__
        for key, arr in self._stats()["LOGGED_NP_DATA"].items():
            np.savetxt(os.path.join(path, "%s.%s.%s.txt" % (name, key, postfix)), arr, delimiter=",")

    def save(self, path, postfix):
        pass

    def status(self):
        if self.DEBUG:
            import pprint
            pprint.pprint(vars(self))

    def _print(self, *args, **kwargs):
        if self.DEBUG and not STATS_OBJECT.DISABLE_ALL_DEBUG and (self.N_STEP % self.DEBUG_STEP) in (0, 1, 2):
            print(("%6d\t" % self.N_STEP) + " ".join(map(str, args)), **kwargs)

    def _printalltime(self, *args, **kwargs):
        print(("%6d\t" % self.N_STEP) + ("%10s\t" % self.__class__.__name__) + " ".join(map(str, args)), **kwargs)

    def _debug(self, ENABLE, debug_step=100):
        self.DEBUG = ENABLE
        self.DEBUG_STEP = debug_step

    def _get_tic(self):
        st = self._stats()
        st["_ntimer"] += 1
        st["_timers"][st["_ntimer"]] = _time.time()
        return st["_ntimer"]

    def _get_tim(self, tic_id):
        t0 = self._stats()["_timers"].pop(tic_id, None)
        if t0 is None:
            raise Exception("no timer is found.")
        return (_time.time() - t0) * 1e6
