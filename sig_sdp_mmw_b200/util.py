"""Log writers with the reference's surface (sim_src/util.py:221-272): the drivers log one CSV
row `[g_iteration, iteration, *values]` per data point into a per-run folder."""
import csv
import os
import time


class CSV_WRITER_OBJECT:
    def __init__(self, path=None):
        self.path = path
        if path is not None:
            os.makedirs(path, exist_ok=True)
        self.files = {}
        self.writers = {}

    def _writer(self, data_name):
        if data_name not in self.files:
            self.files[data_name] = open(os.path.join(self.path, data_name), "w", newline="")
            self.writers[data_name] = csv.writer(self.files[data_name])
        return self.writers[data_name]

    def log_one_scalar(self, data_name, iteration, value, g_iteration=0):
        if self.path is None:
            return
        self._writer(data_name).writerow([g_iteration, iteration, value])
        self.files[data_name].flush()

    def log_mul_scalar(self, data_name, iteration, values, g_iteration=0):
        if self.path is None:
            return
        self._writer(data_name).writerow([g_iteration, iteration] + [v for v in values])
        self.files[data_name].flush()

    def close(self):
        for f in self.files.values():
            f.close()


def GET_LOG_PATH_FOR_SIM_SCRIPT(sim_script_path):
    base = os.path.splitext(os.path.basename(sim_script_path))[0]
    folder = os.path.join(os.path.dirname(os.path.realpath(sim_script_path)), base)
    os.makedirs(folder, exist_ok=True)
    return os.path.join(folder, base + "-" + time.strftime("%Y-%B-%d-%H-%M-%S") + "-ail")
