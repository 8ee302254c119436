"""Builds libsigsdp_mmw.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m sig_sdp_mmw_b200.build [--force] [-v]

The kernels are instantiated per (sketch dtype, lanes per row) in eight translation units
(csrc/mmw_inst.cu with -DSIGSDP_T / -DSIGSDP_G) that compile in parallel next to the C ABI
(csrc/mmw_api.cu), the host plan builder (csrc/plan_host.cpp) and the device plan builder
(csrc/plan_device.cu); objects go to csrc/_obj/.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
OUT = os.path.join(HERE, "libsigsdp_mmw.so")
DEPS = ["mmw_api.cu", "mmw_inst.cu", "plan_host.cpp", "plan_device.cu", "numpy_stream.cpp", "mmw_device.cuh", "mmw_kernels.cuh", "plan_host.h",
        "plan_device.h",
        os.path.join("..", "..", "include", "sigsdp_mmw.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
INSTANCES = [(t, g) for t in ("double", "float") for g in (4, 8, 16, 32)]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return os.path.getmtime(os.path.abspath(__file__)) > t or any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def _units():
    units = [("mmw_api.o", ["mmw_api.cu"]), ("plan_host.o", ["plan_host.cpp"]), ("plan_device.o", ["plan_device.cu"]), ("numpy_stream.o", ["-Xcompiler", "-ffp-contract=off", "numpy_stream.cpp"])]   # (no fused multiply-add: numpy rounds every product)
    for t, g in INSTANCES:
        tag = "%s_g%d" % ("f64" if t == "double" else "f32", g)
        units.append(("inst_%s.o" % tag, ["-DSIGSDP_T=%s" % t, "-DSIGSDP_G=%d" % g, "-DSIGSDP_NAME=ks_%s" % tag,
                                          "mmw_inst.cu"]))
    return units


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    os.makedirs(OBJ, exist_ok=True)
    log = []

    def compile_one(unit):
        obj, args = unit
        cmd = [nvcc] + COMMON + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", os.path.join(OBJ, obj)] + args
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        return unit, cmd, r

    with ThreadPoolExecutor(max_workers=min(len(_units()), os.cpu_count() or 4)) as ex:
        for unit, cmd, r in ex.map(compile_one, _units()):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed: " + " ".join(cmd))
            if verbose:
                log.append("==== %s\n%s" % (unit[0], r.stderr))
    objs = [os.path.join(OBJ, u[0]) for u in _units()]
    cmd = [nvcc] + ARCH + ["-shared", "-Xcompiler", "-fPIC", "-o", OUT] + objs
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed: " + " ".join(cmd))
    if verbose:
        sys.stderr.write("\n".join(log))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
