"""ctypes loader of oracle/cpu_scan.c — TEST / BASELINE INFRASTRUCTURE ONLY.

Compiles the C restatement for the host it runs on (-march=native; the prebuilt oracle/_build copy is a
fallback when no compiler is present) and exposes the per-constraint scans plus `numeric_suite`, the
reference's one-scan-per-constraint schedule for the C2 business-rules suite."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(HERE, "cpu_scan.c")
    out = os.path.join(tempfile.gettempdir(), f"libcpu_scan_{os.getuid()}_{int(os.path.getmtime(src))}.so")
    if not os.path.exists(out):
        try:
            subprocess.run(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", "-o", out, src, "-lm"],
                           check=True, capture_output=True)
        except Exception:
            out = os.path.join(HERE, "_build", "libcpu_scan.so")
    L = C.CDLL(out)
    P, I64, D = C.c_void_p, C.c_int64, C.c_double
    L.to_num_threads.restype = C.c_int
    L.to_set_num_threads.argtypes = [C.c_int]
    L.to_count_valid.restype = I64
    L.to_count_valid.argtypes = [P, I64]
    L.to_min_max_f64.argtypes = [P, P, I64, C.POINTER(D), C.POINTER(D), C.POINTER(I64)]
    L.to_min_max_sum_i64.argtypes = [P, P, I64, C.POINTER(I64), C.POINTER(I64), C.POINTER(I64), C.POINTER(D), C.POINTER(I64)]
    L.to_sum_f64.argtypes = [P, P, I64, C.POINTER(D), C.POINTER(I64)]
    L.to_var_f64.argtypes = [P, P, I64, C.POINTER(D), C.POINTER(I64)]
    L.to_corr_f64.argtypes = [P, P, P, P, I64, C.POINTER(D), C.POINTER(D), C.POINTER(I64)]
    L.to_fused_c2.argtypes = [P, P, P, P, P, P, P, P, D, I64, I64, C.POINTER(D)]
    L.to_pred_gt_lt.restype = I64
    L.to_pred_gt_lt.argtypes = [P, P, D, P, P, I64, I64]
    _lib = L
    return L


def num_threads():
    return int(lib().to_num_threads())


def use_all_host_threads():
    """All the CPUs this process may run on, whatever OMP_NUM_THREADS says (torchrun sets it to 1 per rank).
    TG_BENCH_CPU_THREADS overrides."""
    n = int(os.environ.get("TG_BENCH_CPU_THREADS", "0")) or len(os.sched_getaffinity(0))
    lib().to_set_num_threads(n)
    return num_threads()


def _p(a):
    return None if a is None else a.ctypes.data


def count_valid(validity, n):
    return int(lib().to_count_valid(_p(validity), n))


def min_max_f64(v, validity):
    mn, mx, c = C.c_double(), C.c_double(), C.c_int64()
    lib().to_min_max_f64(_p(v), _p(validity), len(v), C.byref(mn), C.byref(mx), C.byref(c))
    return mn.value, mx.value, c.value


def min_max_sum_i64(v, validity):
    mn, mx, s, c, fs = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_double()
    lib().to_min_max_sum_i64(_p(v), _p(validity), len(v), C.byref(mn), C.byref(mx), C.byref(s), C.byref(fs), C.byref(c))
    return mn.value, mx.value, s.value, fs.value, c.value


def sum_f64(v, validity):
    s, c = C.c_double(), C.c_int64()
    lib().to_sum_f64(_p(v), _p(validity), len(v), C.byref(s), C.byref(c))
    return s.value, c.value


def var_f64(v, validity):
    s, c = C.c_double(), C.c_int64()
    lib().to_var_f64(_p(v), _p(validity), len(v), C.byref(s), C.byref(c))
    return s.value, c.value


def corr_f64(x, vx, y, vy):
    r, cv, c = C.c_double(), C.c_double(), C.c_int64()
    lib().to_corr_f64(_p(x), _p(vx), _p(y), _p(vy), len(x), C.byref(r), C.byref(cv), C.byref(c))
    return r.value, cv.value, c.value


def pred_gt_lt(f, vf, a, i, vi, b):
    return int(lib().to_pred_gt_lt(_p(f), _p(vf), float(a), _p(i), _p(vi), int(b), len(f)))


def numeric_suite(cols, n):
    """The literal C2 suite as run_sequential runs it: has_size, has_min(f0), has_mean(f1),
    has_correlation(f0,f1), satisfies(f2 > 0 AND i0 < 1000000) — one full scan per constraint.
    cols[name] = (values ndarray, validity uint8 bitmap | None). Returns the metrics."""
    out = {"size": float(n)}
    out["min_f0"] = min_max_f64(*cols["f0"])[0]
    s, c = sum_f64(*cols["f1"])
    out["mean_f1"] = s / c if c else float("nan")
    out["corr_f0_f1"] = corr_f64(cols["f0"][0], cols["f0"][1], cols["f1"][0], cols["f1"][1])[0]
    out["satisfies"] = pred_gt_lt(cols["f2"][0], cols["f2"][1], 0.0, cols["i0"][0], cols["i0"][1], 1000000) / n
    return out


def numeric_suite_fused(cols, n):
    """The same five constraints in ONE scan over the four referenced columns (the fused CPU variant of BASELINE.md
    §2.2: what a working optimizer/ would hand DataFusion). Same keys as numeric_suite."""
    out = (C.c_double * 4)()
    lib().to_fused_c2(_p(cols["f0"][0]), _p(cols["f0"][1]), _p(cols["f1"][0]), _p(cols["f1"][1]), _p(cols["f2"][0]), _p(cols["f2"][1]),
                      _p(cols["i0"][0]), _p(cols["i0"][1]), 0.0, 1000000, n, out)
    return {"size": float(n), "min_f0": out[0], "mean_f1": out[1], "corr_f0_f1": out[2], "satisfies": out[3]}


def pack_validity(mask: np.ndarray) -> np.ndarray:
    """bool mask -> Arrow validity bitmap (LSB first), padded to 64 bytes"""
    b = np.packbits(mask.astype(np.uint8), bitorder="little")
    return np.concatenate([b, np.zeros((-len(b)) % 64 + 64, dtype=np.uint8)])
