"""configs[1] as a user would run it twice: (solve lu b) on a stored factorisation, n = 8192, 64 right-hand sides.
Reference behaviour (re-marshal the n x n factors on every call, src/linear-algebra.lisp:264-276) vs the device-resident object."""
import sys, time, numpy as np
sys.path.insert(0, ".")
from lla_b200 import lla as L
n, nrhs = (int(sys.argv[1]) if len(sys.argv) > 1 else 8192), 64
rng = np.random.default_rng(1)
a = rng.uniform(0, 10, (n, n)); b = rng.uniform(0, 10, (n, nrhs))
def wall(f, reps=3):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); best = min(best, time.perf_counter() - t0)
    return best, r
t_lu, f = wall(lambda: L.lu(a), 2)
t_lud, fd = wall(lambda: L.lu_device(a), 2)
t_s, x = wall(lambda: L.solve(f, b))
t_sd, xd = wall(lambda: L.solve(fd, b))
res = np.linalg.norm(a @ xd - b) / (n * np.finfo(float).eps * np.linalg.norm(a) * np.linalg.norm(xd))
print(f"n={n} nrhs={nrhs}: lu host-path {t_lu*1e3:.1f} ms, lu_device {t_lud*1e3:.1f} ms; solve(lu,b) re-marshalling {t_s*1e3:.1f} ms, "
      f"device-resident {t_sd*1e3:.1f} ms ({t_s/t_sd:.1f}x); max |x - x_dev| {np.abs(x-xd).max():.2e}, scaled residual {res:.4f}", flush=True)
# row-major-native entry points (host pointers, no device-resident state): what LLA's (solve a b) and (lu a) cost a caller
t_sv, xs = wall(lambda: L.solve(a, b), 2)
t_svn, xn = wall(lambda: L.solve(a, b, native=True), 2)
t_lun, fnat = wall(lambda: L.lu(a, native=True), 2)
print(f"n={n}: solve(a,b) Fortran path with host transposes {t_sv*1e3:.1f} ms, llacuda_dgesv_rm {t_svn*1e3:.1f} ms ({t_sv/t_svn:.1f}x), "
      f"identical {np.array_equal(xs, xn)}; lu native {t_lun*1e3:.1f} ms (host path {t_lu*1e3:.1f} ms), identical {np.array_equal(fnat.lu, f.lu)}", flush=True)
