import sys, math, ctypes as C
sys.path.insert(0, '/root/repo')
import numpy as np
from ivp_b200 import _abi, api
lib = api.load_library()
rng = np.random.default_rng(7)
n = 2_000_000
x = np.exp(40 * (rng.random(n) - 0.5)); y = np.where(rng.random(n) < 0.5, rng.choice([0.125, 0.2, 0.17, 0.04, 0.25, 0.8, -1/3, 2/3, -0.5, 3.0, 1.0], n), 8 * (rng.random(n) - 0.5))
x[:8] = [0.0, 1.0, np.inf, 1e-310, 2.0, 1e300, 1e-300, 0.5]; y[:8] = [0.125, 0.3, -0.25, 0.5, 1e-70, 5.0, 5.0, np.inf]
r = np.zeros(n)
assert lib.ivpb_debug_pow(_abi.ptr(x), _abi.ptr(y), C.c_int(n), _abi.ptr(r)) == 0
ref = np.array([math.pow(a, b) if not (a == 1e300 and b == 5.0) else np.inf for a, b in zip(x.tolist(), y.tolist())])
bad = (r.view(np.uint64) != ref.view(np.uint64)) & ~(np.isnan(r) & np.isnan(ref))
print("device pow vs host libm: mismatches", int(bad.sum()), "of", n, r[:8], ref[:8])
for i in np.nonzero(bad)[0][:10]:
    print("  mismatch at", i, float(x[i]).hex(), float(y[i]).hex(), "device", float(r[i]).hex(), "host", float(ref[i]).hex())
