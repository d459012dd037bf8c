"""Development aid (GPU): time the headline shapes with every library in build_abl/ (one subprocess each, PIORAN_B200_LIB) and
check each against the oracle on a few parameter vectors.  usage: python tools/variants_gpu.py [lib ...]"""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import json, os, sys
sys.path.insert(0, %r)
import numpy as np
import pioran_b200 as pb
from oracle import oracle as orc
ctx = pb.get_context(0)
rng = np.random.default_rng(5)
N = 1000
t = np.cumsum(0.05 + rng.exponential(1.0, N)); y = rng.normal(0, 1, N); s2 = rng.uniform(0.01, 0.05, N) ** 2
fm, fx = 1 / (t[-1] - t[0]), 1 / np.min(np.diff(t)) / 2
ser = ctx.upload_series(t, y, s2)
out = {}
for basis, B in (("DRWCelerite", 65536), ("SHO", 65536), ("DRWCelerite", 4096)):
    spec = pb.make_spec("SingleBendingPowerLaw", fm, fx, 20, basis_function=basis)
    th = np.column_stack([rng.uniform(0, 1.5, B), np.exp(rng.uniform(np.log(fm / 5), np.log(fx * 5), B)), rng.uniform(1.5, 4, B),
                          np.exp(rng.normal(-3, 1.4, B)), rng.gamma(2, 0.5, B), rng.normal(0, 1, B)])
    R = 40 if basis == "SHO" else 60
    flops = B * N * (4 * R * R + 13 * R + 40)
    got = ctx.approx_logl(ser, spec, th)[0]
    ms = []
    for _ in range(4):
        ctx.approx_logl(ser, spec, th); ms.append(ctx.last_kernel_ms())
    o = orc.approx_logl_batch("SBPL", th[:32], fm, fx, 20, t, y, s2, basis=basis, nthreads=8)
    err = float(np.nanmax(np.abs(got[:32] - o) / np.maximum(1, np.abs(o))))
    out[f"{basis}_{B}"] = {"ms": round(min(ms), 3), "frac": round(flops / min(ms) / 1e9 / 36.97, 4), "err": err}
print(json.dumps(out))
''' % ROOT
libs = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "build_abl", "lib_*.so")))
for lib in libs:
    env = dict(os.environ, PIORAN_B200_LIB=os.path.abspath(lib))
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("ERR " + r.stderr[-400:])
    print(os.path.basename(lib), line, flush=True)
