import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests._golden import run_oracle_case
from tests._cuda import run_cuda_case
algo, case, opt = sys.argv[1], sys.argv[2], sys.argv[3]
want = run_oracle_case(case, algo, opt)
got = run_cuda_case(case, algo, opt, form="frames")
for k in sorted(want):
    if k.startswith("loss."):
        print(k, "rel", np.abs(got[k] - want[k]) / (np.abs(want[k]) + 0.1))
for k in sorted(want):
    if k.startswith("grad_") and k.endswith(".sample"):
        sc = np.abs(want[k]).max()
        print(k, "max err/scale %.3g" % (np.abs(got[k] - want[k]).max() / sc))
for k in sorted(want):
    if k.startswith("final."):
        sc = np.abs(want[k]).max()
        d = np.abs(got[k] - want[k])
        print(k, "max abs %.3g (scale %.3g) frac>1e-6*scale %.3f" % (d.max(), sc, (d > 1e-6 * sc).mean()))
