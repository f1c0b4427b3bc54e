"""A few launches of the materialising gather (for ncu): `python scripts/prof_gather.py [rows]`, default the bench
shape (4096 rows); 65536 rows is the HBM-bound regime (710 MB of output per launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from recnn_b200 import _lib
L = _lib.lib(); dev = "cuda:0"
N, NI, D, F = (int(sys.argv[1]) if len(sys.argv) > 1 else 4096), 26744, 128, 10
rng = np.random.default_rng(0)
table = torch.from_numpy(rng.standard_normal((NI, D), dtype=np.float32)).to(dev)
items = torch.from_numpy(rng.integers(0, NI, size=(N, F + 1), dtype=np.int64)).to(dev)
ratings = torch.from_numpy(rng.integers(-4, 6, size=(N, F + 1)).astype(np.float32)).to(dev)
S = F * D + F
state = torch.empty(N, S, device=dev); nxt = torch.empty(N, S, device=dev)
act = torch.empty(N, D, device=dev); rew = torch.empty(N, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    flush.zero_()
    _lib.check(L.recnn_frame_gather(table.data_ptr(), NI, D, items.data_ptr(), ratings.data_ptr(), N, F, state.data_ptr(),
                                    nxt.data_ptr(), act.data_ptr(), rew.data_ptr(), None, st))
torch.cuda.synchronize(); print("done")
