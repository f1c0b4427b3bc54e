import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import recnn_b200
from oracle import cases as C, recnn_oracle as O
from tests._cuda import build_nets, build_optimizers, dump_grad, dump_net
tag = sys.argv[1]; opt = sys.argv[2]
spec = C.CASES["canon"]; inp = C.make_inputs(spec, "ddpg"); dev = torch.device("cuda:0")
nets = build_nets(spec, inp, dev); opts = build_optimizers(opt, nets, "ddpg")
params = dict(C.DDPG_PARAMS)
base = {"items": torch.from_numpy(inp["items"]), "ratings": torch.from_numpy(inp["ratings"]),
        "sizes": torch.from_numpy(inp["sizes"]), "table": torch.from_numpy(inp["table"]).to(dev)}
out = {}
for step in range(3):
    batch = dict(base); batch["dropout_masks"] = [torch.from_numpy(m) for m in inp["masks"][step]]
    loss = recnn_b200.nn.ddpg_update(batch, params, nets, opts, dev, {}, recnn_b200.utils.DummyWriter(), learn=True, step=step)
    for k, v in dump_grad(nets["value_net"]).items(): out["s%d.g.%s" % (step, k)] = v
    for k, v in dump_net(nets["value_net"]).items(): out["s%d.p.%s" % (step, k)] = v
    out["s%d.loss" % step] = np.array([loss["value"], loss["policy"]])
np.savez("gpurun_out/grads_%s_%s.npz" % (tag, opt), **out)
