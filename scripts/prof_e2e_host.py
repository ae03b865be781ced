"""Host-side time budget of one end-to-end training step (wall clock per phase, CUDA-synchronised only at the end of the step)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd, torch
import bench
from reactranker_b200 import _lib
from reactranker_b200.data.load_reactions import DataProcessor, Parsing_features

wl = bench.WORKLOADS["c5"]
dev = torch.device("cuda:0")
_lib.require_device(0)
model, opt, sched, loss_fn = bench.build(wl, 0, 1)
pool = bench.make_pool(wl, 3, seed=1)
fz = Parsing_features()
for ds in pool:
    for tok, m in ds.mols.items():
        fz.add(tok, m)
frame = pd.concat([ds.to_dataframe().assign(flag=lambda d, i=i: d.flag + i * wl["groups"]) for i, ds in enumerate(pool)], ignore_index=True)
planner = DataProcessor(frame)
rows = wl["group"] * wl["groups"]
T = {}


def tick(name, t0):
    T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
    return time.perf_counter()


epoch = 0
n = 0
while n < 24:
    for b in planner.generate_batch_reactions(smiles_list=["rsmi_mapped", "psmi_mapped"], target_name="lgk", batch_size=rows, seed=epoch, add_features_name="temp"):
        if sum(b[2]) != rows:
            continue
        t0 = time.perf_counter()
        reactions, tg, sc, feats = b
        r_b, p_b = fz.parsing_reactions(reactions)
        t0 = tick("parse", t0)
        rg = r_b.to_device(dev)
        pg = p_b.to_device(dev)
        t0 = tick("to_device", t0)
        out = model(rg, pg, gpu=0, add_features=feats)
        t0 = tick("forward(launch)", t0)
        loss = loss_fn(out, sc, torch.FloatTensor(tg).squeeze())
        t0 = tick("loss", t0)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        t0 = tick("backward(launch)", t0)
        opt.step()
        sched.step()
        t0 = tick("opt+sched", t0)
        v = float(loss.detach().cpu().reshape(-1)[0])
        t0 = tick("sync+d2h", t0)
        n += 1
    epoch += 1
for k, v in T.items():
    v = v[6:]
    print(f"{k:18s} median {np.median(v):7.2f} ms   mean {np.mean(v):7.2f}   max {np.max(v):7.2f}")
