"""Validation-pass throughput: ranking_metrics over a synthetic validation frame at the c5 model size, wall clock
per pass with a final device synchronise, and the share of the device-side metrics kernel.  The reference runs one forward per group and
sorts on the host; the figure it would be compared with is groups/s of that loop (CPU oracle: bench.py --impl reference)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from reactranker_b200 import _lib, synthetic
from reactranker_b200.data.load_reactions import DataProcessor, Parsing_features
from reactranker_b200.models.base_model import build_model
from reactranker_b200.train import eval as E

G, N = (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "410,50").split(","))
_lib.require_device(0)
ds = synthetic.make_dataset(7, [N] * G)
fz = Parsing_features(ds.mols)
dp = DataProcessor(ds.to_dataframe())
torch.manual_seed(0)
model = build_model(hidden_size=300, mpnn_depth=3, mpnn_diff_depth=3, ffn_depth=3, use_bias=True, dropout=0.1, task_num=1,
                    ffn_last_layer="with_softplus", add_features_dim=1).cuda(0)
kw = dict(smiles_list=["rsmi_mapped", "psmi_mapped"], target_name="lgk", add_features_name="temp")
times = []
for it in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = E.ranking_metrics(model, 0, dp, fz, show_info=False, **kw)
    torch.cuda.synchronize()
    times.append(time.perf_counter() - t0)
ms = np.median(times[2:]) * 1e3
# the metrics kernel alone on the same shapes
scores = torch.randn(G * N, device="cuda:0")
t = np.random.default_rng(0).normal(size=G * N)
for _ in range(3):
    E.group_metrics(scores, [N] * G, t)
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for _ in range(20):
    E.group_metrics(scores, [N] * G, t)
e.record()
torch.cuda.synchronize()
print(f"ranking_metrics: {G} groups x {N} candidates, {ms:.1f} ms per pass (first pass {times[0] * 1e3:.0f} ms incl. store upload) = "
      f"{G * N / ms * 1e3:,.0f} reactions/s, {G / ms * 1e3:,.0f} groups/s; metrics {r[0]:.3f} {r[1]:.3f} {r[2]:.3f} {np.round(r[3], 3)}")
print(f"group_metrics (upload of targets + rr_rank_metrics): {a.elapsed_time(e) / 20 * 1e3:.0f} us per call")
