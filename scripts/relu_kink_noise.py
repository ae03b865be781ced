import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import reactranker_oracle as O
from reactranker_b200 import synthetic
from helpers import grads_close
torch.set_num_threads(8)
groups, n, hidden, depth = int(sys.argv[1]), int(sys.argv[2]), 300, 3
sizes = [n] * groups
ds = synthetic.make_dataset(4242, sizes)
sd = O.init_state_dict(hidden, 1, 1, True, seed=11)
r_o, p_o = O.OracleBatch([ds.mols[t] for t in ds.rsmi]), O.OracleBatch([ds.mols[t] for t in ds.psmi])
res = {}
for dt in (torch.float32, torch.float64):
    t0 = time.time()
    s = {k: v.to(dt) for k, v in sd.items()}
    params = {k: v.clone().requires_grad_(True) for k, v in s.items() if 'cached_zero' not in k}
    full = dict(s); full.update(params)
    out = O.model_forward(full, r_o, p_o, ds.temp.reshape(-1, 1))
    l = O.loss_for_task('mle', out, sizes, torch.tensor(ds.lgk.astype(np.float32)).to(dt)); l.backward(torch.ones_like(l))
    res[dt] = (out.detach().double().numpy(), float(l.detach().sum()), {k: v.grad.double().numpy() for k, v in params.items()})
    print(dt, 'took', time.time() - t0, flush=True)
a, b = res[torch.float32], res[torch.float64]
print('scores rel', np.abs(a[0]-b[0]).max()/np.abs(b[0]).max(), 'loss rel', abs(a[1]-b[1])/abs(b[1]))
gs = max(np.abs(v).max() for v in b[2].values())
for k in b[2]:
    e = np.abs(a[2][k]-b[2][k]).max(); m = np.abs(b[2][k]).max()
    print(f"{k:30s} max-rel {e/max(m,1e-30):.2e}  relL2 {np.linalg.norm(a[2][k]-b[2][k])/max(np.linalg.norm(b[2][k]),1e-30):.2e}  (max {m/gs:.1e} of gscale)")
