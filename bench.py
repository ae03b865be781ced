#!/usr/bin/env python
"""Benchmark of the ReactRanker training hot path (D-MPNN reaction encoder + LTR loss) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c5|c5-500|c2|c3|c4]

One "step" = one training step over one batch of synthetic reaction graphs: forward, loss,
backward, (gradient all-reduce,) Adam, NoamLR.  Prints ONE JSON line (rank 0).

* ``value``  : reactions/s with the batch already resident in HBM (device path).
* ``e2e``    : reactions/s through the reference-shaped public API from HOST buffers -- per step the batch plan
               (``DataProcessor.generate_batch_reactions``), ``Parsing_features.parsing_reactions`` on the warm
               MolGraph cache, the pinned host->device copies (molecule ids, row offsets, extra features, targets),
               on-device batch assembly from the HBM-resident molecule store, forward/loss/backward/Adam, and a
               device->host read of the loss.  Every step does all of these inside the timed region; the plan / featurise /
               upload of batch i+1 is issued between enqueuing step i and reading its loss (data/prefetch.py: Lookahead), the
               way the reference's own loop overlaps them when it does not read the loss back.
* ``roofline``: the dominant kernel class of the step, timed with CUDA events on the launching stream around every
               launch (rr_profile_begin/end of librr_sm100) while the same K steps run a second time; ``value`` comes from the
               first, uninstrumented pass (``roofline.ms_per_step_instrumented`` is the second pass's step time).
* ``cpu_baseline`` / ``--impl reference``: the CPU restatement of the reference path
               (oracle/reactranker_oracle.py; the reference is Python and /root/reference does not
               travel to the GPU box) on the host cores, on a bounded sample of the same workload.
               ``--impl reference --ref-device cuda`` prints an extra, informative line instead: the same eager PyTorch code with
               its tensors on cuda:0 (the reference's gpu=0 mode), full batch, with and without the per-step graph build.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[4] (the metric's own: D-MPNN + ListMLE, 1/2/4/8 GPUs), first sweep point: 50 candidates/group
    "c5": dict(task="mle", hidden=300, depth=3, diff_depth=3, task_num=1, task_type=None, last="with_softplus", dropout=0.1,
               group=50, groups=82, desc="ListMLE D-MPNN h300 d3/3/3, 82 groups x 50 candidates = 4100 reactions per GPU per step (configs[4])"),
    "c5-500": dict(task="mle", hidden=300, depth=3, diff_depth=3, task_num=1, task_type=None, last="with_softplus", dropout=0.1,
                   group=500, groups=8, desc="ListMLE D-MPNN h300 d3/3/3, 8 groups x 500 candidates = 4000 reactions per GPU per step (configs[4])"),
    "c2": dict(task="listnet", hidden=300, depth=3, diff_depth=3, task_num=1, task_type=None, last="with_softplus", dropout=0.1,
               group=32, groups=128, desc="ListNet@1 D-MPNN h300 d3/3/3, 128 groups x 32 candidates = 4096 reactions per GPU per step (configs[1])"),
    # configs[2]: RankNet (main_ranknet.py: dropout 0.2, no_softplus), 64 candidates/group, 4096 rows per optimiser step = one accumulation
    # window of 64 groups, every group its own segment (own padding rows and max_num_bonds) as in the reference's one-forward-per-group loop
    "c3": dict(task="ranknet", hidden=300, depth=3, diff_depth=3, task_num=1, task_type=None, last="no_softplus", dropout=0.2,
               group=64, groups=64, desc="RankNet (sum_session) D-MPNN h300 d3/3/3, window of 64 groups x 64 candidates = 4096 reactions, 258k ordered pairs (configs[2])"),
    "c4": dict(task="evidential_ranking", hidden=600, depth=5, diff_depth=5, task_num=2, task_type="evidential_ranking", last="with_softplus",
               dropout=0.1, group=32, groups=128, desc="UC-Listwise D-MPNN h600 d5/5/3, 128 groups x 32 = 4096 reactions per GPU per step (configs[3])"),
}
METRIC = "train reactions/sec, D-MPNN+ListMLE"                      # BASELINE.json's metric: the default workload (c5)
LOSS_NAME = {"mle": "ListMLE", "listnet": "ListNet@1", "ranknet": "RankNet", "evidential_ranking": "UC-Listwise"}


def metric_of(wl):
    return METRIC if wl["task"] == "mle" else f"train reactions/sec, D-MPNN+{LOSS_NAME[wl['task']]}"
UNIT = "reactions/s"


_TRAFFIC_KERNELS = {"gemm_fwd": "k_tc_gemm2", "gemm_dgrad": "k_tc_gemm2", "gemm_wgrad": "k_tc_wgrad", "bond_fwd": "k_rowpipe<0", "bond_bwd": "k_rowpipe<1",
                    "nbr_fwd": "k_rowpipe<2", "nbr_bwd": "k_rowpipe<3"}


def ncu_traffic(cls):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the class's kernel, from the newest committed ncu --set full summary
    (profiles/rNN_traffic.json, written by scripts/summarise_profiles.py); None when no capture is committed."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    pat = _TRAFFIC_KERNELS.get(cls)
    if not files or pat is None:
        return None
    hits = []
    for f in reversed(files):                      # the newest capture that holds this kernel (a partial re-capture only lists what changed)
        try:
            t = json.load(open(f))
        except Exception:
            continue
        hits = [v for k, v in t.items() if k.startswith(pat)]
        if hits:
            break
    if not hits:
        return None
    n = sum(v["launches_captured"] for v in hits)
    return sum(v["dram_bytes_per_launch"] * v["launches_captured"] for v in hits) / n


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return dict(hbm=float(p["hbm_gbs"]), tensor=float(p["bf16_tflops_sustained"]), src="MEASURED_PEAKS.json (measured; bf16 sustained)")
    except Exception:
        return dict(hbm=6650.0, tensor=1400.0, src="fallback of B200_PROFILING.md")


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(self.NAMES, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def make_pool(wl, n_batches, seed):
    """``n_batches`` distinct batches of synthetic reactions (SURVEY.md §8d generator)."""
    from reactranker_b200 import synthetic
    pool = []
    for b in range(n_batches):
        ds = synthetic.make_dataset(seed * 1000 + b, [wl["group"]] * wl["groups"])
        pool.append(ds)
    return pool


def algorithmic_work(wl, rg, pg, n_add=1):
    """Per-step algorithmic work of each kernel class (forward + backward), SURVEY.md §8(d):
    compulsory fp32 activation bytes + int32 indices for the gather kernels, 2MNK for the GEMMs with
    the reference's logical dimensions (83/61/h, not the padded strides)."""
    h, T, Td = wl["hidden"], wl["depth"] - 1, wl["diff_depth"] - 1
    s = i = 4
    N = pg.n_mols
    out = {}
    bond = nbr_f = nbr_b = 0.0
    g_f = g_d = 0.0
    bond_b = 0.0
    for g in (rg, pg):
        A, B, W = g.n_atoms, g.n_bonds, g.c.wmax
        bond += T * (2 * B * h * s + (A * W + 2 * B) * i)
        # backward gathers carry the ReLU/dropout backward that follows them (rr_mp_pipe.cu): besides the gather's own read + write,
        # the mask source y is read and the running sum d(input) is read + written; the last one of a graph writes only that sum
        bond_b += max(T - 1, 0) * (5 * B * h * s + (A * W + 2 * B) * i) + (1 if T >= 1 else 0) * (4 * B * h * s + (A * W + 2 * B) * i)
        agg = (B + A) * h * s + A * W * i
        nbr_f += agg
        nbr_b += (A + 3 * B) * h * s + A * W * i            # dout[A] -> dz[B] masked by m^T (or [inp > 0]), d(input) = dz
        g_f += 2.0 * B * 83 * h + T * 2.0 * B * h * h + 2.0 * A * (61 + h) * h
        g_d += T * 2.0 * B * h * h + 2.0 * A * h * h
    A, B, W = pg.n_atoms, pg.n_bonds, pg.c.wmax
    a2a = 2 * A * h * s + A * W * i
    nbr_f += (Td + 1) * a2a + ((B + A) * 83 * s + A * W * i if Td > 0 else 0)
    nbr_b += 2 * (4 * A * h * s + A * W * i) + max(Td - 1, 0) * (5 * A * h * s + A * W * i) if Td > 0 else (4 * A * h * s + A * W * i)
    g_f += 2.0 * A * h * h + Td * 2.0 * A * (h + 83) * h + 2.0 * A * 2 * h * h
    g_d += 2.0 * A * h * h + Td * 2.0 * A * h * h + 2.0 * A * 2 * h * h
    ffn = 2.0 * N * ((h + n_add) * h + h * h + h * wl["task_num"])
    g_f += ffn
    g_d += ffn
    out["bond_fwd"] = (bond, "B")
    out["bond_bwd"] = (bond_b, "B")
    out["nbr_fwd"] = (nbr_f, "B")
    out["nbr_bwd"] = (nbr_b, "B")
    out["gemm_fwd"] = (g_f, "FLOP")
    out["gemm_dgrad"] = (g_d, "FLOP")
    out["gemm_wgrad"] = (g_f, "FLOP")
    return out


def build(wl, dev_index, world):
    from reactranker_b200.models.base_model import build_model
    from reactranker_b200.train import loss as RL
    from reactranker_b200.train.utils import build_lr_scheduler, build_optimizer
    torch.manual_seed(0)
    model = build_model(hidden_size=wl["hidden"], mpnn_depth=wl["depth"], mpnn_diff_depth=wl["diff_depth"], ffn_depth=3, use_bias=True,
                        dropout=wl["dropout"], task_num=wl["task_num"], ffn_last_layer=wl["last"], task_type=wl["task_type"], add_features_dim=1)
    model = model.cuda(dev_index).train()
    opt = build_optimizer(model)
    rows = wl["group"] * wl["groups"]
    sched = build_lr_scheduler(opt, warmup_epochs=2, total_epochs=30, train_data_size=100 * rows, batch_size=rows, init_lr=1e-4, max_lr=1e-3, final_lr=1e-4)
    G, N = wl["groups"] * world, rows * world
    task = wl["task"]
    if task == "mle":
        lm = RL.MLEloss(global_norm=G if world > 1 else None)
        loss_fn = lambda out, scope, t: lm(out, scope, t, dev_index)  # noqa: E731
    elif task == "listnet":
        lm = RL.ListnetLoss(global_norm=N if world > 1 else None)
        loss_fn = lambda out, scope, t: lm(out, scope, t, dev_index)  # noqa: E731
    elif task == "ranknet":
        # every group of the synthetic pool has distinct targets: n (n - 1) ordered pairs each; under DP the divisor is the global window's
        pairs = float(world * wl["groups"] * wl["group"] * (wl["group"] - 1))
        loss_fn = lambda out, scope, t: RL.ranknet_window_loss(out, scope, t, pairs, sigma=1.0, gpu=dev_index)  # noqa: E731
    else:
        lm = RL.evidential_ranking(global_norm=G if world > 1 else None)
        loss_fn = lambda out, scope, t: lm(out, scope, t, 0.0001, 0, 1, dev_index)  # noqa: E731
    return model, opt, sched, loss_fn


def ours(args):
    from reactranker_b200 import _lib
    from reactranker_b200.data.load_reactions import Parsing_features
    from reactranker_b200.features.featurization import BatchMolGraph

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")       # stdout carries exactly one JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    wl = dict(WORKLOADS[args.workload])
    if args.dropout is not None:
        wl["dropout"] = args.dropout
        wl["desc"] += f", dropout {args.dropout:g}"
    dedup = wl["dropout"] == 0 and not args.no_dedup       # exact only without dropout (rr_model_cfg.r_atom_map)
    if dedup:
        wl["desc"] += ", repeated reactants encoded once"
    _lib.require_device(local)
    _lib.check(_lib.lib().rr_set_gemm_mode(1 if args.gemm == "tc" else 0))
    model, opt, sched, loss_fn = build(wl, local, world)
    params = [p for p in model.parameters() if p.requires_grad]
    scope = [wl["group"]] * wl["groups"]
    rows = sum(scope)

    pool = make_pool(wl, args.pool, seed=1 + rank)
    fz = Parsing_features()
    for ds in pool:
        for tok, m in ds.mols.items():
            fz.add(tok, m)
    # the pool as ONE data set for the end-to-end leg: DataProcessor plans the batches like train() does
    import pandas as pd
    from reactranker_b200.data.load_reactions import DataProcessor
    frame = pd.concat([ds.to_dataframe().assign(flag=lambda d, i=i: d.flag + i * wl["groups"]) for i, ds in enumerate(pool)], ignore_index=True)
    planner = DataProcessor(frame)
    # device-resident copies for the device-path measurement
    from reactranker_b200.features.featurization import DeviceGraph
    per_group = wl["task"] == "ranknet"          # RankNet: one segment per group (train_pairwise.py: each group is its own forward)

    def split(batch_of, tokens):
        g = wl["group"] if per_group else len(tokens)
        return [batch_of(tokens[i:i + g]) for i in range(0, len(tokens), g)]

    def to_dev_pair(batch_of, r_tokens, p_tokens):
        if dedup:
            return DeviceGraph.from_batches_dedup(split(fz.parsing_smiles, r_tokens), split(fz.parsing_smiles, p_tokens), dev)
        return DeviceGraph.from_batches(split(batch_of, r_tokens), dev), DeviceGraph.from_batches(split(batch_of, p_tokens), dev)

    resident = []
    for ds in pool:
        mk = lambda toks, ds=ds: BatchMolGraph([ds.mols[t] for t in toks])  # noqa: E731
        resident.append(to_dev_pair(mk, list(ds.rsmi), list(ds.psmi)) + (torch.tensor(ds.temp.reshape(-1, 1), dtype=torch.float32, device=dev),
                                                                          torch.tensor(ds.lgk, dtype=torch.float32, device=dev)))
    torch.cuda.synchronize()

    def reduce_grads():
        if dist is None:
            return
        flat = torch._utils._flatten_dense_tensors([p.grad for p in params])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        torch._foreach_copy_([p.grad for p in params], torch._utils._unflatten_dense_tensors(flat, [p.grad for p in params]))

    def step_resident(i):
        rg, pg, feats, targets = resident[i % len(resident)]
        out = model(rg, pg, gpu=local, add_features=feats)
        loss = loss_fn(out, scope, targets)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        reduce_grads()
        opt.step()
        sched.step()
        return loss

    h2d = [0]
    d2h = 4

    batches = {"it": None, "epoch": 0}

    def next_batch():
        while True:
            if batches["it"] is None:
                batches["it"] = planner.generate_batch_reactions(smiles_list=["rsmi_mapped", "psmi_mapped"], target_name="lgk", batch_size=rows,
                                                                 seed=batches["epoch"], add_features_name="temp")
                batches["epoch"] += 1
            for b in batches["it"]:
                if sum(b[2]) == rows:
                    return b
            batches["it"] = None

    def endless_plan():
        while True:
            yield next_batch()                                                # batch plan (DataProcessor.generate_batch_reactions)

    def featurise(batch):
        reactions, tg, sc, feats = batch
        # warm MolGraph cache -> store ids; pinned H2D of ids / row offsets + on-device assembly from the HBM-resident molecule store,
        # enqueued behind the running step
        reactions = np.asarray(reactions, dtype=object)
        r_tok, p_tok = reactions[:, 0].tolist(), reactions[:, 1].tolist()
        if per_group or dedup:      # many segments / de-duplicated reactants: the id-vector path of train_pairwise._window_graphs and eval._forward_chunks
            lens = [wl["group"]] * (len(r_tok) // wl["group"]) if per_group else [len(r_tok)]
            rg, pg = DeviceGraph.from_id_groups(fz.store, fz.parsing_ids(r_tok), fz.parsing_ids(p_tok), lens, dev, dedup)
        else:
            rg, pg = to_dev_pair(fz.parsing_smiles, r_tok, p_tok)
        # extra features and targets go up with the graphs (pinned, asynchronous), not between the loss read and the next launch
        feats_d = torch.as_tensor(np.asarray(feats, dtype=np.float32)).reshape(len(tg), -1).pin_memory().to(dev, non_blocking=True)
        targets_d = torch.FloatTensor(tg).squeeze().pin_memory().to(dev, non_blocking=True)                 # train_listwise.py:187
        return rg, pg, targets_d, sc, feats_d, rg.h2d_bytes + pg.h2d_bytes + 4 * (feats_d.numel() + targets_d.numel())

    from reactranker_b200.data.prefetch import Lookahead
    e2e_feed = Lookahead(endless_plan(), featurise)
    adv_ms = []

    def step_e2e(i):
        rg, pg, targets, sc, feats, graph_bytes = e2e_feed.current
        out = model(rg, pg, gpu=local, add_features=feats)
        loss = loss_fn(out, sc, targets)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        reduce_grads()
        opt.step()
        sched.step()
        h2d[0] = graph_bytes
        t_adv = time.perf_counter()
        e2e_feed.advance()                                                    # plan + featurise + upload batch i+1 while step i executes
        adv_ms.append((time.perf_counter() - t_adv) * 1e3)
        return float(loss.detach().cpu().reshape(-1)[0])                     # D2H read of the step's result

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, profile=False, trace=None):
        for i in range(warmup):
            fn(i)
        barrier()
        if profile:
            _lib.profile_begin()
        _lib.lib().rr_launch_count_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            t0 = time.perf_counter()
            fn(warmup + i)
            if trace is not None:
                trace.append((time.perf_counter() - t0) * 1e3)     # host wall per step (the e2e step ends with a D2H read)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = int(_lib.lib().rr_launch_count())
        prof = _lib.profile_end() if profile else None
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, launches, prof

    # The synthetic pool leaves ~10^6 long-lived Python objects behind; a generation-2 collection over them costs 20+ ms and used to land
    # inside a timed step now and then.  Collect once and freeze the survivors (they stay alive for the whole run anyway).
    import gc
    gc.collect()
    gc.freeze()
    with ClockSampler(local) as clocks:
        ms, launches, _ = timed(step_resident, args.steps, args.warmup)
        # the same K steps again with a CUDA event pair around every launch (rr_profile_begin/end) for the per-kernel roofline: the event
        # records sit between the kernels, which costs a few per cent and switches off programmatic dependent launch, so `value` is
        # taken from the plain pass above and the instrumented pass reports its own step time next to the kernel shares
        ms_prof, _, prof = timed(step_resident, args.steps, 1, profile=True)
    clk = clocks.summary()
    e2e_steps = max(3, args.steps)
    e2e_trace = []
    ms_e2e, _, _ = timed(step_e2e, e2e_steps, max(3, args.warmup), trace=e2e_trace)
    if rank == 0:
        print("e2e host wall per step (ms): " + " ".join(f"{t:.1f}" for t in e2e_trace), file=sys.stderr)
        print("  of which preparing the next batch (ms): " + " ".join(f"{t:.1f}" for t in adv_ms[-len(e2e_trace):]), file=sys.stderr)

    value = rows * world * args.steps / (ms / 1e3)
    e2e_value = rows * world * e2e_steps / (ms_e2e / 1e3)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pk = peaks()
    rg, pg = resident[0][0], resident[0][1]
    work = algorithmic_work(wl, rg, pg)
    kernels = {}
    for cls, (tot_ms, cnt) in prof.items():
        if cnt == 0:
            continue
        ent = {"ms_per_step": tot_ms / args.steps, "launches_per_step": cnt / args.steps, "share_of_step": tot_ms / ms_prof}
        if cls in work:
            w, unit = work[cls]
            rate = w / (tot_ms / args.steps / 1e3)
            if unit == "B":
                ent.update(bound="hbm", achieved=rate / 1e9, peak=pk["hbm"], unit="GB/s", frac=rate / 1e9 / pk["hbm"])
            else:
                ent.update(bound="tensor", achieved=rate / 1e12, peak=pk["tensor"], unit="TFLOP/s", frac=rate / 1e12 / pk["tensor"])
        kernels[cls] = ent
    top = max((c for c in kernels if "frac" in kernels[c]), key=lambda c: kernels[c]["ms_per_step"])
    roof = {k: kernels[top][k] for k in ("bound", "achieved", "peak", "unit", "frac")}
    roof.update(kernel=top, ms_per_step_instrumented=ms_prof / args.steps, traffic=ncu_traffic(top), peak_source=pk["src"], avg_launch_ms=kernels[top]["ms_per_step"] / kernels[top]["launches_per_step"])
    if roof["bound"] == "tensor":
        # fp32-class accuracy costs three MMAs per product.  Forward: kind::tf32, which runs at half the bf16 rate the peak was measured with,
        # so the ceiling of that path is peak / 6; backward (dgrad, wgrad): bf16 operands at the full rate, ceiling peak / 3.  frac (of the bf16
        # peak, as the contract asks) and frac_of_split_ceiling say the same thing twice.
        bwd_bf16 = top in ("gemm_dgrad", "gemm_wgrad") and bool(_lib.lib().rr_get_backward_bf16())
        fwd_bf16 = top == "gemm_fwd" and bool(_lib.lib().rr_get_forward_bf16())
        bf = bwd_bf16 or fwd_bf16
        roof.update(mma_passes_per_product=3, operand_kind="bf16" if bf else "tf32", frac_of_split_ceiling=roof["frac"] * (3.0 if bf else 6.0),
                    achieved_tensor_pipe_tflops=roof["achieved"] * 3.0)
    mp = [c for c in ("bond_fwd", "bond_bwd", "nbr_fwd", "nbr_bwd") if c in kernels]
    mp_bytes = sum(work[c][0] for c in mp)
    mp_ms = sum(kernels[c]["ms_per_step"] for c in mp)
    roof_mp = {"bound": "hbm", "achieved": mp_bytes / (mp_ms / 1e3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
               "frac": mp_bytes / (mp_ms / 1e3) / 1e9 / pk["hbm"], "kernels": mp, "ms_per_step": mp_ms}

    cpu = cpu_baseline(wl, steps=8, warmup=1) if world == 1 and not args.no_cpu else None
    line = {
        "metric": metric_of(wl), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (dense layers on tcgen05: forward 3 x TF32 operand split, backward 3 x bf16 split, fp32 accumulation in TMEM)" if args.gemm == "tc" else "f32",
        "data": f"synthetic reaction graphs (SURVEY.md 8d generator), pool of {len(pool)} distinct batches per rank cycled; random-init weights",
        "config": {"workload": wl["desc"], "name": args.workload, "reactions_per_gpu_per_step": rows, "global_batch": rows * world,
                   "parallelism": f"dp{world}" if world > 1 else "single", "optimizer": "Adam(fused)+NoamLR",
                   "l2": "inputs and activations (>1 GB per step) exceed the 126 MB L2; no explicit flush"},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d[0]), "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "ms_per_step": ms_e2e / e2e_steps},
        "gpu_launches": launches,
        "roofline": roof, "roofline_message_passing": roof_mp, "kernels": kernels,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU legs (the oracle port of the reference path -- the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_training_steps(wl, groups, steps, warmup, seed=99, device="cpu"):
    """Training steps of the oracle port.  ``device="cuda"`` is the informative secondary comparator of SURVEY.md 8(d): the same eager
    PyTorch code with its tensors on the GPU, as the reference runs with ``gpu=0`` (graphs still built on the host every step)."""
    from oracle import reactranker_oracle as O
    from reactranker_b200 import synthetic
    torch.manual_seed(0)
    dev = torch.device(device)
    sd = O.init_state_dict(wl["hidden"], wl["task_num"], 1, True, seed=0, mpnn_depth=wl["depth"], mpnn_diff_depth=wl["diff_depth"])
    sd = {k: v.to(dev) for k, v in sd.items()}

    def put(g):                                         # featurization tensors follow the model (mpn.py:76-77)
        if dev.type != "cpu":
            for name in ("f_atoms", "f_bonds", "a2b", "b2a", "b2revb"):
                setattr(g, name, getattr(g, name).to(dev))
        return g
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if "cached_zero" not in k}
    full = dict(sd)
    full.update(params)
    opt = torch.optim.Adam([{"params": list(params.values()), "lr": 1e-4, "weight_decay": 0}])
    head = O.resolve_task_type(wl["task_num"], wl["last"], wl["task_type"])
    scope = [wl["group"]] * groups
    pool = [synthetic.make_dataset(seed + b, scope) for b in range(2)]
    for ds in pool:                                     # warm MolGraph cache, as the reference's Parsing_features
        for m in ds.mols.values():
            m._mk_lists()
    times, model_times = [], []
    for i in range(warmup + steps):
        ds = pool[i % len(pool)]
        t0 = t1 = time.perf_counter()
        if wl["task"] == "ranknet":
            # train_pairwise.py:81-160: one forward per group, summed pairwise cost / ordered pairs of the window
            cost, pairs, o = 0.0, 0.0, 0
            for n in scope:
                r_g = put(O.OracleBatch([ds.mols[t] for t in ds.rsmi[o:o + n]]))
                p_g = put(O.OracleBatch([ds.mols[t] for t in ds.psmi[o:o + n]]))
                y = O.model_forward(full, r_g, p_g, ds.temp[o:o + n].reshape(-1, 1), mpnn_depth=wl["depth"], mpnn_diff_depth=wl["diff_depth"], head=head,
                                    dropout=wl["dropout"], training=True)
                c, p = O.ranknet_group_cost(y, ds.lgk[o:o + n])
                cost, pairs, o = cost + c, pairs + p, o + n
            loss = cost / pairs
        else:
            r_g = put(O.OracleBatch([ds.mols[t] for t in ds.rsmi]))     # BatchMolGraph build per step (featurization.py:246-290)
            p_g = put(O.OracleBatch([ds.mols[t] for t in ds.psmi]))
            t1 = time.perf_counter()
            out = O.model_forward(full, r_g, p_g, ds.temp.reshape(-1, 1), mpnn_depth=wl["depth"], mpnn_diff_depth=wl["diff_depth"], head=head,
                                  dropout=wl["dropout"], training=True)
            loss = O.loss_for_task(wl["task"], out, scope, torch.tensor(ds.lgk.astype(np.float32)).to(dev))
        opt.zero_grad()
        loss.backward()
        opt.step()
        if dev.type == "cuda":
            torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
        model_times.append(time.perf_counter() - t1)
    cpu_training_steps.model_only = model_times[warmup:]       # forward + loss + backward + Adam without the per-step graph build
    return times[warmup:], sum(scope)


def cpu_baseline(wl, steps, warmup):
    groups = max(2, min(wl["groups"], 1000 // wl["group"]))      # bounded sample: ~1000 reactions per step, 10-20 s of CPU work in all
    times, rows = cpu_training_steps(wl, groups, steps, warmup)
    return {"value": rows * len(times) / sum(times), "unit": UNIT, "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"{len(times)} training steps of {groups} groups x {wl['group']} = {rows} reactions (same model/loss/optimizer; "
                      f"oracle restatement of the reference PyTorch CPU path incl. per-step BatchMolGraph build), {sum(times):.1f} s"}


def reference(args):
    """``--impl reference``: the reference's own CPU implementation of the path (oracle port; the reference is
    Python and cannot travel to the GPU box), all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    groups = max(2, min(wl["groups"], 500 // wl["group"]))
    if args.ref_device == "cuda":          # informative only, never the reference arm: eager PyTorch on the GPU, full workload batch
        times, rows = cpu_training_steps(wl, wl["groups"], args.steps, args.warmup, device="cuda")
        mo = cpu_training_steps.model_only
        print(json.dumps({"impl": "reference-eager-cuda", "metric": metric_of(wl), "unit": UNIT, "value": rows * len(times) / sum(times),
                          "value_model_only": rows * len(mo) / sum(mo), "ms_per_step": 1e3 * sum(times) / len(times),
                          "ms_per_step_model_only": 1e3 * sum(mo) / len(mo), "steps": args.steps, "warmup": args.warmup,
                          "note": "oracle port of the reference's PyTorch path with tensors on cuda:0 (the reference's gpu=0 mode); graphs are "
                                  "built on the host every step as the reference does; model_only excludes that build"}))
        return
    times, rows = cpu_training_steps(wl, groups, args.steps, args.warmup)
    value = rows * len(times) / sum(times)
    cpu = {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": "port",
           "sample": f"each step = {groups} groups x {wl['group']} = {rows} reactions of the workload on the host CPU"}
    print(json.dumps({
        "impl": "reference", "metric": metric_of(wl), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic reaction graphs (SURVEY.md 8d generator); random-init weights",
        "config": {"workload": wl["desc"], "name": args.workload, "reactions_per_step_sample": rows},
        "cpu_baseline": cpu, "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--pool", type=int, default=3, help="distinct synthetic batches per rank")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--ref-device", default="cpu", choices=["cpu", "cuda"],
                    help="with --impl reference: cuda = informative extra line, the same eager PyTorch code with its tensors on the GPU")
    ap.add_argument("--dropout", type=float, default=None, help="override the workload's dropout (the entry scripts' 0.1 / 0.2 by default)")
    ap.add_argument("--no-dedup", action="store_true", help="at dropout 0: still encode one reactant graph per candidate like the reference")
    ap.add_argument("--gemm", default="tc", choices=["tc", "simt"], help="dense layers: tcgen05 3xTF32 (default) or exact-fp32 SIMT")
    a = ap.parse_args()
    if a.impl == "reference":
        reference(a)
    else:
        if a.warmup < 3:
            a.warmup = 3
        ours(a)
