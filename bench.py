#!/usr/bin/env python
"""Benchmark of the ReactRanker training hot path (D-MPNN reaction encoder + LTR loss) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c5|c5-100|c5-200|c5-500|c2|c3|c4]

One "step" = one training step over one batch of synthetic reaction graphs: forward, loss, backward, (gradient all-reduce,) Adam,
NoamLR.  Prints ONE JSON line (rank 0).  The step is the PRODUCT's: ``reactranker_b200.train.step.TrainStep`` (what ``train()`` runs;
RankNet: ``train_pairwise.prepare_window`` / ``run_window``), single-GPU or, under torchrun, data-parallel -- every rank plans the same
GLOBAL batches with ``DataProcessor``, takes its shard of whole reactant groups packed with the global ``max_num_bonds``, divides by the
global normaliser and all-reduces the flat gradient buffer in place (weak scaling: 4100 reactions per GPU per step).

* ``value``  : reactions/s with every rank's shard already resident in HBM (device path), K steps.
* ``e2e``    : reactions/s from HOST buffers through the same public calls -- per step the batch plan
               (``DataProcessor.generate_batch_reactions``), the shard selection, ``Parsing_features.parsing_reactions`` on the warm
               MolGraph cache, the pinned host->device copies (molecule ids, row offsets, extra features, targets), on-device batch
               assembly from the HBM-resident molecule store, forward/loss/backward/all-reduce/Adam, and a device->host read of the loss
               (pinned, asynchronous, consumed one step later so that the GPU queue never drains; the last one before the closing event).
               Every step does all of these inside the timed region; batch i+1 is prepared while step i executes.
* ``sustained``: the device path again for >= 10 s back to back with the clock trace (``value`` itself is a burst of K steps).
* ``roofline``: the dominant kernel class of the step, timed with CUDA events on the launching stream around every launch
               (rr_profile_begin/end of librr_sm100) while the same K steps run a second time; ``roofline_message_passing`` gives the
               gather kernels in both byte accountings (SURVEY.md 8d's literal formulas, and those plus the ReLU/dropout-backward and
               ``d(input) +=`` traffic the fused backward gathers absorbed).
* ``other_workloads``: short runs (device path) of the other BASELINE.json configs in the same process.
* ``cpu_baseline`` / ``--impl reference``: the reference's own modules (oracle/_ref: its byte code compiled by oracle/build_ref.py; the
               oracle port only if that is absent) training on the host CPUs, all threads, on the same workload (bounded per step only
               if a full batch would not finish in minutes); ``eager_pytorch_b200``: the same reference code with ``gpu=0`` (stock PyTorch
               eager kernels on the same B200) -- the only existing GPU implementation of this path.
"""
import argparse
import json
import math
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

_H300 = dict(hidden=300, depth=3, diff_depth=3, task_num=1, task_type=None, last="with_softplus", dropout=0.1)
WORKLOADS = {
    # BASELINE.json configs[4] (the metric's own: D-MPNN + ListMLE, 1/2/4/8 GPUs), sweep over candidates per group
    "c5": dict(_H300, task="mle", group=50, groups=82, desc="ListMLE D-MPNN h300 d3/3/3, 82 groups x 50 candidates = 4100 reactions per GPU per step (configs[4])"),
    "c5-100": dict(_H300, task="mle", group=100, groups=41, desc="ListMLE D-MPNN h300 d3/3/3, 41 groups x 100 candidates = 4100 reactions per GPU per step (configs[4])"),
    "c5-200": dict(_H300, task="mle", group=200, groups=21, desc="ListMLE D-MPNN h300 d3/3/3, 21 groups x 200 candidates = 4200 reactions per GPU per step (configs[4])"),
    "c5-500": dict(_H300, task="mle", group=500, groups=8, desc="ListMLE D-MPNN h300 d3/3/3, 8 groups x 500 candidates = 4000 reactions per GPU per step (configs[4])"),
    "c2": dict(_H300, task="listnet", group=32, groups=128, desc="ListNet@1 D-MPNN h300 d3/3/3, 128 groups x 32 candidates = 4096 reactions per GPU per step (configs[1])"),
    # configs[2]: RankNet (main_ranknet.py: dropout 0.2, no_softplus), 64 candidates/group, 4096 rows per optimiser step = one accumulation
    # window of 64 groups, every group its own segment (own padding rows and max_num_bonds) as in the reference's one-forward-per-group loop
    "c3": dict(_H300, task="ranknet", last="no_softplus", dropout=0.2, group=64, groups=64,
               desc="RankNet (sum_session) D-MPNN h300 d3/3/3, window of 64 groups x 64 candidates = 4096 reactions, 258k ordered pairs (configs[2])"),
    "c4": dict(task="evidential_ranking", hidden=600, depth=5, diff_depth=5, task_num=2, task_type="evidential_ranking", last="with_softplus",
               dropout=0.1, group=32, groups=128, desc="UC-Listwise D-MPNN h600 d5/5/3, 128 groups x 32 = 4096 reactions per GPU per step (configs[3])"),
}
METRIC = "train reactions/sec, D-MPNN+ListMLE"                      # BASELINE.json's metric: the default workload (c5)
LOSS_NAME = {"mle": "ListMLE", "listnet": "ListNet@1", "ranknet": "RankNet", "evidential_ranking": "UC-Listwise"}
UNIT = "reactions/s"
COLS = ["rsmi_mapped", "psmi_mapped"]


def metric_of(wl):
    return METRIC if wl["task"] == "mle" else f"train reactions/sec, D-MPNN+{LOSS_NAME[wl['task']]}"


_TRAFFIC_KERNELS = {"gemm_fwd": "k_tc_gemm2<16, 0", "gemm_dgrad": "k_tc_gemm2<16, 1", "gemm_wgrad": "k_tc_wgrad", "bond_fwd": "k_rowpipe<0", "bond_bwd": "k_rowpipe<1",
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
        return dict(hbm=float(p["hbm_gbs"]), tensor=float(p["bf16_tflops_sustained"]), tensor_burst=float(p.get("bf16_tflops", 0.0)) or None,
                    src="MEASURED_PEAKS.json (measured; bf16 sustained)")
    except Exception:
        return dict(hbm=6650.0, tensor=1400.0, tensor_burst=None, src="fallback of B200_PROFILING.md")


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw"
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

    @staticmethod
    def _num(x):
        try:
            return float(x)
        except Exception:
            return None

    def summary(self):
        rows = [r for r in self.rows if len(r) >= 6]
        sm = [v for v in (self._num(r[0]) for r in rows) if v is not None]
        mx = [v for v in (self._num(r[1]) for r in rows) if v is not None]
        pw = [v for v in (self._num(r[6]) for r in rows if len(r) >= 7) if v is not None]
        reasons = sorted({n for r in rows for n, v in zip(self.NAMES, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "sm_mhz_min": min(sm) if sm else None, "power_w_median": statistics.median(pw) if pw else None}


# ------------------------------------------------------------------------------------------------
# algorithmic work
# ------------------------------------------------------------------------------------------------
def algorithmic_work(wl, rg, pg, n_add=1):
    """Per-step algorithmic work of each kernel class (forward + backward): compulsory fp32 activation bytes + int32 indices for the
    gather kernels, 2MNK for the GEMMs with the reference's logical dimensions (83/61/h, not the padded strides).  Bytes come in two
    accountings: ``survey`` = SURVEY.md 8(d)'s literal formulas (a backward gather moves what its forward moves), ``fused`` = those plus
    the mask read and the ``d(input)`` read/write of the ReLU / dropout backward the fused backward gathers carry (rr_mp_pipe.cu)."""
    h, T, Td = wl["hidden"], wl["depth"] - 1, wl["diff_depth"] - 1
    s = i = 4
    N = pg.n_mols
    out = {}
    bond = nbr_f = nbr_b = nbr_b_s = 0.0
    g_f = g_d = 0.0
    bond_b = 0.0
    for g in (rg, pg):
        A, B, W = g.n_atoms, g.n_bonds, g.c.wmax
        bond += T * (2 * B * h * s + (A * W + 2 * B) * i)
        # backward gathers carry the ReLU/dropout backward that follows them: besides the gather's own read + write, the mask source y is
        # read and the running sum d(input) is read + written; the last one of a graph writes only that sum
        bond_b += max(T - 1, 0) * (5 * B * h * s + (A * W + 2 * B) * i) + (1 if T >= 1 else 0) * (4 * B * h * s + (A * W + 2 * B) * i)
        agg = (B + A) * h * s + A * W * i
        nbr_f += agg
        nbr_b += (A + 3 * B) * h * s + A * W * i            # dout[A] -> dz[B] masked by m^T (or [inp > 0]), d(input) = dz
        nbr_b_s += agg
        g_f += 2.0 * B * 83 * h + T * 2.0 * B * h * h + 2.0 * A * (61 + h) * h
        g_d += T * 2.0 * B * h * h + 2.0 * A * h * h
    A, B, W = pg.n_atoms, pg.n_bonds, pg.c.wmax
    a2a = 2 * A * h * s + A * W * i
    nbr_f += (Td + 1) * a2a + ((B + A) * 83 * s + A * W * i if Td > 0 else 0)
    nbr_b += 2 * (4 * A * h * s + A * W * i) + max(Td - 1, 0) * (5 * A * h * s + A * W * i) if Td > 0 else (4 * A * h * s + A * W * i)
    nbr_b_s += (Td + 1) * a2a
    g_f += 2.0 * A * h * h + Td * 2.0 * A * (h + 83) * h + 2.0 * A * 2 * h * h
    g_d += 2.0 * A * h * h + Td * 2.0 * A * h * h + 2.0 * A * 2 * h * h
    ffn = 2.0 * N * ((h + n_add) * h + h * h + h * wl["task_num"])
    g_f += ffn
    g_d += ffn
    out["bond_fwd"] = (bond, "B", bond)
    out["bond_bwd"] = (bond_b, "B", bond)                   # survey: the backward of an iteration moves what its forward moves
    out["nbr_fwd"] = (nbr_f, "B", nbr_f)
    out["nbr_bwd"] = (nbr_b, "B", nbr_b_s)
    out["gemm_fwd"] = (g_f, "FLOP", g_f)
    out["gemm_dgrad"] = (g_d, "FLOP", g_d)
    out["gemm_wgrad"] = (g_f, "FLOP", g_f)
    return out


# ------------------------------------------------------------------------------------------------
# one workload on this rank
# ------------------------------------------------------------------------------------------------
class Job:
    def __init__(self, name, local, dropout=None, no_dedup=False, pool=3, seed=1):
        from reactranker_b200 import parallel, synthetic
        from reactranker_b200.data.load_reactions import DataProcessor, Parsing_features
        from reactranker_b200.models.base_model import build_model
        from reactranker_b200.train.step import TrainStep
        from reactranker_b200.train.utils import build_lr_scheduler, build_optimizer
        import pandas as pd
        self.name, self.local = name, local
        self.dev = torch.device("cuda", local)
        wl = dict(WORKLOADS[name])
        if dropout is not None:
            wl["dropout"] = dropout
            wl["desc"] += f", dropout {dropout:g}"
        self.wl = wl
        self.rows = wl["group"] * wl["groups"]                    # reactions per GPU per step
        torch.manual_seed(0)
        model = build_model(hidden_size=wl["hidden"], mpnn_depth=wl["depth"], mpnn_diff_depth=wl["diff_depth"], ffn_depth=3, use_bias=True,
                            dropout=wl["dropout"], task_num=wl["task_num"], ffn_last_layer=wl["last"], task_type=wl["task_type"], add_features_dim=1)
        self.model = model.cuda(local).train()
        if no_dedup:
            self.model.dedup_reactants = False
        self.dedup = wl["dropout"] == 0 and not no_dedup          # exact only without dropout (rr_model_cfg.r_atom_map)
        if self.dedup:
            wl["desc"] += ", repeated reactants encoded once"
        self.opt = build_optimizer(self.model)
        self.sched = build_lr_scheduler(self.opt, warmup_epochs=2, total_epochs=30, train_data_size=100 * self.rows, batch_size=self.rows,
                                        init_lr=1e-4, max_lr=1e-3, final_lr=1e-4)
        self.ranknet = wl["task"] == "ranknet"
        # joins torchrun's process group, broadcasts rank 0's weights, owns the gradient all-reduce
        self.step = TrainStep(self.model, self.opt, self.sched, "mle" if self.ranknet else wl["task"], local)
        self.rank, self.world = self.step.rank, self.step.world
        if self.world > 1:
            torch.manual_seed(1000003 * self.rank)                # own dropout masks per rank, as train() sets them
        # the synthetic data set: `pool` global batches; every rank generates the same frame (the plan is global) and uploads the molecules
        # once into the HBM-resident molecule store
        self.fz = Parsing_features()
        frames = []
        for b in range(pool * self.world):
            ds = synthetic.make_dataset(seed * 1000 + b, [wl["group"]] * wl["groups"])
            for tok, m in ds.mols.items():
                self.fz.add(tok, m)
            frames.append(ds.to_dataframe().assign(flag=lambda d, b=b: d.flag + b * wl["groups"]))
        self.planner = DataProcessor(pd.concat(frames, ignore_index=True))
        self.global_rows = self.rows * self.world
        self._plan_it, self._epoch = None, 0
        self.resident = [self._prepare(self._next_plan()) for _ in range(pool)]
        torch.cuda.synchronize()
        self.h2d = 0
        self.adv_ms = []
        self._feed = None
        _ = parallel

    # ---- the global batch plan: DataProcessor, as train() / factorized_training_loop drive it ----
    def _next_plan(self):
        from reactranker_b200.train.train_pairwise import iter_windows
        while True:
            if self._plan_it is None:
                if self.ranknet:
                    self._plan_it = iter_windows(self.planner, self._epoch, self.global_rows, COLS, "lgk", "temp")
                else:      # train() iterates exactly this: generate_batch_reactions at the planner level (rows + scope), seed = epoch
                    self._plan_it = self.planner.plan_batch_reactions(batch_size=self.global_rows, seed=self._epoch)
                self._epoch += 1
            for b in self._plan_it:
                if self.ranknet:
                    if not b[2]:                                   # full windows only (the tail flush is smaller)
                        return b
                elif sum(b[1]) == self.global_rows:
                    return b
            self._plan_it = None

    def _endless(self):
        while True:
            yield self._next_plan()

    def _prepare(self, plan):
        if self.ranknet:
            from reactranker_b200.train.train_pairwise import prepare_window
            window, pairs, _ = plan
            prepared, sync = prepare_window(self.model, window, self.fz, self.local)
            h2d = 0 if prepared is None else prepared[0].h2d_bytes + prepared[1].h2d_bytes + 8 * len(prepared[2])
            return ("ranknet", prepared, sync, pairs, h2d)
        rows, scope = plan
        return self.step.prepare_rows(self.planner, rows, scope, self.fz, COLS, "lgk", "temp")

    def _run(self, prepared):
        if self.ranknet:
            from reactranker_b200.train.train_pairwise import run_window
            _, prep, sync, pairs, _ = prepared
            loss = run_window(self.model, prep, sync, pairs, self.opt, self.local, 1.0, "sum_session")
            self.sched.step()
            return loss
        return self.step.run(prepared)

    def step_resident(self, i):
        return self._run(self.resident[i % len(self.resident)])

    def step_e2e(self, i):
        """One end-to-end step.  Every step's loss IS read back to the host inside the timed region, through a pinned buffer and one step
        late: the copy of step i's loss is enqueued behind step i, batch i+1 is prepared, and only then the host waits -- for step i-1's
        value, which has long arrived.  The GPU queue never drains, which is how a training loop that logs its loss is written."""
        from reactranker_b200.data.prefetch import Lookahead
        if self._feed is None:
            self._feed = Lookahead(self._endless(), self._prepare)
            self._pin = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._pending = None
            self.losses = []
        cur = self._feed.current
        loss = self._run(cur)
        buf = self._pin[i & 1]
        buf.copy_(loss.detach().reshape(-1)[:1], non_blocking=True)       # D2H of this step's result, asynchronous
        ev = torch.cuda.Event()
        ev.record()
        self.h2d = cur[4] if self.ranknet else cur.h2d_bytes
        t = time.perf_counter()
        self._feed.advance()                                          # plan + shard + ids + upload batch i+1 while step i executes
        self.adv_ms.append((time.perf_counter() - t) * 1e3)
        self.e2e_finish()                                             # the previous step's loss
        self._pending = (ev, buf)

    def e2e_finish(self):
        if getattr(self, "_pending", None) is not None:
            ev, buf = self._pending
            ev.synchronize()
            self.losses.append(float(buf[0]))
            self._pending = None

    def graphs(self):
        r = self.resident[0]
        return (r[1][0], r[1][1]) if self.ranknet else (r.r, r.p)

    def close(self):
        self.resident, self._feed, self.model, self.opt, self.step = None, None, None, None, None
        import gc
        gc.collect()
        torch.cuda.empty_cache()


def barrier(dist):
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, steps, warmup, dist, dev, profile=False, trace=None, finish=None):
    """W untimed steps, then K steps bracketed by barrier + synchronize, CUDA events on the launching stream, MAX over ranks."""
    from reactranker_b200 import _lib
    for i in range(warmup):
        fn(i)
    barrier(dist)
    if profile:
        _lib.profile_begin()
    _lib.lib().rr_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        t0 = time.perf_counter()
        fn(warmup + i)
        if trace is not None:
            trace.append((time.perf_counter() - t0) * 1e3)     # host wall per step
    if finish is not None:
        finish()                                               # e.g. the last step's device->host read: inside the timed region
    e1.record()
    barrier(dist)
    ms = e0.elapsed_time(e1)
    launches = int(_lib.lib().rr_launch_count())
    prof = _lib.profile_end() if profile else None
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms, launches, prof


def ours(args):
    from reactranker_b200 import _lib, parallel
    rank, world, local = parallel.world_from_env()
    if world > 1:
        args.gpus = world
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")       # stdout carries exactly one JSON line
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.require_device(local)
    _lib.check(_lib.lib().rr_set_gemm_mode(1 if args.gemm == "tc" else 0))
    pool = args.pool if world <= 2 else max(2, args.pool - 1)
    job = Job(args.workload, local, args.dropout, args.no_dedup, pool)
    dist = torch.distributed if world > 1 else None
    wl, rows = job.wl, job.rows

    # The synthetic pool leaves ~10^6 long-lived Python objects behind; a generation-2 collection over them costs 20+ ms and used to land
    # inside a timed step now and then.  Collect once and freeze the survivors (they stay alive for the whole run anyway).
    import gc
    gc.collect()
    gc.freeze()
    with ClockSampler(local) as clocks:
        ms, launches, _ = timed(job.step_resident, args.steps, args.warmup, dist, dev)
        # the same K steps again with a CUDA event pair around every launch (rr_profile_begin/end) for the per-kernel roofline: the event
        # records sit between the kernels, which costs a few per cent and switches off programmatic dependent launch, so `value` is
        # taken from the plain pass above and the instrumented pass reports its own step time next to the kernel shares
        ms_prof, _, prof = timed(job.step_resident, args.steps, 1, dist, dev, profile=True)
    clk = clocks.summary()
    e2e_steps = max(3, args.steps)
    e2e_trace = []
    ms_e2e, _, _ = timed(job.step_e2e, e2e_steps, max(3, args.warmup), dist, dev, trace=e2e_trace, finish=job.e2e_finish)
    if rank == 0:
        print("e2e host wall per step (ms): " + " ".join(f"{t:.1f}" for t in e2e_trace), file=sys.stderr)
        print("  of which preparing the next batch (ms): " + " ".join(f"{t:.1f}" for t in job.adv_ms[-len(e2e_trace):]), file=sys.stderr)
    value = rows * world * args.steps / (ms / 1e3)
    e2e_value = rows * world * e2e_steps / (ms_e2e / 1e3)

    # sustained: the device path back to back for >= args.sustain seconds, three consecutive thirds timed separately (">= 10 s,
    # median of 3"), clocks sampled throughout
    sustained = None
    if args.sustain > 0:
        n3 = max(args.steps, int(math.ceil(args.sustain * 1e3 / (ms / args.steps) / 3)))
        with ClockSampler(local) as sclk:
            thirds = [timed(job.step_resident, n3, 0, dist, dev)[0] for _ in range(3)]
        sc = sclk.summary()
        vals = sorted(rows * world * n3 / (t / 1e3) for t in thirds)
        sustained = {"value": vals[1], "unit": UNIT, "runs": vals, "steps": 3 * n3, "seconds": sum(thirds) / 1e3, "ms_per_step": statistics.median(thirds) / n3,
                     "clocks": sc, "note": "median of three consecutive back-to-back runs of the device path"}

    sync_stats = None
    if job.step.sync is not None:
        sync_stats = {"inplace_flat_allreduce_steps": job.step.sync.fast_path_steps, "staged_copy_steps": job.step.sync.copy_path_steps}

    pk = peaks()
    rg, pg = job.graphs()
    work = algorithmic_work(wl, rg, pg)
    n_steps_prof = args.steps
    job_name = job.name
    job.close()

    # the other BASELINE.json configs, device path, short runs (every rank takes part; rank 0 reports)
    others = {}
    if not args.no_others and args.workload == "c5":
        names = ["c2", "c3", "c4", "c5-100", "c5-200", "c5-500"] if world == 1 else ["c4", "c5-500"]
        for name in names:
            try:
                j = Job(name, local, args.dropout, args.no_dedup, pool=2 if world == 1 else 1, seed=7)
                o_ms, _, _ = timed(j.step_resident, args.other_steps, 3, dist, dev)
                e_ms, _, _ = timed(j.step_e2e, args.other_steps, 3, dist, dev, finish=j.e2e_finish)
                others[name] = {"metric": metric_of(j.wl), "value": j.rows * world * args.other_steps / (o_ms / 1e3), "unit": UNIT,
                                "e2e": j.rows * world * args.other_steps / (e_ms / 1e3), "ms_per_step": o_ms / args.other_steps,
                                "steps": args.other_steps, "warmup": 3, "n_gpus": world, "workload": j.wl["desc"]}
                j.close()
            except Exception as exc:          # a side measurement must not lose the headline
                others[name] = {"error": repr(exc)[:300]}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    kernels = {}
    for cls, (tot_ms, cnt) in prof.items():
        if cnt == 0:
            continue
        ent = {"ms_per_step": tot_ms / n_steps_prof, "launches_per_step": cnt / n_steps_prof, "share_of_step": tot_ms / ms_prof}
        if cls in work:
            w, unit, w_survey = work[cls]
            rate = w / (tot_ms / n_steps_prof / 1e3)
            if unit == "B":
                ent.update(bound="hbm", achieved=rate / 1e9, peak=pk["hbm"], unit="GB/s", frac=rate / 1e9 / pk["hbm"],
                           frac_survey_8d=w_survey / (tot_ms / n_steps_prof / 1e3) / 1e9 / pk["hbm"])
            else:
                ent.update(bound="tensor", achieved=rate / 1e12, peak=pk["tensor"], unit="TFLOP/s", frac=rate / 1e12 / pk["tensor"])
        kernels[cls] = ent
    top = max((c for c in kernels if "frac" in kernels[c]), key=lambda c: kernels[c]["ms_per_step"])
    roof = {k: kernels[top][k] for k in ("bound", "achieved", "peak", "unit", "frac")}
    roof.update(kernel=top, ms_per_step_instrumented=ms_prof / n_steps_prof, traffic=ncu_traffic(top), peak_source=pk["src"],
                avg_launch_ms=kernels[top]["ms_per_step"] / kernels[top]["launches_per_step"])
    if roof["bound"] == "tensor":
        # fp32-class accuracy costs three MMAs per product.  Forward: kind::tf32, which runs at half the bf16 rate the peak was measured with,
        # so the ceiling of that path is peak / 6; backward (dgrad, wgrad): bf16 operands at the full rate, ceiling peak / 3.  frac (of the bf16
        # peak, as the contract asks) and frac_of_split_ceiling say the same thing twice.
        bwd_bf16 = top in ("gemm_dgrad", "gemm_wgrad") and bool(_lib.lib().rr_get_backward_bf16())
        fwd_bf16 = top == "gemm_fwd" and bool(_lib.lib().rr_get_forward_bf16())
        bf = bwd_bf16 or fwd_bf16
        roof.update(mma_passes_per_product=3, operand_kind="bf16" if bf else "tf32", frac_of_split_ceiling=roof["frac"] * (3.0 if bf else 6.0),
                    achieved_tensor_pipe_tflops=roof["achieved"] * 3.0, frac_of_burst_peak=(roof["achieved"] / pk["tensor_burst"]) if pk["tensor_burst"] else None)
    mp = [c for c in ("bond_fwd", "bond_bwd", "nbr_fwd", "nbr_bwd") if c in kernels]
    mp_bytes = sum(work[c][0] for c in mp)
    mp_bytes_survey = sum(work[c][2] for c in mp)
    mp_ms = sum(kernels[c]["ms_per_step"] for c in mp)
    roof_mp = {"bound": "hbm", "achieved": mp_bytes / (mp_ms / 1e3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
               "frac": mp_bytes / (mp_ms / 1e3) / 1e9 / pk["hbm"], "frac_fused_bytes": mp_bytes / (mp_ms / 1e3) / 1e9 / pk["hbm"],
               "frac_survey_8d": mp_bytes_survey / (mp_ms / 1e3) / 1e9 / pk["hbm"], "bytes_per_step_fused": mp_bytes, "bytes_per_step_survey_8d": mp_bytes_survey,
               "kernels": mp, "ms_per_step": mp_ms,
               "note": "frac_survey_8d charges every backward gather what its forward moves (SURVEY.md 8d); frac_fused_bytes adds the mask read and the "
                       "d(input) read/write of the ReLU/dropout backward those kernels absorbed"}

    cpu = eager = None
    if world == 1 and not args.no_cpu:
        cpu = cpu_baseline(wl, args.cpu_seconds)
        try:
            eager = eager_cuda(wl, steps=3, warmup=1)
        except Exception as exc:
            eager = {"error": repr(exc)[:300]}
    line = {
        "metric": metric_of(wl), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (dense layers on tcgen05: forward 3 x TF32 operand split, backward 3 x bf16 split, fp32 accumulation in TMEM)" if args.gemm == "tc" else "f32",
        "data": f"synthetic reaction graphs (SURVEY.md 8d generator): one data set of {pool} global batches planned by DataProcessor, "
                "every rank trains on its shard; random-init weights",
        "config": {"workload": wl["desc"], "name": job_name, "reactions_per_gpu_per_step": rows, "global_batch": rows * world,
                   "parallelism": (f"dp{world}: global batch plan, whole groups sharded by atom count, global max_num_bonds and normalisers, "
                                   "in-place all-reduce of the flat gradient buffer (train/step.py)") if world > 1 else "single",
                   "optimizer": "Adam(fused)+NoamLR", "step": "reactranker_b200.train.step.TrainStep (the step train() runs)",
                   "l2": "inputs and activations (>1 GB per step) exceed the 126 MB L2; no explicit flush"},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(job.h2d), "d2h_bytes_per_step": 4, "steps": e2e_steps,
                "ms_per_step": ms_e2e / e2e_steps},
        "gpu_launches": launches,
        "roofline": roof, "roofline_message_passing": roof_mp, "kernels": kernels,
    }
    if sustained is not None:
        line["sustained"] = sustained
    if sync_stats is not None:
        line["gradient_sync"] = sync_stats
    if others:
        line["other_workloads"] = others
    if cpu is not None:
        line["cpu_baseline"] = cpu
    if eager is not None:
        line["eager_pytorch_b200"] = eager
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# reference legs: the reference's own modules (oracle/_ref or /root/reference via oracle/ref_loader), else the oracle port.
# The only place bench.py executes anything under oracle/.
# ------------------------------------------------------------------------------------------------
def _ref_loss(L, task, out, scope, targets, gpu):
    """Loss dispatch of the reference's step body (train_listwise.py:196-285) for the benchmarked keys, on the reference's own loss modules."""
    if task == "mle":
        return L.MLEloss()(out, scope, targets, gpu)
    if task == "listnet":
        return L.ListnetLoss()(out, scope, targets, gpu)
    if task == "evidential_ranking":
        return L.evidential_ranking()(out, scope, targets, 0.0001, 0, 1, gpu)
    raise ValueError(task)


def reference_training_steps(wl, groups, steps, warmup, device="cpu", seed=99, budget_s=None):
    """Training steps of the reference path on ``device``.  Listwise keys with oracle/_ref (or /root/reference) present: the REFERENCE's
    own build_model / BatchMolGraph / loss modules / Adam, fed like its train() (warm MolGraph cache, BatchMolGraph built every step,
    the per-step NaN check's host copy) -> kind "reference".  Otherwise (RankNet's per-group loop, or no reference at hand): the oracle
    port -> kind "port".  Returns (step times, model-only times, reactions per step, kind)."""
    from oracle import ref_loader as R
    from reactranker_b200 import synthetic
    scope = [wl["group"]] * groups
    pool = [synthetic.make_dataset(seed + b, scope) for b in range(2)]
    dev = torch.device(device)
    gpu = None if dev.type == "cpu" else (dev.index or 0)
    times, model_times = [], []
    t_start = time.perf_counter()
    if R.available() and wl["task"] != "ranknet":
        bm, ut, LS = R.ref("models.base_model"), R.ref("train.utils"), R.ref("train.loss")
        torch.manual_seed(0)
        model = bm.build_model(hidden_size=wl["hidden"], mpnn_depth=wl["depth"], mpnn_diff_depth=wl["diff_depth"], ffn_depth=3, use_bias=True,
                               dropout=wl["dropout"], task_num=wl["task_num"], ffn_last_layer=wl["last"], task_type=wl["task_type"], add_features_dim=1)
        if gpu is not None:
            model = model.cuda(gpu)
        model.train()
        opt = ut.build_optimizer(model)
        fzs = [R.RefFeaturizer(ds.mols) for ds in pool]
        for fz, ds in zip(fzs, pool):                       # warm MolGraph cache, as the reference's Parsing_features after its first epoch
            fz.parsing_smiles(list(ds.mols.keys()))
        for i in range(warmup + steps):
            ds, fz = pool[i % 2], fzs[i % 2]
            reactions = np.stack([ds.rsmi, ds.psmi], 1)
            t0 = time.perf_counter()
            targets = torch.FloatTensor(ds.lgk.reshape(-1, 1)).squeeze()                  # train_listwise.py:187
            r_in, p_in = fz.parsing_reactions(reactions)                                   # 188: BatchMolGraph build per step
            t1 = time.perf_counter()
            out = model(r_in, p_in, gpu=gpu, add_features=ds.temp.reshape(-1, 1))
            np.any(np.isnan(model.state_dict()['encoder.W_i.weight'].cpu().tolist()))      # 190: the per-step NaN check and its host copy
            loss = _ref_loss(LS, wl["task"], out, scope, targets, gpu)
            opt.zero_grad()
            loss.backward()
            opt.step()
            if gpu is not None:
                torch.cuda.synchronize()
            t2 = time.perf_counter()
            times.append(t2 - t0)
            model_times.append(t2 - t1)
            if budget_s is not None and i >= warmup and time.perf_counter() - t_start > budget_s:
                break
        return times[warmup:], model_times[warmup:], sum(scope), "reference"
    # oracle port
    from oracle import reactranker_oracle as O
    torch.manual_seed(0)
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
    for ds in pool:                                     # warm MolGraph cache, as the reference's Parsing_features
        for m in ds.mols.values():
            m._mk_lists()
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
        if budget_s is not None and i >= warmup and time.perf_counter() - t_start > budget_s:
            break
    return times[warmup:], model_times[warmup:], sum(scope), "port"


def _all_threads():
    """The reference arm uses every host thread it can: torchrun exports OMP_NUM_THREADS=1 for its workers, which would time a
    single-threaded CPU baseline."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def pick_groups(wl, per_step_s):
    """Largest number of whole groups per step (<= the workload's) whose training step on the host CPUs stays within ``per_step_s``.
    The reference's CPU cost per reaction GROWS with the batch (4100 reactions: 134 s per step = 31 reactions/s in the build container,
    against ~350 reactions/s at 400), so the size is found by doubling probes, not extrapolated from one small step; a bounded sample
    therefore flatters the reference."""
    g, best = min(4, wl["groups"]), min(4, wl["groups"])
    while True:
        t, _, _, _ = reference_training_steps(wl, g, steps=1, warmup=0)
        if t[0] > per_step_s:
            return best if best < g else g
        best = g
        if g >= wl["groups"]:
            return wl["groups"]
        if t[0] > per_step_s / 2.5:                      # the next doubling would overshoot (cost is superlinear)
            return g
        g = min(wl["groups"], g * 2)


def cpu_baseline(wl, seconds):
    """The reference on the host CPUs, all threads, on a bounded sample of the workload: whole groups of the workload's size, as many per
    step as keep one step within a third of ``seconds``; >= 2 timed steps after one warm-up."""
    cores = _all_threads()
    groups = pick_groups(wl, seconds / 3.0)
    times, model_times, rows, kind = reference_training_steps(wl, groups, steps=50, warmup=1, budget_s=seconds)
    return {"value": rows * len(times) / sum(times), "unit": UNIT, "cores": cores, "host_cpus": os.cpu_count(), "kind": kind,
            "value_model_only": rows * len(model_times) / sum(model_times), "same_config": groups == wl["groups"],
            "sample": f"{len(times)} training steps of {groups} groups x {wl['group']} = {rows} reactions (same model / loss / optimizer / group size; the "
                      f"reference's build_model / BatchMolGraph / loss modules / Adam on the host CPUs incl. its per-step BatchMolGraph build), {sum(times):.1f} s; "
                      "the reference's cost per reaction grows with the batch, so a bounded sample flatters it"}


def eager_cuda(wl, steps, warmup):
    """SURVEY.md 8(d)'s secondary comparator: the reference's own code with gpu=0 -- stock PyTorch eager kernels on the same B200, full batch;
    graphs are built on the host every step as the reference does (model_only excludes that build)."""
    times, mo, rows, kind = reference_training_steps(wl, wl["groups"], steps, warmup, device="cuda:0")
    return {"value": rows * len(times) / sum(times), "value_model_only": rows * len(mo) / sum(mo), "unit": UNIT, "ms_per_step": 1e3 * sum(times) / len(times),
            "ms_per_step_model_only": 1e3 * sum(mo) / len(mo), "steps": len(times), "kind": kind,
            "note": "the reference's PyTorch path with its tensors on cuda:0 (its gpu=0 mode): the only existing GPU implementation of this path"}


def reference(args):
    """``--impl reference``: the reference's own CPU implementation of the path on the host cores (all threads), same workload; under
    torchrun rank 0 alone runs and prints."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    if args.ref_device == "cuda":          # informative only, never the reference arm
        print(json.dumps(dict(eager_cuda(wl, args.steps, args.warmup), impl="reference-eager-cuda", metric=metric_of(wl))))
        return
    cores = _all_threads()
    # each step = a bounded sample of the workload: as many whole groups as keep the whole --steps/--warmup run within --ref-budget seconds
    groups = pick_groups(wl, args.ref_budget / (args.steps + args.warmup))
    est = float("nan")
    times, mo, rows, kind = reference_training_steps(wl, groups, args.steps, args.warmup)
    value = rows * len(times) / sum(times)
    full = groups == wl["groups"]
    cpu = {"value": value, "unit": UNIT, "cores": cores, "host_cpus": os.cpu_count(), "kind": kind, "same_config": full,
           "value_model_only": rows * len(mo) / sum(mo),
           "sample": (f"each step = the full batch, {groups} groups x {wl['group']} = {rows} reactions" if full else
                      f"each step = {groups} groups x {wl['group']} = {rows} reactions of the workload, sized so that {args.steps + args.warmup} steps fit "
                      f"{args.ref_budget:.0f} s (a full 4100-reaction step of the reference takes minutes on a CPU)") + " on the host CPUs, all threads"}
    print(json.dumps({
        "impl": "reference", "metric": metric_of(wl), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
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
    ap.add_argument("--pool", type=int, default=3, help="distinct synthetic global batches")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / eager_pytorch_b200 legs")
    ap.add_argument("--no-others", action="store_true", help="skip the short runs of the other BASELINE.json configs")
    ap.add_argument("--other-steps", type=int, default=8)
    ap.add_argument("--sustain", type=float, default=10.0, help="seconds of back-to-back steps for the `sustained` block (0 = off)")
    ap.add_argument("--cpu-seconds", type=float, default=25.0, help="time budget of the cpu_baseline leg")
    ap.add_argument("--ref-budget", type=float, default=200.0, help="--impl reference: seconds the whole run may take before the batch is bounded")
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
