"""Host-side batch planning and graph assembly with the reference's names and yields
(data/load_reactions.py:13-195 ``get_data``, 198-537 ``DataProcessor``, 540-586 ``Parsing_features``).

The reference rescans the whole DataFrame once per reactant per epoch (``df[df.rsmi == r]``,
load_reactions.py:368) and grows its outputs with ``np.vstack``.  Here the rows of every reactant
are indexed once and each batch is one fancy-index; the random streams are reproduced exactly:

* ``sklearn.utils.shuffle(x, random_state=s)``  ==  ``x[idx]`` with ``idx = arange(n)``;
  ``RandomState(s).shuffle(idx)``
* ``DataFrame.sample(frac=1, random_state=s)``   ==  rows in ``RandomState(s).permutation(n)``
* ``DataFrame.sample(n=k, random_state=s)``      ==  the first ``k`` of that permutation

(checked against plans recorded from the reference, tests/golden/planner.npz).
"""
from __future__ import annotations

import datetime
from typing import Dict, List, Optional, Sequence

import numpy as np
import pandas as pd

from ..features.featurization import BatchMolGraph, MolGraph, MoleculeStore, _pack_of


def get_time():
    return datetime.datetime.now().strftime('%Y-%m-%d %H:%M:%S')


def _shuffled(values: np.ndarray, seed) -> np.ndarray:
    idx = np.arange(len(values))
    np.random.RandomState(seed).shuffle(idx)
    return values[idx]


def _group_rows(keys: np.ndarray):
    """unique keys in order of first appearance (``Series.unique``) and, per key, its row
    positions in frame order."""
    codes, uniques = pd.factorize(keys, sort=False)
    order = np.argsort(codes, kind="stable")
    counts = np.bincount(codes, minlength=len(uniques))
    bounds = np.concatenate(([0], np.cumsum(counts)))
    return uniques, [order[bounds[i]:bounds[i + 1]] for i in range(len(uniques))]


class get_data:
    """CSV loader / splitter (load_reactions.py:13-195)."""

    def __init__(self, path):
        self.path = path
        self.num_reactions = None
        self.num_reactants = None
        self.df = None

    def get_num(self):
        self.num_reactants = len(self.df.rsmi.unique())
        self.num_reactions = len(self.df)
        print('reaction number is: ', self.num_reactions)
        print('reactant number is: ', self.num_reactants)
        return self.num_reactions, self.num_reactants

    def read_data(self, sep=','):
        print(get_time(), "load file from {}".format(self.path))
        df = pd.read_csv(self.path, sep=sep)
        for c in df.columns:                      # pandas >= 3 Arrow strings break sklearn.shuffle on .unique()
            if not (pd.api.types.is_numeric_dtype(df[c]) or pd.api.types.is_bool_dtype(df[c])):
                df[c] = df[c].astype(object)
        self.df = df
        print(get_time(), "finish loading from {}".format(self.path))

    def filter_bacth(self, filter_szie: int = 3):
        """Drop every reactant with fewer than ``filter_szie`` candidates (load_reactions.py:42-57)."""
        sizes = self.df.groupby('rsmi', sort=False)['rsmi'].transform('size')
        self.df = self.df[sizes >= filter_szie]

    def get_all_data(self):
        return self.df

    @staticmethod
    def shuffle_data(df, seed: int = 0):
        return df.sample(frac=1, random_state=seed)

    def _split_by(self, column: str, split_size, seed):
        keys, rows = _group_rows(self.df[column].values)
        order = _shuffled(np.arange(len(keys)), seed)
        n = len(order)
        i1, i2 = int(n * split_size[1]), int(n * (split_size[2] + split_size[1]))
        take = lambda ids: self.df.iloc[np.concatenate([rows[i] for i in ids])].reset_index(drop=True)  # noqa: E731
        return take(order[i2:]), take(order[:i1]), take(order[i1:i2])       # train, val, test

    def split_data(self, df=None, split_size=(0.8, 0.1, 0.1), split_type='reactants', seed: int = 0):
        """train/val/test by reactant, by ``flag`` column, or by row (load_reactions.py:101-168)."""
        if split_type == 'reactions':
            data = (self.df if df is None else df).sample(frac=1, random_state=seed)
            rows = data.shape[0]
            i1, i2 = int(rows * split_size[1]), int(rows * (split_size[2] + split_size[1]))
            return (data.iloc[i2:rows].reset_index(drop=True), data.iloc[i1:i2].reset_index(drop=True),
                    data.iloc[0:i1].reset_index(drop=True))
        if split_type == 'reactants':
            return self._split_by('rsmi', split_size, seed)
        if split_type == 'flag':
            return self._split_by('flag', split_size, seed)
        raise Exception('Split strategy is unknown')

    def scaffold_split_data(self, *a, **k):
        raise NotImplementedError("scaffold splitting needs RDKit and is outside the hot path (SURVEY.md §2 row 15)")


class DataProcessor:
    """Batch planners (load_reactions.py:198-537) over a once-built row index."""

    def __init__(self, df, num_properties=2):
        self.df = df
        self.num_pairs = None
        self.num_reactants = len(df.rsmi.unique())
        self.num_properties = num_properties
        self.smiles2graph = {}
        self._index_for = None

    def get_num_reactants(self):
        return self.num_reactants

    def _index(self, df):
        if self._index_for is not df or self._index_len != len(df):
            self._reactants, self._rows = _group_rows(df.rsmi.values)
            self._index_for, self._index_len = df, len(df)
            self._cols: Dict[str, np.ndarray] = {}
        return self._reactants, self._rows

    def _col(self, df, name) -> np.ndarray:
        c = self._cols.get(name)
        if c is None:
            c = self._cols[name] = df[name].values
        return c

    def _gather(self, df, rows, smiles_list, target_name, add_features_name):
        cols = ['rsmi', 'psmi'] if smiles_list is None else list(smiles_list)
        smiles = np.stack([self._col(df, c)[rows] for c in cols], axis=1)
        targets = self._col(df, target_name)[rows]
        feats = None
        if add_features_name is not None:
            if isinstance(add_features_name, (list, tuple)):
                feats = np.stack([self._col(df, c)[rows] for c in add_features_name], axis=1)
            else:
                feats = self._col(df, add_features_name)[rows].reshape(-1, 1)
        return smiles, targets, feats

    @staticmethod
    def _perm(cache, n, seed):
        p = cache.get(n)
        if p is None:
            p = cache[n] = np.random.RandomState(seed).permutation(n)
        return p

    # ---- listwise / pointwise training batches -------------------------------------------
    def plan_batch_reactions(self, df=None, batch_size: int = 50, shuffle_query=True, shuffle_batch=True, seed=0):
        """Row positions and scope of every step of ``generate_batch_reactions`` (336-421)."""
        df = self.df if df is None else df
        reactants, rows = self._index(df)
        order = np.arange(len(reactants))
        if shuffle_query:
            order = _shuffled(order, seed)
        perms: dict = {}
        room, chunk, scope = batch_size, [], []
        for g in order:
            r = rows[g]
            n = len(r)
            if room - n >= 0:
                chunk.append(r[self._perm(perms, n, seed)] if shuffle_batch else r)
                scope.append(n)
                room -= n
                if room < 2:
                    yield np.concatenate(chunk), scope
                    room, chunk, scope = batch_size, [], []
            else:
                # DataFrame.sample(n=room) happens even with shuffle_batch=False (load_reactions.py:397)
                chunk.append(r[self._perm(perms, n, seed)[:room]])
                scope.append(room)
                yield np.concatenate(chunk), scope
                room, chunk, scope = batch_size, [], []
        if room < batch_size:
            yield np.concatenate(chunk), scope

    def generate_batch_reactions(self, df=None, batch_size: int = 50, smiles_list=None, target_name='std_targ',
                                 shuffle_query=True, shuffle_batch=True, seed=0, add_features_name=None):
        """Yields ``(smiles [n,2], targets [n,1], scope, add_features [n,f] | None)`` (336-421)."""
        df = self.df if df is None else df
        for rows, scope in self.plan_batch_reactions(df, batch_size, shuffle_query, shuffle_batch, seed):
            smiles, targets, feats = self._gather(df, rows, smiles_list, target_name, add_features_name)
            yield smiles, targets.reshape(-1, 1), scope, feats

    # ---- one group per yield (RankNet, validation) -----------------------------------------
    def plan_batch_per_query(self, df=None, shuffle_query=True, shuffle_batch=True, seed=0):
        df = self.df if df is None else df
        reactants, rows = self._index(df)
        order = np.arange(len(reactants))
        if shuffle_query:
            order = _shuffled(order, seed)
        perms: dict = {}
        for g in order:
            r = rows[g]
            yield r[self._perm(perms, len(r), seed)] if shuffle_batch else r

    def generate_batch_per_query(self, df=None, smiles_list=None, target_name='std_targ', shuffle_query=True,
                                 shuffle_batch=True, seed=0, add_features_name=None):
        """Yields ``(smiles, targets, add_features)`` per reactant (235-273).  As in the reference, the
        additional feature column is read from ``target_name`` (the leak at lines 264-267)."""
        df = self.df if df is None else df
        for rows in self.plan_batch_per_query(df, shuffle_query, shuffle_batch, seed):
            smiles, targets, feats = self._gather(df, rows, smiles_list, target_name,
                                                  target_name if add_features_name is not None else None)
            yield smiles, targets, feats

    def generate_batch_querys(self, df=None, batch_size: int = 2, smiles_list=None, target_name='std_targ', shuffle_query=True,
                              shuffle_batch=True, seed=0, add_features_name=None, use_flag=False):
        """``batch_size`` whole groups per yield (275-334); used by the evaluation routines."""
        df = self.df if df is None else df
        chunk, scope = [], []
        for rows in self.plan_batch_per_query(df, shuffle_query, shuffle_batch, seed):
            chunk.append(rows)
            scope.append(len(rows))
            if len(scope) >= batch_size:
                s, t, f = self._gather(df, np.concatenate(chunk), smiles_list, target_name, add_features_name)
                yield s, t.reshape(-1, 1), scope, f
                chunk, scope = [], []
        if scope:
            s, t, f = self._gather(df, np.concatenate(chunk), smiles_list, target_name, add_features_name)
            yield s, t.reshape(-1, 1), scope, f

    def generate_batch(self, df=None, batch_size: int = 2, smiles_list=None, target_name='std_targ', shuffle_data=True, seed=0):
        """Plain row batches (423-455)."""
        df = self.df if df is None else df
        if shuffle_data:
            df = df.sample(frac=1, random_state=seed)
        cols = ['rsmi', 'psmi'] if smiles_list is None else list(smiles_list)
        for lo in range(0, df.shape[0], batch_size):
            part = df.iloc[lo:lo + batch_size]
            yield part[cols].values, part[target_name].values.reshape(-1, 1)

    def get_num_pairs(self):
        if self.num_pairs is None:
            total = 0
            for _, target, _ in self.generate_batch_per_query(self.df):
                t = np.asarray(target).reshape(-1, 1)
                total += int(np.sum(t - t.T > 0)) * 2
            self.num_pairs = total
        return self.num_pairs


class Parsing_features:
    """SMILES -> MolGraph cache and batch assembly (load_reactions.py:540-586).

    The cache doubles as a ``MoleculeStore``: every molecule is registered once and mirrored in HBM the first time a
    batch goes to the device, after which a batch is a vector of store ids (``parsing_smiles`` = one dict lookup per
    SMILES) assembled on the GPU.  Pre-built graphs (synthetic molecules, graphs featurised elsewhere) can be registered
    with ``add``.  ``smiles2graph`` keeps the reference's attribute name and contents."""

    def __init__(self, graphs: Optional[dict] = None):
        self.smiles2graph = {}
        self.store = MoleculeStore()
        self._sid: Dict[str, int] = {}
        for k, v in (graphs or {}).items():
            self.add(k, v)

    def add(self, smiles: str, graph) -> None:
        self.smiles2graph[smiles] = graph
        self._sid.pop(smiles, None)

    def _register(self, smi) -> int:
        g = self.smiles2graph.get(smi)
        if g is None:
            g = self.smiles2graph[smi] = MolGraph(smi, reaction=True, atom_messages=False)
        sid = self.store.register(_pack_of(g))
        self._sid[smi] = sid
        return sid

    def parsing_smiles(self, smiles: list = None):
        if smiles is None:
            return None
        sid_of = self._sid
        try:
            ids = [sid_of[s] for s in smiles]
        except KeyError:
            ids = [sid_of[s] if s in sid_of else self._register(s) for s in smiles]
        if ids and min(ids) < 0:                    # a molecule the store cannot hold: host path
            return BatchMolGraph([self.smiles2graph[s] for s in smiles])
        return BatchMolGraph.from_store(self.store, np.asarray(ids, dtype=np.int32), smiles)

    def parsing_ids(self, smiles) -> Optional[np.ndarray]:
        """Store ids of the molecules (registering new ones), or None when one of them cannot live in the store.  The batched
        evaluation / RankNet paths hand these to ``DeviceGraph.from_id_groups`` instead of building one BatchMolGraph per group."""
        sid_of = self._sid
        try:
            ids = [sid_of[s] for s in smiles]
        except KeyError:
            ids = [sid_of[s] if s in sid_of else self._register(s) for s in smiles]
        if ids and min(ids) < 0:
            return None
        return np.asarray(ids, dtype=np.int32)

    def frame_ids(self, df, column: str) -> np.ndarray:
        """Store id of ``df[column]``'s molecule for EVERY row of the frame (int32, -1 where a molecule cannot live in the store), computed
        once per (frame, column) and kept while the frame is alive.  With it a training batch is two integer fancy-indexes of the planner's
        row positions instead of one dictionary lookup per SMILES per step (train/step.py: TrainStep.prepare_rows) -- what keeps the
        data-parallel ranks, each of which plans the GLOBAL batch, off the host's critical path."""
        import weakref
        cache = self.__dict__.setdefault("_frame_ids", {})
        key = (id(df), column)
        hit = cache.get(key)
        if hit is not None and hit[0]() is df and hit[1].shape[0] == len(df):
            return hit[1]
        codes, uniq = pd.factorize(df[column].values, sort=False)     # one dictionary lookup per DISTINCT molecule
        ids = self.parsing_ids_or_minus_one(list(uniq))[codes].astype(np.int32)
        cache[key] = (weakref.ref(df), ids)
        return ids

    def parsing_ids_or_minus_one(self, smiles) -> np.ndarray:
        sid_of = self._sid
        out = np.empty(len(smiles), np.int32)
        for i, s in enumerate(smiles):
            sid = sid_of.get(s)
            out[i] = self._register(s) if sid is None else sid
        return out

    def parsing_reactions(self, reactions: list = None):
        if reactions is None:
            return [None, None]
        reactions = np.asarray(reactions, dtype=object)
        return [self.parsing_smiles(reactions[:, 0].tolist()), self.parsing_smiles(reactions[:, 1].tolist())]

    def clear_cache(self):
        self.smiles2graph.clear()
        self._sid.clear()
        self.store = MoleculeStore()
