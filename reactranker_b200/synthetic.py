"""Synthetic reaction graphs of the shape SURVEY.md §8(d) names.

RDKit featurisation is the reference's host-side input stage and is not installed
here, so benchmarks, smoke tests and parity tests feed the path with synthetic
molecules that follow the ``MolGraph`` layout contract of the reference
(features/featurization.py:135-210): atoms in atom-map order, for every bonded pair
a1<a2 two directed bonds ``b1 = a1->a2`` (listed as incoming to a2) and
``b2 = a2->a1``, ``f_bonds[b] = f_atoms[source] || f_bond``, ``b2revb[b] = b ^ 1``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

ATOM_FDIM = 61   # features/featurization.py:63
BOND_FDIM = 22   # features/featurization.py:64
_ATOM_BLOCKS = (16, 6, 6, 5, 6, 6, 6)           # one-hot blocks of atom_features (featurization.py:76-84)
_MASSES = np.array([1.008, 12.011, 14.007, 15.999]) * 0.01


@dataclass
class SynthMol:
    """One molecule in MolGraph layout, numpy-backed (lists on demand)."""
    smiles: str
    n_atoms: int
    n_bonds: int
    f_atoms_np: np.ndarray      # [A, 61] float32
    f_bond_np: np.ndarray       # [B, 22] float32  pure bond features
    b2a_np: np.ndarray          # [B] int32  source atom of each directed bond
    a2b_flat: np.ndarray        # CSR payload: incoming bonds per atom, in MolGraph append order
    a2b_ptr: np.ndarray         # [A+1] int32
    _lists: Optional[dict] = field(default=None, repr=False)

    # --- MolGraph attribute names (python lists, as the reference stores them) ---
    def _mk_lists(self):
        if self._lists is None:
            fb = np.concatenate([self.f_atoms_np[self.b2a_np], self.f_bond_np], axis=1) if self.n_bonds else \
                np.zeros((0, ATOM_FDIM + BOND_FDIM), np.float32)
            self._lists = dict(
                f_atoms=self.f_atoms_np.tolist(),
                f_bonds=fb.tolist(),
                a2b=[self.a2b_flat[self.a2b_ptr[a]:self.a2b_ptr[a + 1]].tolist() for a in range(self.n_atoms)],
                b2a=self.b2a_np.tolist(),
                b2revb=(np.arange(self.n_bonds) ^ 1).tolist(),
            )
        return self._lists

    @property
    def f_atoms(self): return self._mk_lists()["f_atoms"]
    @property
    def f_bonds(self): return self._mk_lists()["f_bonds"]
    @property
    def a2b(self): return self._mk_lists()["a2b"]
    @property
    def b2a(self): return self._mk_lists()["b2a"]
    @property
    def b2revb(self): return self._mk_lists()["b2revb"]


def _random_edges(rng: np.random.Generator, n: int, deg_cap: int = 4, ring_p: float = 0.5) -> List[Tuple[int, int]]:
    deg = np.zeros(n, np.int64)
    edges = set()
    for i in range(1, n):
        cand = np.flatnonzero(deg[:i] < deg_cap)
        j = int(cand[rng.integers(len(cand))])
        edges.add((j, i))
        deg[i] += 1
        deg[j] += 1
    if n >= 4 and rng.random() < ring_p:
        free = np.flatnonzero(deg < deg_cap)
        if len(free) >= 2:
            for _ in range(8):
                a, b = sorted(int(x) for x in rng.choice(free, 2, replace=False))
                if (a, b) not in edges:
                    edges.add((a, b))
                    break
    return sorted(edges)


def make_molecule(rng: np.random.Generator, n_atoms: int, token: str,
                  edges: Optional[List[Tuple[int, int]]] = None) -> SynthMol:
    if edges is None:
        edges = _random_edges(rng, n_atoms)
    fa = np.zeros((n_atoms, ATOM_FDIM), np.float32)
    col = 0
    for blk in _ATOM_BLOCKS:
        fa[np.arange(n_atoms), col + rng.integers(blk, size=n_atoms)] = 1.0
        col += blk
    fa[:, col] = rng.random(n_atoms) < 0.2
    fa[:, col + 1] = _MASSES[rng.integers(4, size=n_atoms)].astype(np.float32)
    fa[:, col + 2:col + 10] = rng.random((n_atoms, 8)) < 0.1

    ne = len(edges)
    fbond = np.zeros((ne, BOND_FDIM), np.float32)
    if ne:
        fbond[np.arange(ne), 1 + rng.integers(4, size=ne)] = 1.0
        fbond[:, 5] = rng.random(ne) < 0.3
        fbond[:, 6] = rng.random(ne) < 0.3
        fbond[:, 7:15] = rng.random((ne, 8)) < 0.1
        fbond[np.arange(ne), 15 + rng.integers(7, size=ne)] = 1.0
    e = np.asarray(edges, np.int32).reshape(-1, 2)
    b2a = e.reshape(-1).astype(np.int32)            # [a1, a2, a1', a2', ...]: bond 2k from a1, 2k+1 from a2
    dst = e[:, ::-1].reshape(-1)                    # bond 2k goes into a2, 2k+1 into a1
    # incoming lists in append order == increasing bond index per destination atom
    order = np.argsort(dst, kind="stable")
    ptr = np.zeros(n_atoms + 1, np.int32)
    np.add.at(ptr, dst + 1, 1)
    ptr = np.cumsum(ptr).astype(np.int32)
    return SynthMol(smiles=token, n_atoms=n_atoms, n_bonds=2 * ne, f_atoms_np=fa,
                    f_bond_np=np.repeat(fbond, 2, axis=0), b2a_np=b2a,
                    a2b_flat=order.astype(np.int32), a2b_ptr=ptr)


def star_molecule(rng: np.random.Generator, n_leaves: int, token: str) -> SynthMol:
    """A centre atom bonded to ``n_leaves`` leaves: forces max_num_bonds = n_leaves in
    whichever batch contains it (the padding-row regression case, SURVEY.md §0 trap 1)."""
    return make_molecule(rng, n_leaves + 1, token, edges=[(0, i) for i in range(1, n_leaves + 1)])


@dataclass
class SynthDataset:
    """Columns named as the reference's CSV (main.py:41-49; load_reactions.py:148)."""
    mols: Dict[str, SynthMol]
    rsmi: np.ndarray          # object [rows]
    psmi: np.ndarray          # object [rows]
    lgk: np.ndarray           # float64 [rows]
    temp: np.ndarray          # float64 [rows]
    flag: np.ndarray          # int64 [rows]  group id

    def to_dataframe(self):
        import pandas as pd
        df = pd.DataFrame({
            "rsmi": pd.Series(self.rsmi, dtype=object), "psmi": pd.Series(self.psmi, dtype=object),
            "rsmi_mapped": pd.Series(self.rsmi, dtype=object), "psmi_mapped": pd.Series(self.psmi, dtype=object),
            "lgk": self.lgk, "temp": self.temp, "flag": self.flag})
        return df


def make_dataset(seed: int, group_sizes, atoms_lo: int = 12, atoms_hi: int = 28,
                 star_leaves_in_group: Optional[Dict[int, int]] = None) -> SynthDataset:
    """``group_sizes``: iterable of candidates per reactant group.  One reactant graph per
    group and one independent product graph per candidate with the same atom count
    (needed for the atom-wise ``p - r``, models/base_model.py:168)."""
    rng = np.random.default_rng(seed)
    mols: Dict[str, SynthMol] = {}
    rs, ps, gid = [], [], []
    for g, n_cand in enumerate(group_sizes):
        leaves = (star_leaves_in_group or {}).get(g)
        n = int(rng.integers(atoms_lo, atoms_hi + 1)) if leaves is None else leaves + 1
        rtok = f"R{seed}_{g}"
        mols[rtok] = make_molecule(rng, n, rtok) if leaves is None else star_molecule(rng, leaves, rtok)
        for c in range(int(n_cand)):
            ptok = f"P{seed}_{g}_{c}"
            mols[ptok] = make_molecule(rng, n, ptok)
            rs.append(rtok)
            ps.append(ptok)
            gid.append(g)
    rows = len(rs)
    return SynthDataset(mols=mols, rsmi=np.array(rs, dtype=object), psmi=np.array(ps, dtype=object),
                        lgk=rng.standard_normal(rows), temp=rng.random(rows), flag=np.asarray(gid, np.int64))
