"""Graph layout contract of the hot path: ``MolGraph`` -> ``BatchMolGraph`` -> device.

Mirrors the public surface of the reference's ``reactranker/features/featurization.py``
(``ATOM_FDIM``, ``BOND_FDIM``, ``MolGraph``, ``BatchMolGraph.get_components()/get_a2a()``,
``mol2graph``) with bit-identical index construction (featurization.py:246-329), but is
built B200-first:

* every molecule keeps numpy arrays already padded to the device row strides, so a batch
  is a handful of ``np.concatenate(..., out=pinned)`` calls instead of a Python loop over
  atoms and bonds (the reference spends ~1.2 ms per reaction there);
* the reference tensors (``f_atoms f_bonds a2b b2a b2revb`` as float32 / int64) are
  materialised lazily, only if somebody asks for them through the reference API;
* ``DeviceGraph`` is the int32, 16-byte-aligned, single-H2D-copy form the CUDA kernels
  read (``rr_graph`` in include/rr_sm100.h).  It carries the "padded a2b gathers row 0"
  semantics as a per-atom (in-degree, pad multiplicity, pad rows) record, which also lets
  several reference batches (RankNet groups, data-parallel shards) share one launch while
  each keeps its own ``max_num_bonds``.

RDKit is only needed by ``MolGraph(smiles)``; it is imported lazily there.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import threading

import numpy as np
import torch

ATOM_FDIM = 61          # 16+6+6+5+6+6+6 one-hot blocks + aromatic + mass + 8 ring sizes (featurization.py:63)
BOND_FDIM = 22          # 14 + 8 (featurization.py:64)
FBOND_TOTAL = ATOM_FDIM + BOND_FDIM
FA_LD = 64              # RR_FA_LD
FB_LD = 88              # RR_FB_LD

ELEM_LIST = ['H', 'C', 'N', 'O', 'S', 'F', 'Si', 'P', 'Cl', 'Br', 'Mg', 'Na', 'I', 'B', 'K']


def get_atom_fdim() -> int:
    return ATOM_FDIM


def get_bond_fdim() -> int:
    return BOND_FDIM


# ------------------------------------------------------------------------------------------
# RDKit featurisation (host input stage, semantics of featurization.py:28-132)
# ------------------------------------------------------------------------------------------
def onek_encoding_unk(value, choices) -> List[int]:
    """One-hot with a trailing 'unknown' slot (featurization.py:28-42)."""
    enc = [0] * (len(choices) + 1)
    enc[choices.index(value) if value in choices else -1] = 1
    return enc


def _rdkit():
    try:
        from rdkit import Chem
    except ImportError as e:  # pragma: no cover - rdkit is not in the build image
        raise ImportError("MolGraph(smiles) needs RDKit for featurisation; synthetic graphs use "
                          "MolGraph.from_arrays / reactranker_b200.synthetic") from e
    return Chem


def str_to_mol(string: str, explicit_hydrogens: bool = True):
    Chem = _rdkit()
    if string.startswith('InChI'):
        mol = Chem.MolFromInchi(string, removeHs=not explicit_hydrogens)
    else:
        params = Chem.SmilesParserParams()
        params.removeHs = not explicit_hydrogens
        mol = Chem.MolFromSmiles(string, params)
    return Chem.AddHs(mol) if explicit_hydrogens else Chem.RemoveHs(mol)


def atom_features(atom) -> List[float]:
    Chem = _rdkit()
    hyb = Chem.rdchem.HybridizationType
    f = (onek_encoding_unk(atom.GetSymbol(), ELEM_LIST)
         + onek_encoding_unk(atom.GetTotalDegree(), [0, 1, 2, 3, 4])
         + onek_encoding_unk(atom.GetFormalCharge(), [-2, -1, 0, 1, 2])
         + onek_encoding_unk(atom.GetChiralTag(), [0, 1, 2, 3])
         + onek_encoding_unk(atom.GetTotalNumHs(), [0, 1, 2, 3, 4])
         + onek_encoding_unk(atom.GetNumRadicalElectrons(), [0, 1, 2, 3, 4])
         + onek_encoding_unk(atom.GetHybridization(), [hyb.SP, hyb.SP2, hyb.SP3, hyb.SP3D, hyb.SP3D2])
         + [1 if atom.GetIsAromatic() else 0]
         + [atom.GetMass() * 0.01])
    f += [atom.IsInRingSize(k) for k in range(3, 11)]
    return f


def bond_features(bond) -> List[float]:
    Chem = _rdkit()
    if bond is None:
        return [1] + [0] * (BOND_FDIM - 1)
    bt = bond.GetBondType()
    f = [0, bt == Chem.BondType.SINGLE, bt == Chem.BondType.DOUBLE, bt == Chem.BondType.TRIPLE,
         bt == Chem.BondType.AROMATIC,
         (bond.GetIsConjugated() if bt is not None else 0), (bond.IsInRing() if bt is not None else 0)]
    f += [(bond.IsInRingSize(k) if bt is not None else 0) for k in range(3, 11)]
    f += onek_encoding_unk(int(bond.GetStereo()), list(range(6)))
    return f


# ------------------------------------------------------------------------------------------
# MolGraph
# ------------------------------------------------------------------------------------------
class MolGraph:
    """One molecule (featurization.py:135-210).  Reference attributes (python lists)
    ``smiles n_atoms n_bonds f_atoms f_bonds a2b b2a b2revb`` are kept; the numpy pack used
    for batching is built once and cached."""

    def __init__(self, smiles: str, reaction: bool = True, atom_messages: bool = False):
        if atom_messages:
            raise NotImplementedError("atom_messages=True is not on the hot path (base_model.py always uses bond messages)")
        self.smiles = smiles
        mol = str_to_mol(smiles, explicit_hydrogens=True)
        self.n_atoms = mol.GetNumAtoms()
        self.n_bonds = 0
        atoms = sorted(mol.GetAtoms(), key=lambda a: a.GetAtomMapNum()) if reaction else list(mol.GetAtoms())
        self.f_atoms = [atom_features(a) for a in atoms]
        self.f_bonds, self.a2b, self.b2a, self.b2revb = [], [[] for _ in range(self.n_atoms)], [], []
        for a1 in range(self.n_atoms):
            for a2 in range(a1 + 1, self.n_atoms):
                bond = mol.GetBondBetweenAtoms(atoms[a1].GetIdx(), atoms[a2].GetIdx())
                if bond is None:
                    continue
                fb = bond_features(bond)
                self.f_bonds.append(self.f_atoms[a1] + fb)
                self.f_bonds.append(self.f_atoms[a2] + fb)
                b1, b2 = self.n_bonds, self.n_bonds + 1
                self.a2b[a2].append(b1)     # b1 = a1 -> a2
                self.b2a.append(a1)
                self.a2b[a1].append(b2)     # b2 = a2 -> a1
                self.b2a.append(a2)
                self.b2revb += [b2, b1]
                self.n_bonds += 2
        self._pack = None

    @classmethod
    def from_arrays(cls, smiles: str, f_atoms: np.ndarray, f_bond: np.ndarray, b2a: np.ndarray,
                    a2b_flat: np.ndarray, a2b_ptr: np.ndarray, b2revb: Optional[np.ndarray] = None) -> "MolGraph":
        """Build from arrays (no RDKit): ``f_atoms`` [A,61], ``f_bond`` [B,22] pure bond
        features, ``b2a`` [B] source atoms, CSR ``a2b`` incoming-bond lists."""
        g = object.__new__(cls)
        g.smiles = smiles
        g.n_atoms, g.n_bonds = int(f_atoms.shape[0]), int(b2a.shape[0])
        g._pack = _MolPack.build(np.asarray(f_atoms, np.float32), np.asarray(f_bond, np.float32),
                                 np.asarray(b2a, np.int32), np.asarray(a2b_flat, np.int32),
                                 np.asarray(a2b_ptr, np.int32),
                                 (np.arange(g.n_bonds, dtype=np.int32) ^ 1) if b2revb is None else np.asarray(b2revb, np.int32))
        return g

    @classmethod
    def from_synthetic(cls, m) -> "MolGraph":
        return cls.from_arrays(m.smiles, m.f_atoms_np, m.f_bond_np, m.b2a_np, m.a2b_flat, m.a2b_ptr)

    def pack(self) -> "_MolPack":
        if self._pack is None:
            A, B = self.n_atoms, self.n_bonds
            fa = np.asarray(self.f_atoms, np.float32).reshape(A, ATOM_FDIM)
            fb = np.asarray(self.f_bonds, np.float32).reshape(B, FBOND_TOTAL)
            ptr = np.zeros(A + 1, np.int32)
            ptr[1:] = np.cumsum([len(x) for x in self.a2b])
            flat = np.asarray([b for row in self.a2b for b in row], np.int32)
            self._pack = _MolPack.build(fa, fb[:, ATOM_FDIM:], np.asarray(self.b2a, np.int32), flat, ptr,
                                        np.asarray(self.b2revb, np.int32), f_bonds_full=fb)
        return self._pack

    def __getattr__(self, name):
        # list views for molecules created from arrays (reference attribute names)
        if name in ("f_atoms", "f_bonds", "a2b", "b2a", "b2revb") and self.__dict__.get("_pack") is not None:
            p = self.__dict__["_pack"]
            if name == "f_atoms":
                return p.f_atoms[:, :ATOM_FDIM].tolist()
            if name == "f_bonds":
                return p.f_bonds[:, :FBOND_TOTAL].tolist()
            if name == "a2b":
                return [p.a2b_flat[p.a2b_ptr[a]:p.a2b_ptr[a + 1]].tolist() for a in range(self.n_atoms)]
            if name == "b2a":
                return p.b2a.tolist()
            return p.b2revb.tolist()
        raise AttributeError(name)


class _MolPack:
    """Per-molecule arrays, already in device row strides."""
    __slots__ = ("f_atoms", "f_bonds", "b2a", "b2revb", "a2b_flat", "a2b_ptr", "deg", "n_atoms", "n_bonds", "sid", "store")

    @staticmethod
    def build(f_atoms, f_bond, b2a, a2b_flat, a2b_ptr, b2revb, f_bonds_full=None) -> "_MolPack":
        p = _MolPack()
        p.sid, p.store = None, None
        A, B = f_atoms.shape[0], b2a.shape[0]
        p.n_atoms, p.n_bonds = A, B
        p.f_atoms = np.zeros((A, FA_LD), np.float32)
        p.f_atoms[:, :ATOM_FDIM] = f_atoms
        p.f_bonds = np.zeros((B, FB_LD), np.float32)
        if f_bonds_full is not None:
            p.f_bonds[:, :FBOND_TOTAL] = f_bonds_full
        elif B:
            p.f_bonds[:, :ATOM_FDIM] = f_atoms[b2a]          # f_bonds[b] = f_atoms[src] || f_bond (featurization.py:198-199)
            p.f_bonds[:, ATOM_FDIM:FBOND_TOTAL] = f_bond
        p.b2a, p.b2revb = b2a, b2revb
        p.a2b_flat, p.a2b_ptr = a2b_flat, a2b_ptr
        p.deg = np.diff(a2b_ptr).astype(np.int32)
        return p


def _pack_of(m) -> _MolPack:
    if isinstance(m, MolGraph):
        return m.pack()
    cached = getattr(m, "_rr_pack", None)
    if cached is None:                                   # duck-typed molecules (synthetic.SynthMol, reference MolGraph)
        if hasattr(m, "f_atoms_np"):
            cached = MolGraph.from_synthetic(m).pack()
        else:
            g = object.__new__(MolGraph)
            g.__dict__.update(smiles=m.smiles, n_atoms=m.n_atoms, n_bonds=m.n_bonds, f_atoms=m.f_atoms, f_bonds=m.f_bonds,
                              a2b=m.a2b, b2a=m.b2a, b2revb=m.b2revb, _pack=None)
            cached = g.pack()
        try:
            m._rr_pack = cached
        except AttributeError:
            pass
    return cached


# ------------------------------------------------------------------------------------------
# MoleculeStore: the MolGraph cache, resident in HBM
# ------------------------------------------------------------------------------------------
class MoleculeStore:
    """Per-molecule arrays of every registered molecule, concatenated; mirrored on one CUDA device on demand.
    With it a batch costs a few bytes per molecule of host->device traffic (ids + row offsets) and one assembly
    kernel (csrc/rr_assemble.cu) instead of a host rebuild plus ~36 kB per reaction of PCIe traffic."""

    def __init__(self):
        self.packs: List[_MolPack] = []
        self.nA = np.zeros(0, np.int32)
        self.nB = np.zeros(0, np.int32)
        self.maxdeg = np.zeros(0, np.int32)
        self.aoff = np.zeros(0, np.int64)
        self.boff = np.zeros(0, np.int64)
        self._tot_a = 0
        self._tot_b = 0
        self._dev = None
        self._uploaded = 0
        self._d = {}
        self.c = None
        self.h2d_bytes_total = 0
        self._lock = threading.RLock()      # the prefetch worker registers molecules while the training thread uploads / reads

    def __len__(self):
        return len(self.packs)

    def register(self, pack: "_MolPack") -> int:
        """Returns the store id, or -1 for a molecule whose reverse-bond map is not the canonical ``b xor 1``."""
        sid = getattr(pack, "sid", None)
        if sid is not None and getattr(pack, "store", None) is self:
            return sid
        if pack.n_bonds and not np.array_equal(pack.b2revb, np.arange(pack.n_bonds, dtype=np.int32) ^ 1):
            return -1
        with self._lock:
            return self._register_locked(pack)

    def _register_locked(self, pack: "_MolPack") -> int:
        sid = getattr(pack, "sid", None)
        if sid is not None and getattr(pack, "store", None) is self:
            return sid
        sid = len(self.packs)
        if sid >= self.nA.shape[0]:
            cap = max(1024, 2 * self.nA.shape[0])
            for name in ("nA", "nB", "maxdeg", "aoff", "boff"):
                old = getattr(self, name)
                grown = np.zeros(cap, old.dtype)
                grown[:old.shape[0]] = old
                setattr(self, name, grown)
        self.nA[sid], self.nB[sid] = pack.n_atoms, pack.n_bonds
        self.maxdeg[sid] = int(pack.deg.max()) if pack.n_atoms else 0
        self.aoff[sid], self.boff[sid] = self._tot_a, self._tot_b
        self._tot_a += pack.n_atoms
        self._tot_b += pack.n_bonds
        self.packs.append(pack)              # last: a reader that sees the pack sees its table entries
        pack.sid, pack.store = sid, self
        return sid

    # ---- device mirror -----------------------------------------------------------------
    _SPECS = (("f_atoms", torch.float32, FA_LD, "a"), ("f_bonds", torch.float32, FB_LD, "b"), ("deg", torch.int32, 1, "a"),
              ("a2b_start", torch.int32, 1, "a"), ("a2b_flat", torch.int32, 1, "b"), ("b2a", torch.int32, 1, "b"),
              ("n_atoms", torch.int32, 1, "m"), ("n_bonds", torch.int32, 1, "m"), ("atom_off", torch.int64, 1, "m"), ("bond_off", torch.int64, 1, "m"))

    def sync(self, device) -> None:
        """Upload molecules registered since the last call (amortised: capacity doubles)."""
        with self._lock:
            self._sync_locked(torch.device(device))

    def _sync_locked(self, dev) -> None:
        if self._dev is not None and self._dev != dev:
            raise RuntimeError(f"MoleculeStore already lives on {self._dev}")
        n = len(self.packs)
        if self._dev == dev and self._uploaded == n:
            return
        from .. import _lib
        self._dev = dev
        lo = self._uploaded
        new = self.packs[lo:n]
        a_lo, b_lo = int(self.aoff[lo]) if lo < n else self._tot_a, int(self.boff[lo]) if lo < n else self._tot_b
        need = {"a": self._tot_a, "b": self._tot_b, "m": n}
        host = {
            "f_atoms": np.concatenate([p.f_atoms for p in new]) if new else np.zeros((0, FA_LD), np.float32),
            "f_bonds": np.concatenate([p.f_bonds for p in new]) if new else np.zeros((0, FB_LD), np.float32),
            "deg": np.concatenate([p.deg for p in new]), "a2b_start": np.concatenate([p.a2b_ptr[:-1] for p in new]).astype(np.int32),
            "a2b_flat": np.concatenate([p.a2b_flat for p in new]).astype(np.int32), "b2a": np.concatenate([p.b2a for p in new]).astype(np.int32),
            "n_atoms": self.nA[lo:n], "n_bonds": self.nB[lo:n], "atom_off": self.aoff[lo:n], "bond_off": self.boff[lo:n],
        }
        start = {"a": a_lo, "b": b_lo, "m": lo}
        for name, dtype, width, kind in self._SPECS:
            cur = self._d.get(name)
            rows = need[kind]
            if cur is None or cur.shape[0] < rows:
                cap = max(rows, 2 * (cur.shape[0] if cur is not None else 0), 1)
                grown = torch.empty((cap, width) if width > 1 else (cap,), dtype=dtype, device=dev)
                if cur is not None and start[kind] > 0:
                    grown[:start[kind]] = cur[:start[kind]]
                self._d[name] = cur = grown
            h = torch.from_numpy(np.ascontiguousarray(host[name]))
            if h.numel():
                if dev.type == "cuda":
                    h = h.pin_memory()
                cur[start[kind]:start[kind] + h.shape[0]].copy_(h, non_blocking=True)
                self.h2d_bytes_total += h.numel() * h.element_size()
        self._uploaded = n
        c = _lib.RRMolStore()
        for name, *_ in self._SPECS:
            setattr(c, name, self._d[name].data_ptr())
        self.c = c


# ------------------------------------------------------------------------------------------
# BatchMolGraph
# ------------------------------------------------------------------------------------------
class BatchMolGraph:
    """Batch of molecules with the reference's index construction (featurization.py:246-290): row 0 of atoms and
    bonds is padding, per-molecule indices are shifted by the running totals, ``a2b`` is right-padded with 0 to
    ``max_num_bonds = max(1, max in-degree)``.

    Only the sizes are computed eagerly.  The reference tensors are built (vectorised) when somebody reads them; the
    training path never does: ``to_device`` assembles the batch on the GPU from the molecule store when the batch came
    from ``Parsing_features`` and packs it on the host otherwise."""

    def __init__(self, mol_graphs: Sequence, atom_messages: bool = False):
        if atom_messages:
            raise NotImplementedError("atom_messages=True is not on the hot path")
        packs = [_pack_of(m) for m in mol_graphs]
        self._setup(packs, [m.smiles for m in mol_graphs], None, None)

    @classmethod
    def from_store(cls, store: MoleculeStore, ids: np.ndarray, smiles) -> "BatchMolGraph":
        obj = cls.__new__(cls)
        obj._setup(None, smiles, store, np.asarray(ids, dtype=np.int32))
        return obj

    def _setup(self, packs, smiles, store, ids):
        self.smiles_batch = list(smiles)
        self.atom_fdim, self.bond_fdim = ATOM_FDIM, FBOND_TOTAL
        self._store, self._ids, self._packs_ = store, ids, packs
        if ids is not None:
            nA, nB = store.nA[ids].astype(np.int64), store.nB[ids].astype(np.int64)
            maxdeg = int(store.maxdeg[ids].max()) if len(ids) else 0
        else:
            nA = np.fromiter((p.n_atoms for p in packs), np.int64, len(packs))
            nB = np.fromiter((p.n_bonds for p in packs), np.int64, len(packs))
            maxdeg = max((int(p.deg.max()) if p.n_atoms else 0 for p in packs), default=0)
        self.n_mols = len(nA)
        a_start = 1 + np.concatenate(([0], np.cumsum(nA)[:-1])) if self.n_mols else np.zeros(0, np.int64)
        b_start = 1 + np.concatenate(([0], np.cumsum(nB)[:-1])) if self.n_mols else np.zeros(0, np.int64)
        self.n_atoms = int(1 + nA.sum())
        self.n_bonds = int(1 + nB.sum())
        self._a_start, self._a_size = a_start.astype(np.int32), nA.astype(np.int32)
        self._b_start, self._b_size = b_start.astype(np.int32), nB.astype(np.int32)
        self._max_deg = maxdeg
        self.max_num_bonds = max(1, maxdeg)
        self._lazy = {}
        self._index = None
        self.b2b = None
        self.a2a = None

    @property
    def _packs(self):
        if self._packs_ is None:
            sp = self._store.packs
            self._packs_ = [sp[i] for i in self._ids.tolist()]
        return self._packs_

    @property
    def a_scope(self):
        return list(zip(self._a_start.tolist(), self._a_size.tolist()))

    @property
    def b_scope(self):
        return list(zip(self._b_start.tolist(), self._b_size.tolist()))

    # ---- index tables (int32, reference values), built on first use ------------------------
    def _build_index(self):
        if self._index is None:
            packs = self._packs
            nB = self._b_size.astype(np.int64)
            deg = np.zeros(self.n_atoms, np.int32)
            b2a = np.zeros(self.n_bonds, np.int32)
            b2revb = np.zeros(self.n_bonds, np.int32)
            a2b = np.zeros((self.n_atoms, self.max_num_bonds), np.int32)
            if packs:
                np.concatenate([p.deg for p in packs], out=deg[1:])
                np.concatenate([p.b2a for p in packs], out=b2a[1:])
                np.concatenate([p.b2revb for p in packs], out=b2revb[1:])
                b2a[1:] += np.repeat(self._a_start, nB)
                b2revb[1:] += np.repeat(self._b_start, nB)
                flat = np.concatenate([p.a2b_flat for p in packs]) + np.repeat(self._b_start, nB)     # one incoming entry per bond
                rows = np.repeat(np.arange(self.n_atoms, dtype=np.int64), deg)
                first = np.concatenate(([0], np.cumsum(deg, dtype=np.int64)[:-1]))
                cols = np.arange(flat.shape[0], dtype=np.int64) - np.repeat(first, deg)
                a2b[rows, cols] = flat
            self._index = (deg, a2b, b2a, b2revb)
        return self._index

    _deg = property(lambda self: self._build_index()[0])
    _a2b = property(lambda self: self._build_index()[1])
    _b2a = property(lambda self: self._build_index()[2])
    _b2revb = property(lambda self: self._build_index()[3])

    # ---- reference tensors, built on demand --------------------------------------------
    def _feature_tensor(self, which: str) -> torch.Tensor:
        if which not in self._lazy:
            if which == "f_atoms":
                out = np.zeros((self.n_atoms, ATOM_FDIM), np.float32)
                if self._packs:
                    np.concatenate([p.f_atoms[:, :ATOM_FDIM] for p in self._packs], out=out[1:])
            else:
                out = np.zeros((self.n_bonds, FBOND_TOTAL), np.float32)
                if self._packs:
                    np.concatenate([p.f_bonds[:, :FBOND_TOTAL] for p in self._packs], out=out[1:])
            self._lazy[which] = torch.from_numpy(out)
        return self._lazy[which]

    f_atoms = property(lambda self: self._feature_tensor("f_atoms"))
    f_bonds = property(lambda self: self._feature_tensor("f_bonds"))
    a2b = property(lambda self: torch.from_numpy(self._a2b.astype(np.int64)))
    b2a = property(lambda self: torch.from_numpy(self._b2a.astype(np.int64)))
    b2revb = property(lambda self: torch.from_numpy(self._b2revb.astype(np.int64)))

    def get_components(self):
        """(f_atoms, f_bonds, a2b, b2a, b2revb, a_scope, b_scope) -- featurization.py:292-301."""
        return self.f_atoms, self.f_bonds, self.a2b, self.b2a, self.b2revb, self.a_scope, self.b_scope

    def get_a2a(self) -> torch.Tensor:
        """``b2a[a2b]`` (featurization.py:320-329)."""
        if self.a2a is None:
            self.a2a = torch.from_numpy(self._b2a[self._a2b].astype(np.int64))
        return self.a2a

    def get_b2b(self):
        raise NotImplementedError("get_b2b is unused by the reference models")

    def get_smiles(self):
        return self.smiles_batch

    # ---- device form -------------------------------------------------------------------
    def to_device(self, device, max_num_bonds: Optional[int] = None, non_blocking: bool = True) -> "DeviceGraph":
        key = (str(device), max_num_bonds)
        dg = self._lazy.get(key)
        if dg is None:
            dg = DeviceGraph.from_batches([self], device, [max_num_bonds] if max_num_bonds else None, non_blocking)
            self._lazy[key] = dg
        return dg


def mol2graph(smiles_batch: List[str]) -> BatchMolGraph:
    return BatchMolGraph([MolGraph(s) for s in smiles_batch])


# ------------------------------------------------------------------------------------------
# DeviceGraph: the rr_graph the kernels read
# ------------------------------------------------------------------------------------------
class RRAtomMeta(ctypes.Structure):
    _fields_ = [("deg_flags", ctypes.c_int32), ("pad_count", ctypes.c_int32), ("pad_bond", ctypes.c_int32), ("pad_atom", ctypes.c_int32)]


class RRGraph(ctypes.Structure):
    _fields_ = [("n_atoms", ctypes.c_int32), ("n_bonds", ctypes.c_int32), ("n_mols", ctypes.c_int32), ("wmax", ctypes.c_int32),
                ("n_segments", ctypes.c_int32),
                ("f_atoms", ctypes.c_void_p), ("f_bonds", ctypes.c_void_p), ("a_meta", ctypes.c_void_p), ("a2b", ctypes.c_void_p),
                ("a2b_rev", ctypes.c_void_p), ("a2a", ctypes.c_void_p), ("mol_start", ctypes.c_void_p), ("mol_size", ctypes.c_void_p),
                ("pad_bonds", ctypes.c_void_p), ("pad_atoms", ctypes.c_void_p)]


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class _BufferPool:
    """Device blobs and pinned control buffers of assembled batches, recycled when their DeviceGraph dies.  Batch sizes wander by a
    per cent from step to step; asking the allocators for a fresh, slightly larger block now and then costs a cudaMalloc / cudaHostAlloc
    (5-25 ms) in the middle of training.  Buffers are handed out best-fit with 30 % head-room (a data-parallel shard of a few
    large groups moves by a whole group from step to step; the blobs are a few MB); reuse is stream-ordered (the kernels that
    read the old contents were enqueued before the assembly that overwrites them)."""

    def __init__(self, keep: int = 8):
        self.free = {}
        self.keep = keep
        self.lock = threading.Lock()

    def take(self, nbytes: int, key, make):
        with self.lock:
            lst = self.free.setdefault(key, [])
            # a pinned staging buffer is reusable only once the asynchronous copy out of it has run (its event has completed)
            # (device blobs taken here are rewritten on the stream their previous readers ran on: stream order is enough)
            fits = [e for e in lst if e[0].numel() >= nbytes and (key != "pinned" or e[1] is None or e[1].query())]
            if fits:
                best = min(fits, key=lambda e: e[0].numel())
                lst[:] = [e for e in lst if e is not best]     # identity, not tensor equality
                return best[0]
        return make(int(nbytes * 1.3) + 256)

    def take_device(self, nbytes: int, key, make):
        """(tensor, event | None) for a DEVICE blob that will be written on another stream than the one its previous owner read it on:
        the event was recorded on that owner's stream when the blob came back; the writer's stream waits for it (no host wait)."""
        with self.lock:
            lst = self.free.setdefault(key, [])
            fits = [e for e in lst if e[0].numel() >= nbytes]
            if fits:
                best = min(fits, key=lambda e: e[0].numel())
                lst[:] = [e for e in lst if e is not best]
                return best
        return make(int(nbytes * 1.3) + 256), None

    def give(self, t, key, event=None):
        with self.lock:
            lst = self.free.setdefault(key, [])
            if len(lst) < self.keep:
                lst.append((t, event))


_POOL = _BufferPool()


class _SharedBlob:
    """Pooled device blob + pinned control buffer shared by the two DeviceGraphs of a pair: returned to the pool when both are gone."""

    def __init__(self, blob, key, host, event, consumer_stream):
        self.p = (blob, key, host, event, consumer_stream)

    def __del__(self):
        p, self.p = self.p, None
        if p is not None:
            try:
                # every kernel that reads the blob was enqueued on the consumer stream (the stream that was current when the pair was
                # assembled) before its graphs died -- recorded there explicitly: a garbage-collector run may call this anywhere, also
                # while the assembly stream is current
                freed = torch.cuda.Event()
                freed.record(p[4])
                _POOL.give(p[0], p[1], freed)
                _POOL.give(p[2], "pinned", p[3])
            except Exception:      # interpreter shutdown
                pass


_ASM_STREAMS: dict = {}


def _assembly_stream(dev: torch.device) -> "torch.cuda.Stream":
    """Side stream on which batches are assembled: the id upload and the two rr_graph_assemble launches of batch i+1 run NEXT TO the
    kernels of step i instead of behind them (~0.15 ms of copies per step that the dense layers leave HBM bandwidth for)."""
    s = _ASM_STREAMS.get(dev)
    if s is None:
        s = _ASM_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return s


class DeviceGraph:
    """One or more BatchMolGraphs ("segments") laid out for the kernels and shipped with a
    single host->device copy out of pinned memory."""

    def __init__(self):
        self.blob = None
        self.host_blob = None
        self._pooled = None
        self.c = RRGraph()
        self.h2d_bytes = 0
        self.n_atoms = self.n_bonds = self.n_mols = 0
        self.real_atoms = self.real_bonds = 0

    def __del__(self):
        p = self._pooled
        if p is not None:
            self._pooled = None
            try:
                _POOL.give(p[0], p[1])
                _POOL.give(p[2], "pinned", p[3])
            except Exception:      # interpreter shutdown
                pass

    @staticmethod
    def _sections(nA, nB, nM, wmax, S):
        return [("f_atoms", nA * FA_LD * 4), ("f_bonds", nB * FB_LD * 4), ("a_meta", nA * 16), ("a2b", nA * wmax * 4),
                ("a2b_rev", nA * wmax * 4), ("a2a", nA * wmax * 4), ("mol_start", nM * 4), ("mol_size", nM * 4),
                ("pad_bonds", S * 4), ("pad_atoms", S * 4)]

    @staticmethod
    def control_block(store: MoleculeStore, ids: np.ndarray, lens: np.ndarray, W: Optional[np.ndarray] = None, out: Optional[np.ndarray] = None):
        """The host side of a device-assembled graph for any number of segments: molecule ``ids`` (store ids, segments back to back),
        ``lens`` molecules per segment, optional per-segment ``W`` (max_num_bonds; default max(1, largest in-degree of the segment)).
        Returns ``(ctl int32, (nA, nB, nM, wmax, S), (A_s, B_s, W_s))`` where ctl = [ids | a_start | b_start | W | pad_bond | pad_atom] per
        molecule + [a0 | b0 | W] per segment, exactly what ``rr_graph_assemble`` reads.  Every segment is laid out as the reference lays out
        a BatchMolGraph of its molecules: one padding row, then the molecules' rows (featurization.py:264-290).  The arithmetic is one pass
        in C++ (``rr_batch_build``, csrc/rr_host.cu) writing into ``out`` (e.g. the pinned staging buffer) when given;
        ``control_block_numpy`` is the vectorised restatement the tests hold it to."""
        from .. import _lib
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        S, nM = int(lens.shape[0]), int(ids.shape[0])
        n = 6 * nM + 3 * S
        ctl = np.empty(n, np.int32) if out is None else out[:n]
        dims = np.zeros(5, np.int64)
        seg = np.zeros((3, max(S, 1)), np.int64)
        W_arr = None if W is None else np.ascontiguousarray(W, dtype=np.int64)
        if W_arr is not None and W_arr.shape[0] != S:
            raise ValueError(f"{W_arr.shape[0]} max_num_bonds overrides for {S} segments")
        n_store = len(store)
        st = _lib.lib().rr_batch_build(nM, ids.ctypes.data, n_store, store.nA.ctypes.data, store.nB.ctypes.data, store.maxdeg.ctypes.data, S,
                                       lens.ctypes.data, None if W_arr is None else W_arr.ctypes.data, ctl.ctypes.data, dims.ctypes.data,
                                       seg[0].ctypes.data, seg[1].ctypes.data, seg[2].ctypes.data)
        if st != 0:
            raise ValueError(_lib.lib().rr_last_error().decode(errors="replace"))
        return ctl, tuple(int(v) for v in dims), (seg[0, :S], seg[1, :S], seg[2, :S])

    @staticmethod
    def control_block_numpy(store: MoleculeStore, ids: np.ndarray, lens: np.ndarray, W: Optional[np.ndarray] = None):
        """Vectorised numpy restatement of ``control_block`` (test reference for rr_batch_build)."""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        lens = np.asarray(lens, dtype=np.int64)
        S, nM = int(lens.shape[0]), int(ids.shape[0])
        if int(lens.sum()) != nM:
            raise ValueError(f"segment lengths sum to {int(lens.sum())}, {nM} molecule ids given")
        seg_of = np.repeat(np.arange(S, dtype=np.int64), lens)
        nA_m, nB_m = store.nA[ids].astype(np.int64), store.nB[ids].astype(np.int64)
        A_s = 1 + np.bincount(seg_of, weights=nA_m, minlength=S).astype(np.int64)
        B_s = 1 + np.bincount(seg_of, weights=nB_m, minlength=S).astype(np.int64)
        deg_s = np.zeros(S, np.int64)
        np.maximum.at(deg_s, seg_of, store.maxdeg[ids].astype(np.int64))
        W_min = np.maximum(1, deg_s)
        if W is None:
            W_s = W_min
        else:
            W_s = np.asarray(W, dtype=np.int64)
            bad = np.nonzero(W_s < W_min)[0]
            if bad.size:
                raise ValueError(f"max_num_bonds override {int(W_s[bad[0]])} < this batch's in-degree {int(W_min[bad[0]])}")
        a0_s = np.concatenate(([0], np.cumsum(A_s)[:-1])) if S else np.zeros(0, np.int64)
        b0_s = np.concatenate(([0], np.cumsum(B_s)[:-1])) if S else np.zeros(0, np.int64)
        ctl = np.empty(6 * nM + 3 * S, np.int32)
        m = ctl[:6 * nM].reshape(6, nM)
        m[0] = ids
        # rows before a molecule: all earlier molecules' rows + one padding row per segment up to and including its own
        m[1] = np.cumsum(nA_m) - nA_m + seg_of + 1
        m[2] = np.cumsum(nB_m) - nB_m + seg_of + 1
        m[3], m[4], m[5] = W_s[seg_of], b0_s[seg_of], a0_s[seg_of]
        seg = ctl[6 * nM:].reshape(3, S)
        seg[0], seg[1], seg[2] = a0_s, b0_s, W_s
        wmax = max(1, int(deg_s.max())) if S else 1
        return ctl, (int(A_s.sum()), int(B_s.sum()), nM, wmax, S), (A_s, B_s, W_s)

    @staticmethod
    def assemble(batches: Sequence[BatchMolGraph], dev, w_override=None) -> "DeviceGraph":
        """Build the graph ON THE DEVICE from the molecule store: the host sends ids and row offsets only."""
        store = batches[0]._store
        ids = np.concatenate([b._ids for b in batches]) if len(batches) > 1 else batches[0]._ids
        W = [b.max_num_bonds if not (w_override and w_override[i]) else int(w_override[i]) for i, b in enumerate(batches)]
        return DeviceGraph.assemble_ids(store, ids, [b.n_mols for b in batches], dev, W)

    @staticmethod
    def assemble_ids(store: MoleculeStore, ids, lens, dev, W=None) -> "DeviceGraph":
        """``assemble`` without BatchMolGraph objects: store ids of all molecules, segments back to back, and the segment lengths.
        The evaluation pass and the RankNet window hand over hundreds of groups at once; their per-group host objects were the cost."""
        from .. import _lib
        dev = torch.device(dev)
        store.sync(dev)
        ctl, (nA, nB, nM, wmax, S), _ = DeviceGraph.control_block(store, ids, lens, W)
        sections = DeviceGraph._sections(nA, nB, nM, wmax, S)
        offs, total = {}, 0
        for name, nbytes in sections:
            offs[name] = total
            total += _align(max(nbytes, 4))
        # pinned staging + device buffers come from a pool (see _BufferPool); the device control block lives at the end of the blob
        ctl_bytes = _align(ctl.nbytes)
        host = _POOL.take(ctl.nbytes, "pinned", lambda n: torch.empty(n, dtype=torch.uint8, pin_memory=True))
        host[:ctl.nbytes].view(torch.int32).numpy()[:] = ctl
        blob = _POOL.take(total + ctl_bytes, str(dev), lambda n: torch.empty(n, dtype=torch.uint8, device=dev))
        d_ctl = blob[total:total + ctl.nbytes]
        d_ctl.copy_(host[:ctl.nbytes], non_blocking=True)
        copied = torch.cuda.Event()
        copied.record(torch.cuda.current_stream(dev))
        g = DeviceGraph()
        g.host_blob = host
        g.blob = blob
        g._pooled = (blob, str(dev), host, copied)
        g.h2d_bytes = ctl.nbytes
        base = g.blob.data_ptr()
        c = g.c
        c.n_atoms, c.n_bonds, c.n_mols, c.wmax, c.n_segments = nA, nB, nM, wmax, S
        for name, _ in sections:
            setattr(c, name, base + offs[name])
        p0 = d_ctl.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().rr_graph_assemble(ctypes.byref(store.c), nM, p0, p0 + 4 * nM, p0 + 8 * nM, p0 + 12 * nM, p0 + 16 * nM, p0 + 20 * nM,
                                                    S, p0 + 24 * nM, p0 + 24 * nM + 4 * S, p0 + 24 * nM + 8 * S, ctypes.byref(c),
                                                    torch.cuda.current_stream().cuda_stream))
        g._ctl = d_ctl
        g.n_atoms, g.n_bonds, g.n_mols = nA, nB, nM
        g.real_atoms, g.real_bonds = nA - S, nB - S
        g._offs = offs
        return g

    @staticmethod
    def assemble_pair_ids(store: MoleculeStore, r_ids, r_lens, p_ids, p_lens, dev, r_W=None, p_W=None):
        """(reactant DeviceGraph, product DeviceGraph) assembled on the device into ONE blob whose feature arrays are adjacent --
        ``[f_atoms(r) | f_atoms(p)]`` and ``[f_bonds(r) | f_bonds(p)]`` -- so that the model can run its shared-weight encoder once over
        both batches (rr_model.cu: joint graph) without copying a feature row.  Otherwise identical to two ``assemble_ids`` calls."""
        from .. import _lib
        dev = torch.device(dev)
        store.sync(dev)
        # both control blocks are written by rr_batch_build straight into one pinned staging buffer
        n_ctl = [4 * (6 * len(i) + 3 * len(l)) for i, l in ((r_ids, r_lens), (p_ids, p_lens))]
        h_off = [0, _align(n_ctl[0])]
        host = _POOL.take(h_off[1] + n_ctl[1], "pinned", lambda n: torch.empty(n, dtype=torch.uint8, pin_memory=True))
        hv = host.numpy()
        blocks = [DeviceGraph.control_block(store, ids, lens, W, out=hv[h_off[k]:h_off[k] + n_ctl[k]].view(np.int32))
                  for k, (ids, lens, W) in enumerate(((r_ids, r_lens, r_W), (p_ids, p_lens, p_W)))]
        dims = [b[1] for b in blocks]
        offs, total = [{}, {}], 0
        for name, ld in (("f_atoms", FA_LD), ("f_bonds", FB_LD)):          # the two feature pairs first, each pair contiguous
            for k in (0, 1):
                offs[k][name] = total
                total += dims[k][0 if name == "f_atoms" else 1] * ld * 4
            total = _align(total)
        for k in (0, 1):
            nA, nB, nM, wmax, S = dims[k]
            for name, nbytes in DeviceGraph._sections(nA, nB, nM, wmax, S)[2:]:
                offs[k][name] = total
                total += _align(max(nbytes, 4))
        ctl_off = [total, total + h_off[1]]
        total_all = ctl_off[1] + _align(n_ctl[1])
        blob, freed = _POOL.take_device(total_all, str(dev), lambda n: torch.empty(n, dtype=torch.uint8, device=dev))
        main, side = torch.cuda.current_stream(dev), _assembly_stream(dev)
        blob.record_stream(side)
        new_upload = store.h2d_bytes_total != getattr(store, "_asm_seen_bytes", -1)
        if new_upload:                           # the molecule store uploaded molecules (store.sync above, main stream) since the last assembly
            uploaded = torch.cuda.Event()
            uploaded.record(main)
        n_copy = h_off[1] + n_ctl[1]
        out = []
        base = blob.data_ptr()
        with torch.cuda.stream(side):
            if new_upload:
                side.wait_event(uploaded)
                store._asm_seen_bytes = store.h2d_bytes_total
            if freed is not None:
                side.wait_event(freed)           # the blob's previous graphs are read by kernels enqueued earlier on the main stream
            else:
                # a block fresh from the caching allocator may be the recycled memory of a tensor that kernels already queued on the main
                # stream still read (the allocator only orders reuse within ONE stream): this assembly waits for them.  Only the first
                # few batches allocate; after that the pool hands blobs back with their `freed` event.
                side.wait_stream(main)
            blob[ctl_off[0]:ctl_off[0] + n_copy].copy_(host[:n_copy], non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(side)
            shared = _SharedBlob(blob, str(dev), host, copied, main)
            for k in (0, 1):
                nA, nB, nM, wmax, S = dims[k]
                g = DeviceGraph()
                g.blob, g.host_blob, g._shared = blob, host, shared
                g.h2d_bytes = n_ctl[k]
                c = g.c
                c.n_atoms, c.n_bonds, c.n_mols, c.wmax, c.n_segments = nA, nB, nM, wmax, S
                for name, _ in DeviceGraph._sections(nA, nB, nM, wmax, S):
                    setattr(c, name, base + offs[k][name])
                p0 = base + ctl_off[0] + h_off[k]
                with torch.cuda.device(dev):
                    _lib.check(_lib.lib().rr_graph_assemble(ctypes.byref(store.c), nM, p0, p0 + 4 * nM, p0 + 8 * nM, p0 + 12 * nM, p0 + 16 * nM, p0 + 20 * nM,
                                                            S, p0 + 24 * nM, p0 + 24 * nM + 4 * S, p0 + 24 * nM + 8 * S, ctypes.byref(c), side.cuda_stream))
                g.n_atoms, g.n_bonds, g.n_mols = nA, nB, nM
                g.real_atoms, g.real_bonds = nA - S, nB - S
                g._offs = offs[k]
                out.append(g)
            ready = torch.cuda.Event()
            ready.record(side)
        # everything enqueued on the main stream FROM NOW ON sees the assembled graphs; what is already queued there (the running step)
        # is what the assembly overlaps with
        main.wait_event(ready)
        return out[0], out[1]

    @staticmethod
    def pair_from_batches(r_batches: Sequence[BatchMolGraph], p_batches: Sequence[BatchMolGraph], device):
        """(reactant, product) DeviceGraphs of store-backed batches with adjacent feature arrays (``assemble_pair_ids``); host-packed
        batches fall back to two independent graphs (the model then makes the contiguous copy itself)."""
        dev = torch.device(device)
        every = list(r_batches) + list(p_batches)
        if dev.type == "cuda" and all(b._ids is not None and b._store is every[0]._store for b in every):
            cat = lambda bs: np.concatenate([b._ids for b in bs]) if len(bs) > 1 else bs[0]._ids  # noqa: E731
            return DeviceGraph.assemble_pair_ids(every[0]._store, cat(r_batches), [b.n_mols for b in r_batches], cat(p_batches),
                                                 [b.n_mols for b in p_batches], dev, [b.max_num_bonds for b in r_batches],
                                                 [b.max_num_bonds for b in p_batches])
        return DeviceGraph.from_batches(r_batches, dev), DeviceGraph.from_batches(p_batches, dev)

    @staticmethod
    def from_batches(batches: Sequence[BatchMolGraph], device, w_override: Optional[Sequence[Optional[int]]] = None,
                     non_blocking: bool = True) -> "DeviceGraph":
        dev = torch.device(device)
        if dev.type == "cuda" and all(b._ids is not None and b._store is batches[0]._store for b in batches):
            return DeviceGraph.assemble(batches, dev, w_override)
        S = len(batches)
        nA = sum(b.n_atoms for b in batches)
        nB = sum(b.n_bonds for b in batches)
        nM = sum(b.n_mols for b in batches)
        wmax = max(1, max(b._max_deg for b in batches))
        sections = DeviceGraph._sections(nA, nB, nM, wmax, S)
        offs, total = {}, 0
        for name, nbytes in sections:
            offs[name] = total
            total += _align(max(nbytes, 4))
        dev = torch.device(device)
        pin = dev.type == "cuda"
        host = torch.empty(total, dtype=torch.uint8, pin_memory=pin)
        hb = host.numpy()

        def view(name, dtype, shape):
            n = int(np.prod(shape)) * np.dtype(dtype).itemsize
            return hb[offs[name]:offs[name] + n].view(dtype).reshape(shape)

        fa, fb = view("f_atoms", np.float32, (nA, FA_LD)), view("f_bonds", np.float32, (nB, FB_LD))
        meta = view("a_meta", np.int32, (nA, 4))
        a2b, a2br, a2a = (view(k, np.int32, (nA, wmax)) for k in ("a2b", "a2b_rev", "a2a"))
        ms, mz = view("mol_start", np.int32, (nM,)), view("mol_size", np.int32, (nM,))
        pb, pa = view("pad_bonds", np.int32, (S,)), view("pad_atoms", np.int32, (S,))
        a0 = b0 = m0 = 0
        for s, b in enumerate(batches):
            W = b.max_num_bonds if not (w_override and w_override[s]) else int(w_override[s])
            if W < b.max_num_bonds:
                raise ValueError(f"max_num_bonds override {W} < this batch's in-degree {b.max_num_bonds}")
            A, B = b.n_atoms, b.n_bonds
            fa[a0] = 0.0
            fb[b0] = 0.0
            if b._packs:
                np.concatenate([p.f_atoms for p in b._packs], out=fa[a0 + 1:a0 + A])
                np.concatenate([p.f_bonds for p in b._packs], out=fb[b0 + 1:b0 + B])
            deg = b._deg
            meta[a0:a0 + A, 0] = deg
            meta[a0, 0] |= 0x100
            meta[a0:a0 + A, 1] = W - deg
            meta[a0:a0 + A, 2] = b0
            meta[a0:a0 + A, 3] = a0
            wb = b._a2b.shape[1]
            valid = np.arange(wb, dtype=np.int32)[None, :] < deg[:, None]
            a2b[a0:a0 + A, :wb] = np.where(valid, b._a2b + b0, 0)
            a2br[a0:a0 + A, :wb] = np.where(valid, b._b2revb[b._a2b] + b0, 0)
            a2a[a0:a0 + A, :wb] = np.where(valid, b._b2a[b._a2b] + a0, 0)
            if wb < wmax:
                a2b[a0:a0 + A, wb:] = 0
                a2br[a0:a0 + A, wb:] = 0
                a2a[a0:a0 + A, wb:] = 0
            ms[m0:m0 + b.n_mols] = b._a_start + a0
            mz[m0:m0 + b.n_mols] = b._a_size
            pb[s], pa[s] = b0, a0
            a0 += A
            b0 += B
            m0 += b.n_mols
        g = DeviceGraph()
        g.host_blob = host
        g.blob = host.to(dev, non_blocking=non_blocking) if pin else host
        g.h2d_bytes = total
        base = g.blob.data_ptr()
        c = g.c
        c.n_atoms, c.n_bonds, c.n_mols, c.wmax, c.n_segments = nA, nB, nM, wmax, S
        for name, _ in sections:
            setattr(c, name, base + offs[name])
        g.n_atoms, g.n_bonds, g.n_mols = nA, nB, nM
        g.real_atoms, g.real_bonds = nA - S, nB - S
        g._offs = offs
        return g

    # ---- optional reactant de-duplication (rr_model_cfg.r_atom_map) ----------------------------
    @staticmethod
    def dedup_plan(r_batches: Sequence[BatchMolGraph], p_batches: Sequence[BatchMolGraph]):
        """The reference repeats the reactant MolGraph once per candidate (load_reactions.py:574-576).  For store-backed batches this
        returns (unique reactant batches, their max_num_bonds, atom_map) where every segment keeps each distinct reactant once (in order of
        first appearance) and ``atom_map[a]`` is the reactant-encoder row that product atom row ``a`` subtracts (segment padding rows map to
        padding rows), or None when nothing repeats / the batches are not store-backed."""
        if not r_batches or len(r_batches) != len(p_batches) or any(b._ids is None for b in r_batches):
            return None
        uniq, maps, a0_p, a0_u, repeats = [], [], 0, 0, False
        for rb, pb in zip(r_batches, p_batches):
            if rb.n_mols != pb.n_mols or not np.array_equal(rb._a_size, pb._a_size):
                raise ValueError("reactant and product batches must list the same molecules' atom counts (p - r is atom-wise, base_model.py:168)")
            ids = rb._ids
            u_sorted, first, inv = np.unique(ids, return_index=True, return_inverse=True)
            order = np.argsort(first, kind="stable")
            rank = np.empty_like(order)
            rank[order] = np.arange(order.shape[0])
            inv = rank[inv.reshape(-1)]
            repeats |= u_sorted.shape[0] < ids.shape[0]
            ub = BatchMolGraph.from_store(rb._store, u_sorted[order], [rb.smiles_batch[i] for i in first[order].tolist()])
            ub.max_num_bonds = rb.max_num_bonds          # same molecules: same maximum in-degree
            m = np.empty(pb.n_atoms, np.int32)
            m[0] = a0_u                                   # the segment's padding atom
            shift = (ub._a_start[inv].astype(np.int64) + a0_u) - (pb._a_start.astype(np.int64) + a0_p)
            m[1:] = np.arange(1, pb.n_atoms, dtype=np.int64) + a0_p + np.repeat(shift, pb._a_size)
            uniq.append(ub)
            maps.append(m)
            a0_p += pb.n_atoms
            a0_u += ub.n_atoms
        if not repeats:
            return None
        return uniq, [b.max_num_bonds for b in r_batches], np.concatenate(maps)

    @staticmethod
    def from_batches_dedup(r_batches: Sequence[BatchMolGraph], p_batches: Sequence[BatchMolGraph], device):
        """(reactant DeviceGraph, product DeviceGraph).  When reactants repeat, the reactant graph is de-duplicated and carries
        ``atom_map`` (device int32) for ``rr_model_cfg.r_atom_map``; exact in eval mode and at dropout 0 only (the model checks)."""
        dev = torch.device(device)
        plan = DeviceGraph.dedup_plan(r_batches, p_batches)
        if plan is None:
            return DeviceGraph.pair_from_batches(r_batches, p_batches, dev)
        uniq, w, amap = plan
        for b, wb in zip(uniq, w):
            b.max_num_bonds = wb
        rg, pg = DeviceGraph.pair_from_batches(uniq, p_batches, dev)
        host = torch.from_numpy(amap).pin_memory()
        rg.atom_map = host.to(dev, non_blocking=True)
        rg._atom_map_host = host
        rg.h2d_bytes += amap.nbytes
        return rg, pg

    @staticmethod
    def dedup_ids(store: MoleculeStore, r_ids, p_ids, lens):
        """``dedup_plan`` on store ids, vectorised over all segments: returns ``(unique reactant ids, their segment lengths, atom_map)`` or
        None when no reactant repeats inside a segment.  Unique reactants keep the order of first appearance inside their segment."""
        r_ids, p_ids = np.asarray(r_ids, dtype=np.int64), np.asarray(p_ids, dtype=np.int64)
        lens = np.asarray(lens, dtype=np.int64)
        S, nM = lens.shape[0], r_ids.shape[0]
        if nM == 0 or p_ids.shape[0] != nM:
            return None
        nA_m = store.nA[r_ids].astype(np.int64)
        if not np.array_equal(nA_m, store.nA[p_ids].astype(np.int64)):
            raise ValueError("reactant and product batches must list the same molecules' atom counts (p - r is atom-wise, base_model.py:168)")
        seg_of = np.repeat(np.arange(S, dtype=np.int64), lens)
        _, first, inv = np.unique(seg_of * (int(r_ids.max()) + 1) + r_ids, return_index=True, return_inverse=True)
        if first.shape[0] == nM:
            return None
        order = np.argsort(first, kind="stable")            # first appearance; segments are contiguous, so this is segment-major too
        rank = np.empty_like(order)
        rank[order] = np.arange(order.shape[0])
        inv = rank[inv.reshape(-1)]
        keep = first[order]
        u_ids, u_lens = r_ids[keep].astype(np.int32), np.bincount(seg_of[keep], minlength=S)
        # global first rows: all earlier molecules' rows + one padding row per segment so far
        u_nA, u_seg = nA_m[keep], seg_of[keep]
        u_start = np.cumsum(u_nA) - u_nA + u_seg + 1
        p_start = np.cumsum(nA_m) - nA_m + seg_of + 1
        A_p = 1 + np.bincount(seg_of, weights=nA_m, minlength=S).astype(np.int64)
        A_u = 1 + np.bincount(u_seg, weights=u_nA, minlength=S).astype(np.int64)
        a0_p = np.concatenate(([0], np.cumsum(A_p)[:-1]))
        a0_u = np.concatenate(([0], np.cumsum(A_u)[:-1]))
        amap = np.empty(int(A_p.sum()), np.int32)
        is_pad = np.zeros(amap.shape[0], bool)
        is_pad[a0_p] = True
        amap[a0_p] = a0_u                                   # a segment's padding atom maps to its padding atom
        rows = np.nonzero(~is_pad)[0]                        # molecule rows, in molecule order
        amap[rows] = rows + np.repeat(u_start[inv] - p_start, nA_m)
        return u_ids, u_lens, amap

    @staticmethod
    def from_id_groups(store: MoleculeStore, r_ids, p_ids, lens, device, dedup: bool, r_W=None, p_W=None):
        """(reactant DeviceGraph, product DeviceGraph) of many segments straight from store ids (``Parsing_features.parsing_ids``): what
        ``from_batches`` / ``from_batches_dedup`` build from one BatchMolGraph per segment, without creating those objects."""
        dev = torch.device(device)
        plan = DeviceGraph.dedup_ids(store, r_ids, p_ids, lens) if dedup else None
        if plan is None:
            return DeviceGraph.assemble_pair_ids(store, r_ids, lens, p_ids, lens, dev, r_W, p_W)
        u_ids, u_lens, amap = plan
        # the unique reactants of a segment keep the segment's max_num_bonds (same molecules: same largest in-degree)
        rg, pg = DeviceGraph.assemble_pair_ids(store, u_ids, u_lens, p_ids, lens, dev, r_W, p_W)
        host = torch.from_numpy(amap).pin_memory()
        rg.atom_map = host.to(dev, non_blocking=True)
        rg._atom_map_host = host
        rg.h2d_bytes += amap.nbytes
        return rg, pg

    def section(self, name: str, dtype: torch.dtype, shape) -> torch.Tensor:
        """Typed view of one section of the device blob (tests / kernels' unit harness)."""
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        return self.blob[self._offs[name]:self._offs[name] + n].view(dtype).reshape(shape)
