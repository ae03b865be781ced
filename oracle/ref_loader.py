"""TEST INFRASTRUCTURE ONLY -- loader for the real reference (``/root/reference``).

The reference imports ``rdkit`` at module top (features/featurization.py:1) and rdkit
is not installed, so stub modules are injected first.  The reference package is
mounted under the private name ``_rr_reference`` so it never collides with this
repo's own ``reactranker`` compatibility package.

``tests/``, ``tests/golden/make_golden.py`` and ``bench.py``'s reference / cpu_baseline legs use this module.  In the build container the
reference is read where it lies (``/root/reference``).  On the GPU box that path does not exist; what travels there is ``oracle/_ref/``,
the reference's byte code compiled by ``oracle/build_ref.py`` (git-ignored compiler output, like the ``.so``), and the loader mounts that.
The live-reference TESTS deliberately stay container-only (``source_available()``): on the box the committed goldens are the pin.
"""
import importlib
import os
import sys
import types

_SOURCE_ROOT = os.environ.get("RR_REFERENCE_ROOT", "/root/reference")
_BUILT_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_PKG = "_rr_reference"


def source_available() -> bool:
    """The reference's source tree is present (the build container)."""
    return os.path.isdir(os.path.join(_SOURCE_ROOT, "reactranker"))


def built_available() -> bool:
    """``oracle/_ref`` (byte code compiled by oracle/build_ref.py) is present."""
    return os.path.isfile(os.path.join(_BUILT_ROOT, "reactranker", "models", "base_model.rrc"))


def available() -> bool:
    return source_available() or built_available()


REFERENCE_ROOT = _SOURCE_ROOT if source_available() else _BUILT_ROOT


def _install_rdkit_stub() -> None:
    if "rdkit" in sys.modules:
        return

    class _Anything:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, name):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

    def mod(name):
        m = types.ModuleType(name)

        def _ga(attr):
            if attr.startswith("__"):
                raise AttributeError(attr)
            return _Anything()
        m.__getattr__ = _ga  # type: ignore[attr-defined]
        sys.modules[name] = m
        return m

    rdkit = mod("rdkit")
    chem = mod("rdkit.Chem")
    rdkit.Chem = chem
    chem.SmilesParserParams = _Anything
    rdchem = mod("rdkit.Chem.rdchem")
    chem.rdchem = rdchem

    class HybridizationType:
        SP, SP2, SP3, SP3D, SP3D2 = range(5)

    class BondType:
        SINGLE, DOUBLE, TRIPLE, AROMATIC = range(4)

    rdchem.HybridizationType = HybridizationType
    rdchem.Atom = _Anything
    rdchem.Bond = _Anything
    chem.BondType = BondType
    chem.Mol = _Anything
    for sub in ("rdkit.Chem.Scaffolds", "rdkit.Chem.Scaffolds.MurckoScaffold", "rdkit.DataStructs",
                "rdkit.Chem.AllChem", "rdkit.Chem.MACCSkeys"):
        mod(sub)
    chem.Scaffolds = sys.modules["rdkit.Chem.Scaffolds"]
    chem.Scaffolds.MurckoScaffold = sys.modules["rdkit.Chem.Scaffolds.MurckoScaffold"]
    chem.AllChem = sys.modules["rdkit.Chem.AllChem"]
    chem.MACCSkeys = sys.modules["rdkit.Chem.MACCSkeys"]
    rdkit.DataStructs = sys.modules["rdkit.DataStructs"]


class _BuiltFinder:
    """Meta-path finder for the byte code under oracle/_ref: ``_rr_reference.a.b`` -> ``oracle/_ref/reactranker/a/b.rrc``."""

    @staticmethod
    def find_spec(fullname, path=None, target=None):
        import importlib.machinery
        import importlib.util
        if not fullname.startswith(_PKG + "."):
            return None
        f = os.path.join(_BUILT_ROOT, "reactranker", *fullname[len(_PKG) + 1:].split(".")) + ".rrc"
        if not os.path.isfile(f):
            return None
        return importlib.util.spec_from_file_location(fullname, f, loader=importlib.machinery.SourcelessFileLoader(fullname, f))


def _mount() -> None:
    if _PKG in sys.modules:
        return
    _install_rdkit_stub()
    if not source_available():
        sys.meta_path.insert(0, _BuiltFinder)
    root = os.path.join(REFERENCE_ROOT, "reactranker")
    pkg = types.ModuleType(_PKG)
    pkg.__path__ = [root]  # namespace-style package
    sys.modules[_PKG] = pkg
    for sub in ("models", "features", "train", "data"):
        m = types.ModuleType(f"{_PKG}.{sub}")
        m.__path__ = [os.path.join(root, sub)]
        sys.modules[f"{_PKG}.{sub}"] = m
        setattr(pkg, sub, m)


class _AliasAsReactranker:
    """Some reference files import ``reactranker.x.y`` absolutely (data/load_reactions.py:7).
    While one of them is being imported, expose the already-mounted ``_rr_reference``
    modules under the ``reactranker`` names too (same module objects, so class identity
    is preserved), then restore whatever was there (this repo's own shim package)."""

    def __enter__(self):
        self.saved = {k: v for k, v in sys.modules.items() if k == "reactranker" or k.startswith("reactranker.")}
        for k in self.saved:
            del sys.modules[k]
        for k, v in list(sys.modules.items()):
            if k == _PKG or k.startswith(_PKG + "."):
                sys.modules["reactranker" + k[len(_PKG):]] = v

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k == "reactranker" or k.startswith("reactranker.")]:
            del sys.modules[k]
        sys.modules.update(self.saved)


def ref(module: str):
    """Import e.g. ``ref('models.base_model')`` from the real reference."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _mount()
    name = f"{_PKG}.{module}"
    if name in sys.modules:
        return sys.modules[name]
    if module == "data.load_reactions":
        ref("features.featurization")
        ref("data.scaffold")
    with _AliasAsReactranker():
        return importlib.import_module(name)


def to_ref_molgraph(mol):
    """Wrap a synthetic molecule (reactranker_b200.synthetic.SynthMol-like: python
    lists ``f_atoms f_bonds a2b b2a b2revb`` + ``smiles n_atoms n_bonds``) as a real
    reference ``MolGraph`` without running RDKit (featurization.py:135-210)."""
    feat = ref("features.featurization")
    g = object.__new__(feat.MolGraph)
    g.smiles = mol.smiles
    g.n_atoms = int(mol.n_atoms)
    g.n_bonds = int(mol.n_bonds)
    g.f_atoms = [list(map(float, r)) for r in mol.f_atoms]
    g.f_bonds = [list(map(float, r)) for r in mol.f_bonds]
    g.a2b = [list(map(int, r)) for r in mol.a2b]
    g.b2a = list(map(int, mol.b2a))
    g.b2revb = list(map(int, mol.b2revb))
    return g


class RefFeaturizer:
    """Duck-typed ``Parsing_features`` (load_reactions.py:540-586) for the REAL
    reference classes, fed from a dict token -> synthetic molecule."""

    def __init__(self, mols_by_token):
        self._src = mols_by_token
        self.smiles2graph = {}

    def parsing_smiles(self, smiles=None):
        if smiles is None:
            return None
        feat = ref("features.featurization")
        graphs = []
        for s in smiles:
            g = self.smiles2graph.get(s)
            if g is None:
                g = to_ref_molgraph(self._src[s])
                self.smiles2graph[s] = g
            graphs.append(g)
        return feat.BatchMolGraph(graphs)

    def parsing_reactions(self, reactions=None):
        if reactions is None:
            return [None, None]
        return [self.parsing_smiles([s[0] for s in reactions]),
                self.parsing_smiles([s[1] for s in reactions])]

    def clear_cache(self):
        self.smiles2graph.clear()
