"""TEST INFRASTRUCTURE ONLY — imports the unmodified reference (taraak/pita) in THIS container.

The reference's hot-path modules import lightning / hydra / bgflow / matplotlib at module
top; none of those are installed here.  This helper registers empty stub
modules for them, puts /root/reference/pita on sys.path and returns the reference modules.
It is used by `oracle/make_golden.py` (fixture generation) and by the optional
`tests/test_oracle_vs_reference.py` (skipped when /root/reference is absent, e.g. on the
GPU box).  Nothing under `pita_b200/` may import this file.

The only arithmetic restated here is bgflow's two geometry helpers, which are NOT under
/root/reference (bgflow is an un-vendored, unpinned git dependency: environment.yaml:56,
`git+https://github.com/atong01/bgflow.git`).  Call sites: lennardjones_energy.py:9-10,125-127.
Published algorithm (bgflow/utils/geometry.py): distance_vectors(x)[b,i,k] = x_i - x_j over
all j != i;  distances_from_vectors(r, eps=1e-6) = sqrt(sum(r^2) + eps).  The structure is
confirmed by the in-repo restatement `energy2` (sampling/sample_lj13.py:24-30); the eps value
is an assumption that cannot be confirmed in-container.
"""
import os
import sys
import types

REF_ROOT = "/root/reference"
REF_PITA = os.path.join(REF_ROOT, "pita")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_PITA, "src"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install_stubs():
    import torch

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    class _Rank:
        rank = 0

        def __call__(self, fn):
            return fn

    rank_zero_only = _Rank()

    def _identity_decorator(*a, **k):
        def deco(fn):
            return fn
        return deco

    # lightning / pytorch_lightning
    _mod("lightning", LightningModule=torch.nn.Module, Callback=_Dummy, Trainer=_Dummy,
         LightningDataModule=_Dummy, seed_everything=lambda *a, **k: None)
    _mod("lightning.pytorch")
    _mod("lightning.pytorch.loggers", WandbLogger=_Dummy, Logger=_Dummy)
    _mod("lightning.pytorch.utilities", rank_zero_only=rank_zero_only)
    _mod("pytorch_lightning")
    _mod("pytorch_lightning.loggers", WandbLogger=_Dummy)
    _mod("pytorch_lightning.utilities")
    _mod("pytorch_lightning.utilities.rank_zero", rank_zero_only=rank_zero_only)
    _mod("lightning_utilities")
    _mod("lightning_utilities.core")
    _mod("lightning_utilities.core.rank_zero", rank_zero_only=rank_zero_only,
         rank_prefixed_message=lambda m, r: m)
    # hydra / omegaconf
    _mod("hydra", main=_identity_decorator)
    _mod("hydra.utils", get_original_cwd=lambda: os.getcwd(), instantiate=None)
    _mod("hydra.core")
    _mod("hydra.core.hydra_config", HydraConfig=_Dummy)
    _mod("omegaconf", DictConfig=dict, OmegaConf=_Dummy, open_dict=_Dummy)
    # plotting
    _mod("matplotlib")
    _mod("matplotlib.pyplot")
    _mod("PIL")
    _mod("rich.prompt", Prompt=_Dummy) if "rich.prompt" not in sys.modules else None

    # bgflow (see module docstring for provenance)
    class Energy(torch.nn.Module):
        def __init__(self, dim):
            super().__init__()
            if isinstance(dim, int):
                dim = [dim]
            self._event_shape = torch.Size(dim)

        @property
        def event_shape(self):
            return self._event_shape

        def energy(self, x):
            return self._energy(x)

    def distance_vectors(x, remove_diagonal=True):
        n = x.shape[1]
        r = x.unsqueeze(2) - x.unsqueeze(1)  # r[b,i,j] = x_i - x_j
        if remove_diagonal:
            mask = ~torch.eye(n, dtype=torch.bool)
            r = r[:, mask].view(-1, n, n - 1, x.shape[2])
        return r

    def distances_from_vectors(r, eps=1e-6):
        return (r.pow(2).sum(dim=-1) + eps).sqrt()

    _mod("bgflow", Energy=Energy)
    _mod("bgflow.utils", distance_vectors=distance_vectors,
         distances_from_vectors=distances_from_vectors)


def import_reference():
    """Returns a namespace of the reference's hot-path modules (unmodified)."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_PITA)
    install_stubs()
    if REF_PITA not in sys.path:
        sys.path.insert(0, REF_PITA)
    # src/utils/__init__.py eagerly imports template utilities needing hydra etc.; pre-register
    # a bare package so `src.utils.data_utils` resolves without executing that __init__.
    import importlib

    pkg = types.ModuleType("src")
    pkg.__path__ = [os.path.join(REF_PITA, "src")]
    sys.modules.setdefault("src", pkg)
    upkg = types.ModuleType("src.utils")
    upkg.__path__ = [os.path.join(REF_PITA, "src", "utils")]
    sys.modules.setdefault("src.utils", upkg)

    ns = types.SimpleNamespace()
    ns.data_utils = importlib.import_module("src.utils.data_utils")
    ns.egnn = importlib.import_module("src.models.components.egnn_temp_conditioned")
    ns.energy_net = importlib.import_module("src.models.components.energy_net")
    ns.score_net = importlib.import_module("src.models.components.score_net")
    ns.noise = importlib.import_module("src.models.components.noise_schedules")
    ns.anneal = importlib.import_module("src.models.components.annealing_factor_schedules")
    ns.utils = importlib.import_module("src.models.components.utils")
    ns.sdes = importlib.import_module("src.models.components.sdes")
    ns.integ = importlib.import_module("src.models.components.sde_integration")
    ns.prior = importlib.import_module("src.energies.base_prior")
    ns.lj = importlib.import_module("src.energies.lennardjones_energy")
    return ns


def import_reference_ad2():
    """The alanine-dipeptide EGNN (egnn_dynamics_ad2_cat.py; SURVEY §8 row a8').  Its module imports mdtraj at the top
    (:3) only to read a topology for the >= 53-atom systems; a bare stub is enough for the 22-atom network."""
    import importlib

    import_reference()
    if "mdtraj" not in sys.modules:
        _mod("mdtraj")
    return importlib.import_module("src.models.components.egnn_dynamics_ad2_cat")
