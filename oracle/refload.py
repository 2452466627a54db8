"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Loader for the unmodified reference (read-only at /root/reference, build container
only).  Four imports the CaC HTDemucs path never reaches are stubbed:
openunmix (Wiener filter), julius (v1/v2 resampling), dora.log / omegaconf (training
and serialisation).  Used by oracle/make_golden.py, the oracle-pinning test and
bench.py's ``--impl reference`` arm when the reference tree is available.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DEMUCS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "demucs"))


def load():
    """Return the reference's ``demucs`` package modules (htdemucs, apply)."""
    if not available():
        raise FileNotFoundError(f"reference tree not found at {REFERENCE_ROOT}")
    for name in ("openunmix", "openunmix.filtering", "julius", "dora", "dora.log", "omegaconf"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["openunmix.filtering"].wiener = None
    sys.modules["dora.log"].fatal = lambda *a, **k: (_ for _ in ()).throw(RuntimeError(*a))
    sys.modules["dora.log"].bold = lambda s: s
    sys.modules["omegaconf"].OmegaConf = object
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import demucs.htdemucs as htdemucs  # noqa
    import demucs.apply as apply  # noqa
    return types.SimpleNamespace(htdemucs=htdemucs, apply=apply,
                                 HTDemucs=htdemucs.HTDemucs, apply_model=apply.apply_model,
                                 BagOfModels=apply.BagOfModels)


def build_reference_model(cfg, state):
    """Instantiate the reference HTDemucs for ``cfg`` and load ``state`` into it."""
    ref = load()
    model = ref.HTDemucs(**cfg.reference_kwargs()).eval()
    model.load_state_dict({k: v.float() for k, v in state.items()}, strict=True)
    return model
