"""Model packages and repositories -- drop-in for the loading side of reference ``demucs/states.py:50-102``,
``demucs/repo.py:26-160`` and ``demucs/pretrained.py:62-85``.

A reference package (``*.th``) is a ``torch.save``d dict ``{klass, args, kwargs, state, training_args}`` whose
``klass`` pickles as a reference class (``demucs.htdemucs.HTDemucs``).  ``load_model`` reads it WITHOUT the reference
package installed: a restricted unpickler maps the class reference onto ``demucs_b200.HTDemucs`` (and refuses any
other global that is not a tensor-rebuild helper), builds the model from the recorded ``kwargs`` and loads the
(fp16 or fp32) state.  Local repositories (a folder of ``<sig>-<sha256[:8]>.th`` files and bag ``*.yaml`` files) work
as in the reference, checksum check included.  The remote zoo needs the network: a signature is served from the torch
hub cache when the file is already there, otherwise ``ModelLoadingError`` says so.  DiffQ-quantised packages
(``state['__quantized']``, the ``*_q`` bags) need the absent ``diffq`` package and are rejected loudly.
"""
from __future__ import annotations

import hashlib
import inspect
import io
import pickle
import typing as tp
import warnings
from pathlib import Path

import torch
import yaml

from .apply import BagOfModels
from .config import UnsupportedConfig
from .hdemucs import HDemucs, HDemucsConfig
from .hdemucs import _PINNED as _H_PINNED, _IGNORED as _H_IGNORED
from .htdemucs import HTDemucs

AnyModel = tp.Union[HTDemucs, HDemucs, BagOfModels]
ROOT_URL = "https://dl.fbaipublicfiles.com/demucs/"
DEFAULT_MODEL = "htdemucs"
# remote/files.txt + remote/*.yaml of the reference, for the models this engine can run (HTDemucs family)
REMOTE_FILES = {"75fc33f5": "hybrid_transformer/75fc33f5-1941ce65.th", "955717e8": "hybrid_transformer/955717e8-8726e21a.th", "f7e0c4bc": "hybrid_transformer/f7e0c4bc-ba3fe64a.th",
                "d12395a8": "hybrid_transformer/d12395a8-e57c48e6.th", "92cfc3b6": "hybrid_transformer/92cfc3b6-ef3bcb9c.th",
                "04573f0d": "hybrid_transformer/04573f0d-f3cf25b2.th", "5c90dfd2": "hybrid_transformer/5c90dfd2-34c22ccb.th"}
REMOTE_BAGS = {"hdemucs_mmi": {"models": ["75fc33f5"]}, "htdemucs": {"models": ["955717e8"]},
               "htdemucs_ft": {"models": ["f7e0c4bc", "d12395a8", "92cfc3b6", "04573f0d"],
                               "weights": [[1., 0., 0., 0.], [0., 1., 0., 0.], [0., 0., 1., 0.], [0., 0., 0., 1.]]},
               "htdemucs_6s": {"models": ["5c90dfd2"]}}


class ModelLoadingError(RuntimeError):
    pass


class _UnsupportedKlass:
    """Stand-in for a reference class this engine does not run (HDemucs v3 / Demucs v1-v2 members of the mdx bags)."""

    def __init__(self, name):
        self.name = name


_SAFE_GLOBALS = {
    ("collections", "OrderedDict"), ("fractions", "Fraction"), ("builtins", "set"), ("builtins", "frozenset"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_parameter"), ("torch", "FloatStorage"),
    ("torch", "HalfStorage"), ("torch", "BFloat16Storage"), ("torch", "LongStorage"), ("torch", "IntStorage"),
    ("torch", "DoubleStorage"), ("torch", "ByteStorage"), ("torch", "BoolStorage"), ("torch", "Size"),
    ("torch.serialization", "_get_layout"), ("torch", "device"), ("numpy.core.multiarray", "scalar"),
    ("numpy", "dtype"), ("omegaconf.dictconfig", "DictConfig"),
}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module in ("demucs.htdemucs", "demucs_b200.htdemucs") and name == "HTDemucs":
            return HTDemucs
        if module in ("demucs.hdemucs", "demucs_b200.hdemucs") and name == "HDemucs":
            return HDemucs
        if module.startswith("demucs.") or module.startswith("demucs_b200."):
            return _UnsupportedKlass(f"{module}.{name}")
        if (module, name) in _SAFE_GLOBALS:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"model package refers to {module}.{name}, which a model package has no use for")


class _PickleModule:
    """The ``pickle_module`` handed to ``torch.load`` (restricted unpickler; everything else is the stdlib's)."""
    Unpickler = _Unpickler
    load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())
    loads = staticmethod(lambda b, **kw: _Unpickler(io.BytesIO(b), **kw).load())
    __name__ = "pickle"


def load_model(path_or_package, strict: bool = False, mode: str = "strict") -> tp.Union[HTDemucs, HDemucs]:
    """states.py:50-80 -- a model from a serialized package (a dict, or a path to a ``.th`` file)."""
    if isinstance(path_or_package, dict):
        package = path_or_package
    elif isinstance(path_or_package, (str, Path)):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            package = torch.load(path_or_package, "cpu", pickle_module=_PickleModule, weights_only=False)
    else:
        raise ValueError(f"Invalid type for {path_or_package}.")
    klass, args, kwargs = package["klass"], package["args"], dict(package["kwargs"])
    if isinstance(klass, _UnsupportedKlass) or not (isinstance(klass, type) and issubclass(klass, (HTDemucs, HDemucs))):
        name = getattr(klass, "name", getattr(klass, "__name__", str(klass)))
        raise ModelLoadingError(f"{name} packages are outside the accelerated path: demucs_b200 runs HTDemucs (v4) and "
                                "HDemucs (v3, hdemucs_mmi family) models")
    if args:
        kwargs["sources"] = args[0]
    state = package["state"]
    if state.get("__quantized"):
        raise ModelLoadingError("DiffQ-quantised packages (the *_q models) need the `diffq` package, which is not available")
    if not strict:   # states.py:70-75: drop what the constructor does not know
        if issubclass(klass, HDemucs):
            known = set(_H_PINNED) | set(_H_IGNORED) | set(HDemucsConfig.__dataclass_fields__)
        else:
            known = set(inspect.signature(_reference_kwargs_probe).parameters)
        for key in list(kwargs):
            if key not in known:
                warnings.warn("Dropping inexistant parameter " + key)
                del kwargs[key]
    sources = kwargs.pop("sources")
    try:
        model = klass(sources, mode=mode, **kwargs)
    except UnsupportedConfig as err:
        raise ModelLoadingError(str(err)) from err
    model.load_state_dict(state)     # fp16 packages are widened to the fp32 parameters here
    return model


def _reference_kwargs_probe(sources, audio_channels=2, channels=48, channels_time=None, growth=2, nfft=4096, wiener_iters=0,
                            end_iters=0, wiener_residual=False, cac=True, depth=4, rewrite=True, multi_freqs=None,
                            multi_freqs_depth=3, freq_emb=0.2, emb_scale=10, emb_smooth=True, kernel_size=8, time_stride=2,
                            stride=4, context=1, context_enc=0, norm_starts=4, norm_groups=4, dconv_mode=1, dconv_depth=2,
                            dconv_comp=8, dconv_init=1e-3, bottom_channels=0, t_layers=5, t_emb="sin", t_hidden_scale=4.0,
                            t_heads=8, t_dropout=0.0, t_max_positions=10000, t_norm_in=True, t_norm_in_group=False,
                            t_group_norm=False, t_norm_first=True, t_norm_out=True, t_max_period=10000.0, t_weight_decay=0.0,
                            t_lr=None, t_layer_scale=True, t_gelu=True, t_weight_pos_embed=1.0, t_sin_random_shift=0,
                            t_cape_mean_normalize=True, t_cape_augment=True, t_cape_glob_loc_scale=[5000.0, 1.0, 1.4],
                            t_sparse_self_attn=False, t_sparse_cross_attn=False, t_mask_type="diag", t_mask_random_seed=42,
                            t_sparse_attn_window=500, t_global_window=100, t_sparsity=0.95, t_auto_sparsity=False,
                            t_cross_first=False, rescale=0.1, samplerate=44100, segment=10, use_train_segment=True):
    """The reference constructor's signature (htdemucs.py:56-135): what ``load_model(strict=False)`` keeps."""


def serialize_model(model: tp.Union[HTDemucs, HDemucs], half: bool = True) -> dict:
    """states.py:118-130 -- the package of a model (fp16 state by default, as the released files)."""
    args, kwargs = model._init_args_kwargs
    dtype = torch.half if half else None
    state = {k: p.data.to(device="cpu", dtype=dtype) for k, p in model.state_dict().items()}
    return {"klass": type(model), "args": args, "kwargs": kwargs, "state": state, "training_args": {}}


def save_with_checksum(content, path: Path) -> Path:
    """states.py:106-115 -- ``<stem>-<sha256[:8]><suffix>`` next to ``path``."""
    buf = io.BytesIO()
    torch.save(content, buf)
    sig = hashlib.sha256(buf.getvalue()).hexdigest()[:8]
    path = Path(path)
    path = path.parent / (path.stem + "-" + sig + path.suffix)
    path.write_bytes(buf.getvalue())
    return path


def check_checksum(path: Path, checksum: str) -> None:
    """repo.py:26-39."""
    sha = hashlib.sha256()
    with open(path, "rb") as file:
        while True:
            buf = file.read(2 ** 20)
            if not buf:
                break
            sha.update(buf)
    actual = sha.hexdigest()[:len(checksum)]
    if actual != checksum:
        raise ModelLoadingError(f"Invalid checksum for file {path}, expected {checksum} but got {actual}")


class ModelOnlyRepo:
    def has_model(self, sig: str) -> bool:
        raise NotImplementedError()

    def get_model(self, sig: str) -> HTDemucs:
        raise NotImplementedError()

    def list_model(self) -> tp.Dict[str, tp.Union[str, Path]]:
        raise NotImplementedError()


class RemoteRepo(ModelOnlyRepo):
    """repo.py:55-73 without the download: serves what ``torch.hub`` has already cached."""

    def __init__(self, models: tp.Optional[tp.Dict[str, str]] = None, mode: str = "strict"):
        self._models = dict(REMOTE_FILES if models is None else models)
        self.mode = mode

    def has_model(self, sig: str) -> bool:
        return sig in self._models

    def get_model(self, sig: str) -> HTDemucs:
        try:
            rel = self._models[sig]
        except KeyError:
            raise ModelLoadingError(f"Could not find a pre-trained model with signature {sig}.")
        cached = Path(torch.hub.get_dir()) / "checkpoints" / Path(rel).name
        if not cached.exists():
            raise ModelLoadingError(f"{ROOT_URL + rel} is not in the torch hub cache ({cached}) and this build does not "
                                    "download: fetch the file there, or pass `repo=` with a local model folder")
        check_checksum(cached, cached.stem.split("-")[1])
        return load_model(cached, mode=self.mode)

    def list_model(self):
        return {k: ROOT_URL + v for k, v in self._models.items()}


class LocalRepo(ModelOnlyRepo):
    """repo.py:76-110."""

    def __init__(self, root: Path, mode: str = "strict"):
        self.root = Path(root)
        self.mode = mode
        self.scan()

    def scan(self):
        self._models, self._checksums = {}, {}
        for file in self.root.iterdir():
            if file.suffix == ".th":
                if "-" in file.stem:
                    xp_sig, checksum = file.stem.split("-")
                    self._checksums[xp_sig] = checksum
                else:
                    xp_sig = file.stem
                if xp_sig in self._models:
                    raise ModelLoadingError(f"Duplicate pre-trained model exist for signature {xp_sig}. "
                                            "Please delete all but one.")
                self._models[xp_sig] = file

    def has_model(self, sig: str) -> bool:
        return sig in self._models

    def get_model(self, sig: str) -> HTDemucs:
        try:
            file = self._models[sig]
        except KeyError:
            raise ModelLoadingError(f"Could not find pre-trained model with signature {sig}.")
        if sig in self._checksums:
            check_checksum(file, self._checksums[sig])
        return load_model(file, mode=self.mode)

    def list_model(self):
        return self._models


class BagOnlyRepo:
    """repo.py:113-144: YAML files ``{models: [sig...], weights: [[...]...], segment: s}``."""

    def __init__(self, root: tp.Optional[Path], model_repo: ModelOnlyRepo):
        self.root = None if root is None else Path(root)
        self.model_repo = model_repo
        self.scan()

    def scan(self):
        self._bags: tp.Dict[str, tp.Any] = {}
        if self.root is None:
            self._bags.update(REMOTE_BAGS)
            return
        for file in self.root.iterdir():
            if file.suffix == ".yaml":
                self._bags[file.stem] = file

    def has_model(self, name: str) -> bool:
        return name in self._bags

    def get_model(self, name: str) -> BagOfModels:
        try:
            entry = self._bags[name]
        except KeyError:
            raise ModelLoadingError(f"{name} is neither a single pre-trained model or a bag of models.")
        bag = entry if isinstance(entry, dict) else yaml.safe_load(open(entry))
        models = [self.model_repo.get_model(sig) for sig in bag["models"]]
        return BagOfModels(models, bag.get("weights"), bag.get("segment"))

    def list_model(self):
        return self._bags


class AnyModelRepo:
    """repo.py:147-160."""

    def __init__(self, model_repo: ModelOnlyRepo, bag_repo: BagOnlyRepo):
        self.model_repo, self.bag_repo = model_repo, bag_repo

    def has_model(self, name_or_sig: str) -> bool:
        return self.model_repo.has_model(name_or_sig) or self.bag_repo.has_model(name_or_sig)

    def get_model(self, name_or_sig: str) -> AnyModel:
        if self.model_repo.has_model(name_or_sig):
            return self.model_repo.get_model(name_or_sig)
        return self.bag_repo.get_model(name_or_sig)

    def list_model(self):
        models = dict(self.model_repo.list_model())
        models.update(self.bag_repo.list_model())
        return models


def get_model(name: str, repo: tp.Optional[Path] = None, mode: str = "strict") -> AnyModel:
    """pretrained.py:62-85 -- a bag name or a signature, from a local folder (``repo``) or the (cached) remote zoo."""
    if repo is None:
        model_repo: ModelOnlyRepo = RemoteRepo(mode=mode)
        bag_repo = BagOnlyRepo(None, model_repo)
    else:
        repo = Path(repo)
        if not repo.is_dir():
            raise ModelLoadingError(f"{repo} must exist and be a directory.")
        model_repo = LocalRepo(repo, mode=mode)
        bag_repo = BagOnlyRepo(repo, model_repo)
    model = AnyModelRepo(model_repo, bag_repo).get_model(name)
    model.eval()
    return model
