"""demucs_b200 -- B200-native (sm_100a) inference path for Demucs v4 / HTDemucs.

Drop-in surface (reference file it mirrors):
    HTDemucs            demucs/htdemucs.py
    apply_model, BagOfModels, TensorChunk, center_trim      demucs/apply.py, demucs/utils.py
    Separator           demucs/api.py
"""
from .config import HTDemucsConfig, UnsupportedConfig, htdemucs_config, htdemucs_6s_config  # noqa
from .weights import init_weights, param_specs, count_params  # noqa
from .htdemucs import HTDemucs, htdemucs  # noqa
from .apply import apply_model, BagOfModels, TensorChunk, tensor_chunk, center_trim  # noqa
from .api import Separator, LoadAudioError, LoadModelError, list_models  # noqa
from .repo import get_model, load_model, ModelLoadingError  # noqa
from .hdemucs import HDemucs, HDemucsConfig, hdemucs_mmi  # noqa
from .streaming import StreamSeparator  # noqa
from ._lib import KernelError  # noqa

__version__ = "0.1.0"
