"""`install()` registers the B200 mirrors under the reference's import paths (``src.models.*``) so that the reference's
recipes (`recipes/**/setting.py`: ``from src.models.passt.passt_sed import PaSST_SED``) pick them up unchanged."""
import importlib
import sys
import types

_MIRRORS = {
    "src.models.sed_model": "transformer4sed_b200.src_models.sed_model",
    "src.models.pooling": "transformer4sed_b200.src_models.pooling",
    "src.models.transformer_decoder": "transformer4sed_b200.src_models.transformer_decoder",
    "src.models.passt.passt_feature_extraction": "transformer4sed_b200.src_models.passt.passt_feature_extraction",
    "src.models.passt.passt": "transformer4sed_b200.src_models.passt.passt",
    "src.models.passt.passt_sed": "transformer4sed_b200.src_models.passt.passt_sed",
    "src.models.transformer.transformerXL": "transformer4sed_b200.src_models.transformer.transformerXL",
    "src.models.transformer.mask": "transformer4sed_b200.src_models.transformer.mask",
    "src.models.encoder_slide_window": "transformer4sed_b200.src_models.encoder_slide_window",
    "src.models.passt.passt_win": "transformer4sed_b200.src_models.passt.passt_win",
    "src.models.passt.passt_lora": "transformer4sed_b200.src_models.passt.passt_lora",
    "src.models.lora": "transformer4sed_b200.src_models.lora",
    "src.models.lora.layers": "transformer4sed_b200.src_models.lora.layers",
    "src.models.cnn": "transformer4sed_b200.src_models.cnn",
    "src.models.cnn.base": "transformer4sed_b200.src_models.cnn.base",
    "src.models.cnn_transformer.passt_cnn": "transformer4sed_b200.src_models.cnn_transformer.passt_cnn",
    "src.models.detect_any_sound.at_adapter": "transformer4sed_b200.src_models.detect_any_sound.at_adapter",
    "src.models.detect_any_sound.detect_any_sound": "transformer4sed_b200.src_models.detect_any_sound.detect_any_sound",
    "src.postprocess.filter": "transformer4sed_b200.src_postprocess.filter",
    "src.preprocess.scaler": "transformer4sed_b200.src_preprocess.scaler",
    # the three names every recipe imports (`mixup, frame_shift, feature_transformation`: recipes/desed/finetune/train.py:18)
    "src.preprocess.data_aug": "transformer4sed_b200.src_preprocess.data_aug",
}


def install(force=True):
    """Make ``import src.models.<...>`` resolve to the CUDA mirrors.  Packages that already exist (a reference checkout on
    sys.path) are kept for everything that is not mirrored (datasets, codec, utils ...)."""
    for name, target in _MIRRORS.items():
        parts = name.split(".")
        for i in range(1, len(parts)):
            pkg = ".".join(parts[:i])
            if pkg not in sys.modules:
                try:
                    importlib.import_module(pkg)
                except Exception:  # noqa: BLE001
                    m = types.ModuleType(pkg)
                    m.__path__ = []
                    sys.modules[pkg] = m
        if force or name not in sys.modules:
            mod = importlib.import_module(target)
            sys.modules[name] = mod
            setattr(sys.modules[".".join(parts[:-1])], parts[-1], mod)
    return sorted(_MIRRORS)
