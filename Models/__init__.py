"""Drop-in `Models` package: the import paths the reference's driver scripts use
(`INFERENCE.py:71`, `TRAIN_FINAL.py:14`, `INFERENCE_TIMER.py:177`).

Checkpoints written by `TRAIN_FINAL.py:391-429` hold, beside the state_dict and the config, a pickled
`Dataset_Preparation.Normalizer.DatasetNormalizer` (sklearn scalers over numpy arrays).  torch >= 2.6 loads with
`weights_only=True` by default and rejects those classes, so importing this package registers them as safe globals
(`torch.serialization.add_safe_globals`): `TRAIN_FINAL.py` imports `Models` at its top (:14) and is covered.
`INFERENCE.py:65` and `INFERENCE_TIMER.py:172` call `torch.load` BEFORE they import `Models` (:71 / :177); for those two
either `import Models` first or set `TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD=1` (a property of the torch version, see
INTEGRATION.md).
"""


def register_checkpoint_globals() -> list:
    """Allow-list what a reference checkpoint's `normalizer` entry unpickles to.  Returns what was registered."""
    import torch
    allowed = []
    try:
        import numpy as np
        from numpy._core import multiarray as _ma
        allowed += [_ma._reconstruct, _ma.scalar, np.ndarray, np.dtype]
        allowed += [type(np.dtype(t)) for t in (np.float64, np.float32, np.int64, np.int32, np.bool_)]
    except Exception:                                             # pragma: no cover - numpy layout differs
        pass
    try:
        from sklearn.preprocessing import MinMaxScaler, RobustScaler, StandardScaler
        allowed += [RobustScaler, StandardScaler, MinMaxScaler]
    except Exception:                                             # pragma: no cover - sklearn absent
        pass
    try:                                                          # present when run from the reference's tree
        from Dataset_Preparation.Normalizer import DatasetNormalizer
        allowed.append(DatasetNormalizer)
    except Exception:
        pass
    if hasattr(torch.serialization, "add_safe_globals"):
        torch.serialization.add_safe_globals(allowed)
    return allowed


register_checkpoint_globals()
