"""Checkpoint helpers (reference: nbdt/models/utils.py:7-103): `module.`-prefix repair,
pretrained-weight lookup by (arch, dataset) key, optional-kwarg model factories."""
from pathlib import Path

import torch
from torch.hub import load_state_dict_from_url

from nbdt.utils import Colors


def _strip_module(state_dict):
    return {key.replace("module.", "", 1): value for key, value in state_dict.items()}


def load_state_dict(net, state_dict):
    """net.load_state_dict that tolerates a DataParallel `module.` prefix in the checkpoint."""
    try:
        net.load_state_dict(state_dict)
    except RuntimeError as e:
        if "Missing key(s) in state_dict:" in str(e):
            net.load_state_dict(_strip_module(state_dict))


def make_kwarg_optional(init, **optional_kwargs):
    """Factory wrapper: try `init(**optional_kwargs, **kwargs)`, retry without the optional ones."""

    def f(**kwargs):
        try:
            return init(**optional_kwargs, **kwargs)
        except TypeError as e:  # likely because `dataset` is not an accepted argument
            print(e)
        try:
            return init(**kwargs)
        except Exception as e:
            Colors.red(f"Fatal error: {e}")
            exit()

    return f


def coerce_state_dict(state_dict, reference_state_dict):
    """Make `state_dict` keys match `reference_state_dict` w.r.t. the `module.` prefix; unwrap {"net": ...}."""
    if "net" in state_dict:
        state_dict = state_dict["net"]
    want = next(iter(reference_state_dict)).startswith("module.")
    have = next(iter(state_dict)).startswith("module.")
    if have and not want:
        return _strip_module(state_dict)
    if want and not have:
        return {"module." + key: value for key, value in state_dict.items()}
    return state_dict


def get_model_device(model):
    return next(model.parameters()).device


def load_state_dict_from_key(keys, model_urls, pretrained=False, progress=True,
                             root=".cache/torch/checkpoints", device="cpu"):
    valid_keys = [key for key in keys if key in model_urls]
    if not valid_keys:
        raise UserWarning(f"None of the keys {keys} correspond to a pretrained model.")
    key = valid_keys[-1]
    url = model_urls[key]
    Colors.green(f"Loading pretrained model {key} from {url}")
    return load_state_dict_from_url(
        url, Path.home() / root, progress=progress, check_hash=False, map_location=torch.device(device)
    )


def get_pretrained_model(arch, dataset, model, model_urls, pretrained=False, progress=True,
                         root=".cache/torch/checkpoints"):
    if pretrained:
        state_dict = load_state_dict_from_key(
            [(arch, dataset)], model_urls, pretrained, progress, root, device=get_model_device(model)
        )
        model.load_state_dict(coerce_state_dict(state_dict, model.state_dict()))
    return model
