"""Checkpoint / weight pipeline next to the boundary: the reference's ``load_pretrain`` contract
(lib/utils/train_utils.py:92-156) for the ``usot_b200.USOT`` façade, plus a content hash of a state_dict that keys the packed
weights an engine holds.

    load_pretrain(model, path)   same call as the reference: strips the 'module.' / 'feature_extractor.' prefixes (:100-105,
                                 133-137), unwraps {'state_dict': ...} checkpoints, embeds the MoCo 1x1 shortcut kernels of
                                 layer2.0 / layer3.0 into the centre of the 3x3 shortcut convs this ResNet uses (:109-124),
                                 reports missing / unused keys like ``check_keys`` (:140-156) and loads with strict=False.
    state_dict_hash(sd)          sha256 over (name, shape, fp32 bytes) of every floating tensor, in key order.

Differences from the reference, all on the host side: tensors are read with map_location='cpu' (the reference maps every
storage to the current CUDA device first, :98-99) and copied to wherever the model's parameters live by load_state_dict; the
engine re-packs lazily on the next forward call because the parameter versions changed (usot_b200/models.py:_weights_key).
"""
import hashlib

import torch

MOCO_EMBED_KEYS = ("encoder_q.layer2.0.downsample.0.weight", "encoder_q.layer3.0.downsample.0.weight")


def remove_prefix(state_dict, prefix, verbose=True):
    """Same contract as lib/utils/train_utils.py:130-137: keys that start with ``prefix`` lose it (once), others are kept."""
    if verbose:
        print("remove prefix '{}'".format(prefix))
    n = len(prefix)
    out = {}
    for key, value in state_dict.items():
        out[key[n:] if key[:n] == prefix else key] = value
    return out


def check_keys(model, pretrained_state_dict, print_unuse=True, verbose=True):
    """lib/utils/train_utils.py:140-156.  Returns (missing, unused) in addition to printing what the reference prints."""
    ckpt_keys = set(pretrained_state_dict.keys())
    model_keys = set(model.state_dict().keys())
    used = model_keys & ckpt_keys
    unused = sorted(ckpt_keys - model_keys)
    missing = sorted(k for k in (model_keys - ckpt_keys) if 'num_batches_tracked' not in k)
    if verbose:
        print('missing keys:{}'.format(missing))
        if print_unuse:
            print('unused checkpoint keys:{}'.format(unused))
    assert len(used) > 0, 'load NONE from pretrained checkpoint'
    return missing, unused


def convert_moco(pretrained_dict):
    """MoCo-v2 ResNet-50 -> this backbone (lib/utils/train_utils.py:109-124): only ``encoder_q.*`` tensors are kept, renamed to
    ``features.features.*``; the 1x1 shortcut kernels of layer2.0 / layer3.0 become the centre tap of zero 3x3 kernels."""
    out = {}
    for key, value in pretrained_dict.items():
        if "encoder_q" not in key:
            continue
        new_key = key.replace("encoder_q", "features.features")
        if key in MOCO_EMBED_KEYS:
            core = torch.zeros((value.shape[0], value.shape[1], 3, 3), dtype=torch.float32, device=value.device)
            core[:, :, 1, 1] = value[:, :, 0, 0]
            out[new_key] = core
        else:
            out[new_key] = value
    return out


def prepare_state_dict(checkpoint, is_moco=False, verbose=True):
    """Everything load_pretrain does between torch.load and load_state_dict."""
    if "state_dict" in checkpoint.keys():
        sd = remove_prefix(checkpoint['state_dict'], 'module.', verbose)
    else:
        sd = remove_prefix(checkpoint, 'module.', verbose)
    sd = remove_prefix(sd, 'feature_extractor.', verbose)
    return convert_moco(sd) if is_moco else sd


def load_pretrain(model, pretrained_path, print_unuse=True, gpus=None, verbose=True):
    """Drop-in for lib.utils.train_utils.load_pretrain (same arguments and return value)."""
    if verbose:
        print('load pretrained model from {}'.format(pretrained_path))
    if gpus is not None and torch.cuda.is_available():
        torch.cuda.set_device(gpus[0])
    checkpoint = torch.load(pretrained_path, map_location='cpu')
    sd = prepare_state_dict(checkpoint, is_moco="moco" in str(pretrained_path), verbose=verbose)
    check_keys(model, sd, print_unuse=print_unuse, verbose=verbose)
    model.load_state_dict(sd, strict=False)
    return model


def state_dict_hash(state_dict):
    """Content hash of the floating tensors of a state_dict (what the engine packs); integer buffers such as
    ``num_batches_tracked`` do not enter the forward path and are skipped."""
    h = hashlib.sha256()
    for k in sorted(state_dict.keys()):
        v = state_dict[k]
        if not torch.is_tensor(v) or not v.dtype.is_floating_point:
            continue
        t = v.detach().to("cpu", torch.float32).contiguous()
        h.update(k.encode())
        h.update(str(tuple(t.shape)).encode())
        h.update(t.numpy().tobytes())
    return h.hexdigest()
