"""Host-side mirror of the reference model façade (lib/models/models.py:16-306).

``USOT`` keeps the reference's public surface -- constructor ``USOT(settings=None)``, attributes ``pr_pool`` / ``zf``,
sub-module names ``features.features.*`` / ``neck.*`` / ``connect_model.*`` (so ``state_dict()`` has the same 444 keys and
``scripts/train_usot.py:74-121`` can still address parameter groups), and the methods ``template`` / ``track`` /
``extract_memory_feature`` / ``forward`` with the reference's argument meaning -- but the nn.Modules here are parameter
containers only.  All arithmetic runs in the sm_100a engine behind the C ABI (usot_b200/csrc); there is no PyTorch or CPU
fallback: calling any of the methods without a CUDA device / built library raises.
"""
import math
import os
import threading

import torch
import torch.nn as nn

from . import ops
from .engine import Engine, feature_size

_LAYERS = (("layer1", 64, 3, 1), ("layer2", 128, 4, 3), ("layer3", 256, 6, 3))  # name, planes, blocks, shortcut kernel


def _bn(c):
    return nn.BatchNorm2d(c)


class _BlockParams(nn.Module):
    """Parameters of one Bottleneck (lib/models/modules.py:11-35)."""

    def __init__(self, inplanes, planes, shortcut_k):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(inplanes, planes, 1, bias=False), _bn(planes)
        self.conv2, self.bn2 = nn.Conv2d(planes, planes, 3, bias=False), _bn(planes)
        self.conv3, self.bn3 = nn.Conv2d(planes, planes * 4, 1, bias=False), _bn(planes * 4)
        if shortcut_k:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, shortcut_k, bias=False), _bn(planes * 4))


class _ResNetParams(nn.Module):
    """Parameters of ResNet_plus2(Bottleneck, [3,4,6,3], used_layers=[3]) (lib/models/modules.py:61-135)."""

    def __init__(self):
        super().__init__()
        self.conv1, self.bn1 = nn.Conv2d(3, 64, 7, stride=2, bias=False), _bn(64)
        inplanes = 64
        for name, planes, blocks, sk in _LAYERS:
            seq = [_BlockParams(inplanes, planes, sk if name != "layer1" else 1)]
            inplanes = planes * 4
            seq += [_BlockParams(inplanes, planes, 0) for _ in range(1, blocks)]
            setattr(self, name, nn.Sequential(*seq))
        for m in self.modules():  # same init law as modules.py:96-102
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))


class _BackboneParams(nn.Module):  # ResNet50 wrapper, lib/models/backbones.py:12-22
    def __init__(self):
        super().__init__()
        self.features = _ResNetParams()


class _NeckParams(nn.Module):  # AdjustLayer, lib/models/connect.py:284-292
    def __init__(self):
        super().__init__()
        self.downsample = nn.Sequential(nn.Conv2d(1024, 256, 1, bias=False), _bn(256))


def _cbr(bias, dilation=1):
    return nn.Sequential(nn.Conv2d(256, 256, 3, bias=bias, dilation=dilation), _bn(256), nn.ReLU(inplace=True))


class _MatrixParams(nn.Module):  # matrix, lib/models/connect.py:12-53
    def __init__(self):
        super().__init__()
        for nm, dil in (("matrix11", 1), ("matrix12", (2, 1)), ("matrix21", (1, 2))):
            setattr(self, nm + "_k", _cbr(False, dil))
            setattr(self, nm + "_s", _cbr(False, dil))


class _GroupDWParams(nn.Module):  # GroupDW, lib/models/connect.py:77-84
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(3))


class _ConfFusionParams(nn.Module):  # Conf_Fusion, lib/models/connect.py:104-121
    def __init__(self):
        super().__init__()
        self.conf_gen = _cbr(True)
        self.value_gen = _cbr(True)


def _tower():
    layers = []
    for _ in range(4):
        layers += [nn.Conv2d(256, 256, 3, padding=1), _bn(256), nn.ReLU()]
    return nn.Sequential(*layers)


class _HeadParams(nn.Module):  # box_tower_reg(256, 256, tower_num=4), lib/models/connect.py:160-219
    def __init__(self):
        super().__init__()
        self.cls_encode, self.reg_encode = _MatrixParams(), _MatrixParams()
        self.cls_dw, self.reg_dw = _GroupDWParams(), _GroupDWParams()
        self.conf_fusion = _ConfFusionParams()
        self.bbox_tower, self.cls_tower, self.cls_memory_tower = _tower(), _tower(), _tower()
        self.bbox_pred = nn.Conv2d(256, 4, 3, padding=1)
        self.cls_pred = nn.Conv2d(256, 1, 3, padding=1)
        self.cls_memory_pred = nn.Conv2d(256, 1, 3, padding=1)
        self.adjust = nn.Parameter(0.1 * torch.ones(1))
        self.bias = nn.Parameter(torch.ones(1, 4, 1, 1))


class USOT_(nn.Module):
    """Same constructor arguments as the reference USOT_ (lib/models/models.py:17-37)."""

    def __init__(self, mem_size=4, pr_pool=True, search_size=255, score_size=25, maximum_batch=16, sf_size=25, precision=None):
        super().__init__()
        self.features = None
        self.connect_model = None
        self.neck = None
        self.zf = None
        self.search_size = search_size
        self.score_size = score_size
        self.search_feature_size = sf_size
        self.maximum_batch = maximum_batch
        self.mem_size = mem_size
        self.pr_pool = pr_pool
        # dense-conv arithmetic: "fp32" (CUDA-core FMA), "fp16x3" (tcgen05, fp32-equivalent), "fp16" (tcgen05 fast mode)
        self.precision = precision or os.environ.get("USOT_B200_PRECISION", "fp16x3")
        self._engines = {}
        self._engine_keys = {}
        self._lock = threading.Lock()

    # ---- engine plumbing -----------------------------------------------------------------------------
    def _apply(self, fn, *args, **kwargs):
        self._tensors = None  # .cuda() / .to() replace buffer objects and move parameter storage
        return super()._apply(fn, *args, **kwargs)

    def _load_from_state_dict(self, *args, **kwargs):
        self._tensors = None
        return super()._load_from_state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate()
        return out

    def invalidate(self):
        """Force the engines to re-pack the weights on the next call.  Needed after edits that autograd's version counters do
        not see: writes through ``p.data`` (``p.data.copy_()``, ``m.weight.data.normal_()`` -- the reference's own init idiom,
        lib/models/modules.py:99-102) and after re-assigning a Parameter / buffer object of a sub-module."""
        self._tensors = None
        self._engine_keys = {}

    repack = invalidate

    def _weights_key(self):
        """Per-tensor (storage address, version counter) of every parameter / buffer: changes when any of them is modified in place
        through the tensor API (version counters only grow), re-allocated or moved.  Edits through ``.data`` do not bump the
        version counter: call ``invalidate()`` after those.  Called on every forward entry point, so it walks a cached tensor list
        instead of rebuilding state_dict() (1.2 ms -> 0.1 ms)."""
        ts = getattr(self, "_tensors", None)
        if ts is None:
            ts = self._tensors = list(self.state_dict(keep_vars=True).values())
        return tuple((v.data_ptr(), v._version) for v in ts)

    def _engine(self, device=None):
        """The engine of the device the parameters live on, (re)packed if any parameter changed since the last call."""
        p = next(self.parameters())
        if not p.is_cuda:
            raise RuntimeError("usot_b200.USOT runs on CUDA only (call .cuda() first); there is no CPU fallback")
        dev = p.device
        if device is not None and torch.device(device) != dev:
            raise RuntimeError(f"input on {device} but model parameters on {dev}")
        with self._lock:
            eng = self._engines.get(dev.index)
            if eng is None:
                eng = self._engines[dev.index] = Engine(dev, self.precision)
            key = self._weights_key()
            if self._engine_keys.get(dev.index) != key:
                eng.load_state_dict(self.state_dict())
                self._engine_keys[dev.index] = key
        return eng

    # ---- reference API -------------------------------------------------------------------------------
    def template(self, z, template_bbox=None):
        """lib/models/models.py:173-177: caches self.zf (N,256,7,7); returns None."""
        eng = self._engine(z.device)
        if self.pr_pool and template_bbox is None:
            raise ValueError("pr_pool=True needs template_bbox (lib/models/connect.py:306-313)")
        zf, _ = eng.template(z, template_bbox if self.pr_pool else None)
        self.zf = ops.nhwc_view(zf)

    def track(self, x, template_mem=None, score_mem=None):
        """lib/models/models.py:179-198.  Returns (cls, bbox, cls_mem, xf) or (cls, bbox, None, None)."""
        if self.zf is None:
            raise RuntimeError("track() before template()")
        eng = self._engine(x.device)
        zf = ops.as_nhwc(self.zf.to(x.device, torch.float32))
        if template_mem is None:
            cls, bbox, _, _ = eng.track(x, zf, want_xf=False)
            return cls, bbox, None, None
        # only the SHAPE of score_mem enters the math (lib/models/connect.py:258)
        batch, nq = score_mem.shape
        assert batch == x.shape[0], "score_mem batch must match the search batch"
        mem = ops.as_nhwc(template_mem.to(x.device, torch.float32))
        cls, bbox, cls_mem, xf = eng.track(x, zf, mem, nq)
        return cls, bbox, cls_mem, ops.nhwc_view(xf)

    def extract_memory_feature(self, ori_x=None, xf=None, search_bbox=None):
        """lib/models/models.py:200-206.  Returns (N,256,7,7)."""
        ref = ori_x if ori_x is not None else xf
        eng = self._engine(ref.device)
        if ori_x is not None:
            out = eng.extract_memory_feature(ori_x=ori_x, search_bbox=search_bbox)
        else:
            out = eng.extract_memory_feature(xf=ops.as_nhwc(xf.float()), search_bbox=search_bbox)
        return ops.nhwc_view(out)

    def forward(self, template, search, label=None, reg_target=None, reg_weight=None, template_bbox=None, search_memory=None,
                search_bbox=None, cls_ratio=0.40, zf_exchange=None):
        """Training forward of lib/models/models.py:208-295.  Returns the reference's triple (cls_loss, cls_memory_loss or None,
        reg_loss) as 0-d CUDA tensors.

        * With autograd enabled (the reference's training loop, scripts/train_usot.py:196-233) or in ``train()`` mode the call runs the
          TRAINING path (usot_b200/train.py): the losses carry an autograd graph whose backward runs this library's dgrad / wgrad /
          BatchNorm / pooling / correlation / loss gradient kernels, and BatchNorm follows ``self.training`` exactly like the reference
          (batch statistics + running-statistics update in train(), running statistics in eval()).
        * Under ``torch.no_grad()`` in ``eval()`` mode it runs the engine's fused forward-only graph (running statistics).

        ``zf_exchange``: optional callable ``zf_local (n,7,7,256) -> closure returning this rank's rows`` used by usot_b200.dist to
        all-gather the template features across ranks while the search / memory backbones run (forward-only path)."""
        if self.pr_pool and template_bbox is None:
            raise ValueError("pr_pool=True needs template_bbox")
        if self.training or torch.is_grad_enabled():
            if not next(self.parameters()).is_cuda:
                raise RuntimeError("usot_b200.USOT runs on CUDA only (call .cuda() first); there is no CPU fallback")
            from . import train
            return train.forward_train(self, template, search, label, reg_target, reg_weight, template_bbox, search_memory, search_bbox,
                                       cls_ratio)   # every BatchNorm follows its own .training flag, like torch
        eng = self._engine(search.device)
        zf, _ = eng.template(template, template_bbox if self.pr_pool else None)
        pending = zf_exchange(zf) if zf_exchange is not None else None
        xf = eng.backbone_neck(search)
        m, xf_mem = 0, None
        if search_memory is not None:
            batch, m, cx, hx, wx = search_memory.shape
            assert batch == search.shape[0]
            xf_mem = eng.backbone_neck(search_memory.reshape(-1, cx, hx, wx))
        if pending is not None:
            zf = pending()
        losses = eng.forward_train_heads(zf, xf, xf_mem, m, label, reg_target, reg_weight, search_bbox, cls_ratio)
        return losses[0], (losses[1] if m > 0 else None), losses[2]

    def backbone_neck(self, x):
        """feature_extractor + neck (lib/models/models.py:181-184): (N,3,S,S) -> (N,256,F,F)."""
        return ops.nhwc_view(self._engine(x.device).backbone_neck(x))


class USOT(USOT_):
    """lib/models/models.py:298-306."""

    def __init__(self, settings=None, precision=None):
        if settings is None:
            settings = {"mem_size": 4, "pr_pool": True}
        super().__init__(mem_size=settings["mem_size"], pr_pool=settings["pr_pool"], search_size=255, score_size=25,
                         maximum_batch=16, sf_size=25, precision=precision)
        self.features = _BackboneParams()
        self.neck = _NeckParams()
        self.connect_model = _HeadParams()
