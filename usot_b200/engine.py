"""Python handle on the C-ABI engine (usot_engine_* in include/usot_b200.h).

All feature maps that cross this boundary are contiguous NHWC fp32 CUDA tensors; images, boxes and score maps are
NCHW like the reference.  torch only provides device memory and the current stream.
"""
import ctypes
import os

import torch

from . import _lib
from .ops import _stream

C_FEAT = 256  # neck / head width (lib/models/models.py:305-306)


def feature_size(image_size):
    """Spatial size of the stride-8 feature map for a square crop (255 -> 31, 271 -> 33, 127 -> 15)."""
    return _lib.load().usot_feature_size(int(image_size))


class Engine:
    def __init__(self, device, precision="fp16x3"):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("usot_b200.Engine needs a CUDA device; there is no CPU fallback")
        self.device = device if device.index is not None else torch.device("cuda", torch.cuda.current_device())
        self.precision = precision
        self._h = ctypes.c_void_p()
        lib = _lib.load()
        _lib.check(lib.usot_engine_create(ctypes.byref(self._h), self.device.index, _lib.PRECISIONS[precision]))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().usot_engine_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict, cache_dir=None):
        """Stage every float tensor of a reference-layout state_dict and (re)pack.  Synchronises the device.

        ``cache_dir`` (default: $USOT_B200_WEIGHT_CACHE, unset = no cache): directory of packed-weight images keyed by
        sha256(state_dict) + precision + ABI version.  A hit restores the engine with one read + upload instead of the
        host-side BN folding / fp16 splitting; a miss packs as usual and writes the image (atomically) for the next start."""
        lib = _lib.load()
        cache_dir = cache_dir if cache_dir is not None else os.environ.get("USOT_B200_WEIGHT_CACHE")
        path = None
        if cache_dir:
            from .checkpoint import state_dict_hash
            key = "{}_{}_abi{}.usotw".format(state_dict_hash(state_dict)[:32], self.precision, lib.usot_abi_version())
            path = os.path.join(cache_dir, key)
            if os.path.exists(path):
                try:
                    self.import_packed(path)
                    return
                except RuntimeError:  # stale / truncated image: fall through to a fresh pack and overwrite it
                    pass
        for k, v in state_dict.items():
            if not torch.is_tensor(v) or not v.dtype.is_floating_point:
                continue  # num_batches_tracked
            h = v.detach().to("cpu", torch.float32).contiguous()
            _lib.check(lib.usot_engine_load_tensor(self._h, k.encode(), _lib.ptr(h), h.numel()))
        with torch.cuda.device(self.device):
            _lib.check(lib.usot_engine_finalize(self._h))
        if path:
            os.makedirs(cache_dir, exist_ok=True)
            self.export_packed(path)

    def export_packed(self, path):
        """Write the packed-weight image of this (finalized) engine to ``path`` (atomic rename)."""
        lib = _lib.load()
        n = int(lib.usot_engine_packed_size(self._h))
        buf = torch.empty(n, dtype=torch.uint8)
        with torch.cuda.device(self.device):
            _lib.check(lib.usot_engine_export_packed(self._h, _lib.ptr(buf), n))
        tmp = "{}.tmp{}".format(path, os.getpid())
        buf.numpy().tofile(tmp)
        os.replace(tmp, path)
        return n

    def import_packed(self, path):
        """Restore the weights from an image written by export_packed (same precision / ABI version; otherwise RuntimeError)."""
        import numpy as np
        buf = torch.from_numpy(np.fromfile(path, dtype=np.uint8))
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().usot_engine_import_packed(self._h, _lib.ptr(buf), buf.numel()))

    def device_bytes(self):
        return int(_lib.load().usot_engine_device_bytes(self._h))

    # ---- forward graphs ------------------------------------------------------------------------------
    def _empty(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.device)

    @staticmethod
    def _img(x):
        assert x.dim() == 4 and x.shape[1] == 3 and x.shape[2] == x.shape[3], "expected (N,3,S,S) crops"
        return x.contiguous().float()

    def backbone_neck(self, x):
        x = self._img(x)
        f = feature_size(x.shape[2])
        xf = self._empty(x.shape[0], f, f, C_FEAT)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().usot_engine_backbone_neck(self._h, _lib.ptr(x), x.shape[0], x.shape[2], _lib.ptr(xf), _stream(x)))
        return xf

    def template(self, z, template_bbox=None, want_x_ori=False):
        z = self._img(z)
        n = z.shape[0]
        f = feature_size(z.shape[2])
        zf = self._empty(n, 7, 7, C_FEAT)
        x_ori = self._empty(n, f, f, C_FEAT) if want_x_ori else None
        bbox = None if template_bbox is None else template_bbox.to(self.device, torch.float32).contiguous()
        if bbox is not None:
            assert tuple(bbox.shape) == (n, 4)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().usot_engine_template(self._h, _lib.ptr(z), n, z.shape[2], _lib.ptr(bbox), _lib.ptr(zf), _lib.ptr(x_ori),
                                                        _stream(z)))
        return zf, x_ori

    def track(self, x, zf, template_mem=None, nq=0, want_xf=True):
        """x (n,3,S,S); zf NHWC (nz,7,7,256); template_mem NHWC (n*nq,7,7,256).  Returns cls, bbox, cls_mem, xf(NHWC)."""
        x = self._img(x)
        n, s = x.shape[0], x.shape[2]
        f = feature_size(s)
        r = f - 6
        assert zf.is_contiguous() and tuple(zf.shape[1:]) == (7, 7, C_FEAT)
        cls, bbox = self._empty(n, 1, r, r), self._empty(n, 4, r, r)
        cls_mem = None
        if template_mem is not None:
            assert nq > 0 and template_mem.is_contiguous() and tuple(template_mem.shape) == (n * nq, 7, 7, C_FEAT)
            cls_mem = self._empty(n, 1, r, r)
        else:
            nq = 0
        xf = self._empty(n, f, f, C_FEAT) if want_xf else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().usot_engine_track(self._h, _lib.ptr(x), n, s, _lib.ptr(zf), zf.shape[0], _lib.ptr(template_mem), nq,
                                                     _lib.ptr(cls), _lib.ptr(bbox), _lib.ptr(cls_mem), _lib.ptr(xf), _stream(x)))
        return cls, bbox, cls_mem, xf

    def extract_memory_feature(self, ori_x=None, xf=None, search_bbox=None):
        bbox = search_bbox.to(self.device, torch.float32).contiguous()
        n = bbox.shape[0]
        out = self._empty(n, 7, 7, C_FEAT)
        with torch.cuda.device(self.device):
            if ori_x is not None:
                ori_x = self._img(ori_x)
                assert ori_x.shape[0] == n
                _lib.check(_lib.load().usot_engine_extract_memory_feature(self._h, _lib.ptr(ori_x), n, ori_x.shape[2], None, 0, _lib.ptr(bbox),
                                                                          _lib.ptr(out), _stream(ori_x)))
            else:
                assert xf.is_contiguous() and xf.shape[0] == n and xf.shape[3] == C_FEAT
                _lib.check(_lib.load().usot_engine_extract_memory_feature(self._h, None, n, 0, _lib.ptr(xf), xf.shape[1], _lib.ptr(bbox),
                                                                          _lib.ptr(out), _stream(xf)))
        return out

    def forward_train_heads(self, zf, xf, xf_mem, m, label, reg_target, reg_weight, search_bbox, cls_ratio, want_aux=False):
        """Head part of USOT_.forward (lib/models/models.py:223-295).  zf (n,7,7,256), xf (n,F,F,256), xf_mem (n*m,F,F,256) NHWC.
        Returns losses (3,) = [cls_loss, cls_memory_loss, reg_loss] on the device (+ backward_map, pool_box if want_aux)."""
        n, f = xf.shape[0], xf.shape[1]
        r = f - 6
        dev = self.device
        f32 = lambda t: None if t is None else t.to(dev, torch.float32).contiguous()
        label, reg_target, reg_weight, search_bbox = f32(label), f32(reg_target), f32(reg_weight), f32(search_bbox)
        assert zf.is_contiguous() and xf.is_contiguous() and tuple(zf.shape) == (n, 7, 7, C_FEAT)
        assert label.numel() == n * r * r and reg_target.numel() == n * r * r * 4 and reg_weight.numel() == n * r * r
        if m > 0:
            assert xf_mem.is_contiguous() and tuple(xf_mem.shape) == (n * m, f, f, C_FEAT) and tuple(search_bbox.shape) == (n, 4)
        losses = self._empty(3)
        back = self._empty(n, 1, r, r) if want_aux else None
        pbox = self._empty(max(n * m, 1), 4) if want_aux else None
        with torch.cuda.device(dev):
            _lib.check(_lib.load().usot_engine_forward_train(self._h, _lib.ptr(zf), _lib.ptr(xf), _lib.ptr(xf_mem), n, m, f, _lib.ptr(label),
                                                             _lib.ptr(reg_target), _lib.ptr(reg_weight), _lib.ptr(search_bbox),
                                                             float(cls_ratio), _lib.ptr(losses), _lib.ptr(back), _lib.ptr(pbox), _stream(xf)))
        return (losses, back, pbox) if want_aux else losses
