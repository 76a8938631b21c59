"""Host-side mirror of the reference interface: state_dict contract, namespace shadowing, loud CPU failure."""
import os
import subprocess
import sys

import pytest
import torch

import usot_oracle as O
from usot_b200 import USOT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_contract_matches_reference_keys():
    net = USOT()
    sd = net.state_dict()
    ref = O.make_state_dict(0)
    assert len(sd) == 444
    assert set(sd.keys()) == set(ref.keys())
    for k in ref:
        assert tuple(sd[k].shape) == tuple(ref[k].shape), k
    net.load_state_dict(ref, strict=True)


def test_parameter_groups_addressable_like_train_script():
    # scripts/train_usot.py:74-121 walks these attributes
    net = USOT({"mem_size": 3, "pr_pool": True})
    assert net.mem_size == 3 and net.pr_pool is True
    for layer in (net.features.features.layer1, net.features.features.layer2, net.features.features.layer3):
        assert sum(p.numel() for p in layer.parameters()) > 0
    assert sum(p.numel() for p in net.neck.parameters()) == 1024 * 256 + 2 * 256
    assert sum(p.numel() for p in net.connect_model.parameters()) > 15e6


def test_cpu_model_fails_loudly():
    net = USOT()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net.template(torch.zeros(1, 3, 127, 127), torch.tensor([[3.0, 3.0, 11.0, 11.0]]))
    net.zf = torch.zeros(1, 256, 7, 7)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net.track(torch.zeros(1, 3, 255, 255))


def test_ops_reject_cpu_tensors():
    from usot_b200 import ops
    with pytest.raises(NotImplementedError):
        ops.prroi_pool2d(torch.zeros(1, 4, 8, 8), torch.zeros(1, 5), 7, 7, 1.0)
    with pytest.raises(AssertionError):
        ops.prroi_pool2d(torch.zeros(1, 4, 8, 8, dtype=torch.float64), torch.zeros(1, 5, dtype=torch.float64), 7, 7, 1.0)


def test_namespace_shadowing():
    """lib.models.models resolves to this repo when it precedes a reference-like tree on PYTHONPATH."""
    code = ("import lib.models.models as m, lib.tracker.usot_tracker as t, usot_b200, usot_b200.tracker as ut; "
            "assert m.USOT is usot_b200.USOT and t.USOTTracker is ut.USOTTracker and t.USOTConfig is ut.USOTConfig; print('shadow-ok')")
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and "shadow-ok" in r.stdout, r.stderr


@pytest.mark.skipif(not os.path.isdir(os.environ.get("USOT_REFERENCE", "/root/reference")), reason="reference tree not present (GPU box)")
def test_namespace_shadowing_leaves_the_rest_of_the_reference_visible():
    """With this repo BEFORE the reference on PYTHONPATH, lib.models.models / lib.tracker.usot_tracker come from
    here and lib.utils.* still comes from the reference, whose load_pretrain / get_subwindow_tracking signatures we mirror."""
    ref = os.environ.get("USOT_REFERENCE", "/root/reference")
    code = ("import inspect, lib.models.models as m, lib.tracker.usot_tracker as t, lib.utils.track_utils as tu, lib.utils.train_utils as tr, "
            "usot_b200.tracker_ops as ops, usot_b200.checkpoint as ck; "
            f"assert m.__file__.startswith({ROOT!r}) and t.__file__.startswith({ROOT!r}) and tu.__file__.startswith({ref!r}); "
            "a = list(inspect.signature(tu.get_subwindow_tracking).parameters); b = list(inspect.signature(ops.get_subwindow_tracking).parameters); "
            "assert b[:len(a)] == a, (a, b); "
            "a = list(inspect.signature(tr.load_pretrain).parameters); b = list(inspect.signature(ck.load_pretrain).parameters); "
            "assert b[:len(a)] == a, (a, b); print('shadow-ok')")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + ref)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and "shadow-ok" in r.stdout, r.stderr


@pytest.mark.parametrize("k,pad,dil", [(1, (0, 0), (1, 1)), (3, (1, 1), (1, 1)), (3, (2, 2), (2, 2)), (3, (0, 0), (2, 1)), (3, (0, 0), (1, 2)),
                                       (3, (0, 0), (1, 1))])
def test_dgrad_weight_transform_matches_autograd(k, pad, dil):
    """ops.dgrad_weights: conv(grad_out, flipped-transposed filter, padding d*(k-1)-p) equals the autograd input gradient of the
    stride-1 conv (every stride-1 geometry of the network: 1x1, 3x3 p1, dilated p2/d2, the three un-padded encoder dilations)."""
    from usot_b200.ops import dgrad_weights
    g = torch.Generator().manual_seed(k * 7 + dil[0])
    x = torch.randn(2, 8, 13, 11, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(6, 8, k, k, generator=g, dtype=torch.float64)
    y = torch.nn.functional.conv2d(x, w, None, 1, pad, dil)
    go = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(go)
    w_t, pad_t = dgrad_weights(w, pad, dil)
    gx = torch.nn.functional.conv2d(go, w_t, None, 1, pad_t, dil)
    assert gx.shape == x.shape and float((gx - x.grad).abs().max()) <= 1e-10
    with pytest.raises(ValueError):
        dgrad_weights(w, (5, 5), (1, 1))


def test_torch_ops_namespace_is_registered_and_has_no_cpu_kernel():
    """torch.ops.usot_b200.* exists (dispatcher registration, CUDA key only): CPU tensors are rejected by the dispatcher itself."""
    import torch
    import usot_b200.torch_ops as T
    for name in T.OPS:
        assert hasattr(torch.ops.usot_b200, name)
    with pytest.raises(NotImplementedError):
        torch.ops.usot_b200.prroi_pooling_forward(torch.zeros(1, 4, 5, 5), torch.zeros(1, 5), 7, 7, 1.0)
    with pytest.raises(NotImplementedError):
        torch.ops.usot_b200.xcorr_depthwise(torch.zeros(1, 4, 9, 9), torch.zeros(1, 4, 3, 3))


def test_cta_pair_tile_mapping_covers_every_tile_once():
    """Model of conv_tc.cu's pair-tile mapping (tile_of() in conv_tc_kernel, num_pair_tiles in launch_conv_tc): CTA rank r of pair tile t owns the
    ordinary tile of image group 2*gp + r at the same (patch, N block).  Every real tile must be owned exactly once, partners must share patch
    row and N block (the MMA issuer skips the same padded K-steps for both), and only an odd group count may produce phantom tiles."""
    import itertools

    def tile_of(t, rank, n_tiles_n, tiles_w, tiles_h):
        nb, r = t % n_tiles_n, t // n_tiles_n
        sp = tiles_w * tiles_h
        s, gp = r % sp, r // sp
        return ((2 * gp + rank) * sp + s) * n_tiles_n + nb

    for img_tiles, tiles_h, tiles_w, n_tiles_n in itertools.product((2, 3, 5, 8, 13), (1, 25, 31), (1, 2), (1, 2, 8)):
        sp = tiles_w * tiles_h
        num_tiles = img_tiles * sp * n_tiles_n
        num_pair_tiles = ((img_tiles + 1) // 2) * sp * n_tiles_n
        owned, phantom = [], 0
        for t in range(num_pair_tiles):
            a, b = (tile_of(t, r, n_tiles_n, tiles_w, tiles_h) for r in (0, 1))
            assert a % n_tiles_n == b % n_tiles_n                                   # same N block
            assert (a // n_tiles_n) % sp == (b // n_tiles_n) % sp                    # same patch (row and column tile)
            assert (b // n_tiles_n) // sp == (a // n_tiles_n) // sp + 1              # neighbouring image groups
            for x in (a, b):
                if (x // n_tiles_n) // sp >= img_tiles:
                    phantom += 1                                                    # decodes to an image index past the batch: TMA zero-fills / clips it
                else:
                    owned.append(x)
        assert sorted(owned) == list(range(num_tiles))
        assert phantom == (sp * n_tiles_n if img_tiles % 2 else 0)
