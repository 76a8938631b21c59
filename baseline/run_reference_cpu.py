"""Run an UNMODIFIED reference script from baseline/_ref on the CPU (fixture generation / reference-arm harness).

    python baseline/run_reference_cpu.py scripts/test_usot.py --arch USOT --resume ckpt.pth --dataset OTB_SYNTH

TEST / BASELINE INFRASTRUCTURE, never on the product path.  The reference model needs a GPU only for two things, both replaced by
harness-side stand-ins that do not touch the arithmetic under test (same as oracle/gen_golden.py):
  * ``.cuda()`` calls (lib/models/models.py:121-122, scripts/test_usot.py:141, lib/tracker/usot_tracker.py:68-71 ...) become no-ops;
  * the GPU-only PrRoIPool extension (lib/models/prroi_pool/functional.py:20-38; its JIT build needs THC headers that no longer
    exist) is replaced by the oracle's numpy restatement, itself pinned on the GPU box against the reference .cu compiled unchanged.
`easydict`, `shapely` and `imgaug` are not installed in this image: baseline/stubs provides the few symbols the script imports.
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def import_reference(name):
    """Import ``lib.<...>`` from the staged reference tree even when this repository's shadow ``lib/`` is on sys.path (both are
    portions of the same PEP-420 namespace package; the first sys.path entry wins per sub-module)."""
    import importlib
    for p in (os.path.join(HERE, "stubs"), REF):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for k in [k for k in sys.modules if k.startswith("lib.")]:
        f = getattr(sys.modules[k], "__file__", None)   # (namespace portions have no __file__; their __path__ follows sys.path)
        if f and not f.startswith(REF):
            del sys.modules[k]                          # a shadow module of this repository imported earlier in this process
    importlib.invalidate_caches()
    mod = importlib.import_module(name)
    assert (mod.__file__ or "").startswith(REF), f"{name} resolved to {mod.__file__}, not to the staged reference"
    return mod


def install(with_prroi=True):
    """Apply the harness-side stand-ins; returns a callable that restores the patched torch attributes."""
    import torch
    saved = {(torch.Tensor, "cuda"): torch.Tensor.cuda, (torch.nn.Module, "cuda"): torch.nn.Module.cuda,
             (torch.UntypedStorage, "cuda"): torch.UntypedStorage.cuda, (torch.cuda, "current_device"): torch.cuda.current_device,
             (torch.cuda, "set_device"): torch.cuda.set_device}
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.UntypedStorage.cuda = lambda self, *a, **k: self   # load_pretrain maps every storage to the GPU (train_utils.py:97-98)
    torch.cuda.current_device = lambda: 0
    torch.cuda.set_device = lambda *a, **k: None
    ref_models = import_reference("lib.models.models")
    if with_prroi:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import usot_oracle as O
        ref_prroi_mod = import_reference("lib.models.prroi_pool.prroi_pool")
        fn = lambda f, r, ph, pw, s: O.prroi_pool2d(f, r, ph, pw, s)
        ref_models.prroi_pool2d = fn
        ref_prroi_mod.prroi_pool2d = fn

    def restore():
        for (obj, name), val in saved.items():
            setattr(obj, name, val)
    return restore


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.join(REF, sys.argv[1])
    install()
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    sys.argv = [script] + sys.argv[2:]
    with torch.no_grad():
        runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
