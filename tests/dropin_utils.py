"""Helpers of the drop-in acceptance test (tests/test_gpu_dropin.py) and of its fixture generator (oracle/gen_dropin_fixture.py):
a two-video synthetic OTB-style dataset inside the staged reference tree, a DataParallel-style checkpoint of the synthetic
weights, and a runner for the reference's UNMODIFIED scripts/test_usot.py."""
import json
import os
import subprocess
import sys

import cv2
import numpy as np
import torch

import tracker_oracle as T
from helpers import GOLD, load_weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
STUBS = os.path.join(ROOT, "baseline", "stubs")
DATASET = "OTB_SYNTH"


def videos():
    """(name, video seed, box) of the two fixture videos: the ones the tracker trace fixture found robust (arg-max and crop
    rounding margins away from their decision boundaries, oracle/gen_tracker_pin.py): a 255-pixel and a 271-pixel search window."""
    g = np.load(os.path.join(GOLD, "tracker_trace.npz"))
    n = int(g["n_frames"])
    return n, [("synth_box", int(g["video_seed"]), tuple(int(v) for v in g["box"])),
               ("synth_small", int(g["small_video_seed"]), tuple(int(v) for v in g["small_box"]))]


def build_dataset(ref_root=REF):
    """Write <ref>/datasets_test/OTB_SYNTH{.json,/} in the layout lib/dataset_loader/benchmark.py:17-26 reads (PNG frames: lossless)."""
    n, vids = videos()
    base = os.path.join(ref_root, "datasets_test", DATASET)
    info = {}
    for name, seed, box in vids:
        frames, pos0, sz0 = T.synthetic_video(seed=seed, n_frames=n, box=box)
        os.makedirs(os.path.join(base, name), exist_ok=True)
        names, gts = [], []
        for t, f in enumerate(frames):
            rel = os.path.join(name, f"{t + 1:04d}.png")
            assert cv2.imwrite(os.path.join(base, rel), f)
            names.append(rel)
            gts.append([120 + 4 * t + 1, 90 + 3 * t + 1, box[0], box[1]])   # 1-based (x, y, w, h); the loader subtracts [1, 1, 0, 0]
        info[name] = {"video_dir": name, "img_names": names, "gt_rect": gts}
    with open(os.path.join(ref_root, "datasets_test", DATASET + ".json"), "w") as f:
        json.dump(info, f)
    return [v[0] for v in vids]


def write_checkpoint(path):
    """{'state_dict': {'module.<key>': tensor}}: what a DataParallel training run saves; load_pretrain strips the prefix
    (lib/utils/train_utils.py:100-105)."""
    sd = load_weights("damp025")
    torch.save({"state_dict": {"module." + k: v for k, v in sd.items()}}, path)


def run_test_usot(cwd, ckpt, host_tracker=False, shadow=True, extra_env=None, timeout=900):
    """Run the unmodified scripts/test_usot.py.  shadow=True: this repository precedes the reference on PYTHONPATH (sm_100a engine);
    shadow=False: the CPU reference through baseline/run_reference_cpu.py.  Returns {video: (n_frames, 4) array} parsed from the
    result files the script wrote under <cwd>/var/result/."""
    env = dict(os.environ)
    if shadow:
        env["PYTHONPATH"] = os.pathsep.join([ROOT, STUBS, REF, env.get("PYTHONPATH", "")])
        env["USOT_B200_HOST_TRACKER"] = "1" if host_tracker else "0"
        cmd = [sys.executable, os.path.join(REF, "scripts", "test_usot.py")]
    else:
        cmd = [sys.executable, os.path.join(ROOT, "baseline", "run_reference_cpu.py"), "scripts/test_usot.py"]
    env.update(extra_env or {})
    cmd += ["--arch", "USOT", "--resume", ckpt, "--dataset", DATASET]
    r = subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"test_usot.py failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}")
    out = {}
    rdir = os.path.join(cwd, "var", "result", DATASET, "USOT")
    for fn in sorted(os.listdir(rdir)):
        out[fn[:-4]] = np.loadtxt(os.path.join(rdir, fn), delimiter=",", ndmin=2)
    return out, r.stdout
