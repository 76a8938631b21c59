"""Stage the UNMODIFIED reference tree into git-ignored baseline/_ref/ (the base contract's place for the reference arm).

    python baseline/stage_reference.py [--reference /root/reference]

The reference (VISION-SJTU/USOT) is a script tree with no setup.py / pyproject.toml, so there is nothing for pip to install: the
"install" is a verbatim copy of its Python packages, scripts and test configuration.  Only ``lib/``, ``scripts/`` and
``experiments/`` are copied (preprocessing/ is the offline ARFlow data pipeline, out of scope).  Every file is copied byte for byte
and its sha256 is recorded in baseline/_ref/MANIFEST.json so that a test can prove the staged tree is the reference, unmodified.
baseline/_ref/ is listed in .gitignore (reference SOURCES never enter this repository's history) but NOT in .gpurunignore, so it
travels to the GPU box like the built .so files.  __graft_entry__.build() calls this when /root/reference exists (this container);
the GPU box only uses the staged copy.
"""
import argparse
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SUBTREES = ("lib", "scripts", "experiments")
KEEP_EXT = (".py", ".yaml", ".yml", ".c", ".h", ".cu", ".cuh", ".md", ".txt", ".json")


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def stage(reference="/root/reference", dest=DEST):
    if not os.path.isdir(reference):
        raise FileNotFoundError(reference)
    manifest = {}
    for sub in SUBTREES:
        src_root = os.path.join(reference, sub)
        for dirpath, dirnames, filenames in os.walk(src_root):
            dirnames[:] = [d for d in dirnames if d not in ("__pycache__", ".git")]
            for fn in filenames:
                if not fn.endswith(KEEP_EXT):
                    continue
                src = os.path.join(dirpath, fn)
                rel = os.path.relpath(src, reference)
                dst = os.path.join(dest, rel)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                shutil.copyfile(src, dst)
                manifest[rel] = sha256(dst)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": reference, "files": manifest}, f, indent=0, sort_keys=True)
    return manifest


def verify(dest=DEST):
    """True iff every staged file still has the sha256 recorded when it was copied from the reference."""
    with open(os.path.join(dest, "MANIFEST.json")) as f:
        files = json.load(f)["files"]
    return all(os.path.exists(os.path.join(dest, rel)) and sha256(os.path.join(dest, rel)) == h for rel, h in files.items())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("USOT_REFERENCE", "/root/reference"))
    a = ap.parse_args()
    m = stage(a.reference)
    print(f"staged {len(m)} reference files into {DEST}")
