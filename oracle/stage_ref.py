#!/usr/bin/env python
"""Stage the UNMODIFIED reference for boxes where /root/reference does not exist (the GPU box).

Test / measurement infrastructure, like everything under oracle/.  The reference is pure
Python, so "building" it means copying - byte for byte, never editing - the four modules of
the path and the fixtures its own pipeline needs

    config_and_setup.py  helpers.py  embed_process.py  extract_process.py
    bob_private_key.pem  bob_public_key.pem
    media/input/image64.png  media/input/image32.png  media/input/cover_1.mp4

from /root/reference into oracle/_ref/.  That directory is git-ignored (nothing reference-owned
is ever committed) but NOT gpurun-ignored: it travels to the GPU box with the snapshot exactly
like the built .so files do.  MANIFEST.json records the SHA-256 of every staged file; users of
the staged copy verify it, so an edited copy is rejected.

Users (and only these): tests/ (the reference's own pipeline functions running on the B200
path after svs_b200.install()), bench.py's `--impl reference` arm and `cpu_baseline` leg (the
real proses_frame_qim_dct timed on the host cores, kind "reference").

    python oracle/stage_ref.py          # stage (needs /root/reference), print the manifest
"""
import hashlib
import importlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = "/root/reference"
DEST = os.path.join(HERE, "_ref")
FILES = ["config_and_setup.py", "helpers.py", "embed_process.py", "extract_process.py",
         "bob_private_key.pem", "bob_public_key.pem",
         "media/input/image64.png", "media/input/image32.png", "media/input/cover_1.mp4"]
MODULES = ("config_and_setup", "helpers", "embed_process", "extract_process")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def stage(force=False):
    """Copy the reference files into oracle/_ref/ (only where /root/reference exists).
    Returns the staged directory, or None when there is neither a source nor a staged copy."""
    if not os.path.isdir(SOURCE):
        return DEST if available() else None
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(SOURCE, rel), os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(src)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": SOURCE, "sha256": manifest}, f, indent=1, sort_keys=True)
    return DEST


def available():
    """True when a complete, unmodified staged copy is present."""
    try:
        with open(os.path.join(DEST, "MANIFEST.json")) as f:
            manifest = json.load(f)["sha256"]
    except Exception:
        return False
    return all(os.path.exists(os.path.join(DEST, rel)) and _sha(os.path.join(DEST, rel)) == manifest.get(rel)
               for rel in FILES)


def path(rel=""):
    return os.path.join(DEST, rel)


def import_reference():
    """Import the staged reference modules (fresh, from oracle/_ref/ only) and return them as a
    dict name -> module.  Raises RuntimeError when no verified staged copy exists."""
    if not available():
        raise RuntimeError("no staged reference under oracle/_ref (run `python oracle/stage_ref.py` where "
                           "/root/reference exists)")
    for name in MODULES:                      # never mix with a copy imported from elsewhere
        mod = sys.modules.get(name)
        if mod is not None and os.path.dirname(os.path.abspath(getattr(mod, "__file__", ""))) != DEST:
            del sys.modules[name]
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True            # keep oracle/_ref byte-identical to the manifest
    try:
        return {name: importlib.import_module(name) for name in MODULES}
    finally:
        sys.dont_write_bytecode = dont


if __name__ == "__main__":
    d = stage(force="--force" in sys.argv)
    if d is None:
        raise SystemExit("no /root/reference here and nothing staged")
    print(open(os.path.join(d, "MANIFEST.json")).read())
    print("staged:", d, "verified:", available())
