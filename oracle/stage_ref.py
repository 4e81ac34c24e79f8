"""TEST INFRASTRUCTURE ONLY -- stage the UNMODIFIED Python reference into oracle/_ref/ so that it travels to the GPU box.

    python oracle/stage_ref.py            # or `make -C oracle ref`; __graft_entry__.build() runs it when /root/reference exists

The reference is pure Python (no build step): "building" it means copying the files the hot path's callers need, byte for
byte, from where they lie under /root/reference into the git-ignored oracle/_ref/ (kept out of history, NOT gpurun-ignored,
so a snapshot carries it like any other built artefact):

    model.py pack.py tools.py generate.py rolling.py trainer.py pack_net/*.py      the reference's own modules
    pretrain_model/{2d,3d}-bot-C+P+S-lb-soft-width-5-note-sh-R-diff/actor.pt       the shipped pretrained actors (G6)

Nothing under oracle/_ref/ is ever imported by the product (tap-net_b200/); the users are tests/ (model-in-the-loop
parity), bench.py's `--impl reference` arm and `cpu_baseline` leg, through oracle/refshim.py.  A MANIFEST with the sha256
of every staged file is written next to them so a test can tell that the copy is the unmodified reference.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("TAPNET_REFERENCE_SRC", "/root/reference")

MODULES = ["model.py", "pack.py", "tools.py", "generate.py", "rolling.py", "trainer.py"]
CHECKPOINTS = ["2d-bot-C+P+S-lb-soft-width-5-note-sh-R-diff", "3d-bot-C+P+S-lb-soft-width-5-note-sh-R-diff"]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def file_list(src=SRC):
    out = list(MODULES)
    pn = os.path.join(src, "pack_net")
    if os.path.isdir(pn):
        out += sorted(os.path.join("pack_net", f) for f in os.listdir(pn) if f.endswith(".py"))
    out += [os.path.join("pretrain_model", c, "actor.pt") for c in CHECKPOINTS]
    return out


def stage(src=SRC, dest=DEST, verbose=False):
    """Copy the reference files; returns the manifest dict, or None when the reference tree is absent (GPU box)."""
    if not os.path.isfile(os.path.join(src, "tools.py")):
        return None
    manifest = {}
    for rel in file_list(src):
        s, d = os.path.join(src, rel), os.path.join(dest, rel)
        if not os.path.isfile(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        digest = _sha(s)
        if not (os.path.isfile(d) and _sha(d) == digest):
            shutil.copyfile(s, d)
            os.chmod(d, 0o644)
        manifest[rel] = digest
        if verbose:
            print("staged", rel)
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "files": manifest}, f, indent=1, sort_keys=True)
    return manifest


def verify(dest=DEST):
    """True when every file of the manifest is present with the recorded digest."""
    mf = os.path.join(dest, "MANIFEST.json")
    if not os.path.isfile(mf):
        return False
    files = json.load(open(mf))["files"]
    return all(os.path.isfile(os.path.join(dest, rel)) and _sha(os.path.join(dest, rel)) == dg for rel, dg in files.items())


if __name__ == "__main__":
    m = stage(verbose=True)
    if m is None:
        print("reference tree not found at %s -- nothing staged" % SRC)
        sys.exit(1)
    print("%d files -> %s" % (len(m), DEST))
