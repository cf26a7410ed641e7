"""Install the UNMODIFIED reference (PyChebyshev v0.21.1, pure Python) under oracle/_ref/.

TEST INFRASTRUCTURE ONLY (same rules as the rest of oracle/).  The reference is a pure-Python
wheel (``pyproject.toml:56-57``: ``packages = ["src/pychebyshev"]``).  ``pip install --target``
cannot build it in this image (build backend ``hatchling`` is not installed and there is no
network), so this script does what installing that wheel does: it places the package's files,
byte for byte, under ``oracle/_ref/site/pychebyshev`` and writes a ``.dist-info``.  It also places
the reference's own hot-path test files (and their fixtures) under ``oracle/_ref/ref_tests`` so the
GPU box can run them against the swapped-in backend (SURVEY.md §4 "free regression suite").

``oracle/_ref/`` is git-ignored (nothing of the reference enters this repository's history) but NOT
gpurun-ignored, so the install travels to the GPU box, where ``/root/reference`` does not exist.
``MANIFEST.json`` records the sha256 of every installed file; ``verify()`` re-hashes them against
the source tree when it is present.

Run:  python oracle/ref_install.py            (idempotent; called by __graft_entry__.build())
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("PCB_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")
SITE = os.path.join(DEST, "site")
TESTS = os.path.join(DEST, "ref_tests")

#: the reference's hot-path test files + the rest of its suite, which also calls
#: the evaluation methods
TEST_FILES = None  # None = every tests/test_*.py


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def _copy_tree(src, dst, manifest, rel_base):
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        for name in files:
            if name.endswith((".pyc", ".nbi", ".nbc")):
                continue
            s = os.path.join(root, name)
            rel = os.path.relpath(s, src)
            d = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            manifest[os.path.join(rel_base, rel)] = {
                "sha256": _sha(d), "source": os.path.relpath(s, REF_ROOT)}


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "pychebyshev"))


def installed() -> bool:
    return os.path.exists(os.path.join(SITE, "pychebyshev", "__init__.py")) and \
        os.path.exists(os.path.join(DEST, "MANIFEST.json"))


def install(force: bool = False) -> str:
    """Place the reference package + its tests under oracle/_ref/.  Returns the site dir."""
    if not available():
        if installed():
            return SITE
        raise RuntimeError(f"reference tree not found at {REF_ROOT} and no prior install")
    if installed() and not force and not verify(quiet=True):
        return SITE
    for d in (SITE, TESTS):
        shutil.rmtree(d, ignore_errors=True)
    manifest: dict = {}
    _copy_tree(os.path.join(REF_ROOT, "src", "pychebyshev"), os.path.join(SITE, "pychebyshev"),
               manifest, "site/pychebyshev")
    # the wheel's metadata (what `pip install --target` would have written)
    version = "unknown"
    with open(os.path.join(REF_ROOT, "pyproject.toml")) as f:
        for line in f:
            if line.startswith("version"):
                version = line.split("=")[1].strip().strip('"')
                break
    info = os.path.join(SITE, f"pychebyshev-{version}.dist-info")
    os.makedirs(info, exist_ok=True)
    with open(os.path.join(info, "METADATA"), "w") as f:
        f.write(f"Metadata-Version: 2.1\nName: pychebyshev\nVersion: {version}\n")
    with open(os.path.join(info, "INSTALLER"), "w") as f:
        f.write("oracle/ref_install.py\n")
    # the reference's tests + fixtures
    os.makedirs(TESTS, exist_ok=True)
    tdir = os.path.join(REF_ROOT, "tests")
    for name in sorted(os.listdir(tdir)):
        s = os.path.join(tdir, name)
        if os.path.isfile(s) and name.endswith(".py") and (
                TEST_FILES is None or name in TEST_FILES or name == "conftest.py"):
            shutil.copyfile(s, os.path.join(TESTS, name))
            manifest[f"ref_tests/{name}"] = {"sha256": _sha(s), "source": f"tests/{name}"}
    _copy_tree(os.path.join(tdir, "fixtures"), os.path.join(TESTS, "fixtures"), manifest,
               "ref_tests/fixtures")
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"reference": "0xC000005/PyChebyshev", "version": version,
                   "installed_from": REF_ROOT, "files": manifest}, f, indent=1, sort_keys=True)
    return SITE


def verify(quiet: bool = False) -> list:
    """Installed files whose bytes differ from MANIFEST.json (and from the source tree when it is
    present).  Empty list = unmodified."""
    with open(os.path.join(DEST, "MANIFEST.json")) as f:
        man = json.load(f)
    bad = []
    for rel, rec in man["files"].items():
        p = os.path.join(DEST, rel)
        if not os.path.exists(p) or _sha(p) != rec["sha256"]:
            bad.append(rel)
            continue
        src = os.path.join(REF_ROOT, rec["source"])
        if os.path.exists(src) and _sha(src) != rec["sha256"]:
            bad.append(rel)
    if bad and not quiet:
        print("modified or missing:", bad, file=sys.stderr)
    return bad


if __name__ == "__main__":
    site = install(force="--force" in sys.argv)
    bad = verify()
    print(f"reference installed under {site}; {'UNMODIFIED' if not bad else 'MISMATCH'}")
    sys.exit(1 if bad else 0)
