"""Stage the UNMODIFIED reference modules the video hot path needs under baseline/_ref/ (git-ignored, NOT
gpurun-ignored: it travels to the GPU box with the snapshot, like a built .so).

    python baseline/make_ref.py            # in the build container, where /root/reference exists

Copies `video_module/` and `myutils/` byte for byte (SURVEY.md 7.1 step 0, 8c "GPU oracle") and records their
SHA-256 in baseline/_ref/MANIFEST.json, so that a test can assert the staged files are the reference's own.
Nothing under baseline/_ref/ is committed; nothing in vfloodnet_b200/ imports it.  baseline/refshim.py makes
the staged tree importable (two import shims, no arithmetic of the reference's own code touched).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('VFN_REFERENCE', '/root/reference')
DST = os.path.join(HERE, '_ref')
PACKAGES = ('video_module', 'myutils')


def _sha(path):
    h = hashlib.sha256()
    with open(path, 'rb') as f:
        h.update(f.read())
    return h.hexdigest()


def stage(force=False):
    """Returns the staging directory, or None when the reference tree is not present (GPU box: prebuilt copy only)."""
    if not os.path.isdir(os.path.join(REF, 'video_module')):
        return DST if os.path.isdir(os.path.join(DST, 'video_module')) else None
    manifest = {}
    for pkg in PACKAGES:
        dst = os.path.join(DST, pkg)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REF, pkg), dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        for root, _dirs, files in os.walk(dst):
            for fn in sorted(files):
                p = os.path.join(root, fn)
                manifest[os.path.relpath(p, DST)] = _sha(p)
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': REF, 'files': manifest}, f, indent=1, sort_keys=True)
    return DST


if __name__ == '__main__':
    d = stage(force='--force' in sys.argv)
    print(d or f'{REF} not present and no staged copy: nothing to do')
