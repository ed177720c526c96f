"""Recipe: pack the UNMODIFIED reference modules of the hot path into the git-ignored artefact
oracle/_ref/reference.zip so that they travel to the GPU box with gpurun (the box has no /root/reference) and
`bench.py --impl reference` / its cpu_baseline legs can time the reference itself instead of a port.

TEST / BASELINE INFRASTRUCTURE ONLY.  No reference source enters the tracked tree: oracle/_ref/ is listed in
.gitignore (not in .gpurunignore), like the compiled oracle next to it.  Run by __graft_entry__.build() when
/root/reference is present:

    python oracle/fetch_reference.py [reference_root]

Members: flows/flows.py, flows/utils.py, calibrators.py, utils/__init__.py,
utils/ops.py, utils/metrics.py -- byte-for-byte, sha256 recorded in the archive's MANIFEST.json.  Python imports
them straight from the archive (zipimport), see oracle/ref_loader.py.
"""
import hashlib
import json
import os
import shutil
import sys
import zipfile

FILES = ['flows/flows.py', 'flows/utils.py', 'calibrators.py', 'utils/__init__.py', 'utils/ops.py', 'utils/metrics.py']
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref', 'reference.zip')


def fetch(ref_root='/root/reference'):
    if not os.path.isdir(ref_root):
        return False
    os.makedirs(os.path.dirname(DEST), exist_ok=True)
    old_dir = os.path.join(HERE, '_ref', 'reference')
    if os.path.isdir(old_dir):
        shutil.rmtree(old_dir)
    manifest = {}
    tmp = DEST + '.tmp'
    with zipfile.ZipFile(tmp, 'w', zipfile.ZIP_STORED) as z:
        for d in sorted({os.path.dirname(rel) + '/' for rel in FILES if os.path.dirname(rel)}):
            z.writestr(zipfile.ZipInfo(d, date_time=(2020, 1, 1, 0, 0, 0)), b'')     # flows/ has no __init__.py: a namespace
        for rel in FILES:                                                                # package needs its directory entry
            data = open(os.path.join(ref_root, rel), 'rb').read()
            z.writestr(zipfile.ZipInfo(rel, date_time=(2020, 1, 1, 0, 0, 0)), data)
            manifest[rel] = hashlib.sha256(data).hexdigest()
        z.writestr(zipfile.ZipInfo('MANIFEST.json', date_time=(2020, 1, 1, 0, 0, 0)),
                   json.dumps({'source': ref_root, 'sha256': manifest}, indent=1))
    os.replace(tmp, DEST)
    return True


if __name__ == '__main__':
    ok = fetch(sys.argv[1] if len(sys.argv) > 1 else '/root/reference')
    print('reference packed into %s' % DEST if ok else 'no reference tree found; nothing packed')
