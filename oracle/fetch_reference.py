"""Recipe: copy the UNMODIFIED reference files of the hot path into the git-ignored oracle/_ref/reference/
so that they travel to the GPU box with gpurun (the box has no /root/reference) and `bench.py --impl
reference` / its cpu_baseline legs can time the reference itself instead of a port.

TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing is copied into the tracked tree: oracle/_ref/ is listed in
.gitignore (not in .gpurunignore).  Run by __graft_entry__.build() when /root/reference is present:

    python oracle/fetch_reference.py [reference_root]

Files: flows/flows.py, flows/utils.py, calibrators.py, utils/__init__.py,
utils/ops.py, utils/metrics.py -- byte-for-byte copies, sha256 recorded in MANIFEST.json.
"""
import hashlib
import json
import os
import shutil
import sys

FILES = ['flows/flows.py', 'flows/utils.py', 'calibrators.py', 'utils/__init__.py', 'utils/ops.py', 'utils/metrics.py']
HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, '_ref', 'reference')


def fetch(ref_root='/root/reference'):
    if not os.path.isdir(ref_root):
        return False
    manifest = {}
    for rel in FILES:
        src = os.path.join(ref_root, rel)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        manifest[rel] = hashlib.sha256(open(dst, 'rb').read()).hexdigest()
    with open(os.path.join(DEST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': ref_root, 'sha256': manifest}, f, indent=1)
    return True


if __name__ == '__main__':
    ok = fetch(sys.argv[1] if len(sys.argv) > 1 else '/root/reference')
    print('reference copied to %s' % DEST if ok else 'no reference tree found; nothing copied')
