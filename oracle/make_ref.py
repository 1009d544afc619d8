"""Recipe that stages the UNMODIFIED reference package into oracle/_ref/ (git-ignored; it travels to the GPU box
with the repo snapshot, where /root/reference does not exist).  TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/make_ref.py          # build container only: needs /root/reference

What it does: `pip install --no-index --no-deps --target oracle/_ref` of a scratch copy of /root/reference (the
checkout is read-only and setuptools writes build/ and *.egg-info next to setup.py).  The two trained-weight
files (140 MB) are NOT duplicated: oracle/ref_runner.py hands the reference the copies staged under
dmpfold2_b200/trained_model/ through its own `weights_file` argument (predict.py:75, :93-95).
No reference source is committed; nothing here is imported by the product.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference'
DST = os.path.join(HERE, '_ref')
MARK = os.path.join(DST, 'dmpfold', 'network.py')


def build(force: bool = False) -> str:
    if os.path.isfile(MARK) and not force:
        return DST
    if not os.path.isdir(REF_SRC):
        raise RuntimeError('oracle/_ref is absent and /root/reference is not available to build it from')
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, 'reference')
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns('*.pt', '.git*'))
        for root, dirs, files in os.walk(src):                       # the checkout is read-only; the copy must not be
            for n in dirs + files:
                os.chmod(os.path.join(root, n), 0o755 if n in dirs else 0o644)
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '--find-links',
               '/opt/wheelhouse', '--target', DST, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        how = 'pip install --target'
        if r.returncode != 0 or not os.path.isfile(MARK):
            # same files, copied by hand (pip unavailable / setuptools refused): the package is pure Python
            how = 'file copy (pip failed: %s)' % (r.stderr.strip().splitlines()[-1] if r.stderr.strip() else 'unknown')
            shutil.copytree(os.path.join(src, 'dmpfold'), os.path.join(DST, 'dmpfold'), dirs_exist_ok=True,
                            ignore=shutil.ignore_patterns('*.pt'))
    with open(os.path.join(DST, 'HOW_BUILT.txt'), 'w') as fh:
        fh.write('source: %s\nmethod: %s\n' % (REF_SRC, how))
    return DST


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
