"""In-tree build of libmsb200.so (nvcc, sm_100a only).

``python -m minesweeper_b200.build`` or ``minesweeper_b200.build.build()``.
nvcc cross-compiles without a GPU; the resulting library sits next to this
file so it travels with a snapshot of the repo.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libmsb200.so')
SOURCES = ['capi.cu', 'mist_interp.cu', 'pars.cu', 'prior.cu', 'sampler.cu', 'phot_mlp.cu', 'phot_mlp_tc.cu',
           'spec_mlp.cu', 'spec_mlp_tc.cu', 'spec_smooth.cu', 'spec_fast.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr']


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, extra_flags=()):
    """variant / extra_flags: kernel-tuning experiments only (``libmsb200_<variant>.so`` built with extra -D flags,
    selected at run time with MS_LIB_PATH); the product library is the default build."""
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(HERE, '..', 'include', 'minesweeper_b200.h'))
    objdir = os.path.join(HERE, 'build' if not variant else 'build_' + variant)
    lib = LIB if not variant else os.path.join(HERE, 'libmsb200_%s.so' % variant)
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    objs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace('.cu', '.o'))
        objs.append(obj)
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + list(extra_flags) + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed: %s\n%s\n%s' % (' '.join(cmd), r.stdout, r.stderr))
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(run, jobs):
                if verbose and out:
                    sys.stderr.write(out)
    if force or jobs or _stale(lib, objs):
        run([NVCC, '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'])
    return lib


if __name__ == '__main__':
    var = os.environ.get('MS_BUILD_VARIANT')
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, variant=var,
                extra_flags=os.environ.get('MS_NVCC_EXTRA', '').split() if var else ()))
