"""CPU: the C-ABI library loads and exports every symbol include/minesweeper_b200.h
declares; without a GPU it refuses to run instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'minesweeper_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ms_[a-z_0-9]+)\s*\(', src)))


def test_header_symbols_exported():
    from minesweeper_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), 'libmsb200.so does not export %s' % n
        assert n in _lib.PROTOTYPES, 'no ctypes prototype for %s' % n
    assert lib.ms_version().startswith(b'minesweeper_b200')


def test_struct_layouts_match_header():
    from minesweeper_b200 import _lib
    # ms_flags: seven int32; ms_grid_desc: four int32 then six pointers
    assert ctypes.sizeof(_lib.Flags) == 28
    assert ctypes.sizeof(_lib.GridDesc) == 16 + 6 * ctypes.sizeof(ctypes.c_void_p)
    assert _lib.MS_NPARAM == 24 and _lib.MS_NDERIVED == 13


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from minesweeper_b200 import _lib, synth
    from minesweeper_b200.engine import Engine
    with pytest.raises(_lib.MSError) as e:
        Engine(mist=synth.make_mist(ne=10, nm=6, nf=3, na=3))
    assert 'no CPU fallback' in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'minesweeper_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f


def _build_c_smoke(tmp_path):
    import subprocess
    exe = str(tmp_path / 'abi_smoke')
    libdir = os.path.join(ROOT, 'minesweeper_b200')
    cmd = ['gcc', '-O1', '-o', exe, os.path.join(ROOT, 'tests', 'c', 'abi_smoke.c'), '-L' + libdir, '-lmsb200',
           '-Wl,-rpath,' + libdir, '-lm']
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_plain_c_program_links_against_the_abi(tmp_path):
    """A C program with nothing but include/minesweeper_b200.h compiles and links; without a GPU it
    gets MS_ENODEV (exit code 3) instead of a fallback."""
    import subprocess
    import torch
    from minesweeper_b200 import _lib
    _lib.load()
    exe = _build_c_smoke(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
        assert 'lnL[0]' in r.stdout and 'lnL[2] = -inf' in r.stdout
    else:
        assert r.returncode == 3, r.stdout + r.stderr
        assert 'no CPU fallback' in r.stdout


@pytest.mark.gpu
def test_plain_c_program_runs_on_gpu(tmp_path):
    import subprocess
    exe = _build_c_smoke(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert 'lnL[2] = -inf' in r.stdout and 'launches = 2' in r.stdout
