"""Build the reference's OWN compiled evaluator as a checker (test infrastructure).

Recipe (does not run the reference's build system, copies no reference source
into the repository): cythonize /root/reference/waveforms/_waveform.pyx — the
only natively compiled module of feihoo87/waveforms, holding calc_parts / _calc
/ _apply and the basis functions 1..15 — where it lies, and compile the
generated C with gcc.  Outputs go ONLY to oracle/_ref/ (git-ignored, shipped to
the GPU box with the snapshot):

    oracle/_ref/waveforms/__init__.py          (empty, written here)
    oracle/_ref/waveforms/_waveform.*.so       (the reference evaluator)
    oracle/_ref/build/_waveform.c              (generated C, scratch)

Usage:  python oracle/build_ref.py            (no-op if /root/reference is absent)
"""
from __future__ import annotations

import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF_PYX = Path('/root/reference/waveforms/_waveform.pyx')
OUT = HERE / '_ref'


def so_path():
    suffix = sysconfig.get_config_var('EXT_SUFFIX')
    return OUT / 'waveforms' / f'_waveform{suffix}'


def build(force=False):
    target = so_path()
    if not REF_PYX.exists():
        return target if target.exists() else None
    if target.exists() and not force and target.stat().st_mtime >= REF_PYX.stat().st_mtime:
        return target
    (OUT / 'waveforms').mkdir(parents=True, exist_ok=True)
    (OUT / 'build').mkdir(parents=True, exist_ok=True)
    (OUT / 'waveforms' / '__init__.py').write_text('')
    c_file = OUT / 'build' / '_waveform.c'
    subprocess.run([sys.executable, '-m', 'cython', '-3', '--module-name',
                    'waveforms._waveform', str(REF_PYX), '-o', str(c_file)],
                   check=True)
    inc = sysconfig.get_paths()['include']
    subprocess.run(['gcc', '-O2', '-fPIC', '-shared', '-fwrapv', '-DNDEBUG',
                    f'-I{inc}', str(c_file), '-o', str(target)], check=True)
    return target


def load():
    """Import the reference evaluator if it has been built; else None."""
    target = so_path()
    if not target.exists():
        return None
    import importlib
    if str(OUT) not in sys.path:
        sys.path.insert(0, str(OUT))
    try:
        return importlib.import_module('waveforms._waveform')
    except Exception:
        return None


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
