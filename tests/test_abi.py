"""The C-ABI library loads without a GPU and exports every symbol that
include/wfm_b200.h declares; no compute call is made here."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / 'include' / 'wfm_b200.h').read_text()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    names = re.findall(r'^\s*(?:const\s+)?(?:int|int64_t|char\*|const char\*)\s+\**\s*(wfm_\w+)\s*\(', text, flags=re.M)
    return sorted(set(names))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    for must in ('wfm_program_create', 'wfm_sample', 'wfm_sample_host', 'wfm_sosfilt', 'wfm_fft_filter'):
        assert must in names


def test_library_exports_every_declared_symbol():
    from waveforms_b200 import engine
    lib = engine.load_library()
    for name in declared_functions():
        assert hasattr(lib, name), name
    assert sorted(engine.EXPORTS) == declared_functions()
    assert lib.wfm_abi_version() == 2


def test_struct_layouts_match_header():
    from waveforms_b200 import lowering as L
    assert L.WAVE_DT.itemsize == 112 and L.WAVE_DT.fields['n'][1] == 56
    assert L.WAVE_DT.fields['out_off2'][1] == 96 and L.WAVE_DT.fields['offset2'][1] == 104
    assert L.WAVE_DT.fields['seg_begin'][1] == 80 and L.WAVE_DT.fields['flags'][1] == 88
    assert L.SEGPTR_DT.itemsize == 8
    assert L.FACTOR_DT.itemsize == 32 and L.FACTOR_DT.fields['shift'][1] == 8
    assert L.TERM_DT.itemsize == 32 and L.TERM_DT.fields['ref_begin'][1] == 16
    assert L.REF_DT.itemsize == 16 and L.REF_DT.fields['slot'][1] == 8


def test_invalid_program_is_rejected_without_gpu():
    """Validation happens on the host before any CUDA call."""
    from waveforms_b200 import engine
    lib = engine.load_library()
    d = engine._ProgramDesc()
    d.n_waves = -1
    h = ctypes.c_void_p()
    assert lib.wfm_program_create(ctypes.byref(d), 0, ctypes.byref(h)) == -1
    assert b'negative' in lib.wfm_last_error()


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from waveforms_b200 import cosPulse, engine
    with pytest.raises(engine.EngineUnavailable):
        cosPulse(1.0)(np.linspace(-1, 1, 11))


def test_product_never_imports_the_oracle():
    pkg = ROOT / 'waveforms_b200'
    for path in pkg.rglob('*.py'):
        text = path.read_text()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), path
        assert 'wfm_oracle' not in text, path
