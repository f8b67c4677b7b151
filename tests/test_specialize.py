"""Host side of the mechanism-specialised kernel images (openccm_b200/specialize.py): the generated term code, eligibility,
the content-addressed image cache and the nvcc build (cross-compiles without a GPU)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import cases as C
from _util import setup_case


def _host(name, tmp_path):
    from openccm_b200.system_solvers.export import export_system
    cp, cmesh, net, mn = setup_case(name, 'B200-RK45', tmp_path)
    return export_system(C.CASES[name]['model'], mn, cp, cmesh)


def test_table_hash_is_fnv1a():
    from openccm_b200.specialize import table_hash
    assert table_hash(np.frombuffer(b'a', dtype=np.uint8)) == 0xaf63dc4c8601ec8c
    assert table_hash(np.frombuffer(b'foobar', dtype=np.uint8)) == 0x85944171f73967e8
    assert table_hash(np.zeros(0, dtype=np.uint8)) == 0xcbf29ce484222325


def test_header_lists_every_term_in_storage_order(tmp_path):
    """One statement per stoichiometric term, species-major like MechanismTables.pack(); a factor of order n is n products;
    coefficients are exact (hex floats)."""
    from openccm_b200.specialize import mechanism_header
    host = _host('cstr_net12_nacl', tmp_path)                 # 2 NaCl + CaCO3 -> Na2CO3 + CaCl2, reversible: orders 2 and 1
    tables = host.tables
    text = mechanism_header(tables)
    terms = sorted(tables.terms, key=lambda tm: tm.specie)
    assert f'#define OCCM_SPEC_NTERM {len(terms)}' in text
    stmts = [ln.strip().rstrip('\\').strip() for ln in text.splitlines() if ln.strip().startswith('r[')]
    assert len(stmts) == len(terms)
    for k, (tm, st) in enumerate(zip(terms, stmts)):
        assert st.startswith(f'r[{tm.specie}] = __dadd_rn(r[{tm.specie}], ')
        assert st.count('__dmul_rn(') == sum(o for _, o in tm.factors)
        assert f'sc[{k}]' in st
        for spi, order in tm.factors:
            assert st.count(f'c[{spi}])') == order
    coef_line = [ln for ln in text.splitlines() if 'occm_spec_coef' in ln][0]
    vals = [float.fromhex(v) for v in coef_line.split('{')[1].split('}')[0].split(', ')]
    assert vals == [tm.coeff for tm in terms]
    rxn_line = [ln for ln in text.splitlines() if 'occm_spec_rxn' in ln][0]
    assert [int(v) for v in rxn_line.split('{')[1].split('}')[0].split(', ')] == [tm.rxn for tm in terms]


@pytest.mark.parametrize('name,ok', [('cstr_net12_nacl', True), ('cstr_reversible', True), ('cstr_zeroth', True),
                                     ('cstr_net20_timebc', False),       # three species: odd
                                     ('cstr_monod', False),              # Monod kinetics: byte-code terms
                                     ('pfr_net8', True),                 # PFR: one upwind entry per point
                                     ('pfr_net6_pointmass', False),      # one species: odd
                                     ('cstr_net9_tracer', False)])       # no reactions: nothing to specialise
def test_eligibility(name, ok, tmp_path):
    from openccm_b200.specialize import eligible
    host = _host(name, tmp_path)
    assert eligible(host, 1 if host.model == 'pfr' else 2) is ok


@pytest.mark.skipif(shutil.which('nvcc') is None and not os.path.exists('/usr/local/cuda/bin/nvcc'), reason='needs nvcc')
def test_image_builds_and_is_cached(tmp_path, monkeypatch):
    from openccm_b200 import specialize as sp
    monkeypatch.setenv('OCCM_B200_SPEC_DIR', str(tmp_path / 'spec'))
    monkeypatch.delenv('OCCM_B200_SPEC_FLAGS', raising=False)
    host = _host('cstr_net12_nacl', tmp_path)
    blob = host.tables.pack()
    path = sp.build_image(host.tables, blob, host.S, 2)
    assert os.path.dirname(path) == str(tmp_path / 'spec') and os.path.getsize(path) > 1000
    with open(path, 'rb') as fh:
        assert fh.read(4) == b'\x7fELF'
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if os.path.exists(cuobjdump):
        out = subprocess.run([cuobjdump, '-elf', path], capture_output=True, text=True).stdout
        for sym in ('occm_spec_stage0', 'occm_spec_stage1', 'occm_spec_stage2', 'occm_spec_stage7', 'occm_spec_info'):
            assert sym in out
        sass = subprocess.run([cuobjdump, '-sass', path], capture_output=True, text=True).stdout
        assert 'UBLKCP' in sass and 'SYNCS' in sass          # bulk copies + mbarriers of the ring
    t0 = os.path.getmtime(path)
    assert sp.build_image(host.tables, blob, host.S, 2) == path and os.path.getmtime(path) == t0      # cache hit: no rebuild
    # another ELL width, other flags or another mechanism are other images
    assert sp.build_image(host.tables, blob, host.S, 3) != path
    assert sp.image_path(sp.mechanism_header(host.tables), host.S, 2, sp.table_hash(blob), ['-DOCCM_ROW_W0=5']) != path
    other = _host('cstr_reversible', tmp_path)
    assert sp.image_path(sp.mechanism_header(other.tables), other.S, 2, sp.table_hash(other.tables.pack()), []) != path


def test_missing_nvcc_is_an_error_only_when_asked_for(tmp_path, monkeypatch):
    from openccm_b200 import specialize as sp
    monkeypatch.setenv('OCCM_B200_SPEC_DIR', str(tmp_path / 'spec'))
    monkeypatch.setenv('NVCC', str(tmp_path / 'no_such_nvcc'))
    host = _host('cstr_net12_nacl', tmp_path)
    with pytest.raises(FileNotFoundError):
        sp.build_image(host.tables, host.tables.pack(), host.S, 2)
    assert sp.mode() == 'auto'
    monkeypatch.setenv('OCCM_B200_SPECIALIZE', '0')
    assert sp.mode() == 'off'
    monkeypatch.setenv('OCCM_B200_SPECIALIZE', '1')
    assert sp.mode() == 'on'
