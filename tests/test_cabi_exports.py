"""The C-ABI library loads and exports every symbol include/sucem_b200.h declares (no GPU needed)."""
import os
import re
import subprocess

from conftest import ROOT


def test_header_symbols_exported():
    import sucemfem._native as N
    hdr = open(os.path.join(ROOT, 'include', 'sucem_b200.h')).read()
    declared = sorted(set(re.findall(r'SFEM_API\s+[\w\s\*]+?\b(sfem_\w+)\s*\(', hdr)))
    assert len(declared) >= 20
    nm = subprocess.run(['nm', '-D', '--defined-only', N.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r' T (sfem_\w+)', nm))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert set(declared) == set(N.EXPORTED)          # the ctypes binding covers the whole header
    assert N.version().startswith('sucem_b200')


def test_sm100a_sass_present():
    import sucemfem._native as N
    out = subprocess.run(['cuobjdump', '-lelf', N.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps every exported entry point to the reference interface it replaces."""
    hdr = open(os.path.join(ROOT, 'include', 'sucem_b200.h')).read()
    doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    declared = sorted(set(re.findall(r'SFEM_API\s+[\w\s\*]+?\b(sfem_\w+)\s*\(', hdr)))
    # families are listed as `sfem_x_create/run/destroy`
    expanded = set(re.findall(r'sfem_\w+', doc))
    for fam in re.findall(r'`(sfem_\w+(?:/\w+)+)`', doc):
        head, *rest = fam.split('/')
        stem = head[:head.rindex('_') + 1]
        expanded.update(stem + r for r in rest)
    missing = [n for n in declared if n not in expanded]
    assert not missing, missing
