"""bench.py's reference arm (runs on the host cores, no GPU): ONE JSON line on stdout with the keys of the contract."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS='2')
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.split('\n') if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'tets/s' and d['higher_is_better'] is True
    assert d['metric'] == 'N1curl order-2 S+T elements assembled/s (fp64)'
    assert d['value'] > 0 and d['ms_per_step'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'tets/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config']


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=120, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ''


def test_closed_form_workload_sizes_match_the_oracle():
    """bench.cube_counts (the `config` both arms print) against the oracle's numbering and pattern on small cubes."""
    import importlib.util
    import numpy as np
    from oracle import fem, meshes
    spec = importlib.util.spec_from_file_location('bench', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for n in (2, 3, 5):
        co, ce = meshes.unit_cube(n)
        num = fem.number_entities(ce)
        cd, ndof = fem.cell_dofs(num, 2)
        rowptr, cols = fem.pattern(cd, ndof)
        c = bench.cube_counts(n)
        assert (c["tets"], c["vertices"], c["ndof"], c["nnz"]) == (len(ce), len(co), ndof, len(cols))
    assert bench.workload_config(119, 1) == bench.workload_config(119, 1)
    assert bench.slab_layers(119, 700e3) == 8 and bench.slab_layers(203, 700e3) == 3
