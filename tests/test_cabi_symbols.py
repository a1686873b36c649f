"""The C-ABI library loads and exports every symbol include/basicrta_b200.h declares (no
compute calls: this runs without a GPU), and the ctypes mirrors match the header."""
import ctypes as C
import os
import re

import pytest

from basicrta_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, 'include', 'basicrta_b200.h')).read()


def _declared_functions():
    body = HEADER[HEADER.index('typedef struct brta_batch'):]
    return sorted(set(re.findall(r'^\s*(?:const\s+)?(?:int|char\*?)\s*\*?\s*(brta_\w+)\s*\(', body, flags=re.M)))


def test_library_exports_every_declared_symbol():
    _cabi.build()
    lib = _cabi.load()
    declared = _declared_functions()
    assert set(declared) == set(_cabi.EXPORTS), (declared, _cabi.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.brta_abi_version() == int(re.search(r'#define BRTA_ABI_VERSION (\d+)', HEADER).group(1))
    assert lib.brta_last_error() == b''


def _struct_fields(name):
    m = re.search(r'typedef struct %s \{(.*?)\} %s;' % (name, name), HEADER, flags=re.S)
    body = re.sub(r'/\*.*?\*/', '', m.group(1), flags=re.S)
    fields = []
    for decl in body.split(';'):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(',')
        fields.append(re.findall(r'(\w+)\s*$', names[0])[0])
        for extra in names[1:]:
            fields.append(extra.strip().lstrip('*').strip())
    return fields


@pytest.mark.parametrize('cname,cls', [('brta_caps', _cabi.Caps), ('brta_launch_info', _cabi.LaunchInfo),
                                       ('brta_task', _cabi.Task), ('brta_batch', _cabi.Batch),
                                       ('brta_gmm_batch', _cabi.GmmBatch)])
def test_ctypes_mirror_matches_header(cname, cls):
    assert _struct_fields(cname) == [f[0] for f in cls._fields_]


def test_constants_match_header():
    def macro(name):
        return re.search(r'#define\s+%s\s+\(?([0-9a-fx]+)u?' % name, HEADER).group(1)
    assert int(macro('BRTA_THREADS')) == _cabi.THREADS
    assert int(macro('BRTA_MAX_NCOMP')) == _cabi.MAX_NCOMP
    assert int(macro('BRTA_LANE_MAX_NCOMP')) == _cabi.LANE_MAX_NCOMP
    assert int(macro('BRTA_MAILBOX_MAX_TEAM')) == _cabi.MAILBOX_MAX_TEAM
    for flag in ('EXACT', 'INJECT_COEF', 'INJECT_U', 'TRACE', 'NO_TABLE', 'CTAS3'):
        assert int(macro('BRTA_FLAG_' + flag)) == getattr(_cabi, 'FLAG_' + flag)
    assert int(macro('BRTA_GMM_MAX_COMPONENTS')) == _cabi.GMM_MAX_COMPONENTS
    assert int(macro('BRTA_GMM_MAX_POINTS')) == _cabi.GMM_MAX_POINTS
    assert [int(macro('BRTA_GMM_' + n)) for n in ('CONVERGED', 'NOT_CONVERGED', 'ILL_DEFINED')] == \
        [_cabi.GMM_CONVERGED, _cabi.GMM_NOT_CONVERGED, _cabi.GMM_ILL_DEFINED]
    from basicrta_b200.plan import TASK_DTYPE
    assert TASK_DTYPE.itemsize == C.sizeof(_cabi.Task)
    assert list(TASK_DTYPE.names) == [f[0] for f in _cabi.Task._fields_]


def test_argument_errors_do_not_need_a_gpu():
    lib = _cabi.load()
    assert lib.brta_gibbs_run_batch(None, None) == -1                     # BRTA_E_NULL
    assert b'NULL' in lib.brta_last_error()
    b = _cabi.Batch()
    b.ncomp = 256
    assert lib.brta_gibbs_run_batch(C.byref(b), None) == -2               # BRTA_E_NCOMP
    assert lib.brta_query(0, None) == -1
    assert lib.brta_pindicator_counts(None, 0, 0, 0, None, 1, 1, None, None) == -1
    one = C.c_int(0)
    p = C.cast(C.byref(one), C.c_void_p)
    assert lib.brta_pindicator_counts(p, 4, 2, 8, p, 15, 3, p, None) == -3   # row_stride < n_data: BRTA_E_RANGE
    assert lib.brta_pindicator_counts(p, 8, 2, 8, p, 15, 33, p, None) == -2  # n_clusters > 32
    assert lib.brta_pindicator_counts(p, 8, 0, 8, p, 15, 3, p, None) == 0    # nothing to do
    assert lib.brta_gmm_fit_batch(None, None) == -1
    g = _cabi.GmmBatch()
    assert lib.brta_gmm_fit_batch(C.byref(g), None) == -1                 # NULL pointers inside
    for f in ('x', 'offsets', 'n_components', 'lower_bound', 'n_iter', 'status', 'params'):
        setattr(g, f, p.value)
    g.n_init, g.max_points = 1, _cabi.GMM_MAX_POINTS + 1
    assert lib.brta_gmm_fit_batch(C.byref(g), None) == -3                 # too many points: BRTA_E_RANGE
    g.max_points, g.n_problems = 10, 0
    assert lib.brta_gmm_fit_batch(C.byref(g), None) == 0                  # nothing to do
    assert lib.brta_gmm_predict(None, None, 0, 0, None, None, None, None) == -1


def test_product_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from basicrta_b200.engine import GibbsEngine
    with pytest.raises(_cabi.BrtaError):
        GibbsEngine(0)
