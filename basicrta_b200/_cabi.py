"""ctypes binding of ``include/basicrta_b200.h`` (the C-ABI drop-in boundary).

This is the stub a basicrta maintainer would add.  There is no CPU
fallback: if ``libbrta_gibbs.so`` is missing or the device is not a CUDA GPU the calls
raise :class:`BrtaError`.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('BRTA_LIB', os.path.join(HERE, 'libbrta_gibbs.so'))   # BRTA_LIB: developer override
CSRC = os.path.join(HERE, 'csrc')
OBJ_DIR = os.path.join(CSRC, '_obj')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

ABI_VERSION = 6
THREADS = 128
MAX_NCOMP = 255
LANE_MAX_NCOMP = 32
TICK_LIMIT = 1 << 23
MAILBOX_MAX_TEAM = 32
MAX_SHARDS = 16
IPC_HANDLE_BYTES = 64


def shard_mailbox_bytes(g):
    return 2 * g * 32 * 32


def exch_bytes(team):
    """BRTA_EXCH_BYTES(team) of include/basicrta_b200.h."""
    return 2 * team * 32 * 16 if team <= MAILBOX_MAX_TEAM else 1280

FLAG_EXACT = 1
FLAG_INJECT_COEF = 2
FLAG_INJECT_U = 4
FLAG_TRACE = 8
FLAG_NO_TABLE = 16
FLAG_CTAS3 = 32

GMM_MAX_COMPONENTS = 16
GMM_MAX_POINTS = 26624
GMM_CONVERGED, GMM_NOT_CONVERGED, GMM_ILL_DEFINED = 0, 1, 2

STATUS_OK = 0
STATUS_NONFINITE = 1
STATUS_TIMEOUT = 2


class BrtaError(RuntimeError):
    pass


class Caps(C.Structure):
    _fields_ = [('abi_version', C.c_int32), ('cc_major', C.c_int32), ('cc_minor', C.c_int32),
                ('sm_count', C.c_int32), ('max_smem_per_cta', C.c_int32),
                ('threads_per_cta', C.c_int32), ('max_ncomp', C.c_int32),
                ('mailbox_max_team', C.c_int32)]


class LaunchInfo(C.Structure):
    _fields_ = [('ctas_per_sm', C.c_int32), ('regs_per_thread', C.c_int32),
                ('static_smem', C.c_int32), ('kernel_ncomp', C.c_int32)]


class Task(C.Structure):
    _fields_ = [('chain', C.c_int32), ('team_size', C.c_int32), ('team_rank', C.c_int32),
                ('quad_begin', C.c_int32), ('quad_count', C.c_int32), ('order', C.c_int32)]


class Batch(C.Structure):
    _fields_ = [
        ('n_chains', C.c_int32), ('ncomp', C.c_int32), ('niter', C.c_int32), ('thin', C.c_int32),
        ('tick_bytes', C.c_int32), ('flags', C.c_uint32), ('seed', C.c_uint64),
        ('ticks', C.c_void_p), ('tick_offset', C.c_void_p), ('perm', C.c_void_p), ('perm_offset', C.c_void_p),
        ('n_data', C.c_void_p), ('max_tick', C.c_void_p),
        ('chain_id', C.c_void_p), ('ts', C.c_void_p), ('whyper', C.c_void_p),
        ('rhyper', C.c_void_p), ('init_c', C.c_void_p), ('init_a', C.c_void_p),
        ('mcweights', C.c_void_p), ('mcrates', C.c_void_p), ('indicator', C.c_void_p),
        ('ind_offset', C.c_void_p), ('ind_stride', C.c_void_p), ('status', C.c_void_p),
        ('inj_c', C.c_void_p), ('inj_a', C.c_void_p), ('inj_u', C.c_void_p),
        ('inj_u_offset', C.c_void_p), ('trace_nk', C.c_void_p), ('trace_tk', C.c_void_p),
        ('tasks', C.c_void_p), ('cta_task_begin', C.c_void_p), ('grid_ctas', C.c_int32),
        ('slice_cap_quads', C.c_int32),
        ('n_shards', C.c_int32), ('shard_rank', C.c_int32), ('shard_mailbox', C.c_void_p),
        ('exchange', C.c_void_p), ('exch_offset', C.c_void_p),
        ('task_cycles', C.c_void_p),
        ('iter_begin', C.c_int32), ('iter_end', C.c_int32), ('final_c', C.c_void_p), ('final_a', C.c_void_p),
        ('watchdog_ns', C.c_uint64), ('device', C.c_int32), ('progress_rows', C.c_int32), ('progress', C.c_void_p),
    ]


class GmmBatch(C.Structure):
    _fields_ = [('n_problems', C.c_int32), ('n_init', C.c_int32), ('max_iter', C.c_int32), ('kmeans_max_iter', C.c_int32),
                ('max_points', C.c_int32), ('class_mask', C.c_int32),
                ('tol', C.c_double), ('reg_covar', C.c_double), ('kmeans_tol', C.c_double), ('seed', C.c_uint64),
                ('x', C.c_void_p), ('offsets', C.c_void_p), ('n_components', C.c_void_p), ('problem_id', C.c_void_p),
                ('init_params', C.c_void_p), ('lower_bound', C.c_void_p), ('n_iter', C.c_void_p), ('status', C.c_void_p),
                ('params', C.c_void_p), ('init_out', C.c_void_p)]


EXPORTS = ('brta_abi_version', 'brta_last_error', 'brta_query', 'brta_gibbs_launch_info',
           'brta_gibbs_run_batch', 'brta_philox_fill', 'brta_mufu_probe', 'brta_enable_peer_access',
           'brta_pindicator_counts', 'brta_gamma_fill', 'brta_shard_mailbox_create', 'brta_shard_mailbox_open',
           'brta_shard_mailbox_close', 'brta_shard_mailbox_destroy', 'brta_shard_mailbox_clear',
           'brta_gmm_fit_batch', 'brta_gmm_predict')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC']
SWEEP_NCOMPS = (2, 3, 4, 5, 6, 8, 10, 12, 15, 16, 20, 24, 30, 32)      # BRTA_FOR_EACH_K of csrc/brta_gibbs.cu
HEADERS = [os.path.join(CSRC, f) for f in ('brta_sweep.cuh', 'brta_math.cuh', 'brta_rng.cuh', 'brta_host.h')] + \
          [os.path.join(INCLUDE, 'basicrta_b200.h')]


def _newer(target, deps):
    return os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(d) for d in deps)


def build(force=False, verbose=False, extra_flags=(), lib_path=None, ncomps=SWEEP_NCOMPS, jobs=None):
    """Compile the library for sm_100a into ``libbrta_gibbs.so`` (in-tree): the host side
    (``brta_gibbs.cu``), ``brta_pindicator.cu`` and one instantiation unit of the sweep kernel per
    supported K (``brta_sweep_inst.cu -DBRTA_INST_K=k``), the units in parallel."""
    from concurrent.futures import ThreadPoolExecutor
    lib_path = lib_path or LIB_PATH
    tag = ('_' + str(abs(hash(tuple(extra_flags))) % 10 ** 8)) if extra_flags else ''
    units = [('host' + tag, os.path.join(CSRC, 'brta_gibbs.cu'), []),
             ('pindicator' + tag, os.path.join(CSRC, 'brta_pindicator.cu'), []),
             ('wide' + tag, os.path.join(CSRC, 'brta_wide.cu'), []),
             ('gmm' + tag, os.path.join(CSRC, 'brta_gmm.cu'), ['-fmad=false'])]
    units += [(f'sweep_k{k}' + tag, os.path.join(CSRC, 'brta_sweep_inst.cu'), [f'-DBRTA_INST_K={k}']) for k in ncomps]
    objs = [os.path.join(OBJ_DIR, name + '.o') for name, _, _ in units]
    if not force and _newer(lib_path, HEADERS + [u[1] for u in units]):
        return lib_path
    os.makedirs(OBJ_DIR, exist_ok=True)
    flags = NVCC_FLAGS + list(extra_flags) + (['-Xptxas', '-v'] if verbose else [])
    if tuple(ncomps) != SWEEP_NCOMPS:
        flags = flags + ['-DBRTA_ONLY_K15']

    def compile_unit(unit_obj):
        (name, src, defs), obj = unit_obj
        if not force and _newer(obj, HEADERS + [src]):
            return ''
        res = subprocess.run(['nvcc'] + flags + defs + ['-c', '-o', obj, src], capture_output=True, text=True)
        if res.returncode != 0:
            raise BrtaError(f'nvcc failed on {name}:\n' + res.stdout + res.stderr)
        return res.stderr

    with ThreadPoolExecutor(max_workers=jobs or min(len(units), os.cpu_count() or 1)) as pool:
        logs = list(pool.map(compile_unit, zip(units, objs)))
    res = subprocess.run(['nvcc', '-shared', '-o', lib_path] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a'],
                         capture_output=True, text=True)
    if res.returncode != 0:
        raise BrtaError('nvcc link failed:\n' + res.stdout + res.stderr)
    if verbose:
        print('\n'.join(l for l in logs if l))
    return lib_path


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BrtaError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; '
                        f'g.build()"` (nvcc, sm_100a). basicrta_b200 has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    lib.brta_abi_version.restype = C.c_int
    lib.brta_last_error.restype = C.c_char_p
    lib.brta_query.argtypes = [C.c_int, C.POINTER(Caps)]
    lib.brta_gibbs_launch_info.argtypes = [C.c_int, C.c_int, C.c_uint32, C.c_int, C.POINTER(LaunchInfo)]
    lib.brta_gibbs_run_batch.argtypes = [C.POINTER(Batch), C.c_void_p]
    lib.brta_philox_fill.argtypes = [C.c_void_p, C.c_int64, C.c_uint32, C.c_uint32, C.c_uint32,
                                     C.c_uint32, C.c_uint64, C.c_void_p]
    lib.brta_enable_peer_access.argtypes = [C.c_int, C.c_int]
    lib.brta_mufu_probe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.brta_pindicator_counts.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                           C.c_int32, C.c_void_p, C.c_void_p]
    lib.brta_gamma_fill.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint64,
                                    C.c_void_p]
    lib.brta_gmm_fit_batch.argtypes = [C.POINTER(GmmBatch), C.c_void_p]
    lib.brta_gmm_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]
    lib.brta_shard_mailbox_create.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_char_p]
    lib.brta_shard_mailbox_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.brta_shard_mailbox_close.argtypes = [C.c_int, C.c_void_p]
    lib.brta_shard_mailbox_clear.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.brta_shard_mailbox_destroy.argtypes = [C.c_int, C.c_void_p]
    for name in EXPORTS:
        getattr(lib, name)
    if lib.brta_abi_version() != ABI_VERSION:
        raise BrtaError('libbrta_gibbs.so ABI version mismatch; rebuild')
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().brta_last_error().decode()
        raise BrtaError(f'{what} failed (rc={rc}): {msg}')


def query(device):
    caps = Caps()
    check(load().brta_query(int(device), C.byref(caps)), 'brta_query')
    return caps


def launch_info(device, ncomp, flags, slice_cap_quads):
    info = LaunchInfo()
    check(load().brta_gibbs_launch_info(int(device), int(ncomp), int(flags), int(slice_cap_quads),
                                        C.byref(info)), 'brta_gibbs_launch_info')
    return info
