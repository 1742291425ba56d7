"""ctypes binding of libipp_b200.so (include/ipp_b200.h) + the torch plumbing around it.

torch is used for device memory, streams and (elsewhere) torch.distributed only; every
numerical step goes through the C ABI.  There is NO CPU fallback: importing this module
without the built library, or calling into it without a CUDA device, raises.
"""
import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('IPP_B200_LIB') or os.path.join(_HERE, 'csrc', 'libipp_b200.so')   # env override: kernel-variant experiments

IPP_MAX_DIM = 3
IPP_APPEND_MAX_K = 32
IPP_ROLLOUT_MAX_POINTS = 128
IPP_ROLLOUT_MAX_LEVELS = 32
REWARD_UCB, REWARD_EI, REWARD_MES, REWARD_NAIVE, REWARD_NAIVE_VALUE, REWARD_INFO, REWARD_HOTSPOT = 0, 1, 2, 3, 4, 5, 6
IPP_ROLLOUT_INFO_MAX_POINTS = 32
VAR_WINV_SYM, VAR_WINV_FULL, VAR_CHOL = 0, 1, 2


class IppRbf(ctypes.Structure):
    _fields_ = [('dim', ctypes.c_int32), ('variance', ctypes.c_double),
                ('inv_lengthscale', ctypes.c_double * IPP_MAX_DIM)]


class IppMtState(ctypes.Structure):
    """MT19937 state of CPython's `random` or numpy's legacy global RandomState (include/ipp_b200.h: ipp_mt_state)."""
    _fields_ = [('key', ctypes.c_uint32 * 624), ('pos', ctypes.c_int32), ('has_gauss', ctypes.c_int32),
                ('gauss', ctypes.c_double)]


class IppTreeParams(ctypes.Structure):
    _fields_ = [('budget', ctypes.c_int32), ('max_depth', ctypes.c_int32), ('kind', ctypes.c_int32),
                ('frontier_size', ctypes.c_int32), ('keep_every', ctypes.c_int32), ('n_obstacles', ctypes.c_int32),
                ('c', ctypes.c_double), ('reward_param', ctypes.c_double), ('time', ctypes.c_double),
                ('horizon', ctypes.c_double), ('turning_radius', ctypes.c_double), ('dense_step', ctypes.c_double),
                ('border', ctypes.c_double), ('extent', ctypes.c_double * 4),
                ('angles', ctypes.c_void_p), ('obstacles', ctypes.c_void_p), ('prior_scale', ctypes.c_void_p),
                ('naive_locs', ctypes.c_void_p), ('n_naive', ctypes.c_int32), ('rff_features', ctypes.c_int32),
                ('rff_W', ctypes.c_void_p), ('rff_b', ctypes.c_void_p), ('rff_theta', ctypes.c_void_p),
                ('naive_maxes', ctypes.c_void_p), ('rff_vals_dev', ctypes.c_void_p), ('rff_vals_host', ctypes.c_void_p),
                ('rff_work', ctypes.c_void_p), ('generator', ctypes.c_int32), ('tree_type', ctypes.c_int32),
                ('sample_step', ctypes.c_double), ('log_table', ctypes.c_void_p),
                ('info_kern', ctypes.c_void_p), ('info_winv', ctypes.c_void_p), ('info_ldw', ctypes.c_int64),
                ('info_param', ctypes.c_double)]


IPP_TREE_STAGE_DOUBLES = 128 * (IPP_MAX_DIM + 1) + 6 * 32 + 32


def py_rng_state():
    """CPython `random` state -> IppMtState (and the trailing gauss_next, which the search never touches)."""
    import random
    version, internal, gauss_next = random.getstate()
    st = IppMtState()
    st.key[:] = internal[:624]
    st.pos = internal[624]
    return st, (version, gauss_next)


def set_py_rng_state(st, extra):
    import random
    random.setstate((extra[0], tuple(st.key) + (int(st.pos),), extra[1]))


def np_rng_state():
    """numpy legacy global RandomState -> IppMtState."""
    name, key, pos, has_gauss, cached = np.random.get_state()
    st = IppMtState()
    ctypes.memmove(st.key, np.ascontiguousarray(key, dtype=np.uint32).ctypes.data, 624 * 4)
    st.pos, st.has_gauss, st.gauss = int(pos), int(has_gauss), float(cached)
    return st


def set_np_rng_state(st):
    key = np.frombuffer(bytes(st.key), dtype=np.uint32).copy()
    np.random.set_state(('MT19937', key, int(st.pos), int(st.has_gauss), float(st.gauss)))


class IppError(RuntimeError):
    pass


_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_D = ctypes.c_double
_INT = ctypes.c_int
_RBF = ctypes.POINTER(IppRbf)

# name -> (restype, argtypes); must list every symbol include/ipp_b200.h declares
SIGNATURES = {
    'ipp_last_error': (ctypes.c_char_p, []),
    'ipp_version': (_INT, []),
    'ipp_sm_count': (_INT, []),
    'ipp_rbf_cross': (_INT, [_RBF, _P, _I64, _P, _I64, _P, _I64, _P]),
    'ipp_gp_factor_workspace': (_I64, [_I64]),
    'ipp_gp_factor': (_INT, [_RBF, _P, _P, _I64, _D, _P, _P, _I64, _P, _P, _P, _P, _P]),
    'ipp_pdinv_workspace': (_I64, [_I64]),
    'ipp_pdinv': (_INT, [_P, _I64, _I64, _P, _P, _I64, _P, _P, _P, _P]),
    'ipp_gp_append_workspace': (_I64, [_I64, _I64]),
    'ipp_gp_append': (_INT, [_RBF, _P, _P, _I64, _I64, _D, _P, _I64, _P, _I64, _P, _P, _P, _P]),
    'ipp_gp_predict_workspace': (_I64, [_I64, _I64]),
    'ipp_gp_predict': (_INT, [_RBF, _P, _I64, _P, _P, _I64, _INT, _D, _P, _I64, _P, _P, _P, _P]),
    'ipp_i8_slice_bytes': (_I64, [_I64, _I64, _INT]),
    'ipp_i8_slice': (_INT, [_P, _I64, _I64, _I64, _D, _INT, _P, _P]),
    'ipp_gp_predict_i8_workspace': (_I64, [_I64, _I64]),
    'ipp_gp_predict_i8': (_INT, [_RBF, _P, _I64, _P, _P, _D, _INT, _D, _P, _I64, _P, _P, _P, _P]),
    'ipp_profile_enable': (_INT, [_INT]),
    'ipp_profile_read': (_INT, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
    'ipp_gp_predict_cov_workspace': (_I64, [_I64, _I64]),
    'ipp_gp_predict_cov': (_INT, [_RBF, _P, _I64, _P, _P, _I64, _D, _P, _I64, _P, _P, _I64, _P, _P]),
    'ipp_reward_paths': (_INT, [_INT, _P, _P, _P, _I64, _P, _INT, _P, _I64, _P, _P, _I64, _P]),
    'ipp_rollout_workspace': (_I64, [_I64, _I64]),
    'ipp_rollout_rewards': (_INT, [_RBF, _P, _I64, _P, _P, _I64, _D, _D, _INT, _P, _INT, _P, _I64, _P, _P, _P, _P, _INT, _P, _P,
                                   _P, _P]),
    'ipp_rollout_mode': (_INT, [_INT]),
    'ipp_rollout_debug_clocks': (_INT, [_P]),
    'ipp_rff_workspace': (_I64, [_I64, _I64, _I64, _INT]),
    'ipp_rff_eval_argmax': (_INT, [_P, _P, _P, _I64, _I64, _INT, _INT, _P, _I64, _P, _P, _P, _P, _P]),
    'ipp_rff_mesh_workspace': (_I64, [_I64, _I64, _I64, _I64]),
    'ipp_rff_eval_argmax_mesh': (_INT, [_P, _P, _P, _I64, _I64, _INT, _P, _I64, _P, _I64, _D, _P, _P, _P, _P, _P]),
    'ipp_rff_features': (_INT, [_P, _P, _I64, _INT, _P, _I64, _D, _P, _I64, _P]),
    'ipp_rff_theta_workspace': (_I64, [_I64, _I64, _I64]),
    'ipp_rff_theta_draw': (_INT, [_P, _P, _I64, _INT, _P, _P, _I64, _D, _D, _P, _I64, _P, _P, _P, _P]),
    'ipp_dgemm_nt': (_INT, [_I64, _I64, _I64, _D, _P, _I64, _P, _I64, _D, _P, _I64, _P, _I64, _P]),
    'ipp_potrf_workspace': (_I64, [_I64]),
    'ipp_potrf_debug_clocks': (_INT, [_P]),
    'ipp_potrf_lower': (_INT, [_P, _I64, _I64, _P, _P, _P, _P]),
    'ipp_info_gain_workspace': (_I64, [_I64, _I64]),
    'ipp_dubins_path_set': (_INT, [_P, _P, _INT, _D, _D, _INT, _P, _D, _P, _INT, _P, _P, _INT, _P, _P, _INT]),
    'ipp_dubins_paths_device': (_INT, [_P, _P, _I64, _D, _D, _INT, _P, _D, _P, _INT, _P, _P, _INT, _P, _P]),
    'ipp_tree_search_dpw': (_INT, [ctypes.POINTER(IppTreeParams), _P, _RBF, _P, _I64, _P, _P, _I64, _D, _D, _P, _I64,
                                   ctypes.POINTER(IppMtState), ctypes.POINTER(IppMtState), _P, _P, _P, _P, _P, _P, _P, _P,
                                   _P, _P]),
    'ipp_abi_struct_sizes': (_INT, [_P]),
    'ipp_tree_last_timing': (_INT, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    'ipp_tree_last_launch_us': (_INT, [ctypes.POINTER(ctypes.c_double)]),
    'ipp_rng_py_choices': (_INT, [ctypes.POINTER(IppMtState), _P, _INT, _P]),
    'ipp_rng_np_normals': (_INT, [ctypes.POINTER(IppMtState), _INT, _P]),
    'ipp_tree_path_set': (_INT, [ctypes.POINTER(IppTreeParams), _P, _P, _P, _P, _P, _INT]),
    'ipp_info_gain_paths': (_INT, [_RBF, _P, _I64, _P, _I64, _P, _P, _I64, _I64, _P, _P, _P, _P]),
}


def source_hash():
    """SHA-256 over the files the library is built from, in the Makefile's order (csrc/Makefile: HASHED)."""
    import glob
    import hashlib
    csrc = os.path.join(_HERE, 'csrc')
    files = sorted(os.path.basename(f) for f in glob.glob(os.path.join(csrc, '*.cu')) + glob.glob(os.path.join(csrc, '*.cuh')))
    paths = [os.path.join(csrc, f) for f in files] + [os.path.join(csrc, 'Makefile'),
                                                     os.path.join(_HERE, '..', 'include', 'ipp_b200.h')]
    h = hashlib.sha256()
    for path in paths:
        with open(path, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def _check_source_stamp():
    """Refuse a library that was not built from the sources beside it (the .so is git-ignored and travels as a built
    artefact: a stale one would otherwise ship silently).  Skipped for an explicit IPP_B200_LIB override."""
    if os.environ.get('IPP_B200_LIB'):
        return
    stamp = LIB_PATH + '.srchash'
    built = open(stamp).read().strip() if os.path.exists(stamp) else None
    if built != source_hash():
        raise ImportError(
            'informative_path_planning_b200: %s was not built from the current csrc/*.cu, *.cuh, Makefile and '
            'include/ipp_b200.h (%s). Rebuild it with `python -c \'import __graft_entry__ as g; g.build()\'`.'
            % (LIB_PATH, 'no source stamp beside it' if built is None else 'source stamp differs'))


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "informative_path_planning_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    _check_source_stamp()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def _check_struct_sizes():
    """The ctypes mirrors above against sizeof() in the library that was actually loaded (a stale .so or an edited
    header would otherwise corrupt memory silently)."""
    from .obstacles import IppObstacleBox
    out = (ctypes.c_int32 * 4)()
    if lib.ipp_abi_struct_sizes(out) != 0:
        raise ImportError('libipp_b200: ipp_abi_struct_sizes failed')
    want = [ctypes.sizeof(IppRbf), ctypes.sizeof(IppObstacleBox), ctypes.sizeof(IppMtState), ctypes.sizeof(IppTreeParams)]
    if list(out) != want:
        raise ImportError('libipp_b200.so and its Python binding disagree on struct sizes (library %s, binding %s): '
                          'rebuild with __graft_entry__.build()' % (list(out), want))


_check_struct_sizes()


def check(rc):
    if rc != 0:
        raise IppError('libipp_b200 error %d: %s' % (rc, lib.ipp_last_error().decode('utf-8', 'replace')))


def device():
    if not torch.cuda.is_available():
        raise IppError('informative_path_planning_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def round_up(x, m):
    return (x + m - 1) // m * m


def make_rbf(dim, variance, lengthscale):
    k = IppRbf()
    k.dim = int(dim)
    k.variance = float(np.asarray(variance, dtype=float).reshape(-1)[0])
    ls = np.asarray(lengthscale, dtype=float).reshape(-1)
    for d in range(IPP_MAX_DIM):
        if d < dim:
            k.inv_lengthscale[d] = 1.0 / float(ls[d] if ls.size > 1 else ls[0])
        else:
            k.inv_lengthscale[d] = 0.0
    return k


_workspaces = {}


def workspace(n_doubles, slot='main'):
    """Grow-only float64 scratch on the current device.  Calls are stream-ordered on the current
    stream, so successive library calls may reuse it."""
    dev = device()
    key = (dev.index, slot)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < n_doubles:
        buf = None
        _workspaces.pop(key, None)
        buf = torch.empty(int(n_doubles), dtype=torch.float64, device=dev)
        _workspaces[key] = buf
    return buf


def to_device(a):
    """numpy (any float layout) -> contiguous float64 CUDA tensor (through pinned memory for big arrays)."""
    if isinstance(a, torch.Tensor):
        return a.to(device=device(), dtype=torch.float64).contiguous()
    a = np.ascontiguousarray(a, dtype=np.float64)
    t = torch.from_numpy(a)
    if t.numel() < (1 << 21):
        return t.to(device())                        # < 16 MB: a pageable copy beats pinning a fresh buffer (~5 ms per cudaHostAlloc)
    return _staged_upload(t)


_staging = {}


def _staged_upload(t):
    """Large host array -> device through a reused pinned staging buffer (grow-only, two slots, each guarded by the
    event of its last DMA so a slot is never overwritten while the copy engine still reads it)."""
    dev = device()
    st = _staging.setdefault(dev.index, {'bufs': [None, None], 'events': [None, None], 'next': 0})
    i = st['next']
    st['next'] = 1 - i
    if st['events'][i] is not None:
        st['events'][i].synchronize()
    buf = st['bufs'][i]
    if buf is None or buf.numel() < t.numel():
        buf = torch.empty(t.numel(), dtype=torch.float64).pin_memory()
        st['bufs'][i] = buf
    view = buf[:t.numel()].view(t.shape)
    view.copy_(t)
    out = view.to(dev, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    st['events'][i] = ev
    return out
