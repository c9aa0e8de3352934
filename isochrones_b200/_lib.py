"""ctypes binding of ``libisochrones_b200.so`` (C ABI: ``include/isochrones_b200.h``).

There is no CPU fallback: if the shared library is missing or no CUDA device is present, every compute
entry point raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"`` or
``make -C isochrones_b200/csrc``.
"""
import ctypes as C
import os
import threading
import weakref

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# ISO_B200_LIB selects another build of the same library (kernel tuning variants); never a different backend
LIB_PATH = os.environ.get("ISO_B200_LIB") or os.path.join(HERE, "lib", "libisochrones_b200.so")

ISO_MAX_BANDS = 16
ISO_MAX_COMP = 3
ISO_MAX_DIM = 4
ISO_MP_NCOLS = 8

ISO_PRIOR_FLAT, ISO_PRIOR_FLATLOG, ISO_PRIOR_POWERLAW, ISO_PRIOR_GAUSSIAN = 1, 2, 3, 4
ISO_PRIOR_LOGNORMAL, ISO_PRIOR_FEH, ISO_PRIOR_BROKEN = 5, 6, 7
ISO_PF_BOUNDED, ISO_PF_HAS_BOUNDS, ISO_PF_LOCAL = 1, 2, 4

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class IsoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("isochrones_b200 error %d: %s" % (code, msg))
        self.code = code


class IsoPriorLeaf(C.Structure):
    _fields_ = [("kind", C.c_int32), ("flags", C.c_int32), ("lo", C.c_double), ("hi", C.c_double),
                ("norm", C.c_double), ("a", C.c_double * 4), ("k", C.c_double * 4)]


class IsoPrior(C.Structure):
    _fields_ = [("self", IsoPriorLeaf), ("n_comp", C.c_int32), ("pad_", C.c_int32),
                ("breakpoints", C.c_double * (ISO_MAX_COMP - 1)), ("norms", C.c_double * ISO_MAX_COMP),
                ("lognorms", C.c_double * ISO_MAX_COMP), ("inv_norms", C.c_double * ISO_MAX_COMP),
                ("inv_norm", C.c_double), ("comp", IsoPriorLeaf * ISO_MAX_COMP)]


class IsoModel(C.Structure):
    _fields_ = [
        ("n_stars", C.c_int32), ("eep_replaces_age", C.c_int32), ("index_order", C.c_int32 * 5),
        ("n_bands", C.c_int32), ("band_col", C.c_int32 * ISO_MAX_BANDS),
        ("has_plax", C.c_int32), ("has_nu_max", C.c_int32), ("has_delta_nu", C.c_int32), ("pad_", C.c_int32),
        ("spec_val", C.c_double * 3), ("spec_unc", C.c_double * 3),
        ("mag_val", C.c_double * ISO_MAX_BANDS), ("mag_unc", C.c_double * ISO_MAX_BANDS),
        ("plax", C.c_double), ("plax_unc", C.c_double),
        ("nu_max", C.c_double), ("nu_max_unc", C.c_double), ("delta_nu", C.c_double), ("delta_nu_unc", C.c_double),
        ("eep_lo", C.c_double), ("eep_hi", C.c_double), ("eep_norm", C.c_double),
        ("eep_has_bounds", C.c_int32), ("pad2_", C.c_int32),
        ("eep_orig", IsoPrior), ("mass", IsoPrior), ("age", IsoPrior), ("feh", IsoPrior),
        ("distance", IsoPrior), ("AV", IsoPrior),
    ]


# every symbol include/isochrones_b200.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SIGNATURES = {
    "iso_abi_version": (C.c_int, []),
    "iso_struct_size": (C.c_int64, [C.c_int]),
    "iso_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "iso_ctx_create": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "iso_ctx_destroy": (C.c_int, [_VP]),
    "iso_ctx_sync": (C.c_int, [_VP]),
    "iso_last_error": (C.c_char_p, [_VP]),
    "iso_ctx_info": (C.c_int, [_VP, C.c_char_p, C.POINTER(C.c_int), c_int64_p, c_int64_p, C.POINTER(C.c_int)]),
    "iso_dev_alloc": (C.c_int, [_VP, C.c_int64, C.POINTER(_VP)]),
    "iso_dev_free": (C.c_int, [_VP, _VP]),
    "iso_host_alloc": (C.c_int, [_VP, C.c_int64, C.POINTER(_VP)]),
    "iso_host_free": (C.c_int, [_VP, _VP]),
    "iso_memcpy_h2d": (C.c_int, [_VP, _VP, _VP, C.c_int64]),
    "iso_memcpy_d2h": (C.c_int, [_VP, _VP, _VP, C.c_int64]),
    "iso_memset": (C.c_int, [_VP, _VP, C.c_int, C.c_int64]),
    "iso_timer_start": (C.c_int, [_VP]),
    "iso_timer_stop": (C.c_int, [_VP, C.POINTER(C.c_float)]),
    "iso_launch_count": (C.c_int, [_VP, c_int64_p]),
    "iso_grid_stage": (C.c_int, [_VP, c_double_p, C.c_int, c_int64_p, C.POINTER(c_double_p), C.POINTER(_VP)]),
    "iso_grid_repack": (C.c_int, [_VP, _VP, c_int32_p, C.c_int, C.c_int, C.POINTER(_VP)]),
    "iso_grid_destroy": (C.c_int, [_VP, _VP]),
    "iso_grid_shape": (C.c_int, [_VP, C.POINTER(C.c_int), c_int64_p]),
    "iso_interp_values": (C.c_int, [_VP, _VP, C.POINTER(c_double_p), C.c_int64, c_int32_p, C.c_int, c_double_p]),
    "iso_interp_mags": (C.c_int, [_VP, _VP, _VP, c_int32_p, C.c_int, C.c_int, C.c_int, C.c_int, c_int32_p, C.c_int,
                                  c_double_p, C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p]),
    "iso_interp_mags_cols": (C.c_int, [_VP, _VP, _VP, c_int32_p, C.c_int, C.c_int, C.c_int, C.c_int, c_int32_p, C.c_int,
                                       C.POINTER(c_double_p), C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p]),
    "iso_interp_values_device": (C.c_int, [_VP, _VP, C.POINTER(_VP), C.c_int64, c_int32_p, C.c_int, _VP]),
    "iso_interp_mags_device": (C.c_int, [_VP, _VP, _VP, c_int32_p, C.c_int, C.c_int, C.c_int, C.c_int, c_int32_p, C.c_int,
                                         C.POINTER(_VP), C.c_int64, _VP, _VP, _VP, _VP]),
    "iso_interp_eeps": (C.c_int, [_VP, _VP, C.c_int, c_int32_p, c_double_p, c_double_p, c_double_p, C.c_int64, c_double_p]),
    "iso_prior_eval": (C.c_int, [_VP, C.POINTER(IsoPrior), C.c_int, c_double_p, C.c_int64, c_double_p]),
    "iso_models_stage": (C.c_int, [_VP, C.POINTER(IsoModel), C.c_int, C.POINTER(_VP)]),
    "iso_models_destroy": (C.c_int, [_VP, _VP]),
    "iso_lnpost_batch": (C.c_int, [_VP, _VP, _VP, _VP, c_int32_p, c_double_p, C.c_int64, c_double_p, c_double_p,
                                   c_double_p]),
    "iso_lnpost_batch_device": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, C.c_int64, _VP, _VP, _VP]),
    "iso_mnest_prior": (C.c_int, [_VP, c_double_p, c_double_p, C.c_int, c_double_p, C.c_int64]),
    "iso_mnest_lnpost_batch": (C.c_int, [_VP, _VP, _VP, _VP, c_double_p, c_double_p, c_double_p, C.c_int64, c_double_p,
                                         c_double_p, c_double_p]),
    "iso_lnpost_prior_draws": (C.c_int, [_VP, _VP, _VP, _VP, c_double_p, c_double_p, C.c_uint64, C.c_int64, C.c_int64,
                                         c_double_p, c_double_p]),
    "iso_lnpost_cube_device": (C.c_int, [_VP, _VP, _VP, _VP, c_double_p, c_double_p, _VP, C.c_int, C.c_uint64, C.c_int64,
                                         C.c_int64, _VP, _VP]),
    "iso_sampler_create": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, C.c_int, c_double_p, C.c_uint64, C.c_double,
                                     C.POINTER(_VP)]),
    "iso_sampler_run": (C.c_int, [_VP, _VP, C.c_int, C.c_int, c_double_p, c_double_p]),
    "iso_sampler_state": (C.c_int, [_VP, _VP, c_double_p, c_double_p, c_int64_p, c_int64_p]),
    "iso_sampler_reset": (C.c_int, [_VP, _VP]),
    "iso_sampler_set_moments": (C.c_int, [_VP, _VP, C.c_int]),
    "iso_sampler_moments": (C.c_int, [_VP, _VP, c_double_p, C.POINTER(_VP)]),
    "iso_sampler_destroy": (C.c_int, [_VP, _VP]),
    "iso_ensemble_create": (C.c_int, [_VP, _VP, _VP, _VP, C.c_int, c_double_p, C.c_uint64, C.c_double, C.c_int, C.c_int,
                                      C.POINTER(_VP)]),
    "iso_ensemble_export": (C.c_int, [_VP, _VP, _VP]),
    "iso_ensemble_connect": (C.c_int, [_VP, _VP, _VP]),
    "iso_ensemble_set_timeout": (C.c_int, [_VP, _VP, C.c_double]),
    "iso_ensemble_run": (C.c_int, [_VP, _VP, C.c_int, C.c_int, c_double_p, c_double_p]),
    "iso_ensemble_state": (C.c_int, [_VP, _VP, c_double_p, c_double_p, c_int64_p, c_int64_p]),
    "iso_ensemble_destroy": (C.c_int, [_VP, _VP]),
    "iso_nccl_unique_id": (C.c_int, [_VP]),
    "iso_nccl_init": (C.c_int, [_VP, _VP, C.c_int, C.c_int]),
    "iso_nccl_destroy": (C.c_int, [_VP]),
    "iso_peer_create": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int64, C.POINTER(_VP)]),
    "iso_peer_export": (C.c_int, [_VP, _VP, _VP]),
    "iso_peer_connect": (C.c_int, [_VP, _VP, _VP]),
    "iso_lnpost_allgather_device": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _VP, C.c_int64, _VP, C.POINTER(_VP)]),
    "iso_peer_set_timeout": (C.c_int, [_VP, _VP, C.c_double]),
    "iso_peer_check": (C.c_int, [_VP, _VP]),
    "iso_peer_destroy": (C.c_int, [_VP, _VP]),
    "iso_allgather_f64": (C.c_int, [_VP, _VP, C.c_int64, _VP]),
}

_lib = None
_lock = threading.Lock()


def lib():
    """The loaded shared library (raises if it has not been built — there is no fallback path)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        "%s not found: build it with `make -C isochrones_b200/csrc` (or __graft_entry__.build()); "
                        "isochrones_b200 has no CPU fallback" % LIB_PATH)
                L = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(L, name)
                    fn.restype = res
                    fn.argtypes = args
                for which, struct in enumerate((IsoPriorLeaf, IsoPrior, IsoModel)):
                    if L.iso_struct_size(which) != C.sizeof(struct):
                        raise ImportError("struct layout mismatch between _lib.py and %s (code %d)" % (LIB_PATH, which))
                _lib = L
    return _lib


def dp(a):
    return a.ctypes.data_as(c_double_p)


def ip(a):
    return a.ctypes.data_as(c_int32_p)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One CUDA context wrapper per GPU (``iso_ctx``)."""

    def __init__(self, device=0):
        self.handle = _VP()
        self.device = device
        rc = lib().iso_ctx_create(int(device), C.byref(self.handle))
        if rc != 0:
            raise IsoError(rc, (lib().iso_last_error(None) or b"").decode())

    def check(self, rc):
        if rc != 0:
            raise IsoError(rc, (lib().iso_last_error(self.handle) or b"").decode())

    def info(self):
        name = C.create_string_buffer(256)
        sm, cc = C.c_int(), C.c_int()
        l2, hbm = C.c_int64(), C.c_int64()
        self.check(lib().iso_ctx_info(self.handle, name, C.byref(sm), C.byref(l2), C.byref(hbm), C.byref(cc)))
        return {"name": name.value.decode(), "sm_count": sm.value, "l2_bytes": l2.value, "hbm_bytes": hbm.value,
                "cc": cc.value}

    def sync(self):
        self.check(lib().iso_ctx_sync(self.handle))

    def launch_count(self):
        n = C.c_int64()
        self.check(lib().iso_launch_count(self.handle, C.byref(n)))
        return n.value

    # ---- raw memory (bench / sampler plumbing) ----------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = _VP()
        self.check(lib().iso_dev_alloc(self.handle, int(nbytes), C.byref(p)))
        return p

    def dev_free(self, p):
        self.check(lib().iso_dev_free(self.handle, p))

    def pinned_empty(self, shape, dtype=np.float64):
        """A page-locked numpy array (direct DMA target of the host-pointer entry points)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = _VP()
        self.check(lib().iso_host_alloc(self.handle, max(n, 1), C.byref(p)))
        buf = (C.c_char * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        # the page-locked block lives exactly as long as some numpy view of it: every view keeps `buf` alive through
        # its base chain, and the finaliser of `buf` returns the block to the driver (iso_host_free)
        weakref.finalize(buf, _free_pinned, self, p.value).atexit = False
        return arr

    def h2d(self, d_ptr, arr):
        arr = np.ascontiguousarray(arr)
        self.check(lib().iso_memcpy_h2d(self.handle, d_ptr, arr.ctypes.data_as(_VP), arr.nbytes))

    def d2h(self, arr, d_ptr):
        assert arr.flags["C_CONTIGUOUS"]
        self.check(lib().iso_memcpy_d2h(self.handle, arr.ctypes.data_as(_VP), d_ptr, arr.nbytes))

    def memset(self, d_ptr, value, nbytes):
        self.check(lib().iso_memset(self.handle, d_ptr, int(value), int(nbytes)))

    def timer_start(self):
        self.check(lib().iso_timer_start(self.handle))

    def timer_stop(self):
        ms = C.c_float()
        self.check(lib().iso_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    def close(self):
        if self.handle:
            lib().iso_ctx_destroy(self.handle)
            self.handle = _VP()


def _free_pinned(ctx, address):
    if ctx.handle:      # a destroyed context has released its allocations already
        try:
            lib().iso_host_free(ctx.handle, _VP(address))
        except Exception:
            pass

_contexts = {}


def default_context(device=None):
    """Process-wide context of a device (LOCAL_RANK selects the device under torchrun)."""
    if device is None:
        device = int(os.environ.get("ISO_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx


def device_count():
    n = C.c_int()
    rc = lib().iso_device_count(C.byref(n))
    return n.value if rc == 0 else 0


class DeviceGrid:
    """A dense grid staged in HBM (``iso_grid``)."""

    def __init__(self, ctx, grid=None, axes=None, handle=None):
        self.ctx = ctx
        if handle is not None:
            self.handle = handle
        else:
            g = f64(grid)
            axes = [f64(a) for a in axes]
            ndim = len(axes)
            assert g.ndim == ndim + 1, "grid must be [n0, .., n_{ndim-1}, ncols]"
            for d in range(ndim):
                assert g.shape[d] == len(axes[d]), "axis %d length does not match the grid" % d
            shape = (C.c_int64 * (ndim + 1))(*g.shape)
            ptrs = (c_double_p * ndim)(*[dp(a) for a in axes])
            self.handle = _VP()
            ctx.check(lib().iso_grid_stage(ctx.handle, dp(g), ndim, shape, ptrs, C.byref(self.handle)))
        nd = C.c_int()
        shp = (C.c_int64 * (ISO_MAX_DIM + 1))()
        ctx.check(lib().iso_grid_shape(self.handle, C.byref(nd), shp))
        self.ndim = nd.value
        self.shape = tuple(shp[i] for i in range(self.ndim + 1))
        self.ncols = self.shape[-1]

    def repack(self, cols, ncols_out=None):
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        ncols_out = len(cols) if ncols_out is None else int(ncols_out)
        h = _VP()
        self.ctx.check(lib().iso_grid_repack(self.ctx.handle, self.handle, ip(cols), len(cols), ncols_out, C.byref(h)))
        return DeviceGrid(self.ctx, handle=h)

    def interp_values(self, xx, icols, out=None):
        xx = [f64(a) for a in xx]
        n = len(xx[0])
        icols = np.ascontiguousarray(icols, dtype=np.int32)
        if out is None:
            out = np.empty((n, len(icols)), dtype=np.float64)
        elif out.shape != (n, len(icols)) or out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous float64 array of shape [N, ncols]")
        ptrs = (c_double_p * self.ndim)(*[dp(a) for a in xx])
        self.ctx.check(lib().iso_interp_values(self.ctx.handle, self.handle, ptrs, n, ip(icols), len(icols), dp(out)))
        return out

    def interp_values_device(self, d_xx, n, icols, d_out):
        """Asynchronous interpolation on device buffers: ``d_xx`` = one device pointer per axis (each ``[n]``),
        ``d_out`` = device ``[n, len(icols)]`` (``iso_interp_values_device``)."""
        icols = np.ascontiguousarray(icols, dtype=np.int32)
        ptrs = (_VP * self.ndim)(*d_xx)
        self.ctx.check(lib().iso_interp_values_device(self.ctx.handle, self.handle, ptrs, int(n), ip(icols), len(icols), d_out))

    def close(self):
        if self.handle:
            lib().iso_grid_destroy(self.ctx.handle, self.handle)
            self.handle = _VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
