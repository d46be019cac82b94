"""Host side of the CUDA engine: device buffers (torch tensors), launches through the C ABI.

torch is used for device memory, streams and torch.distributed only; every number is
produced by the kernels in csrc/ behind include/xrsd_b200.h.  There is no CPU fallback:
without a CUDA device (or without the built library) the calls raise.
"""
import ctypes as C

import numpy as np

from . import abi
from . import lowering

_engine = None


def get_engine():
    global _engine
    if _engine is None:
        _engine = Engine()
    return _engine


def _torch():
    import torch
    return torch


SMEM_CELLS = 90000            # candidate cells the table builder keeps in shared memory (2 B each, beside its lists)
GRID_SCRATCH_BYTES = 2 << 30  # global-memory budget for candidate boxes beyond that


class TableStore(object):
    """Device arrays behind xrsd_tables_t: n_views * n_tables own rows followed by n_shared shared rows
    (include/xrsd_b200.h).  Views of one store (TableSet) share the shared rows: the accepted and the trial point of
    a fit are two views, so a table built for one is there for the other."""

    def __init__(self, torch, device, n_tables, cap_shell, n_pairs, cap_refl=0, n_views=1, n_shared=0):
        self.torch, self.device = torch, device
        self.n_tables, self.cap_shell, self.n_pairs, self.cap_refl = n_tables, cap_shell, n_pairs, cap_refl
        self.n_views, self.n_shared = n_views, n_shared
        self.n_rows = n_rows = n_views * n_tables + n_shared
        i32, f64 = torch.int32, torch.float64
        self.n_shell = torch.zeros(n_rows, dtype=i32, device=device)
        self.n_refl = torch.zeros(n_rows, dtype=i32, device=device)
        self.n_all = torch.zeros(n_rows, dtype=i32, device=device)
        self.status = torch.zeros(n_rows, dtype=i32, device=device)
        self.shell_hkl = torch.zeros((n_rows, cap_shell, 4), dtype=i32, device=device)
        self.shell_geo = torch.zeros((n_rows, cap_shell, n_pairs), dtype=f64, device=device)
        self.refl_hkl = torch.zeros((n_rows, cap_refl, 4), dtype=i32, device=device) if cap_refl else None
        self.shared_tag = torch.zeros(max(1, n_shared), dtype=i32, device=device)
        self.grid, self.grid_counter, self.grid_slots, self.grid_cells = None, torch.zeros(1, dtype=i32, device=device), 0, 0
        self.big, self.cap_big, self.cap_list_hint = None, 0, 0
        self.views = [TableSet(self, k) for k in range(n_views)]

    def reset_shared(self):
        """Forget every shared row: the next build rebuilds the tables it needs."""
        self.shared_tag.zero_()

    def attach_grid(self, n_slots, cap_cells):
        self.grid = self.torch.empty(int(n_slots) * int(cap_cells), dtype=self.torch.uint16, device=self.device)
        self.grid_slots, self.grid_cells = int(n_slots), int(cap_cells)
        for v in self.views:
            v._fill()

    def attach_big(self, cap_big, lib):
        self.big = self.torch.empty(self.n_tables * int(lib.xrsd_table_scratch_bytes(cap_big)), dtype=self.torch.uint8,
                                    device=self.device)
        self.cap_big = int(cap_big)
        for v in self.views:
            v._fill()


class TableSet(object):
    """One view of a TableStore: batch*n_xtal tables, table t in storage row table_of[t]."""

    def __init__(self, store, k):
        self.store, self.k = store, k
        self.n_tables, self.cap_shell, self.n_pairs, self.cap_refl = store.n_tables, store.cap_shell, store.n_pairs, store.cap_refl
        self.own_base = k * store.n_tables
        torch = store.torch
        self.table_of = torch.arange(self.own_base, self.own_base + self.n_tables, dtype=torch.int32, device=store.device)
        self.build_flag = torch.zeros(max(1, self.n_tables), dtype=torch.uint8, device=store.device)
        self.c = abi.Tables()
        self._fill()

    def _fill(self):
        st, c = self.store, self.c
        c.cap_shell, c.cap_refl, c.n_pairs_max, c.cap_list_hint = st.cap_shell, st.cap_refl, st.n_pairs, st.cap_list_hint
        c.n_shell, c.n_refl, c.n_all, c.status = (st.n_shell.data_ptr(), st.n_refl.data_ptr(), st.n_all.data_ptr(),
                                                  st.status.data_ptr())
        c.shell_hkl, c.shell_geo = st.shell_hkl.data_ptr(), st.shell_geo.data_ptr()
        c.refl_hkl = st.refl_hkl.data_ptr() if st.cap_refl else None
        c.big_scratch, c.cap_big = (st.big.data_ptr(), st.cap_big) if st.big is not None else (None, 0)
        c.table_of, c.build_flag = self.table_of.data_ptr(), self.build_flag.data_ptr()
        c.own_base, c.shared_base, c.n_shared = self.own_base, st.n_views * st.n_tables, st.n_shared
        c.shared_tag = st.shared_tag.data_ptr()
        c.grid_scratch = st.grid.data_ptr() if st.grid is not None else None
        c.n_grid_slots, c.cap_grid_cells, c.grid_slot_counter = st.grid_slots, st.grid_cells, st.grid_counter.data_ptr()

    def _rows(self):
        return self.table_of.long()

    # per-table values, gathered through table_of (checks and tests; the kernels follow table_of themselves)
    n_shell = property(lambda self: self.store.n_shell[self._rows()])
    n_refl = property(lambda self: self.store.n_refl[self._rows()])
    n_all = property(lambda self: self.store.n_all[self._rows()])
    status = property(lambda self: self.store.status[self._rows()])
    shell_hkl = property(lambda self: self.store.shell_hkl[self._rows()])
    shell_geo = property(lambda self: self.store.shell_geo[self._rows()])
    refl_hkl = property(lambda self: self.store.refl_hkl[self._rows()] if self.store.refl_hkl is not None else None)

    def shared_fraction(self):
        return float((self.table_of >= self.store.n_views * self.store.n_tables).double().mean().item())


class Engine(object):

    def __init__(self, device=None):
        self.lib = abi.load()
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("xrsdkit_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.torch = torch
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.launches = 0     # kernels launched through this engine (bench.py reports it)

    # -- helpers ------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def capture_stream(self):
        """A side stream per host thread for CUDA-graph captures (capture is not allowed on the default stream)."""
        import threading
        pool = self.__dict__.setdefault("_capture_streams", {})
        key = threading.get_ident()
        if key not in pool:
            pool[key] = self.torch.cuda.Stream(device=self.device)
        return pool[key]

    def to_device(self, arr):
        torch = self.torch
        if isinstance(arr, torch.Tensor):
            return arr.to(device=self.device, dtype=torch.float64).contiguous()
        a = np.ascontiguousarray(arr, dtype=np.float64)
        if not a.flags.writeable:      # e.g. arrays that come out of an .npz: torch wants a writable buffer
            a = a.copy()
        return torch.from_numpy(a).to(self.device)

    def check_grid(self, layout, q):
        """The q passed with a layout must be the grid the layout was lowered for (SystemLayout.check_grid).  Host
        arrays are checked on every call; a device tensor is read back once per tensor object and version."""
        if getattr(layout, "grid", None) is None:
            return
        if isinstance(q, self.torch.Tensor):
            # remembered ON the tensor object (an address can be handed to another tensor by the caching allocator)
            mark = (id(layout), q._version)
            if getattr(q, "_xrsd_grid_ok", None) == mark:
                return
            layout.check_grid(q.detach().cpu().numpy())
            q._xrsd_grid_ok = mark
        else:
            layout.check_grid(np.asarray(q, dtype=np.float64))

    def _n_pairs(self, layout):
        n = 1
        for p in layout.crystalline_pops():
            n = max(n, p.n_species * (p.n_species + 1) // 2)
        return n

    # -- reflection tables --------------------------------------------------------------
    @staticmethod
    def tables_shareable(layout):
        """Every crystalline population is cubic and has no atomic coordinates among its parameters: its tables
        depend on the cell edge only through the integer set of passing shells (xrsd_tables_t)."""
        pops = layout.crystalline_pops()
        return bool(pops) and all(p.lattice == abi.LAT_CUBIC and all(p.p_uvw[a][j] < 0 for a in range(p.n_species) for j in range(3))
                                  for p in pops)

    @staticmethod
    def _cap_refl(export_reflections):
        """rows of refl_hkl per table: 0 (no export), 4096 (True) or the number given"""
        if not export_reflections:
            return 0
        return 4096 if export_reflections is True else int(export_reflections)

    def new_table_store(self, layout, batch, max_cells, cap_shell=None, export_reflections=False, n_views=1, share=True,
                        cap_list=0, cap_big=0):
        has_ops = all(p.use_symmetry and p.n_ops > 0 for p in layout.crystalline_pops())
        if cap_shell is None:
            cap_shell = 128 if has_ops else 1024
        n_shared = 0
        if share and self.tables_shareable(layout):
            half = (int(round(float(max_cells) ** (1.0 / 3.0))) + 1) // 2 + 1       # box half-width n: N <= 3 n^2
            n_shared = min(3 * half * half + 2, 1 << 20)
        st = TableStore(self.torch, self.device, batch * layout.n_xtal, cap_shell, self._n_pairs(layout),
                        self._cap_refl(export_reflections), n_views, n_shared)
        st.cap_list_hint = cap_list
        if cap_big:
            st.attach_big(cap_big, self.lib)
        if max_cells > SMEM_CELLS:
            # boxes beyond shared memory: regions in global memory, as many as the budget allows (shared tables need
            # one per distinct table, own tables one each)
            want = st.n_tables if not n_shared else min(st.n_tables, n_shared)
            st.attach_grid(max(1, min(want, GRID_SCRATCH_BYTES // (2 * int(max_cells)))), int(max_cells))
        for v in st.views:
            v._fill()
        return st

    def build_tables(self, layout, d_params, batch, params_host=None, cap_shell=None, export_reflections=False,
                     max_cells=None, check=True, reuse=None, active=None, share=True):
        """Build the reflection tables of all crystalline populations for a batch.

        `params_host` (numpy [batch, n_params]) is only used to bound the candidate box when `max_cells` is not
        given.  With check=True the status words are read back (one small device->host copy) and capacity
        overflows are retried with larger tables.  `reuse` is a TableSet (a view of a TableStore) to overwrite
        instead of allocating a new store; its shared rows built by earlier calls stay valid (reset_shared()
        forgets them).  export_reflections (True = up to 4096 per table, or a number) also writes the reference's
        reduced hkl list and multiplicities (refl_hkl), for parity checks.  share=False gives every system its own table even where tables could be shared."""
        if layout.n_xtal == 0:
            return None
        torch = self.torch
        if max_cells is None:
            if params_host is None:
                params_host = d_params.detach().cpu().numpy()
            max_cells = lowering.max_cells(layout, params_host)
        n_tables = batch * layout.n_xtal
        cap_refl = self._cap_refl(export_reflections)
        T = None
        if reuse is not None and reuse.n_tables == n_tables and reuse.cap_refl == cap_refl and \
                (cap_shell is None or reuse.cap_shell >= cap_shell) and \
                (max_cells <= SMEM_CELLS or reuse.store.grid_cells >= max_cells) and (share or reuse.store.n_shared == 0):
            T = reuse
        cap_list, cap_big = 0, 0
        while True:
            if T is None:
                k = reuse.k if reuse is not None else 0
                nv = reuse.store.n_views if reuse is not None else 1
                T = self.new_table_store(layout, batch, max_cells, cap_shell, export_reflections, nv, share, cap_list, cap_big).views[k]
            cap_shell = T.cap_shell
            mask, n_over_prev = active, None
            while True:
                abi.check(self.lib.xrsd_reflection_tables(C.byref(layout.c), d_params.data_ptr(), d_params.shape[1], batch,
                                                          C.byref(T.c), min(int(max_cells), SMEM_CELLS),
                                                          mask.data_ptr() if mask is not None else None, self._stream()))
                self.launches += 2
                if not check:
                    return T
                st_t = T.status.view(batch, layout.n_xtal)
                if active is not None:
                    st_t = st_t * active.to(torch.int32)[:, None]
                bits = 0
                if n_tables and int(st_t.max().item()) != 0:
                    # OR of the status words (their maximum alone can hide a bit)
                    flat = st_t.reshape(-1)
                    seen = torch.stack([((flat & bit) != 0).any() for bit in (1, 2, 4, 8)]).cpu().tolist()
                    bits = sum(bit for bit, on in zip((1, 2, 4, 8), seen) if on)
                if bits == abi.TABLE_GRID_OVERFLOW and T.store.grid is not None and T.store.grid_cells >= max_cells:
                    # more large boxes than global regions: the rest is built in further rounds
                    over = ((st_t & abi.TABLE_GRID_OVERFLOW) != 0).any(dim=1)
                    n_over = int(over.sum().item())
                    if n_over_prev is None or n_over < n_over_prev:
                        n_over_prev = n_over
                        mask = over.to(torch.uint8)
                        if T.store.n_shared:        # a shared row that found no region is claimed again in the next round
                            base = T.store.n_views * T.store.n_tables
                            rows = T.table_of.long().view(batch, layout.n_xtal)[over].reshape(-1)
                            T.store.shared_tag[rows[rows >= base] - base] = 0
                        continue
                break
            if bits == 0:
                return T
            if bits & abi.TABLE_BAD_LATTICE:
                raise ValueError("degenerate lattice parameters (cell volume or edge is zero / not finite)")
            if bits & abi.TABLE_GRID_OVERFLOW:
                raise abi.XrsdError("candidate hkl box of more than %d cells: beyond the table builder's 2^24-cell limit "
                                    "or the %d-byte global scratch budget" % (int(max_cells), GRID_SCRATCH_BYTES))
            n_max = int(T.n_all.max().item())
            if bits & abi.TABLE_LIST_OVERFLOW:
                if cap_list < 4096:
                    cap_list = 4096
                elif not cap_big and n_max <= (1 << 18):
                    # large cell without usable symmetry: sort / shell arrays of those tables go to global memory
                    cap_big = 1 << max(13, (n_max - 1).bit_length())
                else:
                    raise abi.XrsdError("reduced reflection list too long (%d reflections in the shell)" % n_max)
            elif bits & abi.TABLE_SHELL_OVERFLOW and cap_shell < 65536:
                cap_shell *= 4
            else:
                raise abi.XrsdError("reflection table build failed with status %d" % bits)
            T = None        # new store with the larger capacities

    # -- intensity ----------------------------------------------------------------------
    def intensity(self, layout, q, params, tables=None, out=None, precision="fp64"):
        """I[b, :] on the device.  q: [n_q]; params: [batch, n_params] (numpy or cuda tensors).
        precision="fp32" selects the optional single-precision path (xrsd_intensity_batch_f32: size-distribution
        loops in FP32, <= 1e-5 relative to the default; layouts without crystalline populations only)."""
        if precision not in ("fp64", "fp32"):
            raise ValueError("precision must be 'fp64' or 'fp32'")
        torch = self.torch
        self.check_grid(layout, q)
        d_q = self.to_device(q)
        host_params = None if isinstance(params, torch.Tensor) else np.atleast_2d(np.asarray(params, dtype=np.float64))
        d_p = self.to_device(params if host_params is None else host_params)
        if d_p.dim() == 1:
            d_p = d_p.unsqueeze(0)
        batch, n_q = d_p.shape[0], d_q.shape[0]
        if out is None:
            out = torch.empty((batch, n_q), dtype=torch.float64, device=self.device)
        if layout.n_xtal and tables is None:
            tables = self.build_tables(layout, d_p, batch, params_host=host_params)
        if precision == "fp32":
            abi.check(self.lib.xrsd_intensity_batch_f32(C.byref(layout.c), d_q.data_ptr(), n_q, d_p.data_ptr(), d_p.shape[1],
                                                        batch, out.data_ptr(), out.shape[1], self._stream()))
            self.launches += 1
            return out
        abi.check(self.lib.xrsd_intensity_batch(C.byref(layout.c), d_q.data_ptr(), n_q, d_p.data_ptr(), d_p.shape[1], batch,
                                                C.byref(tables.c) if tables is not None else None,
                                                out.data_ptr(), out.shape[1], self._stream()))
        self.launches += 1
        return out

    def intensity_components(self, layout, q, params, tables=None):
        """(I[b, :], parts[b, n_pops, :]) where parts[:, 0] is the noise term and parts[:, k] population k
        (layout order), all from one launch."""
        torch = self.torch
        self.check_grid(layout, q)
        d_q = self.to_device(q)
        host_params = None if isinstance(params, torch.Tensor) else np.atleast_2d(np.asarray(params, dtype=np.float64))
        d_p = self.to_device(params if host_params is None else host_params)
        if d_p.dim() == 1:
            d_p = d_p.unsqueeze(0)
        batch, n_q, n_pops = d_p.shape[0], d_q.shape[0], layout.c.n_pops
        out = torch.empty((batch, n_q), dtype=torch.float64, device=self.device)
        cum = torch.empty((batch, n_pops, n_q), dtype=torch.float64, device=self.device)
        if layout.n_xtal and tables is None:
            tables = self.build_tables(layout, d_p, batch, params_host=host_params)
        abi.check(self.lib.xrsd_intensity_components_batch(
            C.byref(layout.c), d_q.data_ptr(), n_q, d_p.data_ptr(), d_p.shape[1], batch,
            C.byref(tables.c) if tables is not None else None, out.data_ptr(), out.shape[1], cum.data_ptr(), self._stream()))
        self.launches += 1
        parts = cum.clone()
        parts[:, 1:] -= cum[:, :-1]
        return out, parts

    def intensity_single(self, layout, q):
        """numpy in, numpy out: one system with the parameter values stored in its layout.  Goes through
        the host-buffer C entry point (xrsd_intensity_batch_host): for a single pattern the cost is launch
        and copy latency, and that path has the fewest host-side steps."""
        q = np.ascontiguousarray(q, dtype=np.float64)
        if q.size == 0:
            return np.zeros(0)
        vals = np.ascontiguousarray(layout.values, dtype=np.float64)
        out = np.empty(q.size)
        self.torch.cuda.set_device(self.device)
        rc = self.lib.xrsd_intensity_batch_host(C.byref(layout.c), abi.ptr(q), q.size, abi.ptr(vals), vals.size, 1,
                                                abi.ptr(out), q.size)
        if rc == abi.ECAPACITY and layout.n_xtal:
            # the host path's fixed table capacities were exceeded (very many shells): the device API grows them
            return self.intensity(layout, q, vals[None, :])[0].cpu().numpy()
        abi.check(rc)
        self.launches += 1 + (1 if layout.n_xtal else 0)
        return out

    # -- objective ----------------------------------------------------------------------
    def objective(self, error_weighted=True, logI_weighted=True, q_range=(0., float('inf')), has_dI=False):
        o = abi.Objective()
        o.error_weighted, o.logI_weighted, o.has_dI = int(bool(error_weighted)), int(bool(logI_weighted)), int(bool(has_dI))
        o.q_lo, o.q_hi = float(q_range[0]), float(q_range[1])
        return o

    def prepare_objective(self, obj, d_q, d_I, d_dI, ld_obs, batch):
        """(objective marked prepared, f(I_obs) stream, weight stream) for a fit: the per-point logarithm, default
        error estimate and mask tests of the objective are evaluated once instead of in every launch of the loop
        (xrsd_objective_prepare).  The streams take the place of I_obs / dI in the residual, Jacobian and chi2 calls."""
        n_q = d_q.shape[0]
        rows = batch if ld_obs else 1
        fo = self.torch.empty((rows, n_q) if ld_obs else (n_q,), dtype=self.torch.float64, device=self.device)
        w = self.torch.empty_like(fo)
        abi.check(self.lib.xrsd_objective_prepare(C.byref(obj), d_q.data_ptr(), n_q, d_I.data_ptr(),
                                                  d_dI.data_ptr() if d_dI is not None else None, ld_obs, rows,
                                                  fo.data_ptr(), w.data_ptr(), n_q, self._stream()))
        self.launches += 1
        op = abi.Objective()
        C.memmove(C.byref(op), C.byref(obj), C.sizeof(abi.Objective))
        op.prepared, op.has_dI = 1, 1
        return op, fo, w

    def _obs(self, Iobs, dI, batch, n_q):
        """Device copies of the observed pattern(s) and their error estimates plus the row stride the
        kernels use: 0 when ONE pattern [n_q] is shared by the whole batch, n_q for one row per system.
        I_obs and dI always end up with the same shape."""
        d_I = self.to_device(Iobs)
        d_dI = None if dI is None else self.to_device(np.broadcast_to(np.asarray(dI, dtype=np.float64), (n_q,))
                                                      if np.ndim(dI) == 0 else dI)
        for name, t in (("I", d_I), ("dI", d_dI)):
            if t is not None and (t.shape[-1] != n_q or t.dim() > 2 or (t.dim() == 2 and t.shape[0] not in (1, batch))):
                raise ValueError("%s has shape %s, expected [n_q] or [batch, n_q] with n_q=%d, batch=%d"
                                 % (name, tuple(t.shape), n_q, batch))
        per_system = any(t is not None and t.dim() == 2 and t.shape[0] > 1 for t in (d_I, d_dI))
        shape = (batch, n_q) if per_system else (n_q,)

        def fit_shape(t):
            if t is None:
                return None
            return t.reshape(-1, n_q).expand(batch, n_q).contiguous() if per_system else t.reshape(n_q).contiguous()

        d_I, d_dI = fit_shape(d_I), fit_shape(d_dI)
        assert tuple(d_I.shape) == shape
        return d_I, d_dI, (n_q if per_system else 0)

    def residual_workspace(self, batch, n_q):
        """Zero-initialised partial-sum / ticket area of xrsd_residual_batch, kept per shape (the layout of the area
        depends on it; the kernel leaves the tickets zeroed after every launch)."""
        cache = self.__dict__.setdefault("_res_ws", {})
        key = (batch, n_q, self.torch.cuda.current_stream(self.device).cuda_stream)      # launches on one stream are ordered
        if key not in cache:
            if len(cache) >= 8:
                cache.pop(next(iter(cache)))
            cache[key] = self.torch.zeros(int(self.lib.xrsd_residual_workspace_bytes(batch, n_q)), dtype=self.torch.uint8,
                                          device=self.device)
        return cache[key]

    def residual(self, layout, obj, q, params, Iobs, dI=None, tables=None, active=None, out=None):
        """chi2[b] of System.evaluate_residual for every parameter set of the batch."""
        torch = self.torch
        self.check_grid(layout, q)
        d_q = self.to_device(q)
        host_params = None if isinstance(params, torch.Tensor) else np.atleast_2d(np.asarray(params, dtype=np.float64))
        d_p = self.to_device(params if host_params is None else host_params)
        if d_p.dim() == 1:
            d_p = d_p.unsqueeze(0)
        batch, n_q = d_p.shape[0], d_q.shape[0]
        d_I, d_dI, ld = self._obs(Iobs, dI, batch, n_q)
        if not obj.prepared:
            obj.has_dI = int(d_dI is not None)
        if layout.n_xtal and tables is None:
            tables = self.build_tables(layout, d_p, batch, params_host=host_params)
        chi2 = out if out is not None else torch.empty(batch, dtype=torch.float64, device=self.device)
        ws = self.residual_workspace(batch, n_q)
        abi.check(self.lib.xrsd_residual_batch(C.byref(layout.c), C.byref(obj), d_q.data_ptr(), n_q, d_p.data_ptr(),
                                               d_p.shape[1], batch, C.byref(tables.c) if tables is not None else None,
                                               d_I.data_ptr(), d_dI.data_ptr() if d_dI is not None else None, ld,
                                               chi2.data_ptr(), ws.data_ptr(),
                                               active.data_ptr() if active is not None else None, self._stream()))
        self.launches += 1
        return chi2

    @staticmethod
    def tie_array(ties):
        """[(dst, src[, scale, offset]), ...] -> (int32 slot pairs, float64 (scale, offset) pairs, count) for the C
        calls; (None, None, 0) without constraints."""
        if not ties:
            return None, None, 0
        flat = [int(v) for t in ties for v in t[:2]]
        coef = [float(v) for t in ties for v in ((t[2], t[3]) if len(t) >= 4 else (1.0, 0.0))]
        return (C.c_int32 * len(flat))(*flat), (C.c_double * len(coef))(*coef), len(ties)

    def jacobian(self, layout, obj, q, d_params, Iobs_dev, dI_dev, ld_obs, free_slots, tables, rel_step=1e-6, out=None,
                 active=None, ties=None, resident=None, base_ready=False, patterns_only=False):
        """Normal equations (JTJ, JTr, chi2) of the weighted residual vector by forward differences.

        By default the variant patterns go through a shared scratch, the batch cut into chunks of <= 4 GB.  With
        `resident` (a key) the whole batch gets its own workspace that survives between calls, so the caller can
        read the pattern slabs (`jacobian_slabs`) and, with base_ready=True, have the call trust the base pattern
        and running totals already in it (the LM loop copies an accepted trial evaluation there instead of
        evaluating the base again).  patterns_only=True stops after the pattern slabs (no reduction, nothing is
        written to `out`; Iobs_dev may be None): with one linear free parameter that is the pattern and the running
        totals of every active system."""
        torch = self.torch
        self.check_grid(layout, q)
        batch, nf = d_params.shape[0], len(free_slots)
        if patterns_only and (resident is None or base_ready):
            raise ValueError("patterns_only needs a resident workspace (and excludes base_ready)")
        mode = abi.JAC_PATTERNS_ONLY if patterns_only else (abi.JAC_BASE_READY if base_ready else abi.JAC_FULL)
        if out is None and not patterns_only:
            out = (torch.empty((batch, nf, nf), dtype=torch.float64, device=self.device),
                   torch.empty((batch, nf), dtype=torch.float64, device=self.device),
                   torch.empty(batch, dtype=torch.float64, device=self.device))
        idx = (C.c_int32 * nf)(*free_slots) if nf else None
        tie_c, coef_c, n_tie = self.tie_array(ties)
        n_q = q.shape[0]
        if resident is not None:
            chunk = batch
            ws = self.resident_workspace(resident, n_q, chunk, nf, must_exist=base_ready)
        else:
            if base_ready:
                raise ValueError("base_ready requires a resident workspace")
            # HBM scratch for the variant patterns: chunk the batch so it stays under ~4 GB
            chunk = max(1, min(batch, int(4e9 // (8 * (nf + 1 + abi.MAX_POPS) * n_q))))
            need = int(self.lib.xrsd_jacobian_workspace_bytes(n_q, chunk, nf))
            # one scratch per stream (launches of one stream are ordered; fits on different streams must not share it)
            slot = self.__dict__.setdefault("_jac_ws", {}).setdefault(self.torch.cuda.current_stream(self.device).cuda_stream,
                                                                     {"ws": None, "key": None})
            ws = slot["ws"]
            if ws is None or ws.numel() < need:
                ws = torch.empty(need, dtype=torch.uint8, device=self.device)
                slot["ws"], slot["key"] = ws, None
        nx = layout.n_xtal
        for lo in range(0, batch, chunk):
            hi = min(batch, lo + chunk)
            if resident is None:
                # the partial-sum / ticket area sits behind the pattern slabs of THIS launch's row count and has to
                # start at zero (the kernels leave it zeroed): when the shape changes only that tail is cleared,
                # not the gigabytes of slabs in front of it
                if slot["key"] != (n_q, hi - lo, nf):
                    self._zero_jacobian_tail(ws, n_q, hi - lo, nf)
                    slot["key"] = (n_q, hi - lo, nf)
            elif hi - lo != chunk:
                ws.zero_()
            tc = None
            if tables is not None:
                tc = abi.Tables()
                C.memmove(C.byref(tc), C.byref(tables.c), C.sizeof(abi.Tables))
                tc.table_of = tables.table_of.data_ptr() + 4 * lo * nx      # rows are absolute: only the index moves
                tc.build_flag = tables.build_flag.data_ptr() + lo * nx
            abi.check(self.lib.xrsd_jacobian_batch(
                C.byref(layout.c), C.byref(obj), q.data_ptr(), n_q, d_params[lo:hi].data_ptr(), d_params.shape[1], hi - lo,
                C.byref(tc) if tc is not None else None,
                (Iobs_dev.data_ptr() + 8 * lo * ld_obs) if Iobs_dev is not None else None,
                (dI_dev.data_ptr() + 8 * lo * ld_obs) if dI_dev is not None else None,
                ld_obs, idx, nf, tie_c, coef_c, n_tie, float(rel_step), mode,
                *([None, None, None] if patterns_only else [out[0][lo:hi].data_ptr(), out[1][lo:hi].data_ptr(), out[2][lo:hi].data_ptr()]),
                ws.data_ptr(), (active.data_ptr() + lo) if active is not None else None, self._stream()))
            self.launches += 1 if patterns_only else 2
        return out

    def resident_workspace(self, key, n_q, rows, nf, must_exist=False):
        """The Jacobian workspace kept under `key` (created on first use: pattern slabs, then the zeroed partial-sum /
        ticket area) for `rows` systems and `nf` free parameters."""
        need = int(self.lib.xrsd_jacobian_workspace_bytes(n_q, rows, nf))
        store = self.__dict__.setdefault("_jac_resident", {})
        ent = store.get(key)
        if ent is None or ent[0].numel() < need or ent[1] != (n_q, rows, nf):
            if must_exist:
                raise RuntimeError("base_ready needs the resident workspace of a previous call with the same shape")
            # only the accumulator / ticket area behind the pattern slabs has to start at zero
            ws_new = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
            ws_new[int(rows) * (nf + 1 + abi.MAX_POPS) * n_q * 8:].zero_()
            ent = (ws_new, (n_q, rows, nf))
            store[key] = ent
        return ent[0]

    def _zero_jacobian_tail(self, ws, n_q, rows, nf):
        head = int(rows) * (nf + 1 + abi.MAX_POPS) * int(n_q) * 8
        ws[head:int(self.lib.xrsd_jacobian_workspace_bytes(n_q, rows, nf))].zero_()

    def pattern_chi2(self, obj, q, patterns, Iobs_dev, dI_dev, ld_obs, out=None, active=None):
        """chi2 of patterns that are already on the device: `patterns` is a [batch, n_q] float64 tensor whose rows may
        be strided (e.g. slab 0 of jacobian_slabs())."""
        batch, n_q = patterns.shape
        if patterns.stride(1) != 1:
            raise ValueError("pattern rows must be contiguous")
        if out is None:
            out = self.torch.empty(batch, dtype=self.torch.float64, device=self.device)
        abi.check(self.lib.xrsd_pattern_chi2_batch(C.byref(obj), q.data_ptr(), n_q, patterns.data_ptr(), patterns.stride(0) if batch > 1 else n_q,
                                                   batch, Iobs_dev.data_ptr(), dI_dev.data_ptr() if dI_dev is not None else None,
                                                   ld_obs, out.data_ptr(), active.data_ptr() if active is not None else None,
                                                   self._stream()))
        self.launches += 1
        return out

    def jacobian_variants(self, layout, free_slots, ties=None):
        """Pattern slabs per system that jacobian() evaluates for this free set (base + non-linear parameters)."""
        nf = len(free_slots)
        idx = (C.c_int32 * nf)(*free_slots)
        tie_c, _, n_tie = self.tie_array(ties)
        n = int(self.lib.xrsd_jacobian_n_variants(C.byref(layout.c), idx, nf, tie_c, n_tie))
        if n < 1:
            raise ValueError("invalid free set")
        return n

    def jacobian_slabs(self, resident, layout, batch, n_q, n_var):
        """[batch, n_var + n_pops, n_q] float64 view of a resident Jacobian workspace: slab 0 = base pattern,
        1..n_var-1 = perturbed patterns, the last n_pops = running totals of the base after each population."""
        ws = self._jac_resident[resident][0]
        n_slab = n_var + int(layout.c.n_pops)
        return ws[:batch * n_slab * n_q * 8].view(self.torch.float64).view(batch, n_slab, n_q)

    def take_slab_rows(self, dst_key, src_key, layout, batch, n_q, n_var, mask):
        """Rows flagged in `mask` (bool / uint8 [batch]): base pattern and running totals of the resident workspace
        `dst_key` (n_var variants) are overwritten with those of `src_key` (a one-variant workspace)."""
        n_pops = int(layout.c.n_pops)
        dst, src = self._jac_resident[dst_key][0].data_ptr(), self._jac_resident[src_key][0].data_ptr()
        ld_dst, ld_src = (n_var + n_pops) * n_q, (1 + n_pops) * n_q
        for off_dst, off_src, row_len in ((0, 0, n_q), (n_var * n_q, n_q, n_pops * n_q)):
            abi.check(self.lib.xrsd_copy_rows_masked(dst + 8 * off_dst, ld_dst, src + 8 * off_src, ld_src, row_len, batch,
                                                     mask.data_ptr(), self._stream()))
        self.launches += 2

    def release_resident(self, *keys):
        for k in keys:
            self.__dict__.get("_jac_resident", {}).pop(k, None)

    def peak_profile(self, kind, x, p0, p1=0.0):
        """gaussian / lorentzian / voigt of tools/peak_math.py evaluated on the device."""
        d_x = self.to_device(x)
        out = self.torch.empty_like(d_x)
        abi.check(self.lib.xrsd_peak_profile(abi.PROFILE[kind], float(p0), float(p1), d_x.data_ptr(), d_x.numel(),
                                             out.data_ptr(), self._stream()))
        self.launches += 1
        return out

    def features(self, d_q, d_I, out=None):
        """Featuriser (reference tools/profiler.py:64): d_I [B, n_q] on the device -> [B, 21] features."""
        B, n_q = d_I.shape
        if out is None:
            out = self.torch.empty((B, abi.N_FEATURES), dtype=self.torch.float64, device=self.device)
        abi.check(self.lib.xrsd_feature_batch(d_q.data_ptr(), n_q, d_I.data_ptr(), d_I.stride(0) if B > 1 else n_q, B,
                                              out.data_ptr(), out.stride(0) if B > 1 else abi.N_FEATURES, self._stream()))
        self.launches += 1
        return out

    def probe_fp64_peak(self):
        t, ms = C.c_double(), C.c_double()
        abi.check(self.lib.xrsd_probe_fp64_peak(C.byref(t), C.byref(ms), self._stream()))
        return t.value, ms.value
