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


class TableSet(object):
    """Device arrays of xrsd_tables_t for batch*n_xtal reflection tables."""

    def __init__(self, torch, device, n_tables, cap_shell, n_pairs, cap_refl=0):
        self.n_tables, self.cap_shell, self.n_pairs, self.cap_refl = n_tables, cap_shell, n_pairs, cap_refl
        i32, f64 = torch.int32, torch.float64
        self.n_shell = torch.zeros(n_tables, dtype=i32, device=device)
        self.n_refl = torch.zeros(n_tables, dtype=i32, device=device)
        self.n_all = torch.zeros(n_tables, dtype=i32, device=device)
        self.status = torch.zeros(n_tables, dtype=i32, device=device)
        self.shell_hkl = torch.zeros((n_tables, cap_shell, 4), dtype=i32, device=device)
        self.shell_geo = torch.zeros((n_tables, cap_shell, n_pairs), dtype=f64, device=device)
        self.refl_hkl = torch.zeros((n_tables, cap_refl, 4), dtype=i32, device=device) if cap_refl else None
        c = abi.Tables()
        c.cap_shell, c.cap_refl, c.n_pairs_max = cap_shell, cap_refl, n_pairs
        c.n_shell, c.n_refl, c.n_all, c.status = (self.n_shell.data_ptr(), self.n_refl.data_ptr(),
                                                  self.n_all.data_ptr(), self.status.data_ptr())
        c.shell_hkl, c.shell_geo = self.shell_hkl.data_ptr(), self.shell_geo.data_ptr()
        c.refl_hkl = self.refl_hkl.data_ptr() if cap_refl else None
        self.c = c

    def take_rows_from(self, other, rows):
        """Overwrite the tables of the systems flagged in `rows` (bool [n_tables]) with `other`'s, in place (the
        device pointers in self.c stay valid).  Used by the LM loop: an accepted trial step makes the trial tables
        the system's tables, so nothing has to be rebuilt.  Returns False when the two sets are not congruent."""
        if (other.n_tables, other.cap_shell, other.n_pairs) != (self.n_tables, self.cap_shell, self.n_pairs):
            return False
        for name in ("n_shell", "n_refl", "n_all", "status"):
            a, b = getattr(self, name), getattr(other, name)
            a.copy_(b.where(rows, a))
        m3 = rows[:, None, None]
        self.shell_hkl.copy_(other.shell_hkl.where(m3, self.shell_hkl))
        self.shell_geo.copy_(other.shell_geo.where(m3, self.shell_geo))
        return True


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

    def to_device(self, arr):
        torch = self.torch
        if isinstance(arr, torch.Tensor):
            return arr.to(device=self.device, dtype=torch.float64).contiguous()
        a = np.ascontiguousarray(arr, dtype=np.float64)
        if not a.flags.writeable:      # e.g. arrays that come out of an .npz: torch wants a writable buffer
            a = a.copy()
        return torch.from_numpy(a).to(self.device)

    def _n_pairs(self, layout):
        n = 1
        for p in layout.crystalline_pops():
            n = max(n, p.n_species * (p.n_species + 1) // 2)
        return n

    # -- reflection tables --------------------------------------------------------------
    def build_tables(self, layout, d_params, batch, params_host=None, cap_shell=None, export_reflections=False,
                     max_cells=None, check=True, reuse=None, active=None):
        """Build the reflection tables of all crystalline populations for a batch.

        `params_host` (numpy [batch, n_params]) is only used to bound the candidate box when
        `max_cells` is not given.  With check=True the per-table status is read back (one small
        device->host copy) and capacity overflows are retried with larger tables.  `reuse` is a
        TableSet of the same shape to overwrite instead of allocating a new one."""
        if layout.n_xtal == 0:
            return None
        torch = self.torch
        if max_cells is None:
            if params_host is None:
                params_host = d_params.detach().cpu().numpy()
            max_cells = lowering.max_cells(layout, params_host)
        has_ops = all(p.use_symmetry and p.n_ops > 0 for p in layout.crystalline_pops())
        if cap_shell is None:
            cap_shell = 128 if has_ops else 1024
        n_tables = batch * layout.n_xtal
        cap_refl = 4096 if export_reflections else 0
        cap_list = 0
        cap_big = 0
        while True:
            if reuse is not None and reuse.n_tables == n_tables and reuse.cap_refl == cap_refl and reuse.cap_shell >= cap_shell:
                T, cap_shell = reuse, reuse.cap_shell
            else:
                T = TableSet(torch, self.device, n_tables, cap_shell, self._n_pairs(layout), cap_refl)
            reuse = None
            T.c.cap_list_hint = cap_list
            if cap_big:
                T.big = torch.empty(n_tables * int(self.lib.xrsd_table_scratch_bytes(cap_big)), dtype=torch.uint8, device=self.device)
                T.c.big_scratch, T.c.cap_big = T.big.data_ptr(), cap_big
            abi.check(self.lib.xrsd_reflection_tables(C.byref(layout.c), d_params.data_ptr(), d_params.shape[1], batch,
                                                      C.byref(T.c), max_cells,
                                                      active.data_ptr() if active is not None else None, self._stream()))
            self.launches += 1
            if not check:
                return T
            st = int(T.status.max().item()) if n_tables else 0
            if st == 0:
                return T
            if st & abi.TABLE_BAD_LATTICE:
                raise ValueError("degenerate lattice parameters (cell volume or edge is zero / not finite)")
            if st & abi.TABLE_GRID_OVERFLOW:
                raise abi.XrsdError("candidate hkl box exceeds the shared-memory budget (q_max * cell edge too large)")
            if st & abi.TABLE_LIST_OVERFLOW:
                if cap_list < 4096:
                    cap_list = 4096
                    continue
                if not cap_big and cap_refl == 0:
                    # large cell without usable symmetry: sort / shell arrays of those tables go to global memory
                    n_max = int(T.n_all.max().item())
                    cap_big = 1 << max(13, (n_max - 1).bit_length())
                    if cap_big <= (1 << 18):
                        continue
                raise abi.XrsdError("reduced reflection list too long (%d reflections in the shell)" % int(T.n_all.max().item()))
            if st & abi.TABLE_SHELL_OVERFLOW and cap_shell < 65536:
                cap_shell *= 4
                continue
            raise abi.XrsdError("reflection table build failed with status %d" % st)

    # -- intensity ----------------------------------------------------------------------
    def intensity(self, layout, q, params, tables=None, out=None):
        """I[b, :] on the device.  q: [n_q]; params: [batch, n_params] (numpy or cuda tensors)."""
        torch = self.torch
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
        abi.check(self.lib.xrsd_intensity_batch(C.byref(layout.c), d_q.data_ptr(), n_q, d_p.data_ptr(), d_p.shape[1], batch,
                                                C.byref(tables.c) if tables is not None else None,
                                                out.data_ptr(), out.shape[1], self._stream()))
        self.launches += 1
        return out

    def intensity_components(self, layout, q, params, tables=None):
        """(I[b, :], parts[b, n_pops, :]) where parts[:, 0] is the noise term and parts[:, k] population k
        (layout order), all from one launch."""
        torch = self.torch
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

    def residual_workspace(self, batch):
        """Zero-initialised accumulator / ticket area of xrsd_residual_batch, kept per batch size (the layout
        of the area depends on it; the kernel leaves it zeroed after every launch)."""
        cache = self.__dict__.setdefault("_res_ws", {})
        if batch not in cache:
            cache.clear()
            cache[batch] = self.torch.zeros(int(self.lib.xrsd_residual_workspace_bytes(batch)), dtype=self.torch.uint8,
                                            device=self.device)
        return cache[batch]

    def residual(self, layout, obj, q, params, Iobs, dI=None, tables=None, active=None, out=None):
        """chi2[b] of System.evaluate_residual for every parameter set of the batch."""
        torch = self.torch
        d_q = self.to_device(q)
        host_params = None if isinstance(params, torch.Tensor) else np.atleast_2d(np.asarray(params, dtype=np.float64))
        d_p = self.to_device(params if host_params is None else host_params)
        if d_p.dim() == 1:
            d_p = d_p.unsqueeze(0)
        batch, n_q = d_p.shape[0], d_q.shape[0]
        d_I, d_dI, ld = self._obs(Iobs, dI, batch, n_q)
        obj.has_dI = int(d_dI is not None)
        if layout.n_xtal and tables is None:
            tables = self.build_tables(layout, d_p, batch, params_host=host_params)
        chi2 = out if out is not None else torch.empty(batch, dtype=torch.float64, device=self.device)
        ws = self.residual_workspace(batch)
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
        batch, nf = d_params.shape[0], len(free_slots)
        if patterns_only and (resident is None or base_ready):
            raise ValueError("patterns_only needs a resident workspace (and excludes base_ready)")
        mode = abi.JAC_PATTERNS_ONLY if patterns_only else (abi.JAC_BASE_READY if base_ready else abi.JAC_FULL)
        if out is None and not patterns_only:
            out = (torch.empty((batch, nf, nf), dtype=torch.float64, device=self.device),
                   torch.empty((batch, nf), dtype=torch.float64, device=self.device),
                   torch.empty(batch, dtype=torch.float64, device=self.device))
        idx = (C.c_int32 * nf)(*free_slots)
        tie_c, coef_c, n_tie = self.tie_array(ties)
        n_q = q.shape[0]
        if resident is not None:
            chunk = batch
            need = int(self.lib.xrsd_jacobian_workspace_bytes(n_q, chunk, nf))
            store = self.__dict__.setdefault("_jac_resident", {})
            ent = store.get(resident)
            if ent is None or ent[0].numel() < need or ent[1] != (n_q, chunk, nf):
                if base_ready:
                    raise RuntimeError("base_ready needs the resident workspace of a previous call with the same shape")
                ent = (torch.zeros(need, dtype=torch.uint8, device=self.device), (n_q, chunk, nf))
                store[resident] = ent
            ws = ent[0]
        else:
            if base_ready:
                raise ValueError("base_ready requires a resident workspace")
            # HBM scratch for the variant patterns: chunk the batch so it stays under ~4 GB
            chunk = max(1, min(batch, int(4e9 // (8 * (nf + 1 + abi.MAX_POPS) * n_q))))
            need = int(self.lib.xrsd_jacobian_workspace_bytes(n_q, chunk, nf))
            ws = getattr(self, "_jac_ws", None)
            if ws is None or ws.numel() < need:
                ws = torch.zeros(need, dtype=torch.uint8, device=self.device)
                self._jac_ws = ws
                self._jac_ws_key = None
            if self._jac_ws_key != (n_q, chunk, nf):      # ticket area moves with the shape: re-zero it
                ws.zero_()
                self._jac_ws_key = (n_q, chunk, nf)
        nx = layout.n_xtal
        for lo in range(0, batch, chunk):
            hi = min(batch, lo + chunk)
            if hi - lo != chunk:
                ws.zero_()
                self._jac_ws_key = None
            tc = None
            if tables is not None:
                tc = abi.Tables()
                C.memmove(C.byref(tc), C.byref(tables.c), C.sizeof(abi.Tables))
                tc.n_shell = tables.n_shell.data_ptr() + 4 * lo * nx
                tc.n_refl = tables.n_refl.data_ptr() + 4 * lo * nx
                tc.n_all = tables.n_all.data_ptr() + 4 * lo * nx
                tc.status = tables.status.data_ptr() + 4 * lo * nx
                tc.shell_hkl = tables.shell_hkl.data_ptr() + 16 * lo * nx * tables.cap_shell
                tc.shell_geo = tables.shell_geo.data_ptr() + 8 * lo * nx * tables.cap_shell * tables.n_pairs
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
