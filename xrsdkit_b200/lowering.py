"""Lowering: System / Population objects -> the flat descriptors the CUDA engine consumes.

A *layout* is everything about a system that is not a parameter value: which populations
exist, their settings, and which slot of the flat parameter vector each parameter lives in.
Many systems that share a layout (a batch of parameter sets, the iterates of a fit) are then
evaluated in one launch from a [batch, n_params] matrix.

The slot order is the reference's flatten_params() order (system/__init__.py:211-218):
noise parameters first, then each population's parameters in dict order, keyed
'<population>__<parameter>'.
"""

import numpy as np

from . import abi
from . import definitions as xrsdefs

# point group -> operations the reference applies, in order (scattering/symmetries.py:27-42).
# Every other point group has no operations defined there and reduces nothing.
POINT_GROUP_OPS = {
    '1': [],
    '-1': ['inversion'],
    'm-3m': ['mirror_x', 'mirror_y', 'mirror_z', 'mirror_x_y', 'mirror_y_z', 'mirror_z_x',
             'mirror_nx_y', 'mirror_ny_z', 'mirror_nz_x'],
    '6/mmm': ['inversion', 'mirror_z', 'mirror_x_y', 'mirror_nx_y'],
}

_LATTICE_CLASS = {}
for _l in ('P_cubic', 'I_cubic', 'F_cubic', 'diamond'):
    _LATTICE_CLASS[_l] = abi.LAT_CUBIC
_LATTICE_CLASS['hcp'] = abi.LAT_HCP
_LATTICE_CLASS['hexagonal'] = abi.LAT_HEXAGONAL
for _l in ('P_orthorhombic', 'C_orthorhombic', 'A_orthorhombic', 'I_orthorhombic', 'F_orthorhombic'):
    _LATTICE_CLASS[_l] = abi.LAT_ORTHORHOMBIC
for _l in ('P_tetragonal', 'I_tetragonal'):
    _LATTICE_CLASS[_l] = abi.LAT_TETRAGONAL
for _l in ('P_monoclinic', 'C_monoclinic'):
    _LATTICE_CLASS[_l] = abi.LAT_MONOCLINIC
_LATTICE_CLASS['rhombohedral'] = abi.LAT_RHOMBOHEDRAL
_LATTICE_CLASS['triclinic'] = abi.LAT_TRICLINIC

_LAT_SLOTS = ('a', 'b', 'c', 'alpha', 'beta', 'gamma')


class CapacityError(ValueError, NotImplementedError):
    """An input beyond one of the engine's fixed capacities (include/xrsd_b200.h: XRSD_MAX_POPS, XRSD_MAX_SPECIES,
    XRSD_MAX_PARAMS, XRSD_MAX_FREE, XRSD_MAX_TIES); the message names the limit.  The reference has no such limits,
    so this is a ValueError about the input as far as its callers are concerned (and a NotImplementedError for
    callers of the round-1 interface)."""


class SystemLayout(object):
    """ctypes layout + the parameter bookkeeping that goes with it."""

    def __init__(self, c_layout, names, values, records):
        self.c = c_layout
        self.names = names            # 'pop__param' per slot
        self.values = values          # float64[n_params]
        self.records = records        # the parameter dicts (shared with the System, like flatten_params)
        self.n_params = len(names)
        self.n_xtal = c_layout.n_xtal
        self.grid = None              # fingerprint of the q grid the layout was lowered for (set by lower_system)

    def check_grid(self, q):
        """Raise ValueError when `q` (numpy) is not the grid this layout was lowered for, in a way that matters: the
        layout carries q_step (> 0: the kernels advance phases along a uniform grid and skip the ordering test),
        q0_is_zero (the only zero the reference special-cases) and, for crystalline populations without an explicit
        q_max, q[-1] (scattering/__init__.py:226)."""
        g = self.grid
        if g is None or len(q) == 0:
            return
        if g['q_step'] > 0.0:
            same = len(q) == g['n'] and q[0] == g['q0'] and q[-1] == g['q_last'] and uniform_step(q) == g['q_step']
        else:
            same = bool(q[0] == 0) == g['q0_is_zero'] and \
                (not g['uses_q_last'] or q[-1] == g['q_last'])
        if not same:
            raise ValueError("this layout was lowered for another q grid (%d points from %g to %g); lower the system again "
                             "for the grid in use (batch.layout_of(system, q))" % (g['n'], g['q0'], g['q_last']))

    def slot(self, key):
        return self.names.index(key)

    def free_slots(self):
        """Slots lmfit would vary: not fixed and without a constraint expression
        (reference system/__init__.py:197-209)."""
        return [i for i, r in enumerate(self.records) if not r.get('fixed') and not r.get('constraint_expr')]

    def ties(self):
        """Constraint terms for parameters that carry a constraint_expr, in the convention of xrsd_jacobian_batch
        (include/xrsd_b200.h, XRSD_MAX_TIES).  An expression that is LINEAR in other parameters gives
        (dst, src, scale, offset) terms: value[dst] = sum_k scale_k * value[src_k] + offset, the first term sets the slot,
        further terms carry dst = -(slot + 1) and add to it; 'particle__r' -- the only form found in xrsdkit's fitted
        systems -- gives one term (1, 0); '2*particle__r', '(shell__r - 3) / 2', 'core__r + shell__t' likewise.  Any
        other arithmetic expression of lmfit's asteval language -- products and quotients of parameters, powers, sqrt,
        exp, log, log10, sin, cos, tan, arcsin, arccos, arctan, abs, min, max, pi, e -- is compiled into a postfix program
        (compile_constraint): words (-(TIE_WORD + op), slot, constant, 0) closed by a store term (dst, -1, 1, 0)."""
        out = []
        for i, r in enumerate(self.records):
            expr = r.get('constraint_expr')
            if not expr:
                continue
            try:
                terms, off = parse_linear_constraint(str(expr), self.names)
            except NotLinear:
                words = compile_constraint(str(expr), self.names)
                if any(w[0] == -(abi.TIE_WORD + abi.TIE_OPS.index('pushp')) and w[1] == i for w in words):
                    raise NotImplementedError("constraint_expr %r refers to its own parameter" % expr)
                out.extend(words)
                out.append((i, -1, 1.0, 0.0))
                continue
            if any(src == i for src, _ in terms):
                raise NotImplementedError("constraint_expr %r refers to its own parameter" % expr)
            if not terms:                                   # a constant: 0 * (any slot) + offset
                terms = [(i, 0.0)]
            for k, (src, scale) in enumerate(terms):
                out.append((i if k == 0 else -(i + 1), src, scale, off if k == 0 else 0.0))
        if len(out) > abi.MAX_TIES:
            raise CapacityError("at most %d constraint terms per system (XRSD_MAX_TIES); this system has %d" % (abi.MAX_TIES, len(out)))
        dsts = tied_slots(out)
        if any(tie_reads(t) is not None and tie_reads(t) in dsts and (t[0] <= -abi.TIE_WORD or t[2] != 0.0) for t in out):
            raise NotImplementedError("chained constraints (a constrained parameter used in another constraint_expr)")
        return out

    def bounds(self, slots):
        lo = np.array([-np.inf if self.records[i]['bounds'][0] is None else self.records[i]['bounds'][0] for i in slots], float)
        hi = np.array([np.inf if self.records[i]['bounds'][1] is None else self.records[i]['bounds'][1] for i in slots], float)
        return lo, hi

    def coordinate_slots(self):
        """Slots of the atomic coordinates u_i, v_i, w_i of polyatomic crystalline populations: they enter the
        reflection tables (structure-factor phases), not the per-q evaluation."""
        out = set()
        for p in self.crystalline_pops():
            for a in range(p.n_species):
                for j in range(3):
                    if p.p_uvw[a][j] >= 0:
                        out.add(int(p.p_uvw[a][j]))
        return out

    def crystalline_pops(self):
        return [p for p in self.c.pops[:self.c.n_pops] if p.kind == abi.POP_CRYSTALLINE]


def _blank_pop():
    # a copy of one prepared record: filling the 29 slots field by field was a third of the lowering time
    global _BLANK_POP
    if _BLANK_POP is None:
        _BLANK_POP = bytes(_make_blank_pop())
    return abi.Pop.from_buffer_copy(_BLANK_POP)


_BLANK_POP = None


def _make_blank_pop():
    p = abi.Pop()
    for f in ('p_I0', 'p_r', 'p_sigma', 'p_rg', 'p_D', 'p_r_hard', 'p_v_fraction', 'p_flat_fraction',
              'p_hwhm', 'p_hwhm_g', 'p_hwhm_l'):
        setattr(p, f, -1)
    for i in range(6):
        p.p_lat[i] = -1
    for a in range(abi.MAX_SPECIES):
        for j in range(3):
            p.p_uvw[a][j] = -1
    return p


def _set_atom(pop, idx, symbol):
    rec = xrsdefs.atomic_params[symbol]     # KeyError for an unknown symbol, as the reference
    pop.atom_Z[idx] = float(rec['Z'])
    for j in range(4):
        pop.atom_a[idx][j] = float(rec['a'][j])
        pop.atom_b[idx][j] = float(rec['b'][j])


def _val(rec):
    return rec['value'] if isinstance(rec, dict) else rec


def _lower_noise(model, slots):
    p = _blank_pop()
    if model == 'flat':
        p.kind = abi.POP_NOISE_FLAT
        p.p_I0 = slots['I0']
    elif model == 'low_q_scatter':
        p.kind = abi.POP_NOISE_LOW_Q
        p.p_I0 = slots['I0']
        p.p_flat_fraction = slots['I0_flat_fraction']
        p.p_rg = slots['effective_rg']
        p.p_D = slots['effective_D']
    else:
        raise ValueError('unsupported noise specification: {}'.format(model))
    return p


def _lower_population(structure, form, settings, slots, q_last, table_slot, parameters=None):
    """One population -> abi.Pop.  Mirrors the dispatcher scattering/__init__.py:12-66."""
    p = _blank_pop()
    p.p_I0 = slots['I0']
    if structure == 'crystalline':
        p.kind = abi.POP_CRYSTALLINE
        lattice = settings['lattice']
        p.lattice = _LATTICE_CLASS[lattice]            # KeyError for an unknown lattice
        sites = xrsdefs.lattice_coords(lattice)
        p.n_sites = len(sites)
        for i, s in enumerate(sites):
            for j in range(3):
                p.sites[i][j] = float(s[j])
        sg = settings['space_group']
        if sg and sg not in xrsdefs.lattice_space_groups[lattice].values():
            raise ValueError('space group {} not valid for {} lattice'.format(sg, lattice))
        ops = []
        if sg:
            ops = POINT_GROUP_OPS.get(xrsdefs.sg_point_groups[sg]) or []
        p.n_ops = len(ops)
        for i, nm in enumerate(ops):
            p.ops[i] = abi.OPS[nm]
        p.use_symmetry = 1 if settings['use_symmetry'] else 0
        p.profile = abi.PROFILE[settings['profile']]
        p.sf_radial = 1 if settings['structure_factor_mode'] == 'radial' else 0
        p.polz = 1 if settings['polarization_correction'] else 0
        p.lorentz = 1 if settings['lorentz_correction'] else 0
        p.q_min = float(settings['q_min'])
        q_max = settings['q_max']
        p.q_max = float(q_max) if q_max else float(q_last)     # scattering/__init__.py:226
        for i, nm in enumerate(_LAT_SLOTS):
            if nm in xrsdefs.lattice_param_names(lattice):
                p.p_lat[i] = slots[nm]
        if not xrsdefs.lattice_param_names(lattice):
            # the reference hands lattice_vectors() no parameters for this lattice and fails
            raise TypeError('lattice {} defines no lattice parameters'.format(lattice))
        if settings['profile'] == 'voigt':
            p.p_hwhm_g, p.p_hwhm_l = slots['hwhm_g'], slots['hwhm_l']
        else:
            p.p_hwhm = slots['hwhm']
        if form == 'spherical':
            p.n_species = 1
            p.species_form[0] = abi.FORM_SPHERE
            p.p_r = slots['r']
        elif form == 'atomic':
            p.n_species = 1
            p.species_form[0] = abi.FORM_ATOMIC
            _set_atom(p, 0, settings['symbol'])
        elif form == 'polyatomic':
            n_at = settings['n_atoms']
            if n_at > abi.MAX_SPECIES:
                raise CapacityError('polyatomic populations are limited to {} atoms per cell (XRSD_MAX_SPECIES); '
                                    'this one has {}'.format(abi.MAX_SPECIES, n_at))
            p.n_species = n_at
            for i in range(n_at):
                p.species_form[i] = abi.FORM_ATOMIC
                _set_atom(p, i, settings['symbol_%d' % i])
                for j, ax in enumerate('uvw'):
                    p.p_uvw[i][j] = slots['%s_%d' % (ax, i)]
        p.table_slot = table_slot
        # extinct shells (|S|^2 ~ 1e-31 of the allowed ones) are dropped when the profile has
        # algebraic tails, where their contribution is < 1e-15 relative; kept for 'gaussian',
        # whose tails underflow and would otherwise leave them as the only contribution -- and for a
        # Voigt profile lowered with hwhm_l == 0, which IS that Gaussian (the reference's own lower bound
        # for hwhm_l is 1e-9, definitions.py:292; batches keep the flag of the system they were lowered from)
        gaussian_like = settings['profile'] == 'gaussian' or (
            settings['profile'] == 'voigt' and parameters is not None and float(parameters['hwhm_l']['value']) == 0.0)
        p.prune_extinct = 0 if gaussian_like else 1
        return p
    p.kind = abi.POP_NONCRYSTALLINE
    p.disordered = 1 if structure == 'disordered' else 0
    if p.disordered:
        p.p_r_hard, p.p_v_fraction = slots['r_hard'], slots['v_fraction']
    if form == 'atomic':
        p.form = abi.FORM_ATOMIC
        _set_atom(p, 0, settings['symbol'])
    elif form == 'spherical':
        p.p_r = slots['r']
        if settings['distribution'] == 'r_normal':
            p.form = abi.FORM_SPHERE_R_NORMAL
            p.p_sigma = slots['sigma']
            p.sampling_width = float(settings['sampling_width'])
            p.sampling_step = float(settings['sampling_step'])
        else:
            p.form = abi.FORM_SPHERE
    elif form == 'guinier_porod':
        p.form = abi.FORM_GUINIER_POROD
        p.p_rg, p.p_D = slots['rg'], slots['D']
    else:
        raise ValueError('{} structure does not support {} forms'.format(structure, form))
    return p


def tie_slot(dst):
    """slot of a constraint term's destination (-(slot + 1) marks an accumulating term); None for a program word"""
    if dst <= -abi.TIE_WORD:
        return None
    return -dst - 1 if dst < 0 else dst


def tied_slots(ties):
    """the slots some constraint term sets"""
    return set(tie_slot(t[0]) for t in ties) - {None}


def tie_reads(term):
    """the parameter slot a constraint term reads, or None (constants, operators, the store term of a program)"""
    if term[0] <= -abi.TIE_WORD:
        return term[1] if -term[0] - abi.TIE_WORD == abi.TIE_OPS.index('pushp') else None
    return term[1] if term[1] >= 0 else None


class NotLinear(NotImplementedError):
    """raised by parse_linear_constraint for a well-formed expression that is not linear in the parameters"""


_FUNCS = {'sqrt': 'sqrt', 'exp': 'exp', 'log': 'log', 'ln': 'log', 'log10': 'log10', 'sin': 'sin', 'cos': 'cos', 'tan': 'tan',
          'abs': 'abs', 'fabs': 'abs', 'absolute': 'abs', 'arcsin': 'asin', 'asin': 'asin', 'arccos': 'acos', 'acos': 'acos',
          'arctan': 'atan', 'atan': 'atan'}
_FUNCS2 = {'min': 'min', 'minimum': 'min', 'max': 'max', 'maximum': 'max', 'pow': 'pow', 'power': 'pow'}
_CONSTS = {'pi': np.pi, 'e': np.e}


def compile_constraint(expr, names):
    """A constraint expression as a postfix program: [(-(TIE_WORD + op), slot, constant, 0.0), ...] (abi.TIE_OPS).
    Numbers, parameter names, pi / e, + - * / **, unary minus, and the functions listed in SystemLayout.ties();
    constant sub-expressions are folded.  Raises NotImplementedError for anything else (comparisons, conditionals,
    attribute access, ...: the part of asteval's language that is not arithmetic)."""
    import ast
    import math
    try:
        tree = ast.parse(expr.strip(), mode='eval')
    except SyntaxError:
        raise NotImplementedError("constraint_expr %r is not an arithmetic expression" % expr)
    OP = dict((n, -(abi.TIE_WORD + k)) for k, n in enumerate(abi.TIE_OPS))
    fold1 = {'neg': lambda a: -a, 'sqrt': math.sqrt, 'exp': math.exp, 'log': math.log, 'log10': math.log10, 'sin': math.sin,
             'cos': math.cos, 'tan': math.tan, 'abs': abs, 'asin': math.asin, 'acos': math.acos, 'atan': math.atan}
    fold2 = {'add': lambda a, b: a + b, 'sub': lambda a, b: a - b, 'mul': lambda a, b: a * b, 'div': lambda a, b: a / b,
             'pow': lambda a, b: a ** b, 'min': min, 'max': max}

    def const_of(words):
        return words[0][2] if len(words) == 1 and words[0][0] == OP['pushc'] else None

    def unary(name, a):
        ca = const_of(a)
        if ca is not None:
            try:
                return [(OP['pushc'], 0, float(fold1[name](ca)), 0.0)]
            except (ValueError, OverflowError, ZeroDivisionError):
                pass
        return a + [(OP[name], 0, 0.0, 0.0)]

    def binary(name, a, b):
        ca, cb = const_of(a), const_of(b)
        if ca is not None and cb is not None:
            try:
                return [(OP['pushc'], 0, float(fold2[name](ca, cb)), 0.0)]
            except (ValueError, OverflowError, ZeroDivisionError):
                pass
        return a + b + [(OP[name], 0, 0.0, 0.0)]

    def ev(node):
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)) and not isinstance(node.value, bool):
            return [(OP['pushc'], 0, float(node.value), 0.0)]
        if isinstance(node, ast.Name):
            if node.id in names:
                return [(OP['pushp'], names.index(node.id), 0.0, 0.0)]
            if node.id in _CONSTS:
                return [(OP['pushc'], 0, float(_CONSTS[node.id]), 0.0)]
            raise NotImplementedError("constraint_expr %r: unknown parameter %r" % (expr, node.id))
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.UAdd):
            return ev(node.operand)
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, ast.USub):
            return unary('neg', ev(node.operand))
        if isinstance(node, ast.BinOp):
            name = {ast.Add: 'add', ast.Sub: 'sub', ast.Mult: 'mul', ast.Div: 'div', ast.Pow: 'pow'}.get(type(node.op))
            if name is None:
                raise NotImplementedError("constraint_expr %r: operator %s is not supported" % (expr, type(node.op).__name__))
            return binary(name, ev(node.left), ev(node.right))
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and not node.keywords:
            fn = node.func.id
            if fn in _FUNCS and len(node.args) == 1:
                return unary(_FUNCS[fn], ev(node.args[0]))
            if fn in _FUNCS2 and len(node.args) == 2:
                return binary(_FUNCS2[fn], ev(node.args[0]), ev(node.args[1]))
            raise NotImplementedError("constraint_expr %r: function %s() is not supported" % (expr, fn))
        raise NotImplementedError("constraint_expr %r: only arithmetic on numbers and parameter names is supported" % expr)

    words = ev(tree)
    depth = peak = 0
    for w in words:
        op = abi.TIE_OPS[-w[0] - abi.TIE_WORD]
        depth += 1 if op in ('pushc', 'pushp') else (-1 if op in fold2 else 0)
        peak = max(peak, depth)
    if peak > abi.TIE_STACK:
        raise CapacityError("constraint_expr %r needs an expression stack of %d values (XRSD_TIE_STACK = %d)" % (expr, peak, abi.TIE_STACK))
    return words


def parse_linear_constraint(expr, names):
    """([(source slot, coefficient), ...], offset) of a constraint expression that is linear in parameter names;
    raises NotImplementedError otherwise.  The expression is parsed with `ast` (numbers, names, + - * /, unary
    minus, parentheses) and evaluated symbolically as a linear form, so linearity is established exactly."""
    import ast
    try:
        tree = ast.parse(expr.strip(), mode='eval')
    except SyntaxError:
        raise NotImplementedError("constraint_expr %r is not an arithmetic expression" % expr)

    def ev(node):           # -> (dict name -> coefficient, constant)
        if isinstance(node, ast.Expression):
            return ev(node.body)
        if isinstance(node, ast.Constant) and isinstance(node.value, (int, float)) and not isinstance(node.value, bool):
            return {}, float(node.value)
        if isinstance(node, ast.Name):
            if node.id in _CONSTS and node.id not in names:
                return {}, float(_CONSTS[node.id])
            if node.id not in names:
                raise NotImplementedError("constraint_expr %r: unknown parameter %r" % (expr, node.id))
            return {node.id: 1.0}, 0.0
        if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
            c, k = ev(node.operand)
            sg = -1.0 if isinstance(node.op, ast.USub) else 1.0
            return dict((n, sg * v) for n, v in c.items()), sg * k
        if isinstance(node, ast.BinOp) and isinstance(node.op, (ast.Add, ast.Sub)):
            (ca, ka), (cb, kb) = ev(node.left), ev(node.right)
            sg = -1.0 if isinstance(node.op, ast.Sub) else 1.0
            out = dict(ca)
            for n, v in cb.items():
                out[n] = out.get(n, 0.0) + sg * v
            return out, ka + sg * kb
        if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Mult):
            (ca, ka), (cb, kb) = ev(node.left), ev(node.right)
            if ca and cb:
                raise NotLinear("constraint_expr %r is not linear (a product of parameters)" % expr)
            c, k, f = (ca, ka, kb) if ca else (cb, kb, ka)
            return dict((n, v * f) for n, v in c.items()), k * f
        if isinstance(node, ast.BinOp) and isinstance(node.op, ast.Div):
            (ca, ka), (cb, kb) = ev(node.left), ev(node.right)
            if cb:
                raise NotLinear("constraint_expr %r is not linear (division by a parameter)" % expr)
            if kb == 0.0:
                raise NotImplementedError("constraint_expr %r divides by zero" % expr)
            return dict((n, v / kb) for n, v in ca.items()), ka / kb
        raise NotLinear("constraint_expr %r: only + - * / of numbers and parameter names are linear" % expr)

    coefs, off = ev(tree)
    if not np.isfinite([off] + list(coefs.values())).all():
        raise NotImplementedError("constraint_expr %r does not evaluate to finite coefficients" % expr)
    return [(names.index(n), float(v)) for n, v in coefs.items() if v != 0.0], float(off)


def parse_affine_constraint(expr, names):
    """(source slot, scale, offset) of a constraint expression that is affine in exactly ONE parameter name (kept for
    callers of the round-1 interface); see parse_linear_constraint for the general linear form."""
    terms, off = parse_linear_constraint(expr, names)
    if len(terms) != 1:
        raise NotImplementedError("constraint_expr %r must name exactly one other parameter" % expr)
    return terms[0][0], terms[0][1], off


def uniform_step(q):
    """Step of q if it is a linspace-like grid (every point within 4 ulp of q[0] + i*step, strictly ascending),
    else 0.  Evaluated by the library's host helper (one pass in C: this sits on the single-pattern latency
    path); the numpy statement of the same test serves when the library cannot be loaded."""
    n = len(q)
    if n < 3:
        return 0.0
    try:
        lib = abi.load()
    except ImportError:
        lib = None
    if lib is not None and isinstance(q, np.ndarray) and q.dtype == np.float64 and q.flags.c_contiguous:
        return float(lib.xrsd_uniform_step(abi.ptr(q), n))
    step = (q[-1] - q[0]) / (n - 1)
    if not (step > 0):
        return 0.0
    tol = 4 * np.finfo(float).eps * max(abs(q[0]), abs(q[-1]))
    if np.max(np.abs(q - (q[0] + np.arange(n) * step))) <= tol and np.all(np.diff(q) > 0):
        return float(step)
    return 0.0


def lower_system(system, q):
    """System -> SystemLayout for the grid `q` (q is needed only for q[0]==0 and the q_max fallback)."""
    q = np.asarray(q, dtype=float)
    names, values, records = [], [], []
    L = abi.Layout()

    def take(prefix, params):
        slots = {}
        for nm, rec in params.items():
            slots[nm] = len(names)
            names.append(prefix + '__' + nm)
            values.append(float(_val(rec)))
            records.append(rec)
        return slots

    pops = [_lower_noise(system.noise_model.model, take('noise', system.noise_model.parameters))]
    n_x = 0
    for pop_name, pop in system.populations.items():
        slots = take(pop_name, pop.parameters)
        lp = _lower_population(pop.structure, pop.form, pop.settings, slots, q[-1] if len(q) else 0.0, n_x, pop.parameters)
        if lp.kind == abi.POP_CRYSTALLINE:
            n_x += 1
        pops.append(lp)
    if len(pops) > abi.MAX_POPS:
        raise CapacityError('a system layout holds at most {} populations besides the noise model (XRSD_MAX_POPS = {}); '
                            'this system has {}'.format(abi.MAX_POPS - 1, abi.MAX_POPS, len(pops) - 1))
    if len(names) > abi.MAX_PARAMS:
        raise CapacityError('a system layout holds at most {} parameters (XRSD_MAX_PARAMS); this system has {}'
                            .format(abi.MAX_PARAMS, len(names)))
    L.n_pops = len(pops)
    L.n_params = max(1, len(names))
    L.n_xtal = n_x
    L.q0_is_zero = 1 if (len(q) and q[0] == 0) else 0
    L.q_step = uniform_step(q)
    wl = system.sample_metadata['source_wavelength']
    if n_x and not wl:
        raise ValueError('cannot compute diffraction with source wavelength of {}'.format(wl))
    L.wavelength = float(wl) if wl else 0.0
    for i, p in enumerate(pops):
        L.pops[i] = p
    if not names:
        names, values, records = ['_unused'], [0.0], [dict(value=0.0, fixed=True, bounds=[None, None], constraint_expr=None)]
    lay = SystemLayout(L, names, np.array(values, dtype=float), records)
    uses_q_last = any(not pop.settings.get('q_max') for pop in system.populations.values() if pop.structure == 'crystalline')
    lay.grid = dict(n=len(q), q0=float(q[0]) if len(q) else 0.0, q_last=float(q[-1]) if len(q) else 0.0,
                    q_step=float(L.q_step), q0_is_zero=bool(L.q0_is_zero), uses_q_last=bool(uses_q_last))
    return lay


class _Shell(object):
    """Minimal stand-in for a System around one population or one noise model."""

    def __init__(self, noise_model, populations, wavelength):
        self.noise_model = noise_model
        self.populations = populations
        self.sample_metadata = dict(source_wavelength=wavelength)


class _ZeroNoise(object):
    model = 'flat'
    parameters = {'I0': {'value': 0.0, 'fixed': True, 'bounds': [0., None], 'constraint_expr': None}}


def lower_population_only(structure, form, settings, parameters, q, wavelength):
    class _P(object):
        pass
    pop = _P()
    pop.structure, pop.form, pop.settings, pop.parameters = structure, form, settings, parameters
    return lower_system(_Shell(_ZeroNoise(), {'pop': pop}, wavelength), q)


def lower_noise_only(noise_model, q):
    return lower_system(_Shell(noise_model, {}, 1.0), q)


# ---------------------------------------------------------------------------------------
# host-side bound on the candidate hkl box (sizes the table builder's shared memory)
# ---------------------------------------------------------------------------------------
def box_half_widths(c_pop, params):
    """Upper bound of n1,n2,n3 = ceil(G_max |a_i| / (b_i.a_i)) (scattering/__init__.py:266-268)
    over a [batch, n_params] matrix.  |a_i| are the cell edge lengths for every lattice class
    and b_i.a_i = 1 up to rounding, so n_i <= ceil(G_max * edge_i) + 1."""
    params = np.atleast_2d(params)

    def col(i):
        return params[:, c_pop.p_lat[i]] if c_pop.p_lat[i] >= 0 else None

    a, b, c = col(0), col(1), col(2)
    lat = c_pop.lattice
    if lat == abi.LAT_CUBIC:
        e = [a, a, a]
    elif lat == abi.LAT_HCP:
        e = [a, a, a * np.sqrt(8. / 3.)]
    elif lat in (abi.LAT_HEXAGONAL, abi.LAT_TETRAGONAL):
        e = [a, a, c]
    elif lat == abi.LAT_RHOMBOHEDRAL:
        e = [a, a, a]
    else:
        e = [a, b, c]
    g_max = c_pop.q_max / (2 * np.pi)
    out = []
    for edge in e:
        m = np.nanmax(np.abs(edge)) if edge is not None and len(edge) else 0.0
        n = np.ceil(g_max * m * (1 + 1e-12)) + 1
        out.append(int(max(1, min(n, 10000))))
    return out


def max_cells(layout, params):
    cells = 1
    for p in layout.crystalline_pops():
        n1, n2, n3 = box_half_widths(p, params)
        cells = max(cells, (2 * n1 - 1) * (2 * n2 - 1) * (2 * n3 - 1))
    return int(cells)
