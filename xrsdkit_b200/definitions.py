"""Definition tables of the intensity path: structures, forms, settings, parameters,
lattices, space groups and the atomic form-factor coefficients.

Same *content* as the reference's xrsdkit/definitions.py (and its
scattering/atomic_scattering_params.yml), which the drop-in API has to reproduce as data:
defaults and bounds (definitions.py:87-149, 181-216, 255-318), lattice names (:153-161),
the space-group table (:334-460, including the reference's 'F4-3m' spelling of group 216)
and the lattice geometry (:462-570).  tests/test_definitions.py compares every table with a
dump taken from the unmodified reference (tests/golden/definitions.json).

The tables are stored as compact text blocks and expanded at import.
"""
from collections import OrderedDict
import copy

import numpy as np

structure_names = ['diffuse', 'disordered', 'crystalline']
form_factor_names = ['atomic', 'polyatomic', 'guinier_porod', 'spherical']
noise_model_names = ['flat', 'low_q_scatter']

structures = OrderedDict([
    ('diffuse', 'disordered, non-interacting particles'),
    ('disordered', 'disordered, interacting particles'),
    ('crystalline', 'particles arranged in a lattice')])
form_factors = OrderedDict([
    ('atomic', 'Single atom'),
    ('polyatomic', 'Multiple atoms'),
    ('guinier_porod', 'Scatterer described by Guinier-Porod equations'),
    ('spherical', 'Spherical particle')])
noise_models = OrderedDict([
    ('flat', 'Flat noise floor for all q'),
    ('low_q_scatter', 'Flat noise floor plus a Guinier-Porod-like contribution')])


def _param(value, fixed, lo, hi):
    return {'value': value, 'fixed': fixed, 'bounds': [lo, hi], 'constraint_expr': None}


# ---------------------------------------------------------------------------------------
# settings (reference definitions.py:87-117, 181-216)
# ---------------------------------------------------------------------------------------
structure_settings = dict(
    diffuse=OrderedDict(),
    disordered=OrderedDict([('interaction', 'hard_spheres')]),
    crystalline=OrderedDict([
        ('lattice', 'P_cubic'), ('space_group', ''), ('texture', 'random'),
        ('integration_mode', 'spherical'), ('use_symmetry', True),
        ('structure_factor_mode', 'local'), ('profile', 'voigt'),
        ('polarization_correction', True), ('lorentz_correction', True)]))
form_settings = dict(
    atomic=OrderedDict([('symbol', 'C')]),
    polyatomic=OrderedDict([('n_atoms', 2)]),
    guinier_porod=OrderedDict(),
    spherical=OrderedDict([('distribution', 'single')]))

modelable_structure_settings = dict(diffuse=[], disordered=['interaction'], crystalline=['lattice'])
modelable_form_factor_settings = dict(atomic=['symbol'], polyatomic=['n_atoms'], guinier_porod=[],
                                      spherical=['distribution'])


def secondary_settings(structure, form, primary_settings):
    """Settings implied by the primary ones, with defaults (reference definitions.py:181-216)."""
    out = OrderedDict()
    for name, val in primary_settings.items():
        if name == 'n_atoms':
            for i in range(val):
                out['symbol_%d' % i] = 'H'
        elif name == 'integration_mode' and val == 'spherical':
            out['q_min'] = 0.
            out['q_max'] = 1.
        elif name == 'distribution' and form == 'spherical' and val == 'r_normal':
            out['sampling_width'] = 3.5
            out['sampling_step'] = 0.05
    return out


def setting_datatypes(stg_nm):
    if stg_nm in ('lattice', 'space_group', 'texture', 'profile', 'structure_factor_mode',
                  'integration_mode', 'interaction', 'distribution'):
        return str
    if stg_nm in ('q_min', 'q_max', 'sampling_width', 'sampling_step'):
        return float
    if stg_nm == 'n_atoms':
        return int
    if 'symbol' in stg_nm:
        return str
    if stg_nm in ('polarization_correction', 'lorentz_correction', 'use_symmetry'):
        return bool


def setting_selections(stg_nm, structure=None, form=None, prior_settings={}):
    if setting_datatypes(stg_nm) == bool:
        return [True, False]
    fixed = dict(lattice=all_lattices, texture=['random'], profile=['gaussian', 'lorentzian', 'voigt'],
                 structure_factor_mode=['local', 'radial'], integration_mode=['spherical'],
                 interaction=['hard_spheres'])
    if stg_nm in fixed:
        return fixed[stg_nm]
    if stg_nm == 'space_group':
        if 'lattice' in prior_settings:
            return [''] + list(lattice_space_groups[prior_settings['lattice']].values())
        return ['']
    if stg_nm == 'distribution' and form == 'spherical':
        return ['single', 'r_normal']
    if stg_nm in ('q_min', 'q_max', 'sampling_width', 'sampling_step', 'n_atoms'):
        return []
    if 'symbol' in stg_nm:
        return list(atomic_params.keys())


def validate(structure, form, settings):
    """Legal structure/form/settings combinations (reference definitions.py:64-79)."""
    if structure in ('diffuse', 'disordered'):
        if form == 'polyatomic':
            raise ValueError('{} structure does not support polyatomic forms'.format(structure))
    elif structure == 'crystalline':
        if form == 'guinier_porod':
            raise ValueError('crystalline structure does not support guinier_porod forms')
        if form == 'spherical' and 'distribution' in settings and settings['distribution'] != 'single':
            raise ValueError('crystalline structure does not support size distribution {}'
                             .format(settings['distribution']))


# ---------------------------------------------------------------------------------------
# parameters (reference definitions.py:123-149, 255-318)
# ---------------------------------------------------------------------------------------
form_factor_params = dict(
    atomic=OrderedDict(),
    polyatomic=OrderedDict(),
    guinier_porod=OrderedDict([('rg', _param(10., False, 0.1, None)), ('D', _param(4., True, 0., 4.))]),
    spherical=OrderedDict([('r', _param(20., False, 0.1, None))]))
noise_params = dict(
    flat=OrderedDict([('I0', _param(1.E-3, False, 0., None))]),
    low_q_scatter=OrderedDict([
        ('I0', _param(100, False, 0., None)),
        ('I0_flat_fraction', _param(0.01, False, 0., 1.)),
        ('effective_rg', _param(100., False, 0.1, None)),
        ('effective_D', _param(4., False, 0., 4.))]))

# lattice -> names of its cell parameters (lengths default 10/12/15[/20 for c of a,c cells], angles 90)
_CELL_PARAMS = {}
for _lat in ('P_cubic', 'I_cubic', 'F_cubic', 'diamond', 'hcp'):
    _CELL_PARAMS[_lat] = ['a']
for _lat in ('hexagonal', 'P_tetragonal', 'I_tetragonal'):
    _CELL_PARAMS[_lat] = ['a', 'c']
_CELL_PARAMS['rhombohedral'] = ['a', 'alpha']
for _lat in ('P_orthorhombic', 'C_orthorhombic', 'I_orthorhombic', 'F_orthorhombic'):
    _CELL_PARAMS[_lat] = ['a', 'b', 'c']
for _lat in ('P_monoclinic', 'C_monoclinic'):
    _CELL_PARAMS[_lat] = ['a', 'b', 'c', 'beta']
_CELL_PARAMS['triclinic'] = ['a', 'b', 'c', 'alpha', 'beta', 'gamma']
# NB 'A_orthorhombic' gets no cell parameters from the reference (definitions.py:262-288)


def lattice_param_names(lattice):
    return list(_CELL_PARAMS.get(lattice, []))


def structure_params(structure, prior_settings):
    params = OrderedDict()
    if structure == 'disordered' and prior_settings.get('interaction') == 'hard_spheres':
        params['r_hard'] = _param(20., False, 1.E-1, None)
        params['v_fraction'] = _param(0.5, False, 0.01, 0.7405)
    if structure == 'crystalline':
        names = _CELL_PARAMS.get(prior_settings.get('lattice'), [])
        two = (names == ['a', 'c'])
        for nm in names:
            if nm in ('alpha', 'beta', 'gamma'):
                params[nm] = _param(90., False, 0, 180.)
            else:
                params[nm] = _param({'a': 10., 'b': 12., 'c': 20. if two else 15.}[nm], False, 1.E-1, None)
        prof = prior_settings.get('profile')
        if prof == 'voigt':
            params['hwhm_g'] = _param(1.E-3, False, 1.E-9, None)
            params['hwhm_l'] = _param(1.E-3, False, 1.E-9, None)
        elif prof in ('gaussian', 'lorentzian'):
            params['hwhm'] = _param(1.E-3, False, 1.E-9, None)
    return params


def additional_form_factor_params(form, prior_settings):
    params = OrderedDict()
    if form == 'polyatomic' and 'n_atoms' in prior_settings:
        for i in range(prior_settings['n_atoms']):
            params['occupancy_%d' % i] = _param(1., True, 0., 1.)
            for ax in 'uvw':
                params['%s_%d' % (ax, i)] = _param(0.1 * i, True, -1., 1.)
    if form == 'spherical' and prior_settings.get('distribution') == 'r_normal':
        params['sigma'] = _param(0.05, False, 0., 2.)
    return params


def all_params(structure, form=None, prior_settings={}):
    out = OrderedDict()
    out['I0'] = _param(1., False, 0., None)
    out.update(structure_params(structure, prior_settings))
    if form:
        out.update(copy.deepcopy(form_factor_params[form]))
        out.update(additional_form_factor_params(form, prior_settings))
    return out


# ---------------------------------------------------------------------------------------
# lattices, space groups, point groups (reference definitions.py:151-161, 320-460)
# ---------------------------------------------------------------------------------------
crystal_systems = ['triclinic', 'monoclinic', 'orthorhombic', 'tetragonal', 'trigonal', 'hexagonal', 'cubic']
bravais_lattices = ['triclinic', 'P_monoclinic', 'C_monoclinic', 'P_orthorhombic', 'C_orthorhombic',
                    'A_orthorhombic', 'I_orthorhombic', 'F_orthorhombic', 'P_tetragonal', 'I_tetragonal',
                    'rhombohedral', 'hexagonal', 'P_cubic', 'I_cubic', 'F_cubic']
all_lattices = bravais_lattices + ['hcp', 'diamond']

_low = ['2', '2/m', '222', 'm', 'mm2', 'mmm']
crystal_point_groups = dict(
    triclinic=['1', '-1'], monoclinic=list(_low), orthorhombic=list(_low),
    tetragonal=['4', '-4', '4/m', '422', '4mm', '-42m', '4/mmm'],
    trigonal=['3', '-3', '32', '3m', '-3m'],
    hexagonal=['6', '-6', '6/m', '622', '6mm', '-6m2', '6/mmm'],
    cubic=['23', 'm-3', '432', '-43m', 'm-3m'])
all_point_groups = []
for _xs in crystal_systems:
    all_point_groups.extend(crystal_point_groups[_xs])

# number, symbol, lattice, point group
_SPACE_GROUP_TABLE = """
  1 P1             triclinic       1
  2 P-1            triclinic       -1
  3 P2             P_monoclinic    2
  4 P2(1)          P_monoclinic    2
  5 C2             C_monoclinic    2
  6 Pm             P_monoclinic    m
  7 Pc             P_monoclinic    m
  8 Cm             C_monoclinic    m
  9 Cc             C_monoclinic    m
 10 P2/m           P_monoclinic    2/m
 11 P2(1)/m        P_monoclinic    2/m
 12 C2/m           C_monoclinic    2/m
 13 P2/c           P_monoclinic    2/m
 14 P2(1)/c        P_monoclinic    2/m
 15 C2/c           C_monoclinic    2/m
 16 P222           P_orthorhombic  222
 17 P222(1)        P_orthorhombic  222
 18 P2(1)2(1)2     P_orthorhombic  222
 19 P2(1)2(1)2(1)  P_orthorhombic  222
 20 C222(1)        C_orthorhombic  222
 21 C222           C_orthorhombic  222
 22 F222           F_orthorhombic  222
 23 I222           I_orthorhombic  222
 24 I2(1)2(1)2(1)  I_orthorhombic  222
 25 Pmm2           P_orthorhombic  mm2
 26 Pmc2(1)        P_orthorhombic  mm2
 27 Pcc2           P_orthorhombic  mm2
 28 Pma2           P_orthorhombic  mm2
 29 Pca2(1)        P_orthorhombic  mm2
 30 Pnc2           P_orthorhombic  mm2
 31 Pmn2(1)        P_orthorhombic  mm2
 32 Pba2           P_orthorhombic  mm2
 33 Pna2(1)        P_orthorhombic  mm2
 34 Pnn2           P_orthorhombic  mm2
 35 Cmm2           C_orthorhombic  mm2
 36 Cmc2(1)        C_orthorhombic  mm2
 37 Ccc2           C_orthorhombic  mm2
 38 Amm2           A_orthorhombic  mm2
 39 Aem2           A_orthorhombic  mm2
 40 Ama2           A_orthorhombic  mm2
 41 Aea2           A_orthorhombic  mm2
 42 Fmm2           F_orthorhombic  mm2
 43 Fdd2           F_orthorhombic  mm2
 44 Imm2           I_orthorhombic  mm2
 45 Iba2           I_orthorhombic  mm2
 46 Ima2           I_orthorhombic  mm2
 47 Pmmm           P_orthorhombic  mmm
 48 Pnnn           P_orthorhombic  mmm
 49 Pccm           P_orthorhombic  mmm
 50 Pban           P_orthorhombic  mmm
 51 Pmma           P_orthorhombic  mmm
 52 Pnna           P_orthorhombic  mmm
 53 Pmna           P_orthorhombic  mmm
 54 Pcca           P_orthorhombic  mmm
 55 Pbam           P_orthorhombic  mmm
 56 Pccn           P_orthorhombic  mmm
 57 Pbcm           P_orthorhombic  mmm
 58 Pnnm           P_orthorhombic  mmm
 59 Pmmn           P_orthorhombic  mmm
 60 Pbcn           P_orthorhombic  mmm
 61 Pbca           P_orthorhombic  mmm
 62 Pnma           P_orthorhombic  mmm
 63 Cmcm           C_orthorhombic  mmm
 64 Cmce           C_orthorhombic  mmm
 65 Cmmm           C_orthorhombic  mmm
 66 Cccm           C_orthorhombic  mmm
 67 Cmme           C_orthorhombic  mmm
 68 Ccce           C_orthorhombic  mmm
 69 Fmmm           F_orthorhombic  mmm
 70 Fddd           F_orthorhombic  mmm
 71 Immm           I_orthorhombic  mmm
 72 Ibam           I_orthorhombic  mmm
 73 Ibca           I_orthorhombic  mmm
 74 Imma           I_orthorhombic  mmm
 75 P4             P_tetragonal    4
 76 P4(1)          P_tetragonal    4
 77 P4(2)          P_tetragonal    4
 78 P4(3)          P_tetragonal    4
 79 I4             I_tetragonal    4
 80 I4(1)          I_tetragonal    4
 81 P-4            P_tetragonal    -4
 82 I-4            I_tetragonal    -4
 83 P4/m           P_tetragonal    4/m
 84 P4(2)/m        P_tetragonal    4/m
 85 P4/n           P_tetragonal    4/m
 86 P4(2)/n        P_tetragonal    4/m
 87 I4/m           I_tetragonal    4/m
 88 I4(1)/a        I_tetragonal    4/m
 89 P422           P_tetragonal    422
 90 P42(1)2        P_tetragonal    422
 91 P4(1)22        P_tetragonal    422
 92 P4(1)2(1)2     P_tetragonal    422
 93 P4(2)22        P_tetragonal    422
 94 P4(2)2(1)2     P_tetragonal    422
 95 P4(3)22        P_tetragonal    422
 96 P4(3)2(1)2     P_tetragonal    422
 97 I422           I_tetragonal    422
 98 I4(1)22        I_tetragonal    422
 99 P4mm           P_tetragonal    4mm
100 P4bm           P_tetragonal    4mm
101 P4(2)cm        P_tetragonal    4mm
102 P4(2)nm        P_tetragonal    4mm
103 P4cc           P_tetragonal    4mm
104 P4nc           P_tetragonal    4mm
105 P4(2)mc        P_tetragonal    4mm
106 P4(2)bc        P_tetragonal    4mm
107 I4mm           I_tetragonal    4mm
108 I4cm           I_tetragonal    4mm
109 I4(1)md        I_tetragonal    4mm
110 I4(1)cd        I_tetragonal    4mm
111 P-42m          P_tetragonal    -42m
112 P-42c          P_tetragonal    -42m
113 P-42(1)m       P_tetragonal    -42m
114 P-42(1)c       P_tetragonal    -42m
115 P-4m2          P_tetragonal    -42m
116 P-4c2          P_tetragonal    -42m
117 P-4b2          P_tetragonal    -42m
118 P-4n2          P_tetragonal    -42m
119 I-4m2          I_tetragonal    -42m
120 I-4c2          I_tetragonal    -42m
121 I-42m          I_tetragonal    -42m
122 I-42d          I_tetragonal    -42m
123 P4/mmm         P_tetragonal    4/mmm
124 P4/mcc         P_tetragonal    4/mmm
125 P4/nbm         P_tetragonal    4/mmm
126 P4/nnc         P_tetragonal    4/mmm
127 P4/mbm         P_tetragonal    4/mmm
128 P4/mnc         P_tetragonal    4/mmm
129 P4/nmm         P_tetragonal    4/mmm
130 P4/ncc         P_tetragonal    4/mmm
131 P4(2)/mmc      P_tetragonal    4/mmm
132 P4(2)/mcm      P_tetragonal    4/mmm
133 P4(2)/nbc      P_tetragonal    4/mmm
134 P4(2)/nnm      P_tetragonal    4/mmm
135 P4(2)/mbc      P_tetragonal    4/mmm
136 P4(2)/mnm      P_tetragonal    4/mmm
137 P4(2)/nmc      P_tetragonal    4/mmm
138 P4(2)/ncm      P_tetragonal    4/mmm
139 I4/mmm         I_tetragonal    4/mmm
140 I4/mcm         I_tetragonal    4/mmm
141 I4(1)/amd      I_tetragonal    4/mmm
142 I4(1)/acd      I_tetragonal    4/mmm
143 P3             hexagonal       3
144 P3(1)          hexagonal       3
145 P3(2)          hexagonal       3
146 R3             rhombohedral    3
147 P-3            hexagonal       -3
148 R-3            rhombohedral    -3
149 P312           hexagonal       32
150 P321           hexagonal       32
151 P3(1)12        hexagonal       32
152 P3(1)21        hexagonal       32
153 P3(2)12        hexagonal       32
154 P3(2)21        hexagonal       32
155 R32            rhombohedral    32
156 P3m1           hexagonal       3m
157 P31m           hexagonal       3m
158 P3c1           hexagonal       3m
159 P31c           hexagonal       3m
160 R3m            rhombohedral    3m
161 R3c            rhombohedral    3m
162 P-31m          hexagonal       -3m
163 P-31c          hexagonal       -3m
164 P-3m1          hexagonal       -3m
165 P-3c1          hexagonal       -3m
166 R-3m           rhombohedral    -3m
167 R-3c           rhombohedral    -3m
168 P6             hexagonal       6
169 P6(1)          hexagonal       6
170 P6(5)          hexagonal       6
171 P6(2)          hexagonal       6
172 P6(4)          hexagonal       6
173 P6(3)          hexagonal       6
174 P-6            hexagonal       -6
175 P6/m           hexagonal       6/m
176 P6(3)/m        hexagonal       6/m
177 P622           hexagonal       622
178 P6(1)22        hexagonal       622
179 P6(5)22        hexagonal       622
180 P6(2)22        hexagonal       622
181 P6(4)22        hexagonal       622
182 P6(3)22        hexagonal       622
183 P6mm           hexagonal       6mm
184 P6cc           hexagonal       6mm
185 P6(3)cm        hexagonal       6mm
186 P6(3)mc        hexagonal       6mm
187 P-6m2          hexagonal       -6m2
188 P-6c2          hexagonal       -6m2
189 P-62m          hexagonal       -6m2
190 P-62c          hexagonal       -6m2
191 P6/mmm         hexagonal       6/mmm
192 P6/mcc         hexagonal       6/mmm
193 P6(3)/mcm      hexagonal       6/mmm
194 P6(3)/mmc      hexagonal       6/mmm
195 P23            P_cubic         23
196 F23            F_cubic         23
197 I23            I_cubic         23
198 P2(1)3         P_cubic         23
199 I2(1)3         I_cubic         23
200 Pm-3           P_cubic         m-3
201 Pn-3           P_cubic         m-3
202 Fm-3           F_cubic         m-3
203 Fd-3           F_cubic         m-3
204 Im-3           I_cubic         m-3
205 Pa-3           P_cubic         m-3
206 Ia-3           I_cubic         m-3
207 P432           P_cubic         432
208 P4(2)32        P_cubic         432
209 F432           F_cubic         432
210 F4(1)32        F_cubic         432
211 I432           I_cubic         432
212 P4(3)32        P_cubic         432
213 P4(1)32        P_cubic         432
214 I4(1)32        I_cubic         432
215 P-43m          P_cubic         -43m
216 F4-3m          F_cubic         -43m
217 I-43m          I_cubic         -43m
218 P-43n          P_cubic         -43m
219 F-43c          F_cubic         -43m
220 I-43d          I_cubic         -43m
221 Pm-3m          P_cubic         m-3m
222 Pn-3n          P_cubic         m-3m
223 Pm-3n          P_cubic         m-3m
224 Pn-3m          P_cubic         m-3m
225 Fm-3m          F_cubic         m-3m
226 Fm-3c          F_cubic         m-3m
227 Fd-3m          F_cubic         m-3m
228 Fd-3c          F_cubic         m-3m
229 Im-3m          I_cubic         m-3m
230 Ia-3d          I_cubic         m-3m
"""

lattice_space_groups = dict((lat, {}) for lat in bravais_lattices)
all_space_groups = {}
sg_point_groups = OrderedDict()
for _line in _SPACE_GROUP_TABLE.strip().splitlines():
    _num, _sym, _lat, _pg = _line.split()
    lattice_space_groups[_lat][int(_num)] = _sym
    all_space_groups[int(_num)] = _sym
    sg_point_groups[_sym] = _pg
lattice_space_groups['hcp'] = {194: 'P6(3)/mmc'}
lattice_space_groups['diamond'] = {227: 'Fd-3m'}


# ---------------------------------------------------------------------------------------
# lattice geometry (reference definitions.py:462-570)
# ---------------------------------------------------------------------------------------
_SITES = {
    'P': [[0., 0., 0.]],
    'C': [[0., 0., 0.], [0., 0.5, 0.5]],
    'A': [[0., 0., 0.], [0.5, 0.5, 0.]],
    'I': [[0., 0., 0.], [0.5, 0.5, 0.5]],
    'F': [[0., 0., 0.], [0.5, 0.5, 0.], [0.5, 0., 0.5], [0., 0.5, 0.5]],
}
_CENTRING = dict(triclinic='P', rhombohedral='P', hexagonal='P')


def lattice_coords(lattice):
    """Basis sites of the conventional cell, fractional coordinates."""
    if lattice == 'hcp':
        return np.array([[0., 0., 0.], [2. / 3, 1. / 3, 0.5]])
    if lattice == 'diamond':
        return np.array(_SITES['F'] + [[0.25, 0.25, 0.25], [0.75, 0.75, 0.25], [0.75, 0.25, 0.75], [0.25, 0.75, 0.75]])
    if lattice not in bravais_lattices:
        return None
    return np.array(_SITES[_CENTRING.get(lattice, lattice[0])])


def lattice_vectors(lattice, a=None, b=None, c=None, alpha=None, beta=None, gamma=None):
    """Cartesian cell vectors a1, a2, a3 (lists)."""
    rad = np.pi / 180.
    if lattice in ('P_cubic', 'I_cubic', 'F_cubic', 'diamond'):
        b, c = a, a
    elif lattice in ('P_tetragonal', 'I_tetragonal'):
        b = a
    if lattice in ('P_cubic', 'I_cubic', 'F_cubic', 'diamond', 'P_tetragonal', 'I_tetragonal') \
            or lattice.endswith('_orthorhombic'):
        return [a, 0., 0.], [0., b, 0.], [0., 0., c]
    if lattice in ('hcp', 'hexagonal'):
        c3 = np.sqrt(8. / 3.) * a if lattice == 'hcp' else c
        return [a, 0., 0.], [0.5 * a, np.sqrt(3.) / 2 * a, 0.], [0., 0., c3]
    if lattice in ('P_monoclinic', 'C_monoclinic'):
        br = float(beta) * rad
        return [a, 0., 0.], [0., b, 0.], [c * np.cos(br), 0., c * np.sin(br)]
    if lattice == 'rhombohedral':
        b, c, beta, gamma = a, a, alpha, alpha
    if lattice in ('rhombohedral', 'triclinic'):
        ar, br, gr = float(alpha) * rad, float(beta) * rad, float(gamma) * rad
        ca, cb, cg = np.cos(ar), np.cos(br), np.cos(gr)
        omega = a * b * c * np.sqrt(1. - ca ** 2 - cb ** 2 - cg ** 2 + 2 * ca * cb * cg)
        cy = c * (ca - cb * cg) / np.sin(gr)
        cz = omega / (a * b * np.sin(gr))
        return [a, 0., 0.], [b * cg, b * np.sin(gr), 0.], [c * cb, cy, cz]


def reciprocal_lattice_vectors(lat1, lat2, lat3, crystallographic=True):
    """b_i = (a_j x a_k)/V; times 2 pi unless `crystallographic`."""
    x1, x2, x3 = np.cross(lat2, lat3), np.cross(lat3, lat1), np.cross(lat1, lat2)
    vol = np.dot(lat1, x1)
    scale = 1. if crystallographic else 2 * np.pi
    return scale * x1 / vol, scale * x2 / vol, scale * x3 / vol


# ---------------------------------------------------------------------------------------
# atomic scattering coefficients: symbol Z a1..a4 b1..b4
# (reference scattering/atomic_scattering_params.yml, loaded at definitions.py:58-62)
# ---------------------------------------------------------------------------------------
_ATOMIC_TABLE = """
H    1  0.202 0.244 0.082 0.0  30.868 8.544 1.273 0.0
He   2  0.091 0.181 0.11 0.036  18.183 6.212 1.803 0.284
Li   3  1.611 1.246 0.326 0.099  107.638 30.48 4.533 0.495
Be   4  1.25 1.334 0.36 0.106  60.804 18.591 3.653 0.416
B    5  0.945 1.312 0.419 0.116  46.444 14.178 3.223 0.377
C    6  0.731 1.195 0.456 0.125  36.995 11.297 2.814 0.346
N    7  0.572 1.043 0.465 0.131  28.847 9.054 2.421 0.317
O    8  0.455 0.917 0.472 0.138  23.78 7.622 2.144 0.296
F    9  0.387 0.811 0.475 0.146  20.239 6.609 1.931 0.279
Ne  10  0.303 0.72 0.475 0.153  17.64 5.86 1.762 0.266
Na  11  2.241 1.333 0.907 0.286  108.004 24.505 3.391 0.435
Mg  12  2.268 1.803 0.839 0.289  73.67 20.175 3.013 0.405
Al  13  2.276 2.428 0.858 0.317  72.322 19.773 3.08 0.408
Si  14  2.129 2.533 0.835 0.322  57.775 16.476 2.88 0.386
P   15  1.888 2.469 0.805 0.32  44.876 13.538 2.642 0.361
S   16  1.659 2.386 0.79 0.321  36.65 11.488 2.469 0.34
Cl  17  1.452 2.292 0.787 0.322  30.935 9.98 2.234 0.323
Ar  18  1.274 2.19 0.793 0.326  26.682 8.813 2.219 0.307
K   19  3.951 2.545 1.98 0.482  137.075 22.402 4.532 0.434
Ca  20  4.47 2.971 1.97 0.482  99.523 22.696 4.195 0.417
Sc  21  3.966 2.917 1.925 0.48  88.96 20.606 3.856 0.399
Ti  22  3.565 2.818 1.893 0.483  81.982 19.049 3.59 0.386
V   23  3.245 2.698 1.86 0.486  76.379 17.726 3.363 0.374
Cr  24  2.307 2.334 1.823 0.49  78.405 15.785 3.157 0.364
Mn  25  2.747 2.456 1.792 0.498  67.786 15.674 3.0 0.357
Fe  26  2.544 2.343 1.759 0.506  64.424 14.88 2.854 0.35
Co  27  2.367 2.236 1.724 0.515  61.431 14.18 2.725 0.344
Ni  28  2.21 2.134 1.689 0.524  58.727 13.553 2.609 0.339
Cu  29  1.579 1.82 1.658 0.532  62.94 12.453 2.504 0.333
Zn  30  1.942 1.95 1.619 0.543  54.162 12.518 2.416 0.33
Ga  31  2.321 2.486 1.688 0.599  65.602 15.458 2.581 0.351
Ge  32  2.447 2.702 1.616 0.601  55.893 14.393 2.446 0.342
As  33  2.399 2.79 1.529 0.594  45.718 12.817 2.28 0.328
Se  34  2.298 2.854 1.456 0.59  38.83 11.536 2.146 0.316
Br  35  2.166 2.904 1.395 0.589  33.899 10.497 2.041 0.307
Kr  36  2.034 2.927 1.342 0.589  29.999 9.598 1.952 0.299
Rb  37  4.776 3.859 2.234 0.868  140.782 18.991 3.701 0.419
Sr  38  5.848 4.003 2.342 0.88  104.972 19.367 3.737 0.414
Y   39  4.129 3.012 1.179 0.0  27.548 5.088 0.591 0.0
Zr  40  4.105 3.144 1.229 0.0  28.492 5.277 0.601 0.0
Nb  41  4.237 3.105 1.234 0.0  27.415 5.074 0.593 0.0
Mo  42  3.12 3.906 2.361 0.85  72.464 14.642 3.237 0.366
Tc  43  4.318 3.27 1.287 0.0  28.246 5.148 0.59 0.0
Ru  44  4.358 3.298 1.323 0.0  27.881 5.179 0.594 0.0
Rh  45  4.431 3.343 1.345 0.0  27.911 5.153 0.592 0.0
Pd  46  4.436 3.454 1.383 0.0  28.67 5.269 0.595 0.0
Ag  47  2.036 3.272 2.511 0.837  61.497 11.824 2.846 0.327
Cd  48  2.574 3.259 2.547 0.838  55.675 11.838 2.784 0.322
In  49  3.153 3.557 2.818 0.884  66.649 14.449 2.976 0.335
Sn  50  3.45 3.735 2.118 0.877  59.104 14.179 2.855 0.327
Sb  51  3.564 3.844 2.687 0.864  50.487 13.316 2.691 0.316
Te  52  4.785 3.688 1.5 0.0  27.999 5.083 0.581 0.0
I   53  3.473 4.06 2.522 0.84  39.441 11.816 2.415 0.298
Xe  54  3.366 4.147 2.443 0.829  35.509 11.117 2.294 0.289
Cs  55  6.062 5.986 3.303 1.096  155.837 19.695 3.335 0.379
Ba  56  7.821 6.004 3.28 1.103  117.657 18.778 3.263 0.376
La  57  4.94 3.968 1.663 0.0  28.716 5.245 0.594 0.0
Ce  58  5.007 3.98 1.678 0.0  28.283 5.183 0.589 0.0
Pr  59  5.085 4.043 1.684 0.0  28.588 5.143 0.581 0.0
Nd  60  5.151 4.075 1.683 0.0  28.304 5.073 0.571 0.0
Pm  61  5.201 4.094 1.719 0.0  28.079 5.081 0.576 0.0
Sm  62  5.255 4.113 1.743 0.0  28.016 5.037 0.577 0.0
Eu  63  6.267 4.844 3.202 1.2  100.298 16.066 2.98 0.367
Gd  64  5.225 4.314 1.827 0.0  29.158 5.259 0.586 0.0
Tb  65  5.272 4.347 1.844 0.0  29.046 5.226 0.585 0.0
Dy  66  5.332 4.37 1.863 0.0  28.888 5.198 0.581 0.0
Ho  67  5.376 4.403 1.884 0.0  28.773 5.174 0.582 0.0
Er  68  5.436 4.437 1.891 0.0  28.655 5.117 0.577 0.0
Tm  69  5.441 4.51 1.956 0.0  29.149 5.264 0.59 0.0
Yb  70  5.529 4.533 1.945 0.0  28.927 5.144 0.578 0.0
Lu  71  5.553 4.58 1.969 0.0  28.907 5.16 0.577 0.0
Hf  72  5.588 4.619 1.997 0.0  29.001 5.164 0.579 0.0
Ta  73  5.659 4.63 2.014 0.0  28.807 5.114 0.578 0.0
W   74  5.709 4.677 2.019 0.0  28.782 5.084 0.572 0.0
Re  75  5.695 4.74 2.064 0.0  28.968 5.156 0.575 0.0
Os  76  5.75 4.773 2.079 0.0  28.933 5.139 0.573 0.0
Ir  77  5.754 4.851 2.096 0.0  29.159 5.152 0.57 0.0
Pt  78  5.803 4.87 2.127 0.0  29.016 5.15 0.572 0.0
Au  79  2.388 4.226 2.689 1.255  42.866 9.743 2.264 0.307
Hg  80  2.682 4.241 2.755 1.27  42.822 9.856 2.295 0.307
Tl  81  5.932 4.972 2.195 0.0  29.086 5.126 0.572 0.0
Pb  82  3.51 4.552 3.154 1.359  52.914 11.884 2.571 0.321
Bi  83  3.841 4.679 3.192 1.363  50.261 11.999 2.56 0.318
Po  84  6.07 4.997 2.232 0.0  28.075 4.999 0.563 0.0
At  85  6.133 5.031 2.239 0.0  28.047 4.957 0.558 0.0
Rn  86  4.078 4.978 3.096 1.326  38.406 11.02 2.355 0.299
Fr  87  6.201 5.121 2.275 0.0  28.2 4.954 0.556 0.0
Ra  88  6.215 5.17 2.316 0.0  28.382 5.002 0.562 0.0
Ac  89  6.278 5.195 2.321 0.0  28.323 4.949 0.557 0.0
Th  90  6.264 5.263 2.367 0.0  28.651 5.03 0.563 0.0
Pa  91  6.306 5.303 2.386 0.0  28.688 5.026 0.561 0.0
U   92  6.767 6.729 4.014 1.561  85.951 15.642 2.936 0.335
Np  93  6.323 5.414 2.453 0.0  29.142 5.096 0.568 0.0
Pu  94  6.415 5.419 2.449 0.0  28.836 5.022 0.561 0.0
Am  95  6.378 5.495 2.495 0.0  29.156 5.102 0.565 0.0
Cm  96  6.46 5.469 2.471 0.0  28.396 4.97 0.554 0.0
Bk  97  6.502 5.478 2.51 0.0  28.375 4.975 0.561 0.0
Cf  98  6.548 5.526 2.52 0.0  28.461 4.965 0.557 0.0
"""
atomic_params = {}
for _line in _ATOMIC_TABLE.strip().splitlines():
    _f = _line.split()
    atomic_params[_f[0]] = dict(Z=int(_f[1]), a=[float(x) for x in _f[2:6]], b=[float(x) for x in _f[6:10]])

# feature names carried by System.features (reference tools/profiler.py profile_keys); the
# featuriser itself is outside the intensity path
profile_keys = ['Imax_over_Imean', 'Ilowq_over_Imean', 'Imax_sharpness', 'I_fluctuation', 'logI_fluctuation', 'logI_max_over_std', 'r_fftIcentroid', 'q_Icentroid', 'q_logIcentroid', 'pearson_q', 'pearson_q2', 'pearson_expq', 'pearson_invexpq', 'q_best_hump', 'q_best_trough', 'best_hump_qwidth', 'best_trough_qwidth', 'q_best_hump_log', 'q_best_trough_log', 'best_hump_qwidth_log', 'best_trough_qwidth_log']

setting_descriptions = dict(
    lattice='Lattice identifier for crystalline populations',
    space_group='Crystalline space group specification (International symbol)',
    texture='Distribution of orientations for crystalline populations',
    profile='Selection of peak profile for broadening diffraction peaks',
    structure_factor_mode='Strategy for computing off-peak structure factors',
    integration_mode='Strategy for integrating over the reciprocal lattice',
    q_min='minimum q-value for reciprocal space integration',
    q_max='maximum q-value for reciprocal space integration',
    interaction='Interaction potential describing disordered populations',
    symbol='Atomic symbol', n_atoms='Number of atoms',
    distribution='Specifies a distribution for parameter values over a population',
    sampling_width='Number of standard deviations to sample from distribution',
    sampling_step='Resolution of sampling, in units of standard deviations')

parameter_units = dict(
    I0='arbitrary', rg='Angstrom', D='unitless', r='Angstrom', sigma='unitless', r_hard='Angstrom',
    v_fraction='unitless', hwhm='1/Angstrom', hwhm_g='1/Angstrom', hwhm_l='1/Angstrom',
    a='Angstrom', b='Angstrom', c='Angstrom', alpha='degree', beta='degree', gamma='degree')
