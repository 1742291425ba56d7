"""Execute the reference's own Python-2.7 source files under Python 3, in memory.

Used ONLY by tests/golden/make_golden.py, in the build container where
/root/reference is mounted.  Nothing in tests/, smoke() or bench.py imports this
at run time (the GPU box has no /root/reference).

What it does, and nothing more:
  * reads ``/root/reference/<module>.py`` as text (the file is never copied into
    this repository);
  * applies a purely syntactic 2 -> 3 transform: tabs expanded (gpmodel_library.py:307
    mixes a tab into space indentation), ``print x`` statements wrapped into
    ``print(x)``, ``xrange`` -> ``range``, ``d.keys()`` -> ``list(d.keys())``
    (mcts_library.py:504,511 index the key list);
  * provides stand-ins in ``sys.modules`` for third-party imports absent from this
    image: ``GPy`` -> oracle.gpy_restated (restated published algorithm),
    ``sets``/``IPython``/``matplotlib``/``shapely``/``dubins``/``sklearn`` -> inert
    placeholders (plotting / geometry the hot path never reaches in the golden cases);
  * restores the two Python-2 SEMANTICS the arithmetic depends on: ``a / b`` on two integers floors (an AST pass turns
    every ``/`` into ``_py2_div(a, b)``: robot_library.py:328 ``t < T/3``), and the builtin ``round`` rounds halves away
    from zero and returns a float (paths_library.py:92 ``int(round(distance/self.ss))``);
  * ``exec``s the transformed module into a fresh module object.

The arithmetic that produces the golden vectors is therefore the reference's own
numpy code (OnlineGPModel.update_model / predict_value, the aq_library rewards,
Tree.leaf_helper ...), line for line.
"""
import ast
import os
import re
import sys
import types

import numpy as np

REF_ROOT = os.environ.get('IPP_REFERENCE_ROOT', '/root/reference')

_PRINT_RE = re.compile(r'^(\s*)print(\s+(?!\()|\s*$)(.*)$')
_PRINT_PAREN_RE = re.compile(r'^(\s*)print\s*\((.*)\)\s*$')
_KEYS_RE = re.compile(r'([A-Za-z_][A-Za-z0-9_\.]*)\.keys\(\)')


def py2_to_py3(src):
    out = []
    for line in src.expandtabs(8).split('\n'):
        stripped = line.lstrip()
        if not stripped.startswith('#'):
            m = _PRINT_RE.match(line)
            if m and not stripped.startswith('print('):
                body = m.group(3).rstrip()
                # keep any trailing comment outside the call
                line = '%sprint(%s)' % (m.group(1), body)
            line = re.sub(r'\bxrange\b', 'range', line)
            line = _KEYS_RE.sub(r'list(\1.keys())', line)
        out.append(line)
    return '\n'.join(out)


def _is_int(x):
    return (isinstance(x, (int, np.integer)) and not isinstance(x, bool)) or \
        (isinstance(x, np.ndarray) and x.dtype.kind in 'iu')


def _py2_div(a, b):
    """Python 2 `/` without `from __future__ import division` (no reference module imports it): floor division when
    both operands are integers (Python ints, numpy integer scalars or integer arrays), true division otherwise."""
    if _is_int(a) and _is_int(b):
        return a // b
    return a / b


def _py2_round(x, ndigits=0):
    """Python 2 builtin round: a float, halves away from zero (C round())."""
    x = float(x)
    m = 10.0 ** ndigits
    y = abs(x) * m
    f = float(np.floor(y))
    r = f + 1.0 if y - f >= 0.5 else f
    return (-r if x < 0 else r) / m


class _Py2Division(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(ast.Call(func=ast.Name(id='_py2_div', ctx=ast.Load()),
                                              args=[node.left, node.right], keywords=[]), node)
        return node

    def visit_AugAssign(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            load = ast.parse(ast.unparse(node.target), mode='eval').body
            call = ast.Call(func=ast.Name(id='_py2_div', ctx=ast.Load()), args=[load, node.value], keywords=[])
            return ast.copy_location(ast.Assign(targets=[node.target], value=call), node)
        return node


class _Inert(types.ModuleType):
    """A module whose every attribute is a do-nothing callable/placeholder."""

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        val = _InertObj(self.__name__ + '.' + name)
        setattr(self, name, val)
        return val


class _InertObj(object):
    def __init__(self, name='inert'):
        self._name = name

    def __call__(self, *a, **k):
        return _InertObj(self._name + '()')

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return _InertObj(self._name + '.' + name)

    def __iter__(self):
        # "fig, ax = plt.subplots()" in the plotting branches
        return iter((_InertObj(self._name + '[0]'), _InertObj(self._name + '[1]')))


def install_stubs():
    repo = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    from oracle import gpy_restated
    for name, mod in gpy_restated.as_gpy_module().items():
        sys.modules[name] = mod
    sets = types.ModuleType('sets')
    sets.Set = set
    sys.modules['sets'] = sets
    for name in ('IPython', 'IPython.display', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.colors',
                 'matplotlib.cm', 'shapely', 'shapely.geometry', 'dubins', 'sklearn', 'sklearn.mixture',
                 'sklearn.neighbors', 'pdb'):
        if name not in sys.modules or name == 'pdb':
            sys.modules[name] = _Inert(name)
    # parent attribute links so "from matplotlib import pyplot" works
    for name in list(sys.modules):
        if '.' in name and isinstance(sys.modules[name], _Inert):
            parent, child = name.rsplit('.', 1)
            if parent in sys.modules and isinstance(sys.modules[parent], _Inert):
                setattr(sys.modules[parent], child, sys.modules[name])


_loaded = {}


def load(modname):
    """Load reference module ``modname`` (e.g. 'gpmodel_library') and register it in sys.modules
    under its own name so the reference's intra-package imports resolve to the same objects."""
    if modname in _loaded:
        return _loaded[modname]
    install_stubs()
    path = os.path.join(REF_ROOT, modname + '.py')
    with open(path) as f:
        src = f.read()
    tree = ast.fix_missing_locations(_Py2Division().visit(ast.parse(py2_to_py3(src), path)))
    code = compile(tree, path, 'exec')
    mod = types.ModuleType(modname)
    mod.__file__ = path
    mod.__dict__['_py2_div'] = _py2_div
    mod.__dict__['round'] = _py2_round
    sys.modules[modname] = mod
    _loaded[modname] = mod
    # dependencies first (reference modules import each other by bare name)
    for dep in ('obstacles', 'gpmodel_library', 'aq_library', 'paths_library', 'envmodel_library'):
        if dep != modname and re.search(r'^\s*(import|from)\s+%s\b' % dep, src, re.M) and dep not in _loaded:
            load(dep)
    exec(code, mod.__dict__)
    return mod
