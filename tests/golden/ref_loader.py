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
  * ``exec``s the transformed text into a fresh module object.

The arithmetic that produces the golden vectors is therefore the reference's own
numpy code (OnlineGPModel.update_model / predict_value, the aq_library rewards,
Tree.leaf_helper ...), line for line.
"""
import os
import re
import sys
import types

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
    code = compile(py2_to_py3(src), path, 'exec')
    mod = types.ModuleType(modname)
    mod.__file__ = path
    sys.modules[modname] = mod
    _loaded[modname] = mod
    # dependencies first (reference modules import each other by bare name)
    for dep in ('obstacles', 'gpmodel_library', 'aq_library', 'paths_library', 'envmodel_library'):
        if dep != modname and re.search(r'^\s*(import|from)\s+%s\b' % dep, src, re.M) and dep not in _loaded:
            load(dep)
    exec(code, mod.__dict__)
    return mod
