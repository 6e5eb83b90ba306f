"""libconfig text <-> Python values, and attribute-path access into them.

The reference's request codec (``nhd/TriadCfgParser.py``) leans on two third-party packages:
``libconf`` (``libconf>=2.0.0``, deploy/requirements.txt:4) to read and write the pod's libconfig
text, and ``magicattr`` (``>=0.1.4``, :5) to address settings by strings such as
``"RxMod[0].dp[0].rx_cores[1]"``.  Neither is available here, so this module provides the part of
both that the codec needs, from the libconfig grammar (settings ``name = value;`` or ``name : value``,
groups ``{}``, arrays ``[]`` of scalars, lists ``()`` of anything, 32/64-bit integers incl. hex and the
``L`` suffix, floats, booleans, strings with escapes and adjacent-literal concatenation, ``#`` ``//``
``/* */`` comments) and with the same Python shapes libconf produces, because the reference's logic
depends on them:

* group -> ``Group`` (ordered ``dict`` whose keys can also be *read* as attributes; assigning an
  attribute does NOT create a setting — ``TriadCfgParser.SetLibConfigValue`` exists because of that,
  ``TriadCfgParser.py:382-395``);
* array ``[...]`` -> ``list`` (mutable: physical core ids are written into it element by element);
* list ``(...)`` -> ``tuple`` (immutable: the GPU map is replaced whole, ``:441-459``).

``dumps`` lays the text out the way libconf does (4-space indent, ``name =`` + newline before a
composite value, ``L`` on integers outside 32 bits).  PARITY with the package's exact output is
UNPINNED (the package is absent); what is pinned is that the unmodified reference class, run on
top of an independently written stand-in for both packages, and ``nhd_b200.TriadCfgParser`` on top
of this module agree on every topology and every rewritten config (``tests/test_codec.py``).
"""
import re
from typing import Any, List, Tuple

I32_MIN, I32_MAX = -2 ** 31, 2 ** 31 - 1


class ConfigParseError(RuntimeError):
    pass


class ConfigSerializeError(TypeError):
    pass


class Group(dict):
    """A libconfig group.  ``g.name`` reads ``g['name']``; writing goes through ``g['name'] = v``."""

    def __getattr__(self, attr):
        try:
            return self[attr]
        except KeyError:
            raise AttributeError('Attribute %r not found' % attr)


class Int64(int):
    """An integer that was written with the ``L`` suffix and keeps it when dumped."""


# ----------------------------------------------------------------------------------------------
# reading
# ----------------------------------------------------------------------------------------------
_NAME_START = set('ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz*')
_NAME_BODY = _NAME_START | set('0123456789-_')
_NUMBER = re.compile(r'[-+]?(0[xX][0-9A-Fa-f]+|\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+|\d+)(LL?)?')
_SIMPLE_ESCAPES = {'\\': '\\', '"': '"', 'n': '\n', 'r': '\r', 't': '\t', 'f': '\f', 'a': '\a', 'b': '\b', 'v': '\v'}


class _Reader:
    def __init__(self, text: str):
        self.s, self.p, self.n = text, 0, len(text)

    def fail(self, what):
        line = self.s.count('\n', 0, self.p) + 1
        raise ConfigParseError(f'{what} at line {line}: {self.s[self.p:self.p + 24]!r}')

    def blank(self):
        s, n = self.s, self.n
        while self.p < n:
            c = s[self.p]
            if c in ' \t\r\n\f\v':
                self.p += 1
            elif c == '#' or s.startswith('//', self.p):
                e = s.find('\n', self.p)
                self.p = n if e < 0 else e + 1
            elif s.startswith('/*', self.p):
                e = s.find('*/', self.p + 2)
                if e < 0:
                    self.fail('unterminated comment')
                self.p = e + 2
            else:
                return

    def peek(self) -> str:
        self.blank()
        return self.s[self.p] if self.p < self.n else ''

    def eat(self, chars: str) -> str:
        c = self.peek()
        if c == '' or c not in chars:
            self.fail(f'expected one of {chars!r}')
        self.p += 1
        return c

    def name(self) -> str:
        if self.peek() not in _NAME_START:
            self.fail('expected a setting name')
        b = self.p
        while self.p < self.n and self.s[self.p] in _NAME_BODY:
            self.p += 1
        return self.s[b:self.p]

    def settings(self, closer: str) -> Group:
        g = Group()
        while self.peek() != closer:
            if self.peek() == '':
                self.fail('unexpected end of text')
            key = self.name()
            self.eat('=:')
            g[key] = self.value()
            if self.peek() in (';', ','):
                self.p += 1
        return g

    def string(self) -> str:
        out: List[str] = []
        while self.peek() == '"':                 # "a" "b" is one string
            self.p += 1
            s = self.s
            while True:
                if self.p >= self.n:
                    self.fail('unterminated string')
                c = s[self.p]
                if c == '"':
                    self.p += 1
                    break
                if c == '\\' and self.p + 1 < self.n:
                    e = s[self.p + 1]
                    if e == 'x' and re.fullmatch(r'[0-9A-Fa-f]{2}', s[self.p + 2:self.p + 4] or ''):
                        out.append(chr(int(s[self.p + 2:self.p + 4], 16)))
                        self.p += 4
                    else:
                        out.append(_SIMPLE_ESCAPES.get(e, '\\' + e))
                        self.p += 2
                else:
                    out.append(c)
                    self.p += 1
        return ''.join(out)

    def scalar(self):
        c = self.peek()
        if c == '"':
            return self.string()
        m = _NUMBER.match(self.s, self.p)
        if m:
            self.p = m.end()
            body, long_ = m.group(1), m.group(2)
            if body[:2] in ('0x', '0X'):
                v = int(body, 16)
            elif any(ch in body for ch in '.eE'):
                if long_:
                    self.fail('L suffix on a float')
                return float(m.group(0))
            else:
                v = int(body)
            if m.group(0)[0] == '-':
                v = -v
            return Int64(v) if long_ else v
        w = self.s[self.p:self.p + 5].lower()
        for word, val in (('true', True), ('false', False)):
            if w.startswith(word) and (self.p + len(word) >= self.n or self.s[self.p + len(word)] not in _NAME_BODY):
                self.p += len(word)
                return val
        self.fail('expected a value')

    def value(self):
        c = self.peek()
        if c == '{':
            self.p += 1
            g = self.settings('}')
            self.eat('}')
            return g
        if c == '(':
            self.p += 1
            items = []
            while self.peek() != ')':
                items.append(self.value())
                if self.peek() == ',':
                    self.p += 1
            self.eat(')')
            return tuple(items)
        if c == '[':
            self.p += 1
            items = []
            while self.peek() != ']':
                items.append(self.scalar())
                if self.peek() == ',':
                    self.p += 1
            self.eat(']')
            return items
        return self.scalar()


def loads(text: str) -> Group:
    r = _Reader(text)
    cfg = r.settings('')
    return cfg


def load(f) -> Group:
    return loads(f.read())


# ----------------------------------------------------------------------------------------------
# writing
# ----------------------------------------------------------------------------------------------
def _scalar_text(v) -> str:
    if isinstance(v, bool):
        return 'true' if v else 'false'
    if isinstance(v, int):
        return str(int(v)) + ('L' if isinstance(v, Int64) or not I32_MIN <= v <= I32_MAX else '')
    if isinstance(v, float):
        t = str(v)
        return t if any(ch in t for ch in '.eE') else t + '.0'
    if isinstance(v, str):
        out = []
        for ch in v:
            if ch == '\\':
                out.append('\\\\')
            elif ch == '"':
                out.append('\\"')
            elif ch in '\f\n\r\t':
                out.append({'\f': '\\f', '\n': '\\n', '\r': '\\r', '\t': '\\t'}[ch])
            elif ord(ch) < 0x20 or ord(ch) == 0x7f:
                out.append('\\x%02x' % ord(ch))
            else:
                out.append(ch)
        return '"' + ''.join(out) + '"'
    raise ConfigSerializeError(f'cannot write {v!r} ({type(v).__name__}) as a libconfig scalar')


def _array_ok(items) -> bool:
    kinds = set()
    for x in items:
        if isinstance(x, (dict, list, tuple)):
            return False
        kinds.add('bool' if isinstance(x, bool) else 'num' if isinstance(x, (int, float)) else type(x).__name__)
    return len(kinds) <= 1


def _emit(out: List[str], key, v, depth: int):
    pad = '    ' * depth
    if isinstance(v, (dict, tuple, list)):
        if isinstance(v, list) and not _array_ok(v):
            raise ConfigSerializeError(f'a libconfig array holds scalars of one type: {v!r}')
        opener, closer = ('{', '}') if isinstance(v, dict) else ('(', ')') if isinstance(v, tuple) else ('[', ']')
        out.append(pad + (opener if key is None else f'{key} =\n{pad}{opener}') + '\n')
        if isinstance(v, dict):
            _emit_settings(out, v, depth + 1)
        else:
            for i, item in enumerate(v):
                _emit(out, None, item, depth + 1)
                if i < len(v) - 1:
                    out.append(',\n')
            out.append('\n')                         # also after nothing: an empty collection spans three lines
        out.append(pad + closer)
    else:
        out.append(pad + ('' if key is None else key + ' = ') + _scalar_text(v))


def _emit_settings(out: List[str], g, depth: int):
    for key, v in g.items():
        if not isinstance(key, str):
            raise ConfigSerializeError(f'setting names are strings: {key!r}')
        _emit(out, key, v, depth)
        out.append(';\n')


def dumps(cfg) -> str:
    if not isinstance(cfg, dict):
        raise ConfigSerializeError('the top level of a configuration is a group (dict)')
    out: List[str] = []
    _emit_settings(out, cfg, 0)
    return ''.join(out)


# ----------------------------------------------------------------------------------------------
# attribute paths ("Mod[0].dp[0].rx_cores[1]")
# ----------------------------------------------------------------------------------------------
_STEP = re.compile(r'\s*(?:\.?\s*([A-Za-z_\*][A-Za-z0-9_\*]*)|\[\s*(-?\d+)\s*\])')


def path_steps(path: str) -> List[Tuple[str, Any]]:
    steps, p = [], 0
    while p < len(path):
        m = _STEP.match(path, p)
        if not m or (p == 0 and path[0] == '.'):
            raise ValueError(f'not an attribute path: {path!r}')
        steps.append(('attr', m.group(1)) if m.group(1) is not None else ('item', int(m.group(2))))
        p = m.end()
    if not steps or steps[0][0] != 'attr':
        raise ValueError(f'not an attribute path: {path!r}')
    return steps


def _step(obj, step):
    kind, key = step
    return getattr(obj, key) if kind == 'attr' else obj[key]


def path_get(obj, path: str):
    """``magicattr.get``: raises AttributeError / IndexError / KeyError / TypeError like the walk itself."""
    for st in path_steps(path):
        obj = _step(obj, st)
    return obj


def path_set(obj, path: str, value):
    """``magicattr.set``: the last step is ``setattr`` for ``.name`` and item assignment for ``[i]``."""
    steps = path_steps(path)
    for st in steps[:-1]:
        obj = _step(obj, st)
    kind, key = steps[-1]
    if kind == 'attr':
        setattr(obj, key, value)          # on a Group this does NOT create a setting (see module docstring)
    else:
        obj[key] = value
