"""Stand-in for the third-party package ``libconf`` (PyPI, pure Python; the reference pins
``libconf>=2.0.0``, deploy/requirements.txt:4) — TEST INFRASTRUCTURE (oracle side).

The package is not installed here and there is no network, so the UNMODIFIED
``nhd/TriadCfgParser.py`` could not even be imported.  This module restates the published
behaviour of libconf 2.0.x that the reference relies on — ``loads`` / ``load`` / ``dumps`` /
``dump``, ``AttrDict`` (attribute *read* access only: ``setattr`` does not reach the mapping, which
is exactly what ``TriadCfgParser.SetLibConfigValue`` works around, TriadCfgParser.py:382-395),
``LibconfInt64`` — so that the reference class runs and can serve as the oracle for
``nhd_b200.TriadCfgParser``.  PARITY UNPINNED for the text format itself: this file follows the
libconfig grammar and libconf's documented dump layout from memory of the package, it was not
checked against the package.  It is written independently of ``nhd_b200/libconfig.py`` (regex
tokens + recursive descent here, a character scanner there) so the two cross-check each other.
"""
import collections
import io
import re


class AttrDict(collections.OrderedDict):
    def __getattr__(self, attr):
        try:
            return self.__getitem__(attr)
        except KeyError:
            raise AttributeError("Attribute %r not found" % attr)


class LibconfInt64(int):
    pass


class ConfigParseError(RuntimeError):
    pass


class ConfigSerializeError(TypeError):
    pass


_SKIP = re.compile(r'\s+|#[^\n]*|//[^\n]*|/\*.*?\*/', re.S)
_TOKENS = [(n, re.compile(p)) for n, p in (
    ('float', r'([-+]?(\d+)?\.\d*([eE][-+]?\d+)?)|([-+]?(\d+)(\.\d*)?[eE][-+]?\d+)'),
    ('hex64', r'0[Xx][0-9A-Fa-f]+(L(L)?)'),
    ('hex', r'0[Xx][0-9A-Fa-f]+'),
    ('integer64', r'[-+]?[0-9]+L(L)?'),
    ('integer', r'[-+]?[0-9]+'),
    ('boolean', r'(?i)(true|false)\b'),
    ('string', r'"([^"\\]|\\.)*"'),
    ('name', r'[A-Za-z\*][-A-Za-z0-9_\*]*'),
    ('}', r'\}'), ('{', r'\{'), (')', r'\)'), ('(', r'\('), (']', r'\]'), ('[', r'\['),
    (',', r','), (';', r';'), ('=', r'='), (':', r':'))]

_ESC = {'\\': '\\', '"': '"', 'n': '\n', 'r': '\r', 't': '\t', 'f': '\f', 'a': '\a', 'b': '\b', 'v': '\v'}


def _unescape(s):
    out, i = [], 0
    while i < len(s):
        c = s[i]
        if c == '\\' and i + 1 < len(s):
            n = s[i + 1]
            if n == 'x' and i + 3 < len(s) + 0 and re.match(r'[0-9A-Fa-f]{2}', s[i + 2:i + 4]):
                out.append(chr(int(s[i + 2:i + 4], 16)))
                i += 4
                continue
            out.append(_ESC.get(n, '\\' + n))
            i += 2
        else:
            out.append(c)
            i += 1
    return ''.join(out)


def _tokenize(text):
    pos, toks = 0, []
    while pos < len(text):
        m = _SKIP.match(text, pos)
        if m:
            pos = m.end()
            continue
        for name, rx in _TOKENS:
            m = rx.match(text, pos)
            if m and m.end() > pos:
                toks.append((name, m.group(0), pos))
                pos = m.end()
                break
        else:
            raise ConfigParseError('Couldn\'t load config, at offset %d: %r' % (pos, text[pos:pos + 20]))
    toks.append(('eof', '', pos))
    return toks


class _Parser:
    def __init__(self, toks):
        self.t, self.i = toks, 0

    def peek(self):
        return self.t[self.i][0]

    def take(self, *kinds):
        k, txt, pos = self.t[self.i]
        if kinds and k not in kinds:
            raise ConfigParseError('Unexpected %s %r at offset %d, expected %s' % (k, txt, pos, '/'.join(kinds)))
        self.i += 1
        return k, txt

    def settings(self, closer):
        d = AttrDict()
        while self.peek() != closer:
            _, name = self.take('name')
            self.take('=', ':')
            d[name] = self.value()
            if self.peek() in (';', ','):
                self.take()
        return d

    def scalar(self):
        k, txt = self.take('float', 'hex64', 'hex', 'integer64', 'integer', 'boolean', 'string')
        if k == 'float':
            return float(txt)
        if k in ('hex64', 'hex', 'integer64', 'integer'):
            v = int(txt.rstrip('L'), 0)
            return LibconfInt64(v) if txt.endswith('L') else v
        if k == 'boolean':
            return txt[0].lower() == 't'
        s = _unescape(txt[1:-1])
        while self.peek() == 'string':                      # adjacent literals concatenate
            s += _unescape(self.take()[1][1:-1])
        return s

    def value(self):
        k = self.peek()
        if k == '{':
            self.take()
            d = self.settings('}')
            self.take('}')
            return d
        if k == '(':
            self.take()
            items = []
            while self.peek() != ')':
                items.append(self.value())
                if self.peek() == ',':
                    self.take()
            self.take(')')
            return tuple(items)
        if k == '[':
            self.take()
            items = []
            while self.peek() != ']':
                items.append(self.scalar())
                if self.peek() == ',':
                    self.take()
            self.take(']')
            return items
        return self.scalar()


def loads(string, filename=None, includedir=''):
    p = _Parser(_tokenize(string))
    cfg = p.settings('eof')
    return cfg


def load(f, filename=None, includedir=''):
    return loads(f.read())


_I32 = (-2 ** 31, 2 ** 31 - 1)


def _dump_scalar(v):
    if isinstance(v, bool):
        return 'true' if v else 'false'
    if isinstance(v, LibconfInt64):
        return str(int(v)) + 'L'
    if isinstance(v, int):
        return str(v) + ('' if _I32[0] <= v <= _I32[1] else 'L')
    if isinstance(v, float):
        s = str(v)
        return s if ('.' in s or 'e' in s or 'E' in s) else s + '.0'
    if isinstance(v, str):
        s = (v.replace('\\', '\\\\').replace('"', '\\"').replace('\f', r'\f').replace('\n', r'\n')
             .replace('\r', r'\r').replace('\t', r'\t'))
        s = re.sub(r'[\x00-\x1f\x7f]', lambda m: r'\x{:02x}'.format(ord(m.group(0))), s)
        return '"' + s + '"'
    raise ConfigSerializeError('Can not serialize object %r of type %s' % (v, type(v)))


def _dump_value(key, value, f, indent):
    spaces = ' ' * indent
    prefix = '' if key is None else key + ' = '
    prefix_nl = '' if key is None else key + ' =\n' + spaces
    if isinstance(value, dict):
        f.write(u'{}{}{{\n'.format(spaces, prefix_nl))
        _dump_dict(value, f, indent + 4)
        f.write(u'{}}}'.format(spaces))
    elif isinstance(value, tuple):
        f.write(u'{}{}(\n'.format(spaces, prefix_nl))
        _dump_collection(value, f, indent + 4)
        f.write(u'\n{})'.format(spaces))
    elif isinstance(value, list):
        kinds = {('n' if isinstance(x, (int, float)) and not isinstance(x, bool) else type(x).__name__) for x in value}
        if any(isinstance(x, (dict, list, tuple)) for x in value) or len(kinds) > 1:
            raise ConfigSerializeError('libconfig arrays hold scalars of one type: %r' % (value,))
        f.write(u'{}{}[\n'.format(spaces, prefix_nl))
        _dump_collection(value, f, indent + 4)
        f.write(u'\n{}]'.format(spaces))
    else:
        f.write(u'{}{}{}'.format(spaces, prefix, _dump_scalar(value)))


def _dump_collection(cfg, f, indent):
    for i, value in enumerate(cfg):
        _dump_value(None, value, f, indent)
        if i < len(cfg) - 1:
            f.write(u',\n')


def _dump_dict(cfg, f, indent):
    for key in cfg:
        if not isinstance(key, str):
            raise ConfigSerializeError('Dict keys must be strings: %r' % (key,))
        _dump_value(key, cfg[key], f, indent)
        f.write(u';\n')


def dumps(cfg):
    f = io.StringIO()
    dump(cfg, f)
    return f.getvalue()


def dump(cfg, f):
    if not isinstance(cfg, dict):
        raise ConfigSerializeError('dump() requires a dict as input, not %r of type %r' % (cfg, type(cfg)))
    _dump_dict(cfg, f, 0)
