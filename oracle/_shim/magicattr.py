"""Stand-in for the third-party package ``magicattr`` (PyPI; the reference pins ``magicattr>=0.1.4``,
deploy/requirements.txt:5) — TEST INFRASTRUCTURE (oracle side), see libconf.py next to this file.

Published behaviour restated: ``get(obj, 'a.b[0].c')`` / ``set`` / ``delete`` walk an attribute
path written as a Python expression made only of names, attribute access and constant subscripts
(parsed with ``ast``); attributes go through ``getattr`` / ``setattr``, subscripts through
``obj[i]`` / ``obj[i] = v``.  PARITY UNPINNED (package absent).
"""
import ast
import functools


def _steps(attr):
    node = ast.parse(attr, mode='eval').body
    out = []
    while True:
        if isinstance(node, ast.Attribute):
            out.append(('a', node.attr))
            node = node.value
        elif isinstance(node, ast.Subscript):
            idx = node.slice
            if isinstance(idx, ast.UnaryOp) and isinstance(idx.op, ast.USub) and isinstance(idx.operand, ast.Constant):
                key = -idx.operand.value
            elif isinstance(idx, ast.Constant):
                key = idx.value
            else:
                raise ValueError('Only constant subscripts are supported: %r' % attr)
            out.append(('s', key))
            node = node.value
        elif isinstance(node, ast.Name):
            out.append(('a', node.id))
            break
        else:
            raise ValueError('Unsupported expression in attribute path: %r' % attr)
    return out[::-1]


def _walk(obj, step):
    kind, key = step
    return getattr(obj, key) if kind == 'a' else obj[key]


def get(obj, attr, **kwargs):
    try:
        return functools.reduce(_walk, _steps(attr), obj)
    except (AttributeError, KeyError, IndexError):
        if 'default' in kwargs:
            return kwargs['default']
        raise


def set(obj, attr, val):
    steps = _steps(attr)
    parent = functools.reduce(_walk, steps[:-1], obj)
    kind, key = steps[-1]
    if kind == 'a':
        setattr(parent, key, val)
    else:
        parent[key] = val


def delete(obj, attr):
    steps = _steps(attr)
    parent = functools.reduce(_walk, steps[:-1], obj)
    kind, key = steps[-1]
    if kind == 'a':
        delattr(parent, key)
    else:
        del parent[key]
