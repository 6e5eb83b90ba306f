"""Change counters for ``Node`` objects, so that ``Matcher.FindNode`` packs only what changed.

The reference's ``FindNode(nl, top)`` is handed the whole (filtered) node dict for every pod
(``NHDScheduler.py:277-278``) and its ``Node`` objects carry no change counter: taken as they are, each call
would have to pack every node again (about 17 us per node in Python).  ``track_changes(cls)`` instruments a
``Node`` class in place — this package's mirror or the reference's own ``nhd.Node.Node`` — so that every
instance counts its mutations in ``_nhd_version``:

* every attribute assignment on the node (``active``, ``maintenance``, ``busy_time``, ``groups`` ... —
  ``NHDScheduler.py:88-98,541-564``, ``Node.py:845``), through ``__setattr__``;
* every call of a method that is not a plain getter (``SetPhysicalIdsFromMapping``,
  ``RemoveResourcesFromTopology``, ``AddResourcesFromTopology``, ``ClaimPodNICResources``, ``ResetResources``,
  ``SetHugepages``, ``ParseLabels`` ...): these are the only places the reference changes the cores, GPUs, NICs and
  hugepages of a node (``Node.py:144-161,489-493,530-646,663-841``); the counter moves whether or not the call
  raises.

Code that changes a node's sub-objects from outside the class has to call ``bump(node)`` itself
(``packing.apply_binding`` does).  Counting too often only costs a re-pack; an uninstrumented class is simply packed
on every call, as before.
"""
import functools
import inspect

_GETTER_PREFIXES = ('Get', 'Is', 'Print')
_GETTER_NAMES = frozenset(('SMTEnabled', 'PodPresent', 'FormatMac', 'ParseRangeList'))
_UNPACKED = frozenset(('last_busy_time_seconds',))    # written by the getter IsBusy (Node.py:848); no record field comes from it
_FLAG = '_nhd_tracked'
VERSION = '_nhd_version'


def bump(node):
    """One more change of ``node`` (no-op for objects without a ``__dict__``)."""
    d = getattr(node, '__dict__', None)
    if d is not None:
        d[VERSION] = d.get(VERSION, 0) + 1


def is_tracked(cls) -> bool:
    return bool(cls.__dict__.get(_FLAG, False))


def _counting(fn):
    @functools.wraps(fn)
    def call(self, *args, **kw):
        try:
            return fn(self, *args, **kw)
        finally:
            d = self.__dict__
            d[VERSION] = d.get(VERSION, 0) + 1
    return call


def track_changes(cls) -> bool:
    """Instrument ``cls`` (idempotent).  False when the class cannot carry the counter (``__slots__`` without
    ``__dict__``, a type that refuses attribute assignment): its nodes are then packed on every call."""
    if is_tracked(cls):
        return True
    if '__slots__' in cls.__dict__ and '__dict__' not in getattr(cls, '__slots__', ()):
        return False
    try:
        plain_setattr = cls.__setattr__

        def __setattr__(self, name, value, _set=plain_setattr):
            _set(self, name, value)
            if name not in _UNPACKED:
                d = self.__dict__
                d[VERSION] = d.get(VERSION, 0) + 1

        wrapped = {}
        for name, fn in inspect.getmembers(cls, inspect.isfunction):
            if name.startswith('_') or name.startswith(_GETTER_PREFIXES) or name in _GETTER_NAMES:
                continue
            if isinstance(inspect.getattr_static(cls, name), (staticmethod, classmethod)):
                continue
            wrapped[name] = _counting(fn)
        cls.__setattr__ = __setattr__
        for name, fn in wrapped.items():
            setattr(cls, name, fn)
        setattr(cls, _FLAG, True)
        return True
    except (TypeError, AttributeError):
        return False
