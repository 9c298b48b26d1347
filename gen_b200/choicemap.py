"""A small mirror of Gen's ChoiceMap (src/choice_map.jl:599-770, DynamicChoiceMap) for the host
side: hierarchical addresses `a => b => c` are written as tuples ("a", b, "c")."""


def _norm(addr):
    if isinstance(addr, tuple):
        return addr if len(addr) != 1 else addr[0]
    return addr


class ChoiceMap:
    def __init__(self, *pairs):
        self._d = {}
        for addr, value in pairs:
            self[addr] = value

    def __setitem__(self, addr, value):
        self._d[_norm(addr)] = value

    def __getitem__(self, addr):
        try:
            return self._d[_norm(addr)]
        except KeyError:
            raise KeyError("no value at address %r" % (addr,))

    def __contains__(self, addr):
        return _norm(addr) in self._d

    has_value = __contains__

    def get_value(self, addr):
        return self[addr]

    def set_value(self, addr, value):
        self[addr] = value

    def isempty(self):
        return not self._d

    def items(self):
        return self._d.items()

    def keys(self):
        return self._d.keys()

    def __len__(self):
        return len(self._d)

    def __eq__(self, other):
        return isinstance(other, ChoiceMap) and self._d == other._d

    def __repr__(self):
        return "ChoiceMap(%s)" % ", ".join("%r: %r" % kv for kv in self._d.items())


def choicemap(*pairs):
    """choicemap((addr, value), ...) -- src/choice_map.jl:752-761."""
    return ChoiceMap(*pairs)


def merge(a, b):
    """merge(a, b) (src/choice_map.jl:237-269): error if an address has a value in both."""
    out = ChoiceMap(*a.items())
    for k, v in b.items():
        if k in out:
            raise ValueError("choicemaps both have a value at address %r" % (k,))
        out[k] = v
    return out


class UnknownChange:
    """src/diff.jl:32-77 argdiff marker (accepted and ignored: a step always extends by one)."""


class NoChange:
    pass
