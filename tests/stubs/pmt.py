"""TEST STAND-IN for GNU Radio's pmt module: the handful of constructors/accessors gr_adapter.py uses."""


class _Sym(str):
    pass


def string_to_symbol(s): return _Sym(s)
intern = string_to_symbol
def symbol_to_string(p): return str(p)
def is_symbol(p): return isinstance(p, _Sym)
def from_uint64(v): return ("u64", int(v))
def from_double(v): return ("f64", float(v))
def to_uint64(p): return int(p[1])
def to_double(p): return float(p[1])
def make_tuple(*items): return tuple(items)
def tuple_ref(t, k): return t[k]
