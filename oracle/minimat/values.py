"""minimat.values -- run-time values and array semantics of the MATLAB subset (TEST INFRASTRUCTURE ONLY).

Numeric values are Python ``float`` / ``complex`` / ``bool`` scalars or numpy arrays with ndim >= 2 whose shape is the MATLAB
size (linear indexing, reshape and (:) go through order='F').  Character rows are Python ``str``.  Value semantics: nothing
is ever mutated in place -- indexed assignment copies, then rebinds the variable.
"""
import math

import numpy as np


class MatlabError(Exception):
    pass


class MStruct:
    __slots__ = ("f",)

    def __init__(self, fields=None):
        self.f = dict(fields) if fields else {}

    def copy(self):
        return MStruct(self.f)

    def __repr__(self):
        return "struct(" + ", ".join(f"{k}={_short(v)}" for k, v in self.f.items()) + ")"


class MStructArr:
    """m x n struct array: a numpy object array of MStruct (a list makes a 1 x n row, what dir() returns)"""
    __slots__ = ("a",)

    def __init__(self, items):
        if isinstance(items, np.ndarray):
            self.a = items
        else:
            items = list(items)
            self.a = np.empty((1, len(items)), dtype=object)
            for i, v in enumerate(items):
                self.a[0, i] = v

    @property
    def items(self):
        return list(self.a.reshape(-1, order="F"))


class MCell:
    __slots__ = ("a",)

    def __init__(self, a):
        self.a = a                 # numpy object array, ndim 2

    @staticmethod
    def row(items):
        a = np.empty((1, len(items)), dtype=object)
        for i, v in enumerate(items):
            a[0, i] = v
        return MCell(a)


class MObject:
    __slots__ = ("cls", "f")

    def __init__(self, cls, fields):
        self.cls, self.f = cls, dict(fields)

    def copy(self):
        return MObject(self.cls, self.f)


class FuncHandle:
    __slots__ = ("kind", "name", "func", "frame", "params", "body", "captured")

    def __init__(self, kind, name=None, func=None, frame=None, params=None, body=None, captured=None):
        self.kind, self.name, self.func, self.frame = kind, name, func, frame
        self.params, self.body, self.captured = params, body, captured


def _short(v):
    if isinstance(v, np.ndarray):
        return f"<{'x'.join(map(str, v.shape))} {v.dtype}>"
    return repr(v)


EMPTY = np.zeros((0, 0))


def is_numeric(v):
    return isinstance(v, (float, complex, bool, int, np.ndarray))


def to_arr(v):
    """any numeric / char value as an ndarray with ndim >= 2"""
    t = type(v)
    if t is np.ndarray:
        return v
    if t is float or t is complex or t is bool:
        return np.array([[v]])
    if t is int:
        return np.array([[float(v)]])
    if t is str:
        return np.array([[float(ord(c)) for c in v]]).reshape(1, len(v))
    if isinstance(v, (np.floating, np.complexfloating, np.bool_, np.integer)):
        return np.array([[v.item()]])
    raise MatlabError(f"value of class {mclass(v)} is not numeric")


def simplify(a):
    """ndarray -> canonical value: 1x1 becomes a Python scalar, trailing singleton dimensions beyond the second are dropped"""
    if type(a) is not np.ndarray:
        if isinstance(a, (np.floating, np.integer)):
            return float(a)
        if isinstance(a, np.complexfloating):
            return complex(a)
        if isinstance(a, np.bool_):
            return bool(a)
        if type(a) is int:
            return float(a)
        return a
    if a.size == 1:
        x = a.reshape(-1)[0]
        k = a.dtype.kind
        if k == "f":
            return float(x)
        if k == "c":
            return complex(x)
        if k == "b":
            return bool(x)
        if k == "O":
            return a if a.ndim == 2 else a.reshape(1, 1)
        return float(x)
    if a.ndim < 2:
        a = a.reshape(1, -1) if a.ndim == 1 else a.reshape(1, 1)
    while a.ndim > 2 and a.shape[-1] == 1:
        a = a.reshape(a.shape[:-1])
    k = a.dtype.kind
    if k in "iu":
        a = a.astype(np.float64)
    return a


def drop_zero_imag(v):
    """MATLAB demotes a complex result whose imaginary part is all zero to real"""
    if type(v) is complex:
        return v.real if v.imag == 0.0 else v
    if type(v) is np.ndarray and v.dtype.kind == "c" and not v.imag.any():
        return v.real.copy()
    return v


def msize(v):
    t = type(v)
    if t is np.ndarray:
        return v.shape
    if t is str:
        return (1, len(v)) if v else (0, 0)
    if t is MCell:
        return v.a.shape
    if t is MStructArr:
        return v.a.shape
    return (1, 1)


def numel(v):
    n = 1
    for d in msize(v):
        n *= d
    return n


def mclass(v):
    t = type(v)
    if t is float or t is int:
        return "double"
    if t is complex:
        return "double"
    if t is bool:
        return "logical"
    if t is str:
        return "char"
    if t is np.ndarray:
        return "logical" if v.dtype.kind == "b" else "double"
    if t is MStruct or t is MStructArr:
        return "struct"
    if t is MCell:
        return "cell"
    if t is MObject:
        return v.cls.name
    if t is FuncHandle:
        return "function_handle"
    return t.__name__


def truth(v):
    """if / while condition: non-empty and all elements non-zero"""
    t = type(v)
    if t is bool:
        return v
    if t is float:
        return v != 0.0
    if t is np.ndarray:
        return v.size > 0 and bool(np.all(v != 0))
    if t is complex:
        return v != 0
    if t is str:
        return len(v) > 0 and all(ord(c) != 0 for c in v)
    raise MatlabError(f"a value of class {mclass(v)} cannot be a condition")


def align(a, b):
    """pad the array with fewer dimensions with trailing singletons (MATLAB aligns leading dimensions, numpy trailing ones)"""
    if a.ndim < b.ndim:
        a = a.reshape(a.shape + (1,) * (b.ndim - a.ndim))
    elif b.ndim < a.ndim:
        b = b.reshape(b.shape + (1,) * (a.ndim - b.ndim))
    for da, db in zip(a.shape, b.shape):
        if da != db and da != 1 and db != 1:
            raise MatlabError(f"arrays have incompatible sizes for this operation ({'x'.join(map(str, a.shape))} and "
                              f"{'x'.join(map(str, b.shape))})")
    return a, b


def _num(a):
    """operand of arithmetic as float/complex ndarray (logical and char promote to double)"""
    a = to_arr(a)
    if a.dtype.kind == "b":
        return a.astype(np.float64)
    return a


_CMP = {"==": np.equal, "~=": np.not_equal, "<": np.less, "<=": np.less_equal, ">": np.greater, ">=": np.greater_equal}


def binop(op, a, b):
    ta, tb = type(a), type(b)
    if ta is float and tb is float:                       # the scalar fast path the per-packet loops live on
        if op == "+":
            return a + b
        if op == "-":
            return a - b
        if op == "*" or op == ".*":
            return a * b
        if op == "/" or op == "./":
            if b != 0.0:
                return a / b
        elif op == "<":
            return a < b
        elif op == ">":
            return a > b
        elif op == "==":
            return a == b
        elif op == "~=":
            return a != b
        elif op == "<=":
            return a <= b
        elif op == ">=":
            return a >= b
        elif op == "^" or op == ".^":
            if a >= 0.0 or b == math.floor(b):
                try:
                    return a ** b
                except (ZeroDivisionError, OverflowError):
                    pass
    if op == "+" and ta is str and tb is str and len(a) != len(b):
        # "text" + 'more' with a double-quoted (string) operand concatenates in MATLAB; this interpreter keeps both quote styles
        # as character rows, so only the case that would otherwise be a size error (different lengths) is read that way
        return a + b
    if op in _CMP:
        x, y = align(_num(a), _num(b))
        if op not in ("==", "~="):
            x, y = x.real, y.real                              # MATLAB orders by the real part
        return simplify(_CMP[op](x, y))
    if op == "&" or op == "|":
        x, y = align(to_arr(a), to_arr(b))
        x, y = x != 0, y != 0
        return simplify(np.logical_and(x, y) if op == "&" else np.logical_or(x, y))
    x, y = _num(a), _num(b)
    with np.errstate(all="ignore"):
        if op == "+":
            x, y = align(x, y)
            r = x + y
        elif op == "-":
            x, y = align(x, y)
            r = x - y
        elif op == ".*":
            x, y = align(x, y)
            r = x * y
        elif op == "./":
            x, y = align(x, y)
            r = x / y
        elif op == ".\\":
            x, y = align(x, y)
            r = y / x
        elif op == ".^":
            x, y = align(x, y)
            r = _power(x, y)
        elif op == "*":
            if x.size == 1 or y.size == 1:
                x, y = align(x, y)
                r = x * y
            else:
                if x.ndim != 2 or y.ndim != 2 or x.shape[1] != y.shape[0]:
                    raise MatlabError("inner matrix dimensions must agree")
                r = x @ y
        elif op == "/":
            if y.size == 1:
                x, y = align(x, y)
                r = x / y
            else:
                r = np.linalg.lstsq(y.T, x.T, rcond=None)[0].T if y.shape[0] != y.shape[1] else np.linalg.solve(y.T, x.T).T
        elif op == "\\":
            if x.size == 1:
                x, y = align(x, y)
                r = y / x
            else:
                r = np.linalg.solve(x, y) if x.shape[0] == x.shape[1] else np.linalg.lstsq(x, y, rcond=None)[0]
        elif op == "^":
            if x.size == 1 and y.size == 1:
                r = _power(x, y)
            elif y.size == 1 and x.ndim == 2 and x.shape[0] == x.shape[1] and float(y.real.flat[0]).is_integer():
                r = np.linalg.matrix_power(x, int(y.real.flat[0]))
            else:
                raise MatlabError("matrix power needs a square matrix and an integer exponent here")
        else:
            raise MatlabError(f"operator {op} is not implemented")
    return drop_zero_imag(simplify(r))


def _power(x, y):
    if x.dtype.kind != "c" and y.dtype.kind != "c":
        neg = (x < 0) & (y != np.floor(y))
        if neg.any():
            return x.astype(np.complex128) ** y
        if y.size == 1 and y.flat[0] == 2.0:
            return x * x
    return x ** y


def unop(op, a):
    if op == "-":
        if type(a) is float:
            return -a
        return simplify(-_num(a))
    if op == "+":
        return simplify(_num(a))
    if op == "~":
        if type(a) is bool:
            return not a
        return simplify(to_arr(a) == 0)
    raise MatlabError(f"unary {op}")


def transpose(op, a):
    t = type(a)
    if t is float or t is bool:
        return a
    if t is complex:
        return a.conjugate() if op == "'" else a
    if t is MCell:
        return MCell(a.a.T.copy())
    x = to_arr(a)
    if x.ndim != 2:
        raise MatlabError("transpose of an N-D array is not defined")
    r = x.T
    if op == "'" and r.dtype.kind == "c":
        r = r.conj()
    if t is str:
        return simplify(r)
    return simplify(np.asfortranarray(r))


def make_range(a, s, b):
    a, s, b = float(np.real(a)), float(np.real(s)), float(np.real(b))
    if s == 0 or (s > 0 and a > b) or (s < 0 and a < b):
        return np.zeros((1, 0))
    n = int(math.floor((b - a) / s + 1e-10))
    return (a + s * np.arange(n + 1, dtype=np.float64)).reshape(1, n + 1)


# ------------------------------------------------------------------------------------------------------- indexing
def _index_vector(ix, dimlen):
    """one subscript -> (0-based integer positions, shape of the subscript)"""
    t = type(ix)
    if t is float:
        k = int(ix)
        if k != ix or k < 1:
            raise MatlabError(f"subscript {ix!r} is not a positive integer")
        return np.array([k - 1]), (1, 1)
    if ix is COLON:
        return np.arange(dimlen), (dimlen, 1)
    if t is bool:
        ix = np.array([[ix]])
    a = to_arr(ix)
    if a.dtype.kind == "b":
        if a.size > dimlen and a.reshape(-1, order="F")[dimlen:].any():
            raise MatlabError("logical subscript reaches beyond the array")
        pos = np.flatnonzero(a.reshape(-1, order="F"))
        shp = (1, pos.size) if (a.ndim == 2 and a.shape[0] == 1) else (pos.size, 1)
        return pos, shp
    v = a.reshape(-1, order="F")
    k = v.real.astype(np.int64)
    if (k != v).any() or (k < 1).any():
        raise MatlabError("subscripts must be positive integers")
    return k - 1, a.shape


class _Colon:
    def __repr__(self):
        return ":"


COLON = _Colon()


def _fold_shape(shape, k):
    """the array shape as seen through k subscripts (trailing dimensions folded into the last, or padded with ones)"""
    n = len(shape)
    if k == n:
        return shape
    if k < n:
        last = 1
        for d in shape[k - 1:]:
            last *= d
        return tuple(shape[:k - 1]) + (last,)
    return tuple(shape) + (1,) * (k - n)


def _scalar(r, kind):
    if kind == "f":
        return float(r)
    if kind == "c":
        return complex(r)
    if kind == "b":
        return bool(r)
    return simplify(np.array([[r]]))


def index(a, subs):
    """a(subs...) for numeric / char / logical arrays"""
    is_str = type(a) is str
    x = to_arr(a)
    k = len(subs)
    if k == 0:
        return a
    if k == 1:
        s = subs[0]
        if type(s) is float:                                     # scalar fast path
            j = int(s)
            if j != s or j < 1:
                raise MatlabError(f"subscript {s!r} is not a positive integer")
            if j > x.size:
                raise MatlabError(f"index {s!r} out of bounds (numel {x.size})")
            r = x.reshape(-1, order="F")[j - 1]
            return chr(int(r)) if is_str else _scalar(r, x.dtype.kind)
        pos, shp = _index_vector(s, x.size)
        if pos.size and pos.max() >= x.size:
            raise MatlabError(f"index {int(pos.max()) + 1} out of bounds (numel {x.size})")
        flat = x.reshape(-1, order="F")[pos]
        if s is COLON:
            out = flat.reshape(-1, 1)
        else:
            src_vec = x.ndim == 2 and (x.shape[0] == 1 or x.shape[1] == 1)
            idx_vec = len(shp) == 2 and (shp[0] == 1 or shp[1] == 1)
            if src_vec and idx_vec and x.size != 1:
                out = flat.reshape(1, -1) if x.shape[0] == 1 else flat.reshape(-1, 1)
            else:
                out = flat.reshape(shp, order="F")
        if is_str:
            return "".join(chr(int(c)) for c in out.reshape(-1, order="F"))
        return simplify(out)
    shape = _fold_shape(x.shape, k)
    if k == 2 and type(subs[0]) is float and type(subs[1]) is float and x.ndim == 2:
        i, j = int(subs[0]), int(subs[1])
        if i != subs[0] or j != subs[1] or i < 1 or j < 1:
            raise MatlabError(f"subscripts ({subs[0]!r},{subs[1]!r}) are not positive integers")
        if i > shape[0] or j > shape[1]:
            raise MatlabError(f"index ({subs[0]!r},{subs[1]!r}) out of bounds ({shape[0]}x{shape[1]})")
        r = x[i - 1, j - 1]
        return chr(int(r)) if is_str else _scalar(r, x.dtype.kind)
    xv = x.reshape(shape, order="F") if shape != x.shape else x
    pos = []
    for d, s in enumerate(subs):
        p, _ = _index_vector(s, shape[d])
        if p.size and p.max() >= shape[d]:
            raise MatlabError(f"index {int(p.max()) + 1} out of bounds in dimension {d + 1} (size {shape[d]})")
        pos.append(p)
    out = xv[np.ix_(*pos)]
    if is_str:
        return "".join(chr(int(c)) for c in out.reshape(-1, order="F"))
    return simplify(np.asfortranarray(out))


def index_assign(a, subs, v):
    """a(subs...) = v; returns the new array (a may be None = undefined variable)"""
    was_str = type(a) is str and type(v) is str
    x = np.zeros((0, 0)) if a is None else to_arr(a)
    vv = to_arr(v)
    k = len(subs)
    # ---- deletion  a(idx) = []
    if vv.size == 0 and type(v) is np.ndarray:
        return _delete(x, subs)
    if k == 1:
        s = subs[0]
        if s is COLON:
            if vv.size == 1:
                out = np.empty(x.shape, dtype=np.result_type(x.dtype, vv.dtype), order="F")
                out[...] = vv.reshape(-1)[0]
                return simplify(out)
            if vv.size != x.size:
                raise MatlabError("A(:) = B needs as many elements in B as in A")
            return simplify(vv.reshape(-1, order="F").reshape(x.shape, order="F").astype(np.result_type(x.dtype, vv.dtype)))
        pos, shp = _index_vector(s, x.size)
        need = int(pos.max()) + 1 if pos.size else 0
        if need > x.size:
            if x.size == 0:
                x = np.zeros((1, need), dtype=x.dtype if a is not None and x.dtype.kind != "f" else np.float64)
            elif x.ndim == 2 and x.shape[0] == 1:
                x = np.concatenate([x, np.zeros((1, need - x.size), dtype=x.dtype)], axis=1)
            elif x.ndim == 2 and x.shape[1] == 1:
                x = np.concatenate([x, np.zeros((need - x.size, 1), dtype=x.dtype)], axis=0)
            else:
                raise MatlabError("linear-index assignment cannot grow a matrix")
        dt = np.result_type(x.dtype, vv.dtype)
        flat = x.reshape(-1, order="F").astype(dt, copy=True)
        if vv.size == 1:
            flat[pos] = vv.reshape(-1)[0]
        else:
            if vv.size != pos.size:
                raise MatlabError(f"A(I) = B: {pos.size} subscripts but {vv.size} values")
            flat[pos] = vv.reshape(-1, order="F")
        out = flat.reshape(x.shape, order="F")
        if was_str:
            return "".join(chr(int(c)) for c in out.reshape(-1, order="F"))
        return simplify(out)
    # ---- k >= 2 subscripts
    shape = list(_fold_shape(x.shape, k)) if x.size or a is not None else [0] * k
    if a is not None and k < x.ndim:
        raise MatlabError("assignment with fewer subscripts than dimensions is not supported")
    vshape = [d for d in vv.shape]
    # a ':' on a dimension of extent 0 takes its extent from the right-hand side
    vdims = [d for d in vshape if d != 1] if vv.size != 1 else []
    pos = []
    colon_dims = [d for d, s in enumerate(subs) if s is COLON and shape[d] == 0]
    if colon_dims and vv.size != 1:
        # match right-hand-side extents to the subscripts in order
        fixed = {d: _index_vector(s, shape[d])[0].size for d, s in enumerate(subs) if not (s is COLON and shape[d] == 0)}
        rem = list(vv.shape) + [1] * max(0, k - vv.ndim)
        if len(rem) == k and all(rem[d] == n for d, n in fixed.items()):
            for d in colon_dims:
                shape[d] = rem[d]
        else:
            free = [n for n in vdims]
            for d in range(k):
                if d in fixed:
                    if fixed[d] != 1 and free and free[0] == fixed[d]:
                        free.pop(0)
                else:
                    shape[d] = free.pop(0) if free else 1
    elif colon_dims:
        for d in colon_dims:
            shape[d] = 1
    for d, s in enumerate(subs):
        p, _ = _index_vector(s, shape[d])
        pos.append(p)
    newshape = [max(shape[d], (int(p.max()) + 1) if p.size else 0) for d, p in enumerate(pos)]
    dt = np.result_type(x.dtype, vv.dtype) if x.size else vv.dtype
    if tuple(newshape) != tuple(shape) or x.size == 0:
        out = np.zeros(newshape, dtype=dt, order="F")
        if x.size:
            xv = x.reshape(shape, order="F")
            out[tuple(slice(0, n) for n in shape)] = xv
    else:
        out = np.array(x.reshape(shape, order="F"), dtype=dt, order="F", copy=True)
    target = tuple(p.size for p in pos)
    if vv.size == 1:
        out[np.ix_(*pos)] = vv.reshape(-1)[0]
    else:
        if [n for n in target if n != 1] != [n for n in vv.shape if n != 1]:
            raise MatlabError(f"A(...) = B: the subscripted region is {'x'.join(map(str, target))} but B is "
                              f"{'x'.join(map(str, vv.shape))}")
        out[np.ix_(*pos)] = vv.reshape(target, order="F")
    if was_str:
        return "".join(chr(int(c)) for c in out.reshape(-1, order="F"))
    return simplify(out)


def _delete(x, subs):
    if len(subs) == 1:
        pos, _ = _index_vector(subs[0], x.size)
        keep = np.ones(x.size, dtype=bool)
        keep[pos] = False
        flat = x.reshape(-1, order="F")[keep]
        return simplify(flat.reshape(-1, 1) if (x.ndim == 2 and x.shape[1] == 1 and x.shape[0] != 1) else flat.reshape(1, -1))
    nc = [d for d, s in enumerate(subs) if s is not COLON]
    if len(nc) != 1:
        raise MatlabError("A(...) = [] may subscript only one dimension")
    d = nc[0]
    shape = _fold_shape(x.shape, len(subs))
    pos, _ = _index_vector(subs[d], shape[d])
    return simplify(np.delete(x.reshape(shape, order="F"), pos, axis=d))


def concat(rows):
    """[ ... ] from evaluated rows (lists of values)"""
    if not rows:
        return EMPTY
    flat = [v for r in rows for v in r]
    if any(type(v) is MCell for v in flat):
        out_rows = []
        for r in rows:
            parts = [v.a if type(v) is MCell else MCell.row([v]).a for v in r]
            parts = [p for p in parts if p.size]
            if parts:
                out_rows.append(np.concatenate(parts, axis=1))
        return MCell(np.concatenate(out_rows, axis=0)) if out_rows else MCell(np.empty((0, 0), dtype=object))
    if any(type(v) in (MStruct, MStructArr) for v in flat):
        items = []
        for v in flat:
            items.extend(v.items if type(v) is MStructArr else [v])
        return MStructArr(items)
    if flat and all(type(v) is str for v in flat) and len(rows) == 1:
        return "".join(flat)
    any_str = any(type(v) is str for v in flat)
    out_rows = []
    for r in rows:
        parts = [to_arr(v) for v in r]
        parts = [p for p in parts if p.size]
        if not parts:
            continue
        if len(parts) == 1:
            out_rows.append(parts[0])
            continue
        nd = max(p.ndim for p in parts)
        parts = [p.reshape(p.shape + (1,) * (nd - p.ndim)) for p in parts]
        for p in parts[1:]:
            if p.shape[0] != parts[0].shape[0] or p.shape[2:] != parts[0].shape[2:]:
                raise MatlabError("horizontal concatenation: dimensions are not consistent")
        out_rows.append(np.concatenate([_promote_cat(p) for p in parts], axis=1))
    if not out_rows:
        return "" if any_str else EMPTY
    if len(out_rows) == 1:
        res = out_rows[0]
    else:
        nd = max(p.ndim for p in out_rows)
        out_rows = [p.reshape(p.shape + (1,) * (nd - p.ndim)) for p in out_rows]
        for p in out_rows[1:]:
            if p.shape[1:] != out_rows[0].shape[1:]:
                raise MatlabError("vertical concatenation: dimensions are not consistent")
        res = np.concatenate([_promote_cat(p) for p in out_rows], axis=0)
    if any_str and res.ndim == 2 and res.shape[0] == 1:
        return "".join(chr(int(c)) for c in res.reshape(-1))
    if all(to_arr(v).dtype.kind == "b" for v in flat):
        return simplify(res.astype(bool))
    return simplify(np.asfortranarray(res))


def _promote_cat(p):
    return p.astype(np.float64) if p.dtype.kind == "b" else p
