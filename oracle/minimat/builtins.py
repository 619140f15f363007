"""minimat.builtins -- the MATLAB library functions the reference's hot path (and the QG driver around it) calls.
TEST INFRASTRUCTURE ONLY, see oracle/minimat/__init__.py.  Every builtin has the signature f(I, args, nargout, frame)."""
import cmath
import glob
import math
import os
import re
import time

import numpy as np

from .values import (MatlabError, MStruct, MStructArr, MCell, MObject, FuncHandle, EMPTY, to_arr, simplify, binop, truth,
                     msize, numel, mclass, drop_zero_imag, concat, make_range)

TABLE = {}


def reg(*names):
    def deco(f):
        for n in names:
            TABLE[n] = f
        return f
    return deco


def _int(v, what="argument"):
    x = to_arr(v)
    if x.size != 1:
        raise MatlabError(f"{what} must be a scalar")
    f = float(np.real(x.reshape(-1)[0]))
    if f != math.floor(f) and not math.isinf(f):
        raise MatlabError(f"{what} must be an integer")
    return f if math.isinf(f) else int(f)


def _dims(args):
    """size arguments of zeros / ones / rand / cell: (n) -> n x n, (m, n, ...), ([m n ...])"""
    args = [a for a in args if type(a) is not str]
    if not args:
        return (1, 1)
    if len(args) == 1:
        a = to_arr(args[0])
        if a.size == 1:
            n = max(_int(args[0]), 0)
            return (n, n)
        d = [max(int(x), 0) for x in a.reshape(-1, order="F")]
    else:
        d = [max(_int(a), 0) for a in args]
    while len(d) > 2 and d[-1] == 1:
        d.pop()
    return tuple(d)


def _f(a):
    x = to_arr(a)
    return x.astype(np.float64) if x.dtype.kind == "b" else x


def _elementwise(name, real_fn, scalar_fn=None):
    def f(I, args, nargout, frame):
        a = args[0]
        if scalar_fn is not None and type(a) is float:
            try:
                return scalar_fn(a)
            except (ValueError, OverflowError, ZeroDivisionError):
                pass
        with np.errstate(all="ignore"):
            return drop_zero_imag(simplify(real_fn(_f(a))))
    TABLE[name] = f


def _sqrt(x):
    if x.dtype.kind != "c" and (x < 0).any():
        return np.sqrt(x.astype(np.complex128))
    return np.sqrt(x)


def _log(x):
    if x.dtype.kind != "c" and (x < 0).any():
        return np.log(x.astype(np.complex128))
    return np.log(x)


def _round(x):
    if x.dtype.kind == "c":
        return _round(x.real) + 1j * _round(x.imag)
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def _sqrt_scalar(a):
    if a < 0:
        raise ValueError
    return math.sqrt(a)


_elementwise("sqrt", _sqrt, _sqrt_scalar)
_elementwise("abs", np.abs, abs)
_elementwise("floor", lambda x: np.floor(x.real) + (1j * np.floor(x.imag) if x.dtype.kind == "c" else 0), lambda a: float(math.floor(a)))
_elementwise("ceil", lambda x: np.ceil(x.real) + (1j * np.ceil(x.imag) if x.dtype.kind == "c" else 0), lambda a: float(math.ceil(a)))
_elementwise("fix", lambda x: np.trunc(x.real) + (1j * np.trunc(x.imag) if x.dtype.kind == "c" else 0), lambda a: float(math.trunc(a)))
_elementwise("round", _round)
_elementwise("exp", np.exp, math.exp)
_elementwise("log", _log, math.log)
_elementwise("log2", lambda x: np.log2(x if (x.dtype.kind == "c" or (x >= 0).all()) else x.astype(complex)))
_elementwise("log10", lambda x: np.log10(x if (x.dtype.kind == "c" or (x >= 0).all()) else x.astype(complex)))
_elementwise("sin", np.sin, math.sin)
_elementwise("cos", np.cos, math.cos)
_elementwise("tan", np.tan, math.tan)
_elementwise("asin", np.arcsin)
_elementwise("acos", np.arccos)
_elementwise("atan", np.arctan, math.atan)
_elementwise("sinh", np.sinh)
_elementwise("cosh", np.cosh)
_elementwise("tanh", np.tanh)
_elementwise("sign", np.sign)
_elementwise("real", lambda x: x.real.copy())
_elementwise("imag", lambda x: x.imag.copy() if x.dtype.kind == "c" else np.zeros(x.shape))
_elementwise("conj", np.conj)
_elementwise("angle", np.angle)
_elementwise("isnan", np.isnan)
_elementwise("isinf", np.isinf)
_elementwise("isfinite", np.isfinite)
_elementwise("double", lambda x: x)
_elementwise("gamma", np.vectorize(math.gamma, otypes=[float]))


@reg("logical")
def _logical(I, args, nargout, frame):
    return simplify(to_arr(args[0]) != 0)


@reg("true")
def _true(I, args, nargout, frame):
    return True if not args else simplify(np.ones(_dims(args), dtype=bool, order="F"))


@reg("false")
def _false(I, args, nargout, frame):
    return False if not args else simplify(np.zeros(_dims(args), dtype=bool, order="F"))


@reg("pi")
def _pi(I, args, nargout, frame):
    return math.pi


@reg("inf", "Inf")
def _inf(I, args, nargout, frame):
    return math.inf if not args else simplify(np.full(_dims(args), np.inf, order="F"))


@reg("nan", "NaN")
def _nan(I, args, nargout, frame):
    return math.nan if not args else simplify(np.full(_dims(args), np.nan, order="F"))


@reg("eps")
def _eps(I, args, nargout, frame):
    if not args:
        return 2.220446049250313e-16
    return simplify(np.spacing(np.abs(_f(args[0]))))


@reg("i", "j", "1i")
def _imag_unit(I, args, nargout, frame):
    return 1j


@reg("nargin")
def _nargin(I, args, nargout, frame):
    return float(frame.nargin)


@reg("nargout")
def _nargout(I, args, nargout, frame):
    return float(frame.nargout)


@reg("zeros")
def _zeros(I, args, nargout, frame):
    return simplify(np.zeros(_dims(args), order="F"))


@reg("ones")
def _ones(I, args, nargout, frame):
    return simplify(np.ones(_dims(args), order="F"))


@reg("eye")
def _eye(I, args, nargout, frame):
    d = _dims(args)
    return simplify(np.asfortranarray(np.eye(d[0], d[1])))


@reg("cell")
def _cell(I, args, nargout, frame):
    d = _dims(args)
    a = np.empty(d[:2], dtype=object)
    for q in range(a.size):
        a.reshape(-1)[q] = EMPTY
    return MCell(a)


@reg("rng")
def _rng(I, args, nargout, frame):
    seed = args[0] if args else 0.0
    if type(seed) is str:
        if seed == "default":
            seed = 0.0
        elif seed == "shuffle":
            seed = float(int(time.time()) % 2**31)
        else:
            raise MatlabError(f"rng('{seed}') is not supported")
    if len(args) > 1 and args[1] not in ("twister",):
        raise MatlabError("only the Mersenne twister generator is implemented")
    # mt19937ar; MATLAB maps seed 0 (the start-up state, rng('default')) to the twister's canonical seed 5489
    I.rng = np.random.RandomState(int(seed) if int(seed) != 0 else 5489)


@reg("rand")
def _rand(I, args, nargout, frame):
    # MATLAB's default generator = MT19937, 53-bit doubles, filled in column-major order = numpy RandomState.random_sample
    d = _dims(args)
    n = int(np.prod(d))
    return simplify(I.rng.random_sample(n).reshape(d, order="F"))


@reg("size")
def _size(I, args, nargout, frame):
    shp = list(msize(args[0]))
    if len(args) > 1:
        dsel = to_arr(args[1]).reshape(-1)
        vals = [float(shp[int(d) - 1]) if int(d) <= len(shp) else 1.0 for d in dsel]
        if len(vals) == 1:
            return vals[0]
        return np.array([vals])
    if nargout <= 1:
        return np.array([[float(x) for x in shp]])
    out = [float(x) for x in shp[:nargout]]
    while len(out) < nargout:
        out.append(1.0)
    if len(shp) > nargout:
        out[-1] = float(np.prod(shp[nargout - 1:]))
    return out


@reg("numel")
def _numel(I, args, nargout, frame):
    return float(numel(args[0]))


@reg("length")
def _length(I, args, nargout, frame):
    s = msize(args[0])
    return 0.0 if 0 in s else float(max(s))


@reg("ndims")
def _ndims(I, args, nargout, frame):
    return float(len(msize(args[0])))


@reg("isempty")
def _isempty(I, args, nargout, frame):
    return numel(args[0]) == 0


@reg("isreal")
def _isreal(I, args, nargout, frame):
    a = args[0]
    if type(a) is complex:
        return False
    return not (type(a) is np.ndarray and a.dtype.kind == "c")


@reg("isnumeric")
def _isnumeric(I, args, nargout, frame):
    a = args[0]
    return type(a) in (float, complex) or (type(a) is np.ndarray and a.dtype.kind in "fc")


@reg("islogical")
def _islogical(I, args, nargout, frame):
    return mclass(args[0]) == "logical"


@reg("ischar")
def _ischar(I, args, nargout, frame):
    return type(args[0]) is str


@reg("iscell")
def _iscell(I, args, nargout, frame):
    return type(args[0]) is MCell


@reg("isstruct")
def _isstruct(I, args, nargout, frame):
    return type(args[0]) in (MStruct, MStructArr)


@reg("isscalar")
def _isscalar(I, args, nargout, frame):
    return numel(args[0]) == 1


@reg("isvector")
def _isvector(I, args, nargout, frame):
    s = msize(args[0])
    return len(s) == 2 and (s[0] == 1 or s[1] == 1) and s[0] * s[1] >= 1


@reg("isa")
def _isa(I, args, nargout, frame):
    v, c = args
    if type(v) is MObject:
        cls = v.cls
        stack = [cls]
        while stack:
            k = stack.pop()
            if k.name == c:
                return True
            stack.extend(x for x in (I.get_class(s) for s in k.supers) if x is not None)
        return False
    if c in ("numeric", "float"):
        return mclass(v) == "double"
    return mclass(v) == c


@reg("class")
def _class(I, args, nargout, frame):
    return mclass(args[0])


@reg("isfield")
def _isfield(I, args, nargout, frame):
    s, n = args
    if type(s) is MStructArr:
        return s.a.size > 0 and type(n) is str and n in s.a.reshape(-1)[0].f
    return type(s) in (MStruct, MObject) and type(n) is str and n in s.f


@reg("fieldnames")
def _fieldnames(I, args, nargout, frame):
    s = args[0]
    names = list(s.f.keys())
    a = np.empty((len(names), 1), dtype=object)
    for i, n in enumerate(names):
        a[i, 0] = n
    return MCell(a)


@reg("struct")
def _struct(I, args, nargout, frame):
    s = MStruct()
    for k in range(0, len(args), 2):
        s.f[args[k]] = args[k + 1]
    return s


def _equal(a, b):
    ta, tb = type(a), type(b)
    if ta in (MStruct, MObject) or tb in (MStruct, MObject):
        return ta is tb and set(a.f) == set(b.f) and all(_equal(a.f[k], b.f[k]) for k in a.f)
    if ta is MStructArr or tb is MStructArr:
        return ta is tb and a.a.shape == b.a.shape and all(_equal(x, y) for x, y in zip(a.a.reshape(-1), b.a.reshape(-1)))
    if ta is MCell or tb is MCell:
        return ta is tb and a.a.shape == b.a.shape and all(_equal(x, y) for x, y in zip(a.a.reshape(-1), b.a.reshape(-1)))
    if ta is FuncHandle or tb is FuncHandle:
        return a is b
    try:
        x, y = to_arr(a), to_arr(b)
    except MatlabError:
        return a is b or a == b
    if x.size == 0 and y.size == 0:
        return x.shape == y.shape or (x.size == 0 and y.size == 0 and len(x.shape) == len(y.shape) == 2 and x.shape == y.shape)
    return msize(a) == msize(b) and bool(np.array_equal(x, y))


@reg("isequal")
def _isequal(I, args, nargout, frame):
    return all(_equal(args[0], b) for b in args[1:])


# ------------------------------------------------------------------------------------------------ arithmetic helpers
@reg("mod")
def _mod(I, args, nargout, frame):
    a, m = args
    if type(a) is float and type(m) is float:
        if m == 0.0:
            return a
        if math.isinf(a) or math.isnan(a) or math.isnan(m):
            return math.nan
        if math.isinf(m):
            return a if (a >= 0) == (m > 0) else m
        r = math.fmod(a, m)                              # exact
        if r != 0.0 and (r < 0) != (m < 0):
            r += m
        return r
    x, y = _f(a), _f(m)
    with np.errstate(all="ignore"):
        r = np.mod(x, y)                                 # numpy: fmod-based, sign of the divisor, exact
        r = np.where(y == 0, x, r)
    return simplify(r)


@reg("rem")
def _rem(I, args, nargout, frame):
    x, y = _f(args[0]), _f(args[1])
    with np.errstate(all="ignore"):
        r = np.fmod(x, y)
        r = np.where(y == 0, np.nan, r)
    return simplify(r)


@reg("atan2")
def _atan2(I, args, nargout, frame):
    return simplify(np.arctan2(_f(args[0]), _f(args[1])))


@reg("hypot")
def _hypot(I, args, nargout, frame):
    return simplify(np.hypot(_f(args[0]), _f(args[1])))


@reg("power")
def _power(I, args, nargout, frame):
    return binop(".^", args[0], args[1])


@reg("times")
def _times(I, args, nargout, frame):
    return binop(".*", args[0], args[1])


@reg("plus")
def _plus(I, args, nargout, frame):
    return binop("+", args[0], args[1])


@reg("minus")
def _minus(I, args, nargout, frame):
    return binop("-", args[0], args[1])


@reg("rdivide")
def _rdivide(I, args, nargout, frame):
    return binop("./", args[0], args[1])


@reg("nthroot")
def _nthroot(I, args, nargout, frame):
    x, n = _f(args[0]), _f(args[1])
    with np.errstate(all="ignore"):
        r = np.sign(x) * np.abs(x) ** (1.0 / n)
        # one Newton correction, as MATLAB's nthroot does, to land on the exact root when there is one
        r = r - (r ** n - x) / (n * r ** (n - 1))
    return simplify(r)


def _reduce_dim(x, args, k=1):
    """the dimension a reduction works along: given, or the first non-singleton one"""
    if len(args) > k and type(args[k]) is not str:
        return _int(args[k]) - 1
    for d, n in enumerate(x.shape):
        if n != 1:
            return d
    return 0


def _along(x, d):
    if d >= x.ndim:
        x = x.reshape(x.shape + (1,) * (d + 1 - x.ndim))
    return x


@reg("sum")
def _sum(I, args, nargout, frame):
    x = _f(args[0])
    if x.size == 0:
        return 0.0
    d = _reduce_dim(x, args)
    x = _along(x, d)
    # MATLAB sums sequentially along the dimension for short vectors; numpy's pairwise sum differs in the last bits for long
    # ones.  A plain left-to-right accumulation is used here (bit-identical to a MATLAB loop, and to sum() for n < ~1000).
    xs = np.moveaxis(x, d, 0)
    acc = np.zeros(xs.shape[1:], dtype=xs.dtype)
    if xs.shape[0] <= 4096:
        for row in xs:
            acc = acc + row
    else:
        acc = xs.sum(axis=0)
    return simplify(np.expand_dims(acc, d))


@reg("prod")
def _prod(I, args, nargout, frame):
    x = _f(args[0])
    d = _reduce_dim(x, args)
    x = _along(x, d)
    return simplify(np.prod(x, axis=d, keepdims=True))


@reg("cumsum")
def _cumsum(I, args, nargout, frame):
    x = _f(args[0])
    d = _reduce_dim(x, args)
    return simplify(np.cumsum(_along(x, d), axis=d))


@reg("mean")
def _mean(I, args, nargout, frame):
    x = _f(args[0])
    d = _reduce_dim(x, args)
    s = to_arr(_sum(I, [x, float(d + 1)], 1, frame))
    return simplify(s / _along(x, d).shape[d])


@reg("dot")
def _dot(I, args, nargout, frame):
    a, b = _f(args[0]), _f(args[1])
    if a.shape != b.shape:
        if a.size == b.size and a.ndim == 2 and b.ndim == 2 and 1 in a.shape and 1 in b.shape:
            a, b = a.reshape(-1, 1), b.reshape(-1, 1)
        else:
            raise MatlabError("dot: A and B must be the same size")
    p = np.conj(a) * b
    return _sum(I, [p] + list(args[2:3]), 1, frame)


@reg("norm")
def _norm(I, args, nargout, frame):
    x = _f(args[0])
    p = args[1] if len(args) > 1 else 2.0
    if 1 in x.shape or x.size <= 1:
        v = x.reshape(-1)
        if p == "inf" or p == math.inf:
            return float(np.abs(v).max()) if v.size else 0.0
        if p == "fro" or p == 2.0:
            return float(np.sqrt(np.sum(np.abs(v) ** 2)))
        return float(np.sum(np.abs(v) ** p) ** (1.0 / p))
    return float(np.linalg.norm(x, "fro" if p == "fro" else (np.inf if p in ("inf", math.inf) else int(p))))


def _minmax(fn, argfn):
    def f(I, args, nargout, frame):
        x = _f(args[0])
        if len(args) >= 2 and numel(args[1]) > 0:                   # max(a, b)
            y = _f(args[1])
            from .values import align
            x, y = align(x, y)
            xr, yr = (np.abs(x), np.abs(y)) if (x.dtype.kind == "c" or y.dtype.kind == "c") else (x, y)
            pick = (xr >= yr) if fn is np.max else (xr <= yr)
            pick = pick | np.isnan(yr)
            return simplify(np.where(pick, x, y))
        if x.size == 0:
            return (EMPTY, EMPTY)[:max(nargout, 1)]
        if len(args) > 2 and type(args[2]) is str and args[2] == "all":
            x = x.reshape(-1, 1, order="F")
            d = 0
        elif len(args) > 2 and numel(args[2]) > 1:                   # vecdim: reduce over several dimensions, one after another
            dims = sorted(int(q) - 1 for q in to_arr(args[2]).reshape(-1))
            for dd in dims:
                x = to_arr(f(I, [x, EMPTY, float(dd + 1)], 1, frame))
            return simplify(x)
        else:
            d = _int(args[2]) - 1 if len(args) > 2 else _reduce_dim(x, [])
        x = _along(x, d)
        key = np.abs(x) if x.dtype.kind == "c" else x
        with np.errstate(all="ignore"):
            if np.isnan(key).any():
                idx = (np.nanargmax if fn is np.max else np.nanargmin)(np.where(np.isnan(key).all(axis=d, keepdims=True), 0.0, key), axis=d)
            else:
                idx = argfn(key, axis=d)
        val = np.take_along_axis(x, np.expand_dims(idx, d), axis=d)
        if nargout >= 2:
            return [simplify(val), simplify(np.expand_dims(idx, d).astype(np.float64) + 1.0)]
        return simplify(val)
    return f


TABLE["max"] = _minmax(np.max, np.argmax)
TABLE["min"] = _minmax(np.min, np.argmin)


@reg("any")
def _any(I, args, nargout, frame):
    x = to_arr(args[0]) != 0
    if x.size == 0:
        return False
    d = _reduce_dim(x, args)
    return simplify(np.any(_along(x, d), axis=d, keepdims=True))


@reg("all")
def _all(I, args, nargout, frame):
    x = to_arr(args[0]) != 0
    if x.size == 0:
        return True
    d = _reduce_dim(x, args)
    return simplify(np.all(_along(x, d), axis=d, keepdims=True))


@reg("find")
def _find(I, args, nargout, frame):
    x = to_arr(args[0])
    pos = np.flatnonzero(x.reshape(-1, order="F")).astype(np.float64) + 1.0
    if len(args) > 1:
        pos = pos[:_int(args[1])]
    if nargout >= 2:
        r = (pos - 1) % x.shape[0] + 1
        c = (pos - 1) // x.shape[0] + 1
        shp = (1, -1) if (x.ndim == 2 and x.shape[0] == 1) else (-1, 1)
        return [simplify(r.reshape(shp)), simplify(c.reshape(shp))]
    return simplify(pos.reshape(1, -1) if (x.ndim == 2 and x.shape[0] == 1 and x.shape[1] != 1) else pos.reshape(-1, 1))


@reg("sort")
def _sort(I, args, nargout, frame):
    x = _f(args[0])
    d = _reduce_dim(x, [a for a in args if type(a) is not str])
    desc = any(a == "descend" for a in args if type(a) is str)
    idx = np.argsort(-x if desc else x, axis=d, kind="stable")
    val = np.take_along_axis(x, idx, axis=d)
    if nargout >= 2:
        return [simplify(val), simplify(idx.astype(np.float64) + 1.0)]
    return simplify(val)


@reg("cumprod")
def _cumprod(I, args, nargout, frame):
    x = _f(args[0])
    d = _reduce_dim(x, args)
    return simplify(np.cumprod(_along(x, d), axis=d))


@reg("diff")
def _diff(I, args, nargout, frame):
    x = _f(args[0])
    d = _reduce_dim(x, [])
    return simplify(np.diff(x, axis=d))


# ------------------------------------------------------------------------------------------------ shape
@reg("reshape")
def _reshape(I, args, nargout, frame):
    a = args[0]
    if len(args) == 2:
        d = [int(x) for x in to_arr(args[1]).reshape(-1, order="F")]
    else:
        d = [None if numel(x) == 0 else _int(x) for x in args[1:]]
    n = numel(a)
    if None in d:
        known = 1
        for x in d:
            if x is not None:
                known *= x
        d[d.index(None)] = n // known if known else 0
    if int(np.prod(d)) != n:
        raise MatlabError(f"reshape: the number of elements must not change ({n} -> {'x'.join(map(str, d))})")
    if len(d) == 1:
        d = [d[0], 1]
    if type(a) is MCell:
        return MCell(a.a.reshape(d[:2], order="F"))
    return simplify(to_arr(a).reshape(d, order="F"))


@reg("squeeze")
def _squeeze(I, args, nargout, frame):
    a = args[0]
    if type(a) is not np.ndarray or a.ndim <= 2:
        return a
    d = [n for n in a.shape if n != 1]
    while len(d) < 2:
        d.append(1)
    return simplify(a.reshape(d, order="F"))


@reg("cat")
def _cat(I, args, nargout, frame):
    d = _int(args[0]) - 1
    parts = [to_arr(a) for a in args[1:]]
    parts = [p for p in parts if p.size]
    if not parts:
        return EMPTY
    nd = max(max(p.ndim for p in parts), d + 1)
    parts = [p.reshape(p.shape + (1,) * (nd - p.ndim)) for p in parts]
    return simplify(np.asfortranarray(np.concatenate(parts, axis=d)))


@reg("horzcat")
def _horzcat(I, args, nargout, frame):
    return concat([list(args)])


@reg("vertcat")
def _vertcat(I, args, nargout, frame):
    return concat([[a] for a in args])


@reg("repmat")
def _repmat(I, args, nargout, frame):
    x = to_arr(args[0])
    reps = _dims(args[1:]) if len(args) > 2 or numel(args[1]) > 1 else (_int(args[1]),) * 2
    nd = max(x.ndim, len(reps))
    x = x.reshape(x.shape + (1,) * (nd - x.ndim))
    reps = tuple(reps) + (1,) * (nd - len(reps))
    return simplify(np.asfortranarray(np.tile(x, reps)))


@reg("permute")
def _permute(I, args, nargout, frame):
    x = to_arr(args[0])
    order = [int(v) - 1 for v in to_arr(args[1]).reshape(-1)]
    x = x.reshape(x.shape + (1,) * (len(order) - x.ndim))
    return simplify(np.asfortranarray(np.transpose(x, order)))


@reg("transpose")
def _transpose(I, args, nargout, frame):
    from .values import transpose
    return transpose(".'", args[0])


@reg("flipud")
def _flipud(I, args, nargout, frame):
    return simplify(np.asfortranarray(to_arr(args[0])[::-1]))


@reg("fliplr")
def _fliplr(I, args, nargout, frame):
    return simplify(np.asfortranarray(to_arr(args[0])[:, ::-1]))


@reg("rot90")
def _rot90(I, args, nargout, frame):
    k = _int(args[1]) if len(args) > 1 else 1
    return simplify(np.asfortranarray(np.rot90(to_arr(args[0]), k)))           # counter-clockwise, like MATLAB


@reg("circshift")
def _circshift(I, args, nargout, frame):
    x = to_arr(args[0])
    sh = [int(v) for v in to_arr(args[1]).reshape(-1)]
    if len(args) > 2:
        return simplify(np.roll(x, sh[0], axis=_int(args[2]) - 1))
    if len(sh) == 1:
        d = _reduce_dim(x, [])
        return simplify(np.roll(x, sh[0], axis=d))
    return simplify(np.roll(x, sh, axis=tuple(range(len(sh)))))


@reg("ndgrid")
def _ndgrid(I, args, nargout, frame):
    vs = [to_arr(a).reshape(-1, order="F") for a in args]
    if len(vs) == 1:
        vs = vs * max(nargout, 2)
    g = np.meshgrid(*vs, indexing="ij")
    return [np.asfortranarray(x) for x in g][:max(nargout, 1)]


@reg("meshgrid")
def _meshgrid(I, args, nargout, frame):
    vs = [to_arr(a).reshape(-1, order="F") for a in args]
    if len(vs) == 1:
        vs = vs * 2
    g = np.meshgrid(*vs, indexing="xy")
    return [np.asfortranarray(x) for x in g][:max(nargout, 1)]


@reg("linspace")
def _linspace(I, args, nargout, frame):
    a, b = float(args[0]), float(args[1])
    n = _int(args[2]) if len(args) > 2 else 100
    if n <= 0:
        return np.zeros((1, 0))
    if n == 1:
        return b
    # MATLAB's linspace: a + (0:n-2)*(b-a)/(n-1), last point exactly b, symmetric fill for accuracy
    n1 = n - 1
    c = (b - a) * (n1 - 1)
    if math.isinf(c):
        y = a + (b / n1) * np.arange(n) - (a / n1) * np.arange(n)
    else:
        y = a + np.arange(n) * (b - a) / n1
    y[0], y[-1] = a, b
    return y.reshape(1, n)


@reg("colon")
def _colon(I, args, nargout, frame):
    if len(args) == 2:
        return make_range(args[0], 1.0, args[1])
    return make_range(args[0], args[1], args[2])


# ------------------------------------------------------------------------------------------------ page-wise linear algebra
def _pages_first(x):
    """m x k x p1 x p2 ... -> (p..., m, k) for numpy's stacked linear algebra"""
    if x.ndim == 2:
        return x
    return np.moveaxis(x, (0, 1), (-2, -1))


def _pages_last(x):
    if x.ndim == 2:
        return x
    return np.asfortranarray(np.moveaxis(x, (-2, -1), (0, 1)))


@reg("pagemtimes")
def _pagemtimes(I, args, nargout, frame):
    a, b = _f(args[0]), _f(args[1])
    # numpy stacks pages in front; MATLAB keeps them behind: pages are reversed into numpy order and back
    A = np.transpose(_pages_first(a), tuple(range(a.ndim - 2))[::-1] + (a.ndim - 2, a.ndim - 1)) if a.ndim > 2 else a
    B = np.transpose(_pages_first(b), tuple(range(b.ndim - 2))[::-1] + (b.ndim - 2, b.ndim - 1)) if b.ndim > 2 else b
    C = np.matmul(A, B)
    if C.ndim > 2:
        C = np.transpose(C, tuple(range(C.ndim - 2))[::-1] + (C.ndim - 2, C.ndim - 1))
    return simplify(_pages_last(C))


@reg("pageinv")
def _pageinv(I, args, nargout, frame):
    a = _f(args[0])
    return simplify(_pages_last(np.linalg.inv(_pages_first(a))))


@reg("pageeig")
def _pageeig(I, args, nargout, frame):
    """[V, D] = pageeig(A): eigenvectors (unit 2-norm, as LAPACK returns them) and the diagonal eigenvalue matrix of every page.
    Order and phase of the eigenvectors are LAPACK's in MATLAB and in numpy alike, and arbitrary in both; callers may only rely
    on V*D*inv(V)."""
    a = _f(args[0])
    w, v = np.linalg.eig(_pages_first(a))
    if nargout < 2:
        return simplify(_pages_last(w[..., :, None]))
    d = np.zeros(v.shape, dtype=w.dtype)
    idx = np.arange(w.shape[-1])
    d[..., idx, idx] = w
    return [simplify(_pages_last(v)), simplify(_pages_last(d))]


@reg("eig")
def _eig(I, args, nargout, frame):
    a = _f(args[0])
    w, v = np.linalg.eig(a)
    if nargout < 2:
        return simplify(w.reshape(-1, 1))
    return [simplify(np.asfortranarray(v)), simplify(np.asfortranarray(np.diag(w)))]


@reg("inv")
def _inv(I, args, nargout, frame):
    return simplify(np.asfortranarray(np.linalg.inv(_f(args[0]))))


@reg("diag")
def _diag(I, args, nargout, frame):
    a = _f(args[0])
    if 1 in a.shape:
        return simplify(np.asfortranarray(np.diag(a.reshape(-1))))
    return simplify(np.diag(a).reshape(-1, 1))


# ------------------------------------------------------------------------------------------------ FFT kit
def _fft_axis(x, args):
    n = None
    if len(args) > 1 and numel(args[1]) > 0:
        n = _int(args[1])
    d = _int(args[2]) - 1 if len(args) > 2 else _reduce_dim(x, [])
    return n, d


@reg("fft")
def _fft(I, args, nargout, frame):
    x = _f(args[0])
    n, d = _fft_axis(x, args)
    return simplify(np.asfortranarray(np.fft.fft(x, n=n, axis=d)))


def _conj_symmetric(x, axes):
    """exactly conjugate-symmetric along ``axes`` (the condition under which MATLAB's ifft / ifft2 return a real array)"""
    if x.dtype.kind != "c":
        xr = x
        for ax in axes:
            xr = np.roll(np.flip(xr, axis=ax), 1, axis=ax)
        return np.array_equal(xr, x)
    y = x
    for ax in axes:
        y = np.roll(np.flip(y, axis=ax), 1, axis=ax)
    return np.array_equal(np.conj(y), x)


@reg("ifft")
def _ifft(I, args, nargout, frame):
    x = _f(args[0])
    n, d = _fft_axis(x, args)
    r = np.fft.ifft(x, n=n, axis=d)
    if any(a == "symmetric" for a in args if type(a) is str) or (n is None and _conj_symmetric(x, (d,))):
        r = r.real
    return simplify(np.asfortranarray(r))


@reg("fft2")
def _fft2(I, args, nargout, frame):
    x = _f(args[0])
    return simplify(np.asfortranarray(np.fft.fft2(x, axes=(0, 1))))


@reg("ifft2")
def _ifft2(I, args, nargout, frame):
    x = _f(args[0])
    r = np.fft.ifft2(x, axes=(0, 1))
    if any(a == "symmetric" for a in args if type(a) is str) or _conj_symmetric(x, (0, 1)):
        r = r.real
    return simplify(np.asfortranarray(r))


@reg("fftshift")
def _fftshift(I, args, nargout, frame):
    x = to_arr(args[0])
    if len(args) > 1:
        return simplify(np.fft.fftshift(x, axes=_int(args[1]) - 1))
    return simplify(np.asfortranarray(np.fft.fftshift(x)))


@reg("ifftshift")
def _ifftshift(I, args, nargout, frame):
    x = to_arr(args[0])
    if len(args) > 1:
        return simplify(np.fft.ifftshift(x, axes=_int(args[1]) - 1))
    return simplify(np.asfortranarray(np.fft.ifftshift(x)))


# ------------------------------------------------------------------------------------------------ strings, printing
def _fmt(fmt, args):
    """sprintf semantics: escapes processed, the format recycled until the arguments are used up"""
    fmt = (fmt.replace("\\n", "\n").replace("\\t", "\t").replace("\\r", "\r").replace("\\\\", "\\"))
    flat = []
    for a in args:
        if type(a) is str:
            flat.append(a)
        elif type(a) in (MCell, MStruct, MObject, FuncHandle):
            raise MatlabError("fprintf / sprintf: cell, struct and object arguments are not printable")
        else:
            flat.extend(simplify(np.array([[x]])) for x in to_arr(a).reshape(-1, order="F"))
    spec = re.compile(r"%(%|[-+ 0#]*\d*(?:\.\d+)?[diufeEgGsxXc])")
    pieces = spec.split(fmt)                       # literal, spec, literal, spec, ...
    nspec = sum(1 for k in range(1, len(pieces), 2) if pieces[k] != "%")
    out = []
    pos = 0
    first = True
    while first or (pos < len(flat) and nspec):
        first = False
        for k, p in enumerate(pieces):
            if k % 2 == 0:
                out.append(p)
                continue
            if p == "%":
                out.append("%")
                continue
            if pos >= len(flat):
                if flat or nspec == 0:
                    break                              # out of data: stop at the first spec that has none
                out.append("")
                continue
            v = flat[pos]
            pos += 1
            conv = p[-1]
            flags = p[:-1]
            if conv in "di":
                if type(v) is str:
                    v = float(ord(v[0])) if v else 0.0
                v = float(np.real(v))
                if v == math.floor(v) and not math.isinf(v):
                    out.append(("%" + flags + "d") % int(v))
                else:
                    out.append(("%" + re.sub(r"\.\d+", "", flags) + "e") % v)
            elif conv in "feEgG":
                if type(v) is str:
                    v = float(ord(v[0])) if v else 0.0
                out.append(("%" + flags + conv) % float(np.real(v)))
            elif conv in "xXu":
                out.append(("%" + flags + ("d" if conv == "u" else conv)) % int(np.real(v)))
            elif conv == "c":
                out.append(v if type(v) is str else chr(int(np.real(v))))
            elif conv == "s":
                if type(v) is str:
                    out.append(("%" + flags + "s") % v)
                else:
                    v = float(np.real(v))
                    out.append(("%" + flags + "s") % (("%d" % int(v)) if v == math.floor(v) and not math.isinf(v) else ("%g" % v)))
        if pos >= len(flat):
            break
    return "".join(out)


@reg("sprintf")
def _sprintf(I, args, nargout, frame):
    return _fmt(args[0], args[1:])


@reg("fprintf")
def _fprintf(I, args, nargout, frame):
    if args and type(args[0]) is not str:
        fid = _int(args[0])
        text = _fmt(args[1], args[2:])
        if fid in (1, 2):
            I.write(text)
        else:
            I.files[fid].write(text.encode())
    else:
        I.write(_fmt(args[0], args[1:]))


@reg("disp")
def _disp(I, args, nargout, frame):
    v = args[0]
    I.write((v if type(v) is str else I.fmt_value(v)) + "\n")


@reg("display")
def _display(I, args, nargout, frame):
    I.display("ans", args[0])


@reg("error")
def _error(I, args, nargout, frame):
    if len(args) > 1 and type(args[0]) is str and ":" in args[0] and " " not in args[0]:
        args = args[1:]                                # error('id:part', fmt, ...)
    msg = _fmt(args[0], args[1:]) if args and type(args[0]) is str else "error"
    raise MatlabError(msg)


@reg("warning")
def _warning(I, args, nargout, frame):
    if args and type(args[0]) is str and args[0] in ("on", "off"):
        return
    I.write("Warning: " + (_fmt(args[0], args[1:]) if args and type(args[0]) is str else "") + "\n")


@reg("assert")
def _assert(I, args, nargout, frame):
    if not truth(args[0]):
        raise MatlabError(_fmt(args[1], args[2:]) if len(args) > 1 else "Assertion failed.")


@reg("num2str")
def _num2str(I, args, nargout, frame):
    v = args[0]
    if type(v) is str:
        return v
    if len(args) > 1 and type(args[1]) is str:
        return _fmt(args[1], [v])
    x = to_arr(v)
    if x.size == 1:
        f = float(np.real(x.reshape(-1)[0]))
        if len(args) > 1:
            return "%.*g" % (_int(args[1]), f)
        if f == math.floor(f) and abs(f) < 1e15:
            return "%d" % int(f)
        return ("%11.5g" % f).strip() if abs(f) >= 1e-5 else ("%11.4e" % f).strip()
    return "  ".join(_num2str(I, [simplify(np.array([[q]]))], 1, frame) for q in x.reshape(-1, order="F"))


@reg("str2double", "str2num")
def _str2double(I, args, nargout, frame):
    try:
        return float(args[0])
    except (TypeError, ValueError):
        return math.nan


@reg("int2str")
def _int2str(I, args, nargout, frame):
    return "%d" % int(_round(_f(args[0])).reshape(-1)[0])


@reg("strcat")
def _strcat(I, args, nargout, frame):
    return "".join(a.rstrip(" \t") if type(a) is str else "".join(chr(int(c)) for c in to_arr(a).reshape(-1)) for a in args)


@reg("strcmp")
def _strcmp(I, args, nargout, frame):
    return type(args[0]) is str and type(args[1]) is str and args[0] == args[1]


@reg("strcmpi")
def _strcmpi(I, args, nargout, frame):
    return type(args[0]) is str and type(args[1]) is str and args[0].lower() == args[1].lower()


@reg("upper")
def _upper(I, args, nargout, frame):
    return args[0].upper()


@reg("lower")
def _lower(I, args, nargout, frame):
    return args[0].lower()


@reg("strtrim")
def _strtrim(I, args, nargout, frame):
    return args[0].strip()


@reg("strrep")
def _strrep(I, args, nargout, frame):
    return args[0].replace(args[1], args[2])


@reg("string", "char")
def _string(I, args, nargout, frame):
    v = args[0]
    if type(v) is str:
        return v
    x = to_arr(v)
    return "".join(chr(int(c)) for c in x.reshape(-1, order="F"))


@reg("datestr")
def _datestr(I, args, nargout, frame):
    return time.strftime("%d-%b-%Y %H:%M:%S")


@reg("func2str")
def _func2str(I, args, nargout, frame):
    h = args[0]
    return h.name or "@anonymous"


@reg("str2func")
def _str2func(I, args, nargout, frame):
    return FuncHandle("name", name=args[0].lstrip("@"))


@reg("feval")
def _feval(I, args, nargout, frame):
    f = args[0]
    if type(f) is str:
        f = FuncHandle("name", name=f)
    return I.call_handle(f, list(args[1:]), nargout, frame)


@reg("deal")
def _deal(I, args, nargout, frame):
    n = max(nargout, 1)
    if len(args) == 1:
        return [args[0]] * n
    return list(args[:n])


@reg("cellfun")
def _cellfun(I, args, nargout, frame):
    f, c = args[0], args[1]
    uniform = True
    rest = list(args[2:])
    if "UniformOutput" in rest:
        k = rest.index("UniformOutput")
        uniform = truth(rest[k + 1])
    if type(f) is str:
        f = FuncHandle("name", name=f)
    outs = [I.call_handle(f, [x], 1, frame)[0] for x in c.a.reshape(-1, order="F")]
    if uniform:
        return simplify(np.array([to_arr(o).reshape(-1)[0] for o in outs]).reshape(c.a.shape, order="F"))
    a = np.empty(c.a.shape, dtype=object)
    for q, o in enumerate(outs):
        a.reshape(-1, order="F")[q] = o
    return MCell(a)


@reg("arrayfun")
def _arrayfun(I, args, nargout, frame):
    f = args[0]
    arrs = []
    for a in args[1:]:
        if type(a) is str:
            break
        arrs.append(to_arr(a))
    if any(x.shape != arrs[0].shape for x in arrs):
        raise MatlabError("arrayfun: all inputs must have the same size")
    cols = [x.reshape(-1, order="F") for x in arrs]
    outs = []
    for q in range(arrs[0].size):
        r = I.call_handle(f, [simplify(np.array([[c[q]]])) for c in cols], 1, frame)
        o = to_arr(r[0])
        if o.size != 1:
            raise MatlabError("arrayfun: non-scalar output with UniformOutput = true")
        outs.append(o.reshape(-1)[0])
    return simplify(np.array(outs).reshape(arrs[0].shape, order="F")) if outs else np.zeros(arrs[0].shape)


# ------------------------------------------------------------------------------------------------ files, path, time
@reg("fullfile")
def _fullfile(I, args, nargout, frame):
    parts = [a for a in args if a != ""]
    if not parts:
        return ""
    out = parts[0]
    for p in parts[1:]:
        out = out.rstrip("/") + "/" + p.lstrip("/") if out else p
    return out


@reg("fileparts")
def _fileparts(I, args, nargout, frame):
    p = args[0]
    d, b = os.path.split(p)
    n, e = os.path.splitext(b)
    return [d, n, e][:max(nargout, 1)]


@reg("mfilename")
def _mfilename(I, args, nargout, frame):
    f = frame.func
    if f is None:
        return ""
    p = os.path.splitext(f.unit.path)[0]
    if args and args[0] == "fullpath":
        return p
    return os.path.basename(p)


@reg("pwd")
def _pwd(I, args, nargout, frame):
    return I.cwd


@reg("cd")
def _cd(I, args, nargout, frame):
    if not args:
        return I.cwd
    p = I.abspath(args[0])
    if not os.path.isdir(p):
        raise MatlabError(f"cd: no such folder '{args[0]}'")
    I.cwd = p


@reg("addpath")
def _addpath(I, args, nargout, frame):
    dirs = []
    for a in args:
        if a in ("-begin", "-end"):
            continue
        dirs.extend(x for x in a.split(os.pathsep) if x)
    end = "-end" in args
    for d in reversed(dirs):
        p = I.abspath(d)
        if not os.path.isdir(p):
            I.write(f"Warning: Name is nonexistent or not a directory: {d}\n")
            continue
        if p in I.path:
            I.path.remove(p)
        if end:
            I.path.append(p)
        else:
            I.path.insert(0, p)


@reg("rmpath")
def _rmpath(I, args, nargout, frame):
    for a in args:
        p = I.abspath(a)
        if p in I.path:
            I.path.remove(p)


@reg("genpath")
def _genpath(I, args, nargout, frame):
    root = I.abspath(args[0])
    return os.pathsep.join(d for d, _, _ in os.walk(root))


@reg("exist")
def _exist(I, args, nargout, frame):
    name = args[0]
    kind = args[1] if len(args) > 1 else None
    if kind in (None, "var") and frame is not None and I.getvar(frame, name) is not None:
        return 1.0
    if kind == "var":
        return 0.0
    p = I.abspath(name)
    if kind in (None, "dir") and os.path.isdir(p):
        return 7.0
    if kind in (None, "file"):
        if os.path.isfile(p) or I.find_file(name) is not None:
            return 2.0
        if os.path.isdir(p):
            return 7.0
        if kind is None and name in TABLE:
            return 5.0
    if kind == "builtin":
        return 5.0 if name in TABLE else 0.0
    return 0.0


@reg("mkdir")
def _mkdir(I, args, nargout, frame):
    os.makedirs(I.abspath(args[0] if len(args) == 1 else os.path.join(args[0], args[1])), exist_ok=True)
    return True


@reg("dir")
def _dir(I, args, nargout, frame):
    pat = I.abspath(args[0]) if args else I.cwd
    if os.path.isdir(pat):
        names = sorted(os.listdir(pat))
        folder = pat
        paths = [os.path.join(pat, n) for n in names]
    else:
        paths = sorted(glob.glob(pat))
        folder = os.path.dirname(pat)
    items = []
    for p in paths:
        st = os.stat(p)
        items.append(MStruct({"name": os.path.basename(p), "folder": folder, "bytes": float(st.st_size),
                              "isdir": os.path.isdir(p), "datenum": st.st_mtime / 86400.0 + 719529.0}))
    return MStructArr(items)


@reg("delete")
def _delete(I, args, nargout, frame):
    for a in args:
        for p in glob.glob(I.abspath(a)):
            os.remove(p)


@reg("fopen")
def _fopen(I, args, nargout, frame):
    name = args[0]
    mode = args[1] if len(args) > 1 else "r"
    if len(args) > 2 and args[2] not in ("n", "native", "l", "ieee-le", "ieee-le.l64"):
        raise MatlabError(f"fopen: machine format '{args[2]}' is not supported")
    pmode = {"r": "rb", "w": "wb", "a": "ab", "r+": "r+b", "w+": "w+b", "a+": "a+b", "rt": "rb", "wt": "wb", "at": "ab"}.get(mode)
    if pmode is None:
        raise MatlabError(f"fopen: mode '{mode}'")
    try:
        fh = open(I.abspath(name), pmode)
    except OSError as ex:
        return [-1.0, ex.strerror or "cannot open file"][:max(nargout, 1)]
    fid = I.next_fid
    I.next_fid += 1
    I.files[fid] = fh
    return [float(fid), ""][:max(nargout, 1)]


@reg("fclose")
def _fclose(I, args, nargout, frame):
    if args and args[0] == "all":
        I.close_all()
        return 0.0
    fh = I.files.pop(_int(args[0]), None)
    if fh is None:
        raise MatlabError("fclose: invalid file identifier")
    fh.close()
    return 0.0


_DTYPES = {"real*8": "<f8", "double": "<f8", "float64": "<f8", "real*4": "<f4", "single": "<f4", "float32": "<f4", "int32": "<i4",
           "int64": "<i8", "uint8": "u1", "int8": "i1", "uint32": "<u4", "int16": "<i2", "uint16": "<u2", "char": "u1", "uchar": "u1"}


def _dtype(spec):
    spec = spec.split("=>")[0].strip()
    if spec not in _DTYPES:
        raise MatlabError(f"precision '{spec}' is not supported")
    return np.dtype(_DTYPES[spec])


@reg("fread")
def _fread(I, args, nargout, frame):
    fh = I.files[_int(args[0])]
    size = args[1] if len(args) > 1 and type(args[1]) is not str else math.inf
    prec = next((a for a in args[1:] if type(a) is str), "uint8")
    dt = _dtype(prec)
    s = to_arr(size).reshape(-1)
    if s.size == 1:
        n = s[0]
        raw = fh.read() if math.isinf(n) else fh.read(int(n) * dt.itemsize)
        data = np.frombuffer(raw[:len(raw) // dt.itemsize * dt.itemsize], dtype=dt).astype(np.float64)
        out = data.reshape(-1, 1)
    else:
        m, n = s[0], s[1]
        raw = fh.read() if math.isinf(n) else fh.read(int(m) * int(n) * dt.itemsize)
        data = np.frombuffer(raw[:len(raw) // dt.itemsize * dt.itemsize], dtype=dt).astype(np.float64)
        m = int(m)
        cols = -(-data.size // m) if m else 0
        if cols * m != data.size:
            data = np.concatenate([data, np.zeros(cols * m - data.size)])
        out = data.reshape((m, cols), order="F")
    cnt = float(data.size)
    out = simplify(np.asfortranarray(out)) if out.size else np.zeros(out.shape)
    return [out, cnt][:max(nargout, 1)]


@reg("fwrite")
def _fwrite(I, args, nargout, frame):
    fh = I.files[_int(args[0])]
    prec = args[2] if len(args) > 2 else "uint8"
    dt = _dtype(prec)
    a = args[1]
    x = to_arr(a) if type(a) is not str else np.array([ord(c) for c in a], dtype=np.float64)
    if x.dtype.kind == "c":
        x = x.real
    fh.write(np.ascontiguousarray(x.reshape(-1, order="F").astype(dt)).tobytes())
    return float(x.size)


@reg("fseek")
def _fseek(I, args, nargout, frame):
    fh = I.files[_int(args[0])]
    origin = args[2] if len(args) > 2 else -1.0
    whence = {"bof": 0, "cof": 1, "eof": 2}.get(origin) if type(origin) is str else {-1: 0, 0: 1, 1: 2}[_int(origin)]
    try:
        fh.seek(_int(args[1]), whence)
        return 0.0
    except (OSError, ValueError):
        return -1.0


@reg("ftell")
def _ftell(I, args, nargout, frame):
    return float(I.files[_int(args[0])].tell())


@reg("ferror")
def _ferror(I, args, nargout, frame):
    return ""


@reg("feof")
def _feof(I, args, nargout, frame):
    fh = I.files[_int(args[0])]
    pos = fh.tell()
    end = fh.seek(0, 2)
    fh.seek(pos)
    return pos >= end


@reg("textscan")
def _textscan(I, args, nargout, frame):
    """textscan(fid, format [, N] [, 'delimiter', d] [, 'headerlines', h]) -> 1 x nconv cell of columns.  The subset the
    reference's log parser uses (symplectic_full_fourier.m:68-83): literal text with %f / %d conversions, applied up to N times
    (as often as it matches when N is absent); white space in the format matches any run of white space, including line
    breaks; 'headerlines' first skips that many line ends, the remainder of the current line counting as one; the file
    position is left behind the last character consumed."""
    fh = I.files[_int(args[0])]
    fmt = args[1]
    rest = list(args[2:])
    nrep = None
    if rest and type(rest[0]) is not str:
        nrep = _int(rest.pop(0))
    opts = {rest[i].lower(): rest[i + 1] for i in range(0, len(rest) - 1, 2)}
    start = fh.tell()
    text = fh.read().decode("latin-1")
    pos = 0
    for _ in range(int(opts.get("headerlines", 0))):
        j = text.find("\n", pos)
        pos = len(text) if j < 0 else j + 1
    # format -> regex
    pieces = re.split(r"(%[fd])", fmt.replace("\\n", "\n"))
    pat, nconv = "", 0
    for pc in pieces:
        if pc in ("%f", "%d"):
            pat += r"\s*([-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?)" if pc == "%f" else r"\s*([-+]?\d+)"
            nconv += 1
        else:
            pat += r"\s*".join(re.escape(w) for w in re.split(r"\s+", pc)) if pc.strip() else (r"\s*" if pc else "")
    rx = re.compile(r"\s*" + pat)
    cols = [[] for _ in range(nconv)]
    while nrep is None or len(cols[0]) < nrep:
        m = rx.match(text, pos)
        if not m or m.end() == pos:
            break
        for c, g in zip(cols, m.groups()):
            c.append(float(g))
        pos = m.end()
    fh.seek(start + len(text[:pos].encode("latin-1")))
    return MCell.row([simplify(np.array(c, dtype=np.float64).reshape(-1, 1)) if c else np.zeros((0, 1)) for c in cols])


@reg("vecnorm")
def _vecnorm(I, args, nargout, frame):
    x = _f(args[0])
    p = float(args[1]) if len(args) > 1 else 2.0
    d = _int(args[2]) - 1 if len(args) > 2 else _reduce_dim(x, [])
    if p != 2.0:
        raise MatlabError("vecnorm: only the 2-norm is implemented")
    sq = _sum(I, [np.abs(x) ** 2, float(d + 1)], 1, frame)
    return simplify(np.sqrt(to_arr(sq)))


@reg("tic")
def _tic(I, args, nargout, frame):
    I.tic_time = time.perf_counter()
    if nargout:
        return I.tic_time


@reg("toc")
def _toc(I, args, nargout, frame):
    t0 = args[0] if args else (I.tic_time if I.tic_time is not None else time.perf_counter())
    el = time.perf_counter() - t0
    if nargout:
        return el
    I.write(f"Elapsed time is {el:.6f} seconds.\n")


@reg("clock")
def _clock(I, args, nargout, frame):
    t = time.localtime()
    return np.array([[float(t.tm_year), float(t.tm_mon), float(t.tm_mday), float(t.tm_hour), float(t.tm_min), float(t.tm_sec)]])


@reg("clear")
def _clear(I, args, nargout, frame):
    if frame is None:
        return
    names = [a for a in args if type(a) is str and not a.startswith("-")]
    if not names or names == ["all"] or names == ["variables"]:
        frame.vars.clear()
        return
    for nm in names:
        I.owner(frame, nm).vars.pop(nm, None) if frame.parent is not None else frame.vars.pop(nm, None)


@reg("get")
def _get(I, args, nargout, frame):
    # graphics property query; only the default colormap size is ever looked at (qg_flow_ray_trace/redblue.m:18)
    if len(args) > 1 and type(args[1]) is str and args[1].lower() == "colormap":
        return np.zeros((256, 3), order="F")
    return EMPTY


# graphics and session commands the driver scripts sprinkle around the numerics: accepted, ignored
def _noop(I, args, nargout, frame):
    return [EMPTY] * nargout if nargout else None


for _n in ("figure", "hold", "close", "clc", "clf", "drawnow", "format", "more", "axis", "colorbar", "shading", "plot", "pcolor",
           "title", "xlabel", "ylabel", "legend", "caxis", "colormap", "set", "pause", "subplot", "scatter", "quiver", "contour",
           "imagesc", "print", "saveas", "xlim", "ylim", "grid", "box", "getframe", "writeVideo", "open", "view", "surf",
           "mesh", "loglog", "semilogx", "semilogy", "histogram", "sgtitle", "daspect", "text", "line", "gca", "gcf", "clim"):
    TABLE[_n] = _noop
