"""minimat.interp -- tree-walking evaluator for the MATLAB subset (TEST INFRASTRUCTURE ONLY, see oracle/minimat/__init__.py)."""
import os
import sys

import numpy as np

from .parser import parse_source, FuncDef, ClassDef
from .values import (MatlabError, MStruct, MStructArr, MCell, MObject, FuncHandle, EMPTY, COLON, to_arr, simplify, binop, unop,
                     transpose, make_range, index, index_assign, concat, truth, msize, numel, mclass)


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    pass


class Frame:
    __slots__ = ("vars", "func", "parent", "nargin", "nargout", "globals", "shared_up")

    def __init__(self, func, parent=None, nargin=0, nargout=0):
        self.vars = {}
        self.func = func
        self.parent = parent             # frame of the lexically enclosing function (nested functions only)
        self.nargin, self.nargout = nargin, nargout
        self.globals = None
        # names a nested function shares with its ancestors: used by an ancestor, not a parameter / output of this function
        self.shared_up = (func.ancestors_names() - func.locals_) if (func is not None and parent is not None) else None


def _tame_malloc():
    """Value semantics copy arrays on every indexed assignment; with glibc every copy above 128 KB is an mmap / munmap pair
    and the page faults that follow (a third of the run time of the recipe).  Raising the mmap and trim thresholds keeps such
    blocks on the heap.  Harmless elsewhere; silently skipped where mallopt is not available."""
    try:
        import ctypes
        libc = ctypes.CDLL("libc.so.6")
        libc.mallopt(-3, 1 << 30)      # M_MMAP_THRESHOLD
        libc.mallopt(-1, 1 << 30)      # M_TRIM_THRESHOLD
    except Exception:
        pass


_tame_malloc()


class Interp:
    def __init__(self, cwd=None, out=None):
        self.cwd = os.path.abspath(cwd or os.getcwd())
        self.path = []
        self.out = out or sys.stdout
        self.units = {}                  # file path -> (mtime, Unit)
        self.classes = {}
        self.globals = {}
        self.files = {}                  # fid -> python file object
        self.next_fid = 3
        self.rng = np.random.RandomState(5489)       # MATLAB's start-up generator: Mersenne twister, seed 0 (= mt19937ar's 5489)
        self.tic_time = None
        self.stack = []                  # (function name, line) for error messages
        from . import builtins as B
        self.builtins = B.TABLE
        self.on_output = None            # optional callback(text) -> may raise to stop a run
        self._lookup_key, self._lookup = None, {}
        self.overrides = {}              # name -> builtin-style callable that shadows the path (test shims, see tests/test_minimat.py)

    # ------------------------------------------------------------------ files and lookup
    def abspath(self, p):
        return p if os.path.isabs(p) else os.path.normpath(os.path.join(self.cwd, p))

    def load_unit(self, path):
        hit = self.units.get(path)
        if hit:
            return hit[1]
        mt = os.path.getmtime(path)
        with open(path, "r", errors="replace") as fh:
            src = fh.read()
        u = parse_source(src, path)
        self.units[path] = (mt, u)
        return u

    def find_file(self, name):
        key = (self.cwd, tuple(self.path))
        if key != self._lookup_key:                  # cd / addpath / rmpath invalidate the name cache
            self._lookup_key, self._lookup = key, {}
        hit = self._lookup.get(name, 0)
        if hit != 0:
            return hit
        found = None
        for d in [self.cwd] + self.path:
            p = os.path.join(d, name + ".m")
            if os.path.isfile(p):
                found = p
                break
        self._lookup[name] = found
        return found

    def resolve_function(self, name, frame):
        """-> FuncDef | ClassDef | python builtin | None, following MATLAB's precedence: nested functions, local functions of
        the running file, files on the path (current folder first), builtins"""
        f = frame.func if frame is not None else None
        g = f
        while g is not None:
            if name in g.nested:
                return g.nested[name]
            g = g.parent
        if f is not None and name in f.unit.funcs:
            return f.unit.funcs[name]
        if name in self.overrides:
            return self.overrides[name]
        p = self.find_file(name)
        if p is not None:
            u = self.load_unit(p)
            if u.kind == "class":
                return u.classdef
            if u.kind == "function":
                return u.main
            return u                       # a script
        return self.builtins.get(name)

    def find_method(self, cls, name):
        if name in cls.methods:
            return cls.methods[name]
        for s in cls.supers:
            sc = self.get_class(s)
            if sc is not None:
                m = self.find_method(sc, name)
                if m is not None:
                    return m
        return None

    def get_class(self, name):
        p = self.find_file(name)
        if p is None:
            return None
        u = self.load_unit(p)
        return u.classdef if u.kind == "class" else None

    def all_props(self, cls):
        props = []
        for s in cls.supers:
            sc = self.get_class(s)
            if sc is not None:
                props += self.all_props(sc)
        return props + list(cls.props)

    # ------------------------------------------------------------------ variables
    def owner(self, frame, name):
        f = frame
        while f.parent is not None and name in f.shared_up and name not in f.vars:
            f = f.parent
        return f

    def getvar(self, frame, name):
        if frame.globals and name in frame.globals:
            return self.globals.get(frame.globals[name], EMPTY)
        f = frame
        while True:
            v = f.vars.get(name)
            if v is not None:
                return v
            if f.parent is not None and name in f.shared_up:
                f = f.parent
                continue
            return None

    def setvar(self, frame, name, val):
        if frame.globals and name in frame.globals:
            self.globals[frame.globals[name]] = val
            return
        if frame.parent is None:
            frame.vars[name] = val
        else:
            self.owner(frame, name).vars[name] = val

    # ------------------------------------------------------------------ calling
    def call_function(self, fn, args, nargout, frame=None, bound=None):
        """call a FuncDef / ClassDef / builtin / script Unit with evaluated args -> list of outputs"""
        if isinstance(fn, FuncDef):
            return self.call_funcdef(fn, args, nargout, frame, bound)
        if isinstance(fn, ClassDef):
            return [self.construct(fn, args)]
        if callable(fn):
            r = fn(self, args, nargout, frame)
            if r is None:
                return []
            return list(r) if isinstance(r, (tuple, list)) else [r]
        if hasattr(fn, "script"):
            # a script runs in the caller's workspace; the local functions at the end of its file are visible while it runs
            prev = frame.func
            if prev is None or prev.unit is not fn:
                frame.func = FuncDef("<script>", [], [], fn)
            try:
                self.exec_block(fn.script, frame)
            finally:
                frame.func = prev
            return []
        raise MatlabError(f"cannot call {fn!r}")

    def parent_frame_for(self, fn, frame):
        """the frame a nested function closes over = the running frame of its lexical parent"""
        if fn.parent is None:
            return None
        f = frame
        while f is not None:
            if f.func is fn.parent:
                return f
            f = f.parent
        raise MatlabError(f"nested function {fn.name} called outside its parent")

    def call_funcdef(self, fn, args, nargout, frame=None, bound=None, pre=None):
        if fn.parent is not None:
            parent = bound if bound is not None else self.parent_frame_for(fn, frame)
        else:
            parent = None
        fr = Frame(fn, parent, len(args), nargout)
        params = fn.params
        if params and params[-1] == "varargin":
            nfix = len(params) - 1
            fr.vars["varargin"] = MCell.row(list(args[nfix:]))
            args = args[:nfix]
        elif len(args) > len(params):
            raise MatlabError(f"{fn.name}: too many input arguments ({len(args)} > {len(params)})")
        for p, a in zip(params, args):
            if p != "~":
                fr.vars[p] = a
        if pre:
            fr.vars.update(pre)
        self.stack.append(fn.name)
        try:
            try:
                self.exec_block(fn.body, fr)
            except _Return:
                pass
        finally:
            self.stack.pop()
        outs = []
        want = max(nargout, 1)
        names = fn.outs
        for i, o in enumerate(names):
            if o == "varargout":
                vo = fr.vars.get("varargout")
                if vo is not None:
                    outs.extend(list(vo.a.reshape(-1, order="F")))
                break
            if i >= want:
                break
            v = fr.vars.get(o)
            if v is None:
                if i < nargout:
                    raise MatlabError(f"{fn.name}: output argument '{o}' was not assigned")
                break
            outs.append(v)
        if nargout > len(outs):
            raise MatlabError(f"{fn.name}: {nargout} outputs requested, {len(outs)} available")
        return outs

    def construct(self, cls, args):
        fields = {}
        for name, default in self.all_props(cls):
            fields[name] = EMPTY if default is None else self.eval(default, Frame(None))
        obj = MObject(cls, fields)
        ctor = cls.methods.get(cls.name)
        if ctor is None:
            return obj
        out = ctor.outs[0]
        r = self.call_funcdef(ctor, args, 1, None, None, pre={out: obj})
        return r[0]

    def call_handle(self, h, args, nargout, frame):
        if h.kind == "anon":
            fr = Frame(h.func, None, len(args), nargout)
            fr.vars.update(h.captured)
            if len(args) > len(h.params):
                raise MatlabError("anonymous function: too many input arguments")
            for p, a in zip(h.params, args):
                if p != "~":
                    fr.vars[p] = a
            return self.eval_multi(h.body, fr, nargout)
        if h.kind == "func":
            fn = h.func
            if isinstance(fn, FuncDef):
                return self.call_funcdef(fn, args, nargout, frame, bound=h.frame)
            return self.call_function(fn, args, nargout, frame)
        if h.kind == "name":
            fn = self.resolve_function(h.name, frame)
            if fn is None:
                raise MatlabError(f"undefined function '{h.name}'")
            return self.call_named(h.name, fn, args, nargout, frame)
        raise MatlabError("bad function handle")

    def call_named(self, name, fn, args, nargout, frame):
        """call by name with method dispatch: an object argument whose class defines ``name`` wins"""
        for a in args:
            if type(a) is MObject:
                m = self.find_method(a.cls, name)
                if m is not None:
                    return self.call_funcdef(m, args, nargout, None)
        if fn is None:
            raise MatlabError(f"undefined function or variable '{name}'")
        return self.call_function(fn, args, nargout, frame)

    # ------------------------------------------------------------------ statements
    def exec_block(self, body, fr):
        for st in body:
            self.exec_stmt(st, fr)

    def exec_stmt(self, st, fr):
        k = st[0]
        if k == "assign":
            self.exec_assign(st, fr)
        elif k == "expr":
            e = st[1]
            vals = self.eval_multi(e, fr, 0)
            if vals:
                v = vals[0]
                if not (e[0] == "id" and self.getvar(fr, e[1]) is not None):
                    self.setvar(fr, "ans", v)
                if st[2]:
                    self.display(e[1] if e[0] == "id" else "ans", v)
        elif k == "if":
            for cond, body in st[1]:
                if self.cond(cond, fr):
                    self.exec_block(body, fr)
                    return
            if st[2] is not None:
                self.exec_block(st[2], fr)
        elif k == "for":
            self.exec_for(st, fr)
        elif k == "while":
            while self.cond(st[1], fr):
                try:
                    self.exec_block(st[2], fr)
                except _Break:
                    break
                except _Continue:
                    continue
        elif k == "switch":
            v = self.eval(st[1], fr)
            for ce, body in st[2]:
                c = self.eval(ce, fr)
                if self.case_match(v, c):
                    self.exec_block(body, fr)
                    return
            if st[3] is not None:
                self.exec_block(st[3], fr)
        elif k == "break":
            raise _Break()
        elif k == "continue":
            raise _Continue()
        elif k == "return":
            raise _Return()
        elif k == "global":
            if fr.globals is None:
                fr.globals = {}                   # local name -> key in the interpreter-wide store
            for n in st[1]:
                fr.globals[n] = n
                self.globals.setdefault(n, EMPTY)
        elif k == "persistent":
            if fr.func is None:
                raise MatlabError("persistent outside a function")
            if fr.globals is None:
                fr.globals = {}
            for n in st[1]:
                key = ("persistent", id(fr.func), n)         # one value per function definition, [] until first assigned
                fr.globals[n] = key
                self.globals.setdefault(key, EMPTY)
        elif k == "try":
            try:
                self.exec_block(st[1], fr)
            except MatlabError as ex:
                if st[2]:
                    self.setvar(fr, st[2], MStruct({"message": str(ex), "identifier": ""}))
                self.exec_block(st[3], fr)
        elif k == "cmd":
            fn = self.resolve_function(st[1], fr)
            if fn is None:
                raise MatlabError(f"undefined command '{st[1]}'")
            self.call_function(fn, list(st[2]), 0, fr)
        else:
            raise MatlabError(f"statement {k} not implemented")

    def case_match(self, v, c):
        if type(c) is MCell:
            return any(self.case_match(v, x) for x in c.a.reshape(-1, order="F"))
        if type(v) is str or type(c) is str:
            return type(v) is str and type(c) is str and v == c
        return truth(binop("==", v, c))

    def cond(self, e, fr):
        # '&' and '|' short-circuit inside if / while conditions when the left operand is a scalar (MATLAB does the same)
        while e[0] == "paren":
            e = e[1]
        if e[0] == "bin" and e[1] in ("&", "|"):
            a = self.eval(e[2], fr)
            if numel(a) == 1:
                ta = truth(a)
                if e[1] == "&" and not ta:
                    return False
                if e[1] == "|" and ta:
                    return True
                return truth(self.eval(e[3], fr))
            return truth(binop(e[1], a, self.eval(e[3], fr)))
        return truth(self.eval(e, fr))

    def exec_for(self, st, fr):
        var, body = st[1], st[3]
        it = self.eval(st[2], fr)
        t = type(it)
        if t is np.ndarray:
            if it.ndim == 2 and it.shape[0] == 1:
                kind = it.dtype.kind
                seq = (float(x) if kind == "f" else simplify(np.array([[x]])) for x in it[0])
            else:
                m = it.reshape(it.shape[0], -1, order="F")
                seq = (simplify(m[:, j:j + 1]) for j in range(m.shape[1]))
        elif t is MCell:
            seq = (MCell(it.a[:, j:j + 1]) for j in range(it.a.shape[1]))
        elif t is str:
            seq = iter(it)
        else:
            seq = iter([it])
        plain = fr.parent is None and fr.globals is None
        for v in seq:
            if plain:
                fr.vars[var] = v
            else:
                self.setvar(fr, var, v)
            try:
                self.exec_block(body, fr)
            except _Break:
                break
            except _Continue:
                continue

    def display(self, name, v):
        self.write(f"{name} =\n{self.fmt_value(v)}\n")

    def fmt_value(self, v):
        if type(v) is np.ndarray:
            return np.array2string(v, precision=5)
        if type(v) is MStruct:
            return "\n".join(f"    {k}: {self.fmt_value(x) if numel(x) < 10 else mclass(x)}" for k, x in v.f.items())
        return f"   {v!r}"

    def write(self, text):
        self.out.write(text)
        if self.on_output is not None:
            self.on_output(text)

    # ------------------------------------------------------------------ assignment
    def exec_assign(self, st, fr):
        lvs, rhs, show = st[1], st[2], st[3]
        if len(lvs) == 1:
            vals = [self.eval(rhs, fr)]
        else:
            vals = self.eval_multi(rhs, fr, len(lvs))
            if len(vals) < len(lvs):
                raise MatlabError(f"{len(lvs)} outputs requested, {len(vals)} returned")
        for lv, v in zip(lvs, vals):
            if lv[0] == "tilde":
                continue
            name, acc = lv[1], lv[2]
            if not acc:
                if fr.parent is None and fr.globals is None:
                    fr.vars[name] = v
                else:
                    self.setvar(fr, name, v)
            elif len(acc) == 1 and acc[0][0] == "()" and len(acc[0][1]) == 1 and type(v) is float:
                cur = self.getvar(fr, name)                # fast path: a(i) = scalar inside the array
                done = False
                if type(cur) is np.ndarray and cur.dtype.kind == "f":
                    a0 = acc[0][1][0]
                    if a0[0] != "colon":
                        sub = self.eval(a0, fr, (cur, 0, 1))
                        if type(sub) is float and 1.0 <= sub <= cur.size and sub == int(sub):
                            new = cur.copy(order="F")
                            new.reshape(-1, order="F")[int(sub) - 1] = v
                            self.setvar(fr, name, new)
                            done = True
                        elif not done:
                            self.setvar(fr, name, index_assign(cur, [sub], v))
                            done = True
                if not done:
                    self.setvar(fr, name, self.assign_into(cur, acc, 0, v, fr))
            else:
                cur = self.getvar(fr, name)
                self.setvar(fr, name, self.assign_into(cur, acc, 0, v, fr))
            if show:
                self.display(name, self.getvar(fr, name))

    def eval_subs(self, args, fr, base, as_cell=False):
        """evaluate subscripts; ``base`` is the value 'end' refers to"""
        n = len(args)
        out = []
        for d, a in enumerate(args):
            if a[0] == "colon":
                out.append(COLON)
            elif a[0] == "str" and a[1] == ":":
                out.append(COLON)
            else:
                vals = self.eval_multi(a, fr, 1, end_ctx=(base, d, n))
                out.extend(vals)
        return out

    def assign_into(self, cur, acc, i, v, fr):
        if i == len(acc):
            return v
        kind, arg = acc[i]
        last = i == len(acc) - 1
        if kind == "." or kind == ".()":
            name = arg if kind == "." else self.eval(arg, fr)
            if type(name) is not str:
                raise MatlabError("dynamic field name must be a character row")
            if cur is None or (type(cur) is np.ndarray and cur.size == 0):
                cur = MStruct()
            if type(cur) is MStruct:
                new = cur.copy()
                new.f[name] = self.assign_into(cur.f.get(name), acc, i + 1, v, fr)
                return new
            if type(cur) is MObject:
                if name not in cur.f:
                    raise MatlabError(f"class {cur.cls.name} has no property '{name}'")
                new = cur.copy()
                old = cur.f.get(name)
                new.f[name] = self.assign_into(None if (type(old) is np.ndarray and old.size == 0 and not last) else old,
                                               acc, i + 1, v, fr)
                return new
            raise MatlabError(f"field assignment to a value of class {mclass(cur)}")
        subs = self.eval_subs(arg, fr, cur if cur is not None else EMPTY)
        if kind == "()":
            if type(cur) in (MStructArr, MStruct) or (cur is None and (not last or type(v) is MStruct)):
                return self._struct_array_assign(cur, subs, acc, i, v, fr, last)
            if type(cur) is MCell:
                if not last:
                    raise MatlabError("chained assignment through c(...) is not supported")
                if type(v) is not MCell:
                    raise MatlabError("c(...) = v needs a cell on the right")
                a = self._cell_grow(cur.a, subs)
                a[self._cell_pos(a, subs)] = v.a.reshape(-1)[0] if v.a.size == 1 else v.a
                return MCell(a)
            if not last:
                raise MatlabError("chained assignment through a(...) is not supported for numeric arrays")
            if type(v) in (MStruct, MObject, MCell, FuncHandle, MStructArr):
                if cur is None or numel(cur) == 0:
                    if type(v) is MStruct:
                        j = int(subs[-1]) - 1
                        items = [MStruct() for _ in range(j + 1)]
                        items[j] = v
                        return v if j == 0 else MStructArr(items)
                    if all(s == 1.0 for s in subs):
                        return v
                raise MatlabError(f"cannot store a {mclass(v)} into a numeric array")
            return index_assign(cur, subs, v)
        if kind == "{}":
            if cur is None or (type(cur) is np.ndarray and cur.size == 0):
                cur = MCell(np.empty((0, 0), dtype=object))
            if type(cur) is not MCell:
                raise MatlabError("brace assignment to a value that is not a cell")
            a = self._cell_grow(cur.a, subs)
            pos = self._cell_pos(a, subs)
            a[pos] = self.assign_into(a[pos] if a[pos] is not None else None, acc, i + 1, v, fr)
            return MCell(a)
        raise MatlabError(f"accessor {kind}")

    def _struct_array_assign(self, cur, subs, acc, i, v, fr, last):
        """S(r, c) = struct  /  S(r, c).field... = v on an m x n struct array (grows; all elements share one field list)"""
        if type(cur) is MStructArr:
            a = cur.a.copy()
        else:
            a = np.empty((1, 1) if cur is not None else (0, 0), dtype=object)
            if cur is not None:
                a[0, 0] = cur
        if len(subs) == 1:
            j = int(subs[0]) - 1
            r, c = (0, j) if a.shape[0] <= 1 else (j % a.shape[0], j // a.shape[0])
        else:
            r, c = int(subs[0]) - 1, int(subs[1]) - 1
            if any(int(s) != 1 for s in subs[2:]):
                raise MatlabError("struct arrays with more than two dimensions are not supported")
        if r < 0 or c < 0:
            raise MatlabError("subscripts must be positive integers")
        if r >= a.shape[0] or c >= a.shape[1]:
            template = list(a.reshape(-1)[0].f) if a.size else []
            b = np.empty((max(r + 1, a.shape[0]), max(c + 1, a.shape[1])), dtype=object)
            for q in range(b.size):
                b.reshape(-1)[q] = MStruct({k: EMPTY for k in template})
            b[:a.shape[0], :a.shape[1]] = a
            a = b
        elem = v if last else self.assign_into(a[r, c], acc, i + 1, v, fr)
        if type(elem) is not MStruct:
            raise MatlabError("only structs can be stored in a struct array")
        other = a.reshape(-1)[0] if (r, c) != (0, 0) else (a.reshape(-1)[-1] if a.size > 1 else None)
        a[r, c] = elem
        if other is not None and list(other.f) != list(elem.f):
            if last and other.f and set(other.f) != set(elem.f):
                raise MatlabError("subscripted assignment between dissimilar structures")
            names = list(elem.f) + [k for k in other.f if k not in elem.f]     # a field added to one element exists ([]) in all
            for q in range(a.size):
                s = a.reshape(-1)[q]
                if list(s.f) != names:
                    a.reshape(-1)[q] = MStruct({k: s.f.get(k, EMPTY) for k in names})
        return a[0, 0] if a.size == 1 else MStructArr(a)

    def _cell_grow(self, a, subs):
        a = a.copy()
        if len(subs) == 1:
            j = int(subs[0])
            if j > a.size:
                if a.size == 0 or a.shape[0] == 1:
                    b = np.empty((1, j), dtype=object)
                    b[0, :a.size] = a.reshape(-1)
                else:
                    b = np.empty((j, 1), dtype=object)
                    b[:a.size, 0] = a.reshape(-1)
                for q in range(b.size):
                    if b.reshape(-1)[q] is None:
                        b.reshape(-1)[q] = EMPTY
                a = b
            return a
        i, j = int(subs[0]), int(subs[1])
        if i > a.shape[0] or j > a.shape[1]:
            b = np.empty((max(i, a.shape[0]), max(j, a.shape[1])), dtype=object)
            for q in range(b.size):
                b.reshape(-1)[q] = EMPTY
            b[:a.shape[0], :a.shape[1]] = a
            a = b
        return a

    def _cell_pos(self, a, subs):
        if len(subs) == 1:
            j = int(subs[0]) - 1
            return np.unravel_index(j, a.shape, order="F")
        return (int(subs[0]) - 1, int(subs[1]) - 1)

    # ------------------------------------------------------------------ expressions
    def eval(self, e, fr, end_ctx=None):
        k = e[0]
        if k == "num":
            return e[1]
        if k == "id":
            if fr.globals is None:                         # own variables win anyway: skip the general lookup
                v = fr.vars.get(e[1])
                if v is not None:
                    return v
            v = self.getvar(fr, e[1])
            if v is not None:
                return v
        elif k == "bin":
            return binop(e[1], self.eval(e[2], fr, end_ctx), self.eval(e[3], fr, end_ctx))
        elif k == "str":
            return e[1]
        elif k == "paren":
            return self.eval(e[1], fr, end_ctx)
        elif k == "index" and e[2] == "()" and e[1][0] == "id":
            v = self.getvar(fr, e[1][1])                   # fast path: numeric variable, plain subscripts
            if type(v) is np.ndarray:
                args = e[3]
                n = len(args)
                subs = []
                for d, a in enumerate(args):
                    ak = a[0]
                    if ak == "colon":
                        subs.append(COLON)
                    elif ak == "index" and a[2] == "{}":
                        subs = None
                        break
                    else:
                        subs.append(self.eval(a, fr, (v, d, n)))
                if subs is not None:
                    return index(v, subs)
        vals = self.eval_multi(e, fr, 1, end_ctx)
        if not vals:
            raise MatlabError("an expression that should give a value gave none")
        return vals[0]

    def eval_multi(self, e, fr, nargout, end_ctx=None):
        """evaluate to a list of values (function calls may give several or none, c{:} a comma-separated list)"""
        k = e[0]
        if k == "num" or k == "str":
            return [e[1]]
        if k == "id":
            name = e[1]
            v = self.getvar(fr, name)
            if v is not None:
                return [v]
            fn = self.resolve_function(name, fr)
            if fn is None:
                raise MatlabError(f"undefined function or variable '{name}'")
            return self.call_named(name, fn, [], nargout, fr)
        if k == "bin":
            return [binop(e[1], self.eval(e[2], fr, end_ctx), self.eval(e[3], fr, end_ctx))]
        if k == "un":
            return [unop(e[1], self.eval(e[2], fr, end_ctx))]
        if k == "post":
            return [transpose(e[1], self.eval(e[2], fr, end_ctx))]
        if k == "paren":
            return [self.eval(e[1], fr, end_ctx)]
        if k == "andand":
            a = self.eval(e[1], fr, end_ctx)
            if not truth(a):
                return [False]
            return [truth(self.eval(e[2], fr, end_ctx))]
        if k == "oror":
            a = self.eval(e[1], fr, end_ctx)
            if truth(a):
                return [True]
            return [truth(self.eval(e[2], fr, end_ctx))]
        if k == "range":
            a = self.eval(e[1], fr, end_ctx)
            s = 1.0 if e[2] is None else self.eval(e[2], fr, end_ctx)
            b = self.eval(e[3], fr, end_ctx)
            for q in (a, s, b):
                if numel(q) != 1:
                    if numel(q) == 0:
                        return [np.zeros((1, 0))]
            a, s, b = (to_arr(q).reshape(-1)[0] for q in (a, s, b))
            return [make_range(a, s, b)]
        if k == "end":
            if end_ctx is None:
                raise MatlabError("'end' outside a subscript")
            base, d, n = end_ctx
            shp = msize(base)
            if n == 1:
                return [float(numel(base))]
            if d < n - 1:
                return [float(shp[d]) if d < len(shp) else 1.0]
            r = 1
            for x in shp[d:]:
                r *= x
            return [float(r)]
        if k == "colon":
            return [":"]
        if k == "matrix":
            rows = []
            for r in e[1]:
                vals = []
                for x in r:
                    vals.extend(self.eval_multi(x, fr, 1, end_ctx))
                rows.append(vals)
            return [concat(rows)]
        if k == "cell":
            rows = []
            for r in e[1]:
                vals = []
                for x in r:
                    vals.extend(self.eval_multi(x, fr, 1, end_ctx))
                rows.append(vals)
            if not rows:
                return [MCell(np.empty((0, 0), dtype=object))]
            w = len(rows[0])
            if any(len(r) != w for r in rows):
                raise MatlabError("cell rows of different length")
            a = np.empty((len(rows), w), dtype=object)
            for i, r in enumerate(rows):
                for j, x in enumerate(r):
                    a[i, j] = x
            return [MCell(a)]
        if k == "anon":
            captured = {}
            f = fr
            seen = set()
            while f is not None:                      # snapshot of every variable visible here
                for n, v in f.vars.items():
                    if n not in seen:
                        captured[n] = v
                        seen.add(n)
                f = f.parent
            if fr.globals:
                for n, key in fr.globals.items():
                    captured[n] = self.globals.get(key, EMPTY)
            return [FuncHandle("anon", func=fr.func, frame=fr, params=e[1], body=e[2], captured=captured)]
        if k == "fhandle":
            name = e[1]
            fn = self.resolve_function(name, fr)
            if isinstance(fn, FuncDef) and fn.parent is not None:
                return [FuncHandle("func", name=name, func=fn, frame=self.parent_frame_for(fn, fr))]
            if isinstance(fn, FuncDef):
                return [FuncHandle("func", name=name, func=fn)]
            return [FuncHandle("name", name=name)]
        if k == "field" or k == "dynfield":
            return self.eval_field(e, fr, nargout, end_ctx)
        if k == "index":
            return self.eval_index(e, fr, nargout, end_ctx)
        raise MatlabError(f"expression {k} not implemented")

    def field_name(self, e, fr):
        if e[0] == "field":
            return e[2]
        n = self.eval(e[2], fr)
        if type(n) is not str:
            raise MatlabError("dynamic field name must be a character row")
        return n

    def eval_field(self, e, fr, nargout, end_ctx):
        base = self.eval(e[1], fr, end_ctx)
        name = self.field_name(e, fr)
        return self.get_field(base, name, [], False, nargout, fr)

    def get_field(self, base, name, args, has_args, nargout, fr):
        t = type(base)
        if t is MStruct:
            if name not in base.f:
                raise MatlabError(f"reference to non-existent field '{name}'")
            v = base.f[name]
            return [self.index_value(v, args, nargout, fr)] if has_args else [v]
        if t is MObject:
            if name in base.f:
                v = base.f[name]
                return [self.index_value(v, args, nargout, fr)] if has_args else [v]
            m = self.find_method(base.cls, name)
            if m is None:
                raise MatlabError(f"class {base.cls.name} has no property or method '{name}'")
            return self.call_funcdef(m, [base] + list(args), nargout, None)
        if t is MStructArr:
            return [it.f[name] for it in base.items]
        raise MatlabError(f"dot reference into a value of class {mclass(base)}")

    def index_value(self, v, subs, nargout, fr):
        t = type(v)
        if t is FuncHandle:
            r = self.call_handle(v, subs, max(nargout, 1), fr)
            return r[0] if r else EMPTY
        if t is MCell:
            return self.cell_paren(v, subs)
        if t is MStructArr:
            if len(subs) == 2 and type(subs[0]) is float and type(subs[1]) is float:
                r, c = int(subs[0]) - 1, int(subs[1]) - 1
                if not (0 <= r < v.a.shape[0] and 0 <= c < v.a.shape[1]):
                    raise MatlabError("index exceeds the dimensions of the struct array")
                return v.a[r, c]
            lin = np.arange(1.0, v.a.size + 1).reshape(v.a.shape, order="F")
            pos = to_arr(index(lin, subs))
            flat = v.a.reshape(-1, order="F")
            out = np.empty(pos.shape, dtype=object)
            for q, pp in enumerate(pos.reshape(-1, order="F")):
                out.reshape(-1, order="F")[q] = flat[int(pp) - 1]
            return out.reshape(-1)[0] if out.size == 1 else MStructArr(out)
        if t in (MStruct, MObject):
            if all((s is COLON) or (type(s) is float and s == 1.0) for s in subs):
                return v
            raise MatlabError("index exceeds the dimensions of a 1x1 struct")
        return index(v, subs)

    def cell_paren(self, c, subs):
        if len(subs) == 1:
            flat = c.a.reshape(-1, order="F")
            if subs[0] is COLON:
                return MCell(flat.reshape(-1, 1))
            pos = to_arr(subs[0]).reshape(-1).astype(int) - 1
            return MCell(flat[pos].reshape(1, -1) if c.a.shape[0] <= 1 else flat[pos].reshape(-1, 1))
        pos = []
        for d, s in enumerate(subs):
            pos.append(np.arange(c.a.shape[d]) if s is COLON else to_arr(s).reshape(-1).astype(int) - 1)
        return MCell(c.a[np.ix_(*pos)])

    def cell_brace(self, c, subs):
        if type(c) is not MCell:
            raise MatlabError(f"brace indexing into a value of class {mclass(c)}")
        if len(subs) == 1:
            flat = c.a.reshape(-1, order="F")
            if subs[0] is COLON:
                return list(flat)
            pos = to_arr(subs[0]).reshape(-1, order="F").astype(int) - 1
            if pos.size and pos.max() >= flat.size:
                raise MatlabError("index exceeds the number of cell elements")
            return [flat[p] for p in pos]
        pos = []
        for d, s in enumerate(subs):
            pos.append(np.arange(c.a.shape[d]) if s is COLON else to_arr(s).reshape(-1).astype(int) - 1)
        return list(c.a[np.ix_(*pos)].reshape(-1, order="F"))

    def eval_index(self, e, fr, nargout, end_ctx):
        base_e, kind, args = e[1], e[2], e[3]
        if kind == "{}":
            base = self.eval(base_e, fr, end_ctx)
            subs = self.eval_subs(args, fr, base)
            return self.cell_brace(base, subs)
        # ---- ( )
        if base_e[0] == "id":
            name = base_e[1]
            v = self.getvar(fr, name)
            if v is not None:
                subs = self.eval_subs(args, fr, v)
                if type(v) is FuncHandle:
                    return self.call_handle(v, subs, nargout, fr)
                return [self.index_value(v, subs, nargout, fr)]
            fn = self.resolve_function(name, fr)
            argv = self.eval_args(args, fr)
            return self.call_named(name, fn, argv, nargout, fr)
        if base_e[0] in ("field", "dynfield"):
            owner = self.eval(base_e[1], fr, end_ctx)
            name = self.field_name(base_e, fr)
            if type(owner) is MObject and name not in owner.f:
                m = self.find_method(owner.cls, name)
                if m is None:
                    raise MatlabError(f"class {owner.cls.name} has no property or method '{name}'")
                return self.call_funcdef(m, [owner] + self.eval_args(args, fr), nargout, None)
            if type(owner) in (MStruct, MObject):
                if name not in owner.f:
                    raise MatlabError(f"reference to non-existent field '{name}'")
                v = owner.f[name]
                subs = self.eval_subs(args, fr, v)
                if type(v) is FuncHandle:
                    return self.call_handle(v, subs, nargout, fr)
                return [self.index_value(v, subs, nargout, fr)]
            raise MatlabError(f"dot reference into a value of class {mclass(owner)}")
        base = self.eval(base_e, fr, end_ctx)
        subs = self.eval_subs(args, fr, base)
        if type(base) is FuncHandle:
            return self.call_handle(base, subs, nargout, fr)
        return [self.index_value(base, subs, nargout, fr)]

    def eval_args(self, args, fr):
        out = []
        for a in args:
            if a[0] == "colon":
                out.append(":")
            else:
                out.extend(self.eval_multi(a, fr, 1))
        return out

    # ------------------------------------------------------------------ entry points
    def call(self, name, *args, nargout=1):
        """call a function on the path (or a builtin) from Python; args are converted with ``from_py``"""
        fr = Frame(None)
        fn = self.resolve_function(name, fr)
        if fn is None:
            raise MatlabError(f"undefined function '{name}'")
        r = self.call_named(name, fn, [from_py(a) for a in args], nargout, fr)
        return r[0] if nargout == 1 and r else r

    def run(self, src, frame=None):
        """run statements given as text in a (new) base workspace; returns the frame"""
        fr = frame or Frame(None)
        u = parse_source(src, "<run>")
        try:
            self.exec_block(u.script, fr)
        except _Return:
            pass
        return fr

    def close_all(self):
        for f in self.files.values():
            try:
                f.close()
            except Exception:
                pass
        self.files.clear()


def from_py(v):
    if isinstance(v, (bool, np.bool_)):
        return bool(v)
    if isinstance(v, (int, float, np.integer, np.floating)):
        return float(v)
    if isinstance(v, complex):
        return v
    if isinstance(v, np.ndarray):
        a = v
        if a.dtype.kind in "iu":
            a = a.astype(np.float64)
        if a.ndim == 0:
            return simplify(a.reshape(1, 1))
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        return simplify(np.asfortranarray(a))
    if isinstance(v, dict):
        return MStruct({k: from_py(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return MCell.row([from_py(x) for x in v])
    return v


def to_py(v):
    if type(v) is MStruct:
        return {k: to_py(x) for k, x in v.f.items()}
    if type(v) is MCell:
        return [to_py(x) for x in v.a.reshape(-1, order="F")]
    return v
