"""minimat.parser -- recursive-descent parser for the MATLAB subset (TEST INFRASTRUCTURE ONLY).

AST nodes are plain tuples, first element = node kind.

expressions  ('num', v) ('str', s) ('id', name) ('colon',) ('end',) ('bin', op, a, b) ('un', op, a) ('post', op, a)
             ('andand', a, b) ('oror', a, b) ('range', a, step|None, b) ('index', base, '()'|'{}', [args])
             ('field', base, name) ('dynfield', base, expr) ('matrix', [[e..]..]) ('cell', [[e..]..])
             ('anon', [params], body) ('fhandle', name)
statements   ('expr', e, show, line) ('assign', [lvalues], rhs, show, line) ('if', [(cond, body)..], else|None)
             ('for', var, e, body) ('while', cond, body) ('switch', e, [(case_e, body)..], otherwise|None)
             ('break',) ('continue',) ('return',) ('global', [names]) ('persistent', [names])
             ('try', body, var|None, catch_body) ('cmd', name, [words])
lvalues      ('lv', name, [accessor..]) with accessors ('()', args) ('{}', args) ('.', name) ('.()', expr); ('tilde',)
"""
from .lexer import tokenize

BLOCK_OPENERS = {"if", "for", "while", "switch", "try", "parfor", "classdef", "properties", "methods", "events", "enumeration"}


class ParseError(Exception):
    pass


class FuncDef:
    def __init__(self, name, params, outs, unit, parent=None):
        self.name, self.params, self.outs, self.unit, self.parent = name, params, outs, unit, parent
        self.body = []
        self.nested = {}
        self.names = set(params) | set(outs)          # every identifier the body mentions (nested bodies excluded)
        self.locals_ = set(params) | set(outs)
        self.line = 0

    def ancestors_names(self):
        s, p = set(), self.parent
        while p is not None:
            s |= p.names
            p = p.parent
        return s

    def __repr__(self):
        return f"<function {self.name} in {self.unit.path}>"


class ClassDef:
    def __init__(self, name, supers, unit):
        self.name, self.supers, self.unit = name, supers, unit
        self.props = []                                # (name, default expr | None)
        self.methods = {}
        self.static = set()


class Unit:
    """one parsed .m file: a script, a function file (main + local functions) or a classdef"""

    def __init__(self, path):
        self.path = path
        self.kind = "script"
        self.script = []
        self.funcs = {}
        self.main = None
        self.classdef = None


class Parser:
    def __init__(self, src, path="<string>"):
        self.path = path
        self.toks = tokenize(src, path)
        self.pos = 0
        self.matrix = [False]          # top = directly inside [ ] or { } (white space separates elements)
        self.fn = [None]               # FuncDef whose body is being parsed (collects names)
        n_end = sum(1 for t in self.toks if t.kind == "kw" and t.val == "end" and not t.idx_end)
        n_open = sum(1 for t in self.toks if t.kind == "kw" and t.val in BLOCK_OPENERS)
        n_fun = sum(1 for t in self.toks if t.kind == "kw" and t.val == "function")
        self.end_mode = n_fun > 0 and n_end >= n_open + n_fun
        self.unit = Unit(path)

    # ------------------------------------------------------------------ token helpers
    @property
    def tok(self):
        return self.toks[self.pos]

    def peek(self, k=1):
        return self.toks[min(self.pos + k, len(self.toks) - 1)]

    def adv(self):
        t = self.toks[self.pos]
        self.pos += 1
        return t

    def is_op(self, *vals):
        t = self.tok
        return t.kind == "op" and t.val in vals

    def is_kw(self, *vals):
        t = self.tok
        return t.kind == "kw" and t.val in vals and not t.idx_end

    def expect_op(self, v):
        if not self.is_op(v):
            self.err(f"expected '{v}', found {self.tok.val!r}")
        return self.adv()

    def err(self, msg):
        raise ParseError(f"{self.path}:{self.tok.line}: {msg}")

    def note(self, name):
        f = self.fn[-1]
        if f is not None:
            f.names.add(name)

    def skip_seps(self):
        while self.tok.kind == "nl" or self.is_op(";", ","):
            self.adv()

    # ------------------------------------------------------------------ file level
    def parse_file(self):
        self.skip_seps()
        u = self.unit
        if self.is_kw("classdef"):
            u.kind = "class"
            u.classdef = self.parse_classdef()
            self.skip_seps()
            while self.is_kw("function"):
                f = self.parse_function(None)
                u.funcs[f.name] = f
                self.skip_seps()
        elif self.is_kw("function"):
            u.kind = "function"
            while self.is_kw("function"):
                f = self.parse_function(None)
                if u.main is None:
                    u.main = f
                u.funcs.setdefault(f.name, f)
                self.skip_seps()
        else:
            u.script = self.parse_block(())
            self.skip_seps()
            while self.is_kw("function"):            # local functions at the end of a script
                f = self.parse_function(None)
                u.funcs[f.name] = f
                self.skip_seps()
        if self.tok.kind != "eof":
            self.err(f"unexpected {self.tok.val!r} at file level")
        return u

    def parse_function(self, parent):
        line = self.tok.line
        self.adv()                                    # 'function'
        outs, name = [], None
        if self.is_op("["):
            self.adv()
            while not self.is_op("]"):
                if self.is_op(","):
                    self.adv()
                    continue
                if self.is_op("~"):
                    self.adv()
                    outs.append("~")
                    continue
                outs.append(self.adv().val)
            self.adv()
            self.expect_op("=")
            name = self.parse_dotted_name()
        else:
            first = self.parse_dotted_name()
            if self.is_op("="):
                self.adv()
                outs = [first]
                name = self.parse_dotted_name()
            else:
                name = first
        params = []
        if self.is_op("("):
            self.adv()
            while not self.is_op(")"):
                if self.is_op(","):
                    self.adv()
                    continue
                if self.is_op("~"):
                    self.adv()
                    params.append("~")
                    continue
                params.append(self.adv().val)
            self.adv()
        f = FuncDef(name, params, outs, self.unit, parent)
        f.line = line
        self.fn.append(f)
        try:
            f.body = self.parse_block(("end", "function"), func=f)
        finally:
            self.fn.pop()
        if self.is_kw("end"):
            if not self.end_mode:
                self.err("'end' closes nothing")
            self.adv()
        return f

    def parse_dotted_name(self):
        if self.tok.kind != "id":
            self.err(f"expected a name, found {self.tok.val!r}")
        name = self.adv().val
        while self.is_op(".") and self.peek().kind == "id":
            self.adv()
            name += "." + self.adv().val
        return name

    def parse_classdef(self):
        self.adv()                                    # classdef
        if self.is_op("("):
            self.skip_parens()
        name = self.adv().val
        supers = []
        if self.is_op("<"):
            self.adv()
            supers.append(self.parse_dotted_name())
            while self.is_op("&"):
                self.adv()
                supers.append(self.parse_dotted_name())
        cd = ClassDef(name, supers, self.unit)
        self.skip_seps()
        while not self.is_kw("end"):
            if self.is_kw("properties"):
                self.adv()
                if self.is_op("("):
                    self.skip_parens()
                self.skip_seps()
                while not self.is_kw("end"):
                    pname = self.adv().val
                    default = None
                    if self.is_op("="):
                        self.adv()
                        default = self.parse_expr()
                    cd.props.append((pname, default))
                    self.skip_seps()
                self.adv()
            elif self.is_kw("methods"):
                self.adv()
                attrs = ""
                if self.is_op("("):
                    attrs = self.skip_parens()
                self.skip_seps()
                if "Abstract" in attrs:
                    while not self.is_kw("end"):      # signatures only
                        self.adv()
                    self.adv()
                else:
                    while not self.is_kw("end"):
                        if not self.is_kw("function"):
                            self.err("expected 'function' inside a methods block")
                        f = self.parse_function(None)
                        cd.methods[f.name] = f
                        if "Static" in attrs:
                            cd.static.add(f.name)
                        self.skip_seps()
                    self.adv()
            elif self.is_kw("events", "enumeration"):
                while not self.is_kw("end"):
                    self.adv()
                self.adv()
            else:
                self.err(f"unexpected {self.tok.val!r} in classdef")
            self.skip_seps()
        self.adv()
        return cd

    def skip_parens(self):
        depth, words = 0, []
        while True:
            t = self.adv()
            if t.kind == "op" and t.val == "(":
                depth += 1
            elif t.kind == "op" and t.val == ")":
                depth -= 1
                if depth == 0:
                    return " ".join(words)
            elif t.kind == "eof":
                self.err("unbalanced parentheses")
            else:
                words.append(str(t.val))

    # ------------------------------------------------------------------ statements
    def parse_block(self, stops, func=None):
        """statements until a keyword of ``stops`` (not consumed) or end of file; ``func`` = the function whose body this is
        (a nested 'function' keyword is parsed into it when the file closes its functions with 'end')"""
        out = []
        while True:
            self.skip_seps()
            t = self.tok
            if t.kind == "eof":
                return out
            if t.kind == "kw" and not t.idx_end:
                if t.val == "function":
                    if func is not None and self.end_mode:
                        g = self.parse_function(func)
                        func.nested[g.name] = g
                        continue
                    return out
                if t.val in stops:
                    return out
            out.append(self.parse_statement())

    def end_of_statement(self):
        """consume the statement terminator; returns True when the result is to be displayed (no ';')"""
        if self.is_op(";"):
            self.adv()
            return False
        if self.is_op(",") or self.tok.kind == "nl":
            self.adv()
            return True
        if self.tok.kind == "eof" or self.tok.kind == "kw":
            return True
        self.err(f"unexpected {self.tok.val!r} after a statement")

    def parse_statement(self):
        t = self.tok
        line = t.line
        if t.kind == "kw":
            kw = t.val
            if kw == "if":
                self.adv()
                clauses, other = [], None
                cond = self.parse_expr()
                body = self.parse_block(("elseif", "else", "end"))
                clauses.append((cond, body))
                while True:
                    if self.is_kw("elseif"):
                        self.adv()
                        cond = self.parse_expr()
                        clauses.append((cond, self.parse_block(("elseif", "else", "end"))))
                    elif self.is_kw("else"):
                        self.adv()
                        other = self.parse_block(("end",))
                    elif self.is_kw("end"):
                        self.adv()
                        break
                    else:
                        self.err("unterminated if")
                return ("if", clauses, other)
            if kw in ("for", "parfor"):
                self.adv()
                paren = self.is_op("(")
                if paren:
                    self.adv()
                var = self.adv().val
                self.note(var)
                self.expect_op("=")
                e = self.parse_expr()
                if paren:
                    self.expect_op(")")
                body = self.parse_block(("end",))
                self.adv()
                return ("for", var, e, body)
            if kw == "while":
                self.adv()
                cond = self.parse_expr()
                body = self.parse_block(("end",))
                self.adv()
                return ("while", cond, body)
            if kw == "switch":
                self.adv()
                e = self.parse_expr()
                self.skip_seps()
                cases, other = [], None
                while not self.is_kw("end"):
                    if self.is_kw("case"):
                        self.adv()
                        ce = self.parse_expr()
                        cases.append((ce, self.parse_block(("case", "otherwise", "end"))))
                    elif self.is_kw("otherwise"):
                        self.adv()
                        other = self.parse_block(("case", "otherwise", "end"))
                    else:
                        self.err("expected case / otherwise / end")
                self.adv()
                return ("switch", e, cases, other)
            if kw == "try":
                self.adv()
                body = self.parse_block(("catch", "end"))
                var, cbody = None, []
                if self.is_kw("catch"):
                    self.adv()
                    if self.tok.kind == "id" and self.peek().kind == "nl":
                        var = self.adv().val
                        self.note(var)
                    cbody = self.parse_block(("end",))
                self.adv()
                return ("try", body, var, cbody)
            if kw in ("break", "continue", "return"):
                self.adv()
                return (kw,)
            if kw in ("global", "persistent"):
                self.adv()
                names = []
                while self.tok.kind == "id":
                    names.append(self.adv().val)
                    self.note(names[-1])
                return (kw, names)
            self.err(f"unexpected keyword {kw!r}")
        if t.kind == "id" and self.peek().kind == "cmd":
            name = self.adv().val
            words = self.adv().val
            return ("cmd", name, words)
        # multi-assignment  [a, b] = f(...)
        if self.is_op("["):
            j, depth = self.pos, 0
            while True:
                tt = self.toks[j]
                if tt.kind == "op" and tt.val in "([{":
                    depth += 1
                elif tt.kind == "op" and tt.val in ")]}":
                    depth -= 1
                    if depth == 0:
                        break
                elif tt.kind == "eof":
                    break
                j += 1
            nxt = self.toks[j + 1] if j + 1 < len(self.toks) else None
            if nxt is not None and nxt.kind == "op" and nxt.val == "=":
                self.adv()
                lvs = []
                self.matrix.append(True)
                try:
                    while not self.is_op("]"):
                        if self.is_op(","):
                            self.adv()
                            continue
                        if self.is_op("~"):
                            self.adv()
                            lvs.append(("tilde",))
                            continue
                        lvs.append(self.to_lvalue(self.parse_postfix()))
                finally:
                    self.matrix.pop()
                self.adv()
                self.expect_op("=")
                rhs = self.parse_expr()
                return ("assign", lvs, rhs, self.end_of_statement(), line)
        e = self.parse_expr()
        if self.is_op("="):
            self.adv()
            lv = self.to_lvalue(e)
            rhs = self.parse_expr()
            return ("assign", [lv], rhs, self.end_of_statement(), line)
        return ("expr", e, self.end_of_statement(), line)

    def to_lvalue(self, e):
        acc = []
        while True:
            k = e[0]
            if k == "id":
                return ("lv", e[1], list(reversed(acc)))
            if k == "index":
                acc.append((e[2], e[3]))
                e = e[1]
            elif k == "field":
                acc.append((".", e[2]))
                e = e[1]
            elif k == "dynfield":
                acc.append((".()", e[2]))
                e = e[1]
            else:
                self.err("cannot assign to this expression")

    # ------------------------------------------------------------------ expressions
    def parse_expr(self):
        return self.parse_oror()

    def parse_oror(self):
        a = self.parse_andand()
        while self.is_op("||"):
            self.adv()
            a = ("oror", a, self.parse_andand())
        return a

    def parse_andand(self):
        a = self.parse_or()
        while self.is_op("&&"):
            self.adv()
            a = ("andand", a, self.parse_or())
        return a

    def parse_or(self):
        a = self.parse_and()
        while self.is_op("|"):
            self.adv()
            a = ("bin", "|", a, self.parse_and())
        return a

    def parse_and(self):
        a = self.parse_cmp()
        while self.is_op("&"):
            self.adv()
            a = ("bin", "&", a, self.parse_cmp())
        return a

    def parse_cmp(self):
        a = self.parse_range()
        while self.is_op("==", "~=", "<", "<=", ">", ">="):
            op = self.adv().val
            a = ("bin", op, a, self.parse_range())
        return a

    def parse_range(self):
        a = self.parse_add()
        if self.is_op(":") and not self.range_colon_is_bare():
            self.adv()
            b = self.parse_add()
            if self.is_op(":") and not self.range_colon_is_bare():
                self.adv()
                c = self.parse_add()
                return ("range", a, b, c)
            return ("range", a, None, b)
        return a

    def range_colon_is_bare(self):
        nxt = self.peek()
        return nxt.kind == "op" and nxt.val in (",", ")", "}", "]", ";") or nxt.kind in ("nl", "eof")

    def binary_breaks_element(self):
        """inside [ ] / { }:  'a -b' and 'a +b' are two elements, 'a - b' and 'a-b' one"""
        if not self.matrix[-1]:
            return False
        t = self.tok
        return t.sp and not self.peek().sp

    def parse_add(self):
        a = self.parse_mul()
        while self.is_op("+", "-") and not self.binary_breaks_element():
            op = self.adv().val
            a = ("bin", op, a, self.parse_mul())
        return a

    def parse_mul(self):
        a = self.parse_unary()
        while self.is_op("*", "/", ".*", "./", "\\", ".\\"):
            op = self.adv().val
            a = ("bin", op, a, self.parse_unary())
        return a

    def parse_unary(self):
        if self.is_op("+", "-", "~"):
            op = self.adv().val
            return ("un", op, self.parse_unary())
        return self.parse_power()

    def parse_power(self):
        a = self.parse_postfix()
        while self.is_op("^", ".^"):
            op = self.adv().val
            if self.is_op("+", "-", "~"):
                u = self.adv().val
                b = ("un", u, self.parse_power_operand())
            else:
                b = self.parse_postfix()
            a = ("bin", op, a, b)
        return a

    def parse_power_operand(self):
        if self.is_op("+", "-", "~"):
            u = self.adv().val
            return ("un", u, self.parse_power_operand())
        return self.parse_postfix()

    def parse_args(self, close):
        """arguments of a ( ) or { } subscript / call; a bare ':' is ('colon',)"""
        args = []
        self.matrix.append(False)
        try:
            while not self.is_op(close):
                if self.is_op(","):
                    self.adv()
                    continue
                if self.is_op(":") and self.range_colon_is_bare():
                    self.adv()
                    args.append(("colon",))
                    continue
                if close == ")" and self.tok.kind == "id" and self.peek().kind == "op" and self.peek().val == "=":
                    args.append(("str", self.adv().val))          # name=value argument (R2021a): f(..., FontSize=14)
                    self.adv()
                args.append(self.parse_expr())
        finally:
            self.matrix.pop()
        self.adv()
        return args

    def parse_postfix(self):
        a = self.parse_primary()
        while True:
            t = self.tok
            if t.kind != "op":
                break
            if t.val == "(":
                if self.matrix[-1] and t.sp:
                    break
                self.adv()
                a = ("index", a, "()", self.parse_args(")"))
            elif t.val == "{":
                if self.matrix[-1] and t.sp:
                    break
                self.adv()
                a = ("index", a, "{}", self.parse_args("}"))
            elif t.val == ".":
                nxt = self.peek()
                if nxt.kind in ("id", "kw") and not (self.matrix[-1] and t.sp):
                    self.adv()
                    a = ("field", a, self.adv().val)
                elif nxt.kind == "op" and nxt.val == "(":
                    self.adv()
                    self.adv()
                    self.matrix.append(False)
                    try:
                        e = self.parse_expr()
                    finally:
                        self.matrix.pop()
                    self.expect_op(")")
                    a = ("dynfield", a, e)
                else:
                    break
            elif t.val in ("'", ".'"):
                self.adv()
                a = ("post", t.val, a)
            else:
                break
        return a

    def parse_primary(self):
        t = self.tok
        if t.kind == "num":
            self.adv()
            s = t.val
            if s[-1] in "ij":
                return ("num", complex(0.0, float(s[:-1].replace("d", "e").replace("D", "e"))))
            return ("num", float(s.replace("d", "e").replace("D", "e")))
        if t.kind == "str":
            self.adv()
            return ("str", t.val)
        if t.kind == "id":
            self.adv()
            self.note(t.val)
            return ("id", t.val)
        if t.kind == "kw" and t.val == "end" and t.idx_end:
            self.adv()
            return ("end",)
        if t.kind == "op":
            if t.val == "(":
                self.adv()
                self.matrix.append(False)
                try:
                    e = self.parse_expr()
                finally:
                    self.matrix.pop()
                self.expect_op(")")
                return ("paren", e)
            if t.val == "[":
                self.adv()
                return ("matrix", self.parse_rows("]"))
            if t.val == "{":
                self.adv()
                return ("cell", self.parse_rows("}"))
            if t.val == "@":
                self.adv()
                if self.is_op("("):
                    self.adv()
                    params = []
                    while not self.is_op(")"):
                        if self.is_op(","):
                            self.adv()
                            continue
                        if self.is_op("~"):
                            self.adv()
                            params.append("~")
                            continue
                        params.append(self.adv().val)
                    self.adv()
                    self.matrix.append(False)
                    try:
                        body = self.parse_expr()
                    finally:
                        self.matrix.pop()
                    return ("anon", params, body)
                return ("fhandle", self.parse_dotted_name())
            if t.val == ":":
                self.adv()
                return ("colon",)
        self.err(f"unexpected {t.val!r} in an expression")

    def parse_rows(self, close):
        rows, row = [], []
        self.matrix.append(True)
        try:
            while True:
                if self.is_op(close):
                    break
                if self.is_op(";") or self.tok.kind == "nl":
                    self.adv()
                    if row:
                        rows.append(row)
                    row = []
                    continue
                if self.is_op(","):
                    self.adv()
                    continue
                if self.tok.kind == "eof":
                    self.err("unterminated [ or {")
                row.append(self.parse_expr())
        finally:
            self.matrix.pop()
        self.adv()
        if row:
            rows.append(row)
        return rows


def parse_source(src, path="<string>"):
    return Parser(src, path).parse_file()
