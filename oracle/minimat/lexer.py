"""minimat.lexer -- tokens of the MATLAB subset (TEST INFRASTRUCTURE ONLY, see oracle/minimat/__init__.py)."""
import re

KEYWORDS = {"function", "end", "if", "elseif", "else", "for", "while", "switch", "case", "otherwise", "return", "break",
            "continue", "classdef", "properties", "methods", "global", "persistent", "try", "catch", "parfor", "events",
            "enumeration"}
OPS = ["==", "~=", "!=", "<=", ">=", "&&", "||", ".*", "./", ".\\", ".^", ".'", "+", "-", "*", "/", "\\", "^", "<", ">", "&", "|",
       "~", "!", "=", "(", ")", "[", "]", "{", "}", ",", ";", ":", ".", "@"]
NUM_RE = re.compile(r"(?:\d+(?:\.(?![*/\\^'])\d*)?|\.\d+)(?:[eEdD][+-]?\d+)?[ij]?")
ID_RE = re.compile(r"[A-Za-z_][A-Za-z0-9_]*")


class Tok:
    __slots__ = ("kind", "val", "sp", "line", "idx_end")

    def __init__(self, kind, val, sp, line, idx_end=False):
        self.kind, self.val, self.sp, self.line, self.idx_end = kind, val, sp, line, idx_end

    def __repr__(self):
        return f"Tok({self.kind},{self.val!r},sp={self.sp},l{self.line})"


class LexError(Exception):
    pass


def _strip_block_comments(src):
    """lines between a line holding only %{ and a line holding only %} are blanked (newlines kept for line numbers)"""
    if "%{" not in src:
        return src
    out, depth = [], 0
    for ln in src.split("\n"):
        s = ln.strip()
        if s == "%{":
            depth += 1
            out.append("")
        elif s == "%}" and depth:
            depth -= 1
            out.append("")
        else:
            out.append("" if depth else ln)
    return "\n".join(out)


def tokenize(src, fname="<string>"):
    """Token kinds: num, str, id, kw, op, nl (statement end by newline), cmd (words of a command-syntax statement), eof.
    ``sp`` = white space stood before the token (matters inside [ ] and { }); newlines inside [ ] / { } become ';'."""
    src = _strip_block_comments(src)
    toks = []
    stack = []                     # open brackets
    i, n, line = 0, len(src), 1
    sp = False

    def stmt_start():
        if not toks:
            return True
        t = toks[-1]
        return t.kind == "nl" or (t.kind == "op" and t.val in (";", ",") and not stack)

    while i < n:
        c = src[i]
        if c in " \t":
            i += 1
            sp = True
            continue
        if src.startswith("...", i):
            j = src.find("\n", i)
            i = n if j < 0 else j + 1
            line += 1
            sp = True
            continue
        if c == "%" or c == "#":
            le = src.find("\n", i)
            le = n if le < 0 else le
            i = le
            continue
        if c == "\n" or c == "\r":
            if c == "\r" and i + 1 < n and src[i + 1] == "\n":
                i += 1
            if stack and stack[-1] in "[{":
                toks.append(Tok("op", ";", sp, line))
            elif not stack:
                toks.append(Tok("nl", "\n", sp, line))
            i += 1
            line += 1
            sp = False
            continue
        if c == "'":
            prev = toks[-1] if toks else None
            in_matrix = bool(stack) and stack[-1] in "[{"
            transpose = (prev is not None and not (in_matrix and sp) and
                         (prev.kind in ("num", "id") or (prev.kind == "op" and prev.val in (")", "]", "}", "'", ".'"))
                          or (prev.kind == "kw" and prev.val == "end" and prev.idx_end)))
            if transpose:
                toks.append(Tok("op", "'", sp, line))
                i += 1
                sp = False
                continue
            j = i + 1
            out = []
            while True:
                if j >= n or src[j] == "\n":
                    raise LexError(f"{fname}:{line}: unterminated string")
                if src[j] == "'":
                    if j + 1 < n and src[j + 1] == "'":
                        out.append("'")
                        j += 2
                        continue
                    break
                out.append(src[j])
                j += 1
            toks.append(Tok("str", "".join(out), sp, line))
            i = j + 1
            sp = False
            continue
        if c == '"':
            j = i + 1
            out = []
            while True:
                if j >= n or src[j] == "\n":
                    raise LexError(f"{fname}:{line}: unterminated string")
                if src[j] == '"':
                    if j + 1 < n and src[j + 1] == '"':
                        out.append('"')
                        j += 2
                        continue
                    break
                out.append(src[j])
                j += 1
            toks.append(Tok("str", "".join(out), sp, line))
            i = j + 1
            sp = False
            continue
        m = NUM_RE.match(src, i)
        if m and (c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit())):
            toks.append(Tok("num", m.group(0), sp, line))
            i = m.end()
            sp = False
            continue
        m = ID_RE.match(src, i)
        if m:
            word = m.group(0)
            after_dot = bool(toks) and toks[-1].kind == "op" and toks[-1].val == "." and not sp
            if word in KEYWORDS and not after_dot:
                toks.append(Tok("kw", word, sp, line, idx_end=(word == "end" and bool(stack))))
                i = m.end()
                sp = False
                continue
            if stmt_start():
                # command syntax ("addpath ./rsw/", "hold on", "subplot 211", "clear P"): a name at the start of a statement,
                # white space, then something that cannot continue an expression -- a word, a number, a quote, or one of . / - ~
                # glued to what follows (MATLAB's own rule for names that are not variables)
                j = m.end()
                k = j
                while k < n and src[k] in " \t":
                    k += 1
                if k > j and k < n:
                    c2 = src[k]
                    nxt = src[k + 1] if k + 1 < n else "\n"
                    if c2.isalnum() or c2 in "_'\"":
                        is_cmd = True
                    elif c2 in "./-~":
                        oplen = 2 if (c2 == "." and nxt in "*/^\\") else 1       # "a ./b" and "addpath ./rsw/" alike: operator glued
                        after = src[k + oplen] if k + oplen < n else "\n"          # to what follows, white space before it
                        is_cmd = after not in " \t=\n\r;,"
                    else:
                        is_cmd = False
                    if is_cmd:
                        e = k
                        while e < n and src[e] not in "\n\r;,%":
                            e += 1
                        toks.append(Tok("id", word, sp, line))
                        toks.append(Tok("cmd", [w.strip("'\"") for w in src[k:e].split()], True, line))
                        i = e
                        sp = False
                        continue
            toks.append(Tok("id", word, sp, line))
            i = m.end()
            sp = False
            continue
        for op in OPS:
            if src.startswith(op, i):
                if op in "([{":
                    stack.append(op)
                elif op in ")]}":
                    if not stack:
                        raise LexError(f"{fname}:{line}: unbalanced '{op}'")
                    stack.pop()
                toks.append(Tok("op", "~=" if op == "!=" else ("~" if op == "!" else op), sp, line))
                i += len(op)
                sp = False
                break
        else:
            raise LexError(f"{fname}:{line}: unexpected character {c!r}")
    toks.append(Tok("nl", "\n", False, line))
    toks.append(Tok("eof", None, False, line))
    return toks
