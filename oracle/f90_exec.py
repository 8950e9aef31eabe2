"""Runs a reference PROGRAM from its own Fortran source text (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

The image has no Fortran compiler, so the reference cannot be built here.  This module is the next best pin for the
oracle: it reads a reference `.f90` file where it lies (/root/reference, never copied into the repo), transliterates the
small statement subset those programs use -- declarations, `do` / `if`, assignments to scalars, array elements and array
sections, a handful of intrinsics, and the MPI calls of the slab exchange -- statement by statement into Python, and
executes the result: every arithmetic expression is the reference's own text, evaluated in IEEE double precision in
source order (numpy float64 scalars; no FMA, no reassociation).  `parameter` constants can be overridden (a smaller
grid, fewer steps), which is what the reference expects its users to do by editing those lines.

MPI programs run as NPROC Python threads, one per rank, with MPI_SENDRECV / MPI_REDUCE emulated by queues.

What is NOT taken from the reference: I/O (print / write / open / close and the image and seismogram writers are skipped;
results are read from the program's variables after it ends) and date_and_time.

Used by tests/golden/make_reference_vectors.py to produce the golden vectors the oracle is checked against
(tests/test_reference_vectors.py).  Nothing here runs at test time on a box without /root/reference.
"""
from __future__ import annotations

import ast
import math
import queue
import re
import threading

import numpy as np


class FortranStop(Exception):
    pass


# ------------------------------------------------------------------------------------------------ source -> logical lines

def _strip_comment(line: str) -> str:
    q = None
    for n, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "!":
            return line[:n]
    return line


def _lower_outside_quotes(s: str) -> str:
    out, q = [], None
    for ch in s:
        if q:
            out.append(ch)
            if ch == q:
                q = None
        else:
            if ch in "'\"":
                q = ch
                out.append(ch)
            else:
                out.append(ch.lower())
    return "".join(out)


def logical_lines(path: str) -> list[str]:
    """The statements of the main program (from `program` to `end program`), comments and OpenMP directives removed,
    continuation lines joined, lower-cased outside character strings (Fortran is case-insensitive)."""
    raw = open(path).read().split("\n")
    start = next(n for n, l in enumerate(raw) if re.match(r"^\s*program\b", l, re.I))
    end = next(n for n, l in enumerate(raw) if re.match(r"^\s*end\s+program\b", l, re.I))
    out, cur = [], ""
    for line in raw[start + 1:end]:
        s = _strip_comment(line).rstrip()
        if not s.strip():
            continue
        body = s.strip()
        if cur:
            if body.startswith("&"):
                body = body[1:]
            cur = cur + " " + body
        else:
            cur = body
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        out.append(_lower_outside_quotes(cur))
        cur = ""
    return out


# ------------------------------------------------------------------------------------------------ expressions

def _match_paren(s: str, i: int) -> int:
    """Index of the parenthesis that closes s[i] == '('."""
    depth, q = 0, None
    for n in range(i, len(s)):
        ch = s[n]
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return n
    raise ValueError("unbalanced parentheses: " + s)


def _split_top(s: str, sep: str = ",") -> list[str]:
    parts, depth, q, cur = [], 0, None, []
    for ch in s:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append("".join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    parts.append("".join(cur).strip())
    return parts


_NUM_D = re.compile(r"(?<![\w.])(\d+\.?\d*|\.\d+)d([+-]?\d+)")
_IDENT = re.compile(r"[a-z_]\w*")


_PY_KEYWORDS = {"lambda", "in", "is", "from", "pass", "global", "class", "def", "del", "as", "with", "yield", "try",
                "except", "raise", "import", "assert", "async", "await", "nonlocal", "while", "for", "return", "break",
                "continue", "finally", "none"}


def _rename_keywords(s: str) -> str:
    """Fortran names that are Python keywords (the programs have a variable called lambda) get a trailing underscore."""
    return re.sub(r"(?<![\w.'\"])([a-z_]\w*)(?![\w'\"])", lambda m: m.group(1) + "_" if m.group(1) in _PY_KEYWORDS else m.group(1), s)


class _DivPow(ast.NodeTransformer):
    """a / b -> _div(a, b) (integer division truncates when both operands are integers); a ** b -> _pow(a, b)
    (x**2 and x**2.d0 are x*x, as gfortran compiles them)."""

    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.Call(func=ast.Name(id="_div", ctx=ast.Load()), args=[node.left, node.right], keywords=[])
        if isinstance(node.op, ast.Pow):
            return ast.Call(func=ast.Name(id="_pow", ctx=ast.Load()), args=[node.left, node.right], keywords=[])
        return node


class Translator:
    def __init__(self, overrides: dict[str, str] | None = None, externals=(), single: bool = False, vectorize: bool = False):
        # vectorize: an innermost `do` loop whose body is plain assignments with no dependence between iterations is
        # executed ONCE with the loop variable bound to the vector of its values (numpy integer-array indexing): the
        # same IEEE operations per element, ~100x faster.  `s = s + expr` accumulations keep their order (cumsum).
        self.vectorize = vectorize
        # single: the reference's single-precision build ("declare everything real", 3D-iso :114-116): every
        # `double precision` entity is a 4-byte real; literals keep their own kind (1.d0 stays double), so mixed
        # expressions are evaluated in double and rounded when they are assigned, as the compiled program would
        self.single = single
        self.real_scalars: set[str] = set()
        self.overrides = {k.lower(): v for k, v in (overrides or {}).items()}
        self.externals = {e.lower() for e in externals}          # external subroutines supplied by the caller
        self.bounds: dict[str, list[tuple[str, str]]] = {}      # array name -> [(lo, hi)] as Python expressions
        self.int_scalars: set[str] = set()                       # assignment to these truncates, as in Fortran
        self.code: list[str] = []
        self.depth = 0

    # ---- expressions
    def expr(self, s: str) -> str:
        s = _rename_keywords(s.strip())
        s = _NUM_D.sub(r"_dbl(\1e\2)" if self.single else r"\1e\2", s)
        for a, b in ((".and.", " and "), (".or.", " or "), (".not.", " not "), (".true.", " True "), (".false.", " False "),
                     (".eq.", "=="), (".ne.", "!="), (".lt.", "<"), (".le.", "<="), (".gt.", ">"), (".ge.", ">=")):
            s = s.replace(a, b)
        s = s.replace("/=", "!=")
        s = self._arrays(s)
        tree = ast.parse(s.strip(), mode="eval")
        tree = ast.fix_missing_locations(_DivPow().visit(tree))
        return ast.unparse(tree)

    def _arrays(self, s: str) -> str:
        out, n = [], 0
        while n < len(s):
            m = _IDENT.match(s, n)
            if m and (n == 0 or not (s[n - 1].isalnum() or s[n - 1] in "_.")):
                name = m.group(0)
                k = m.end()
                while k < len(s) and s[k] == " ":
                    k += 1
                if name in self.bounds and k < len(s) and s[k] == "(":
                    close = _match_paren(s, k)
                    out.append(self._subscript(name, s[k + 1:close]))
                    n = close + 1
                    continue
                out.append(name)
                n = m.end()
                continue
            out.append(s[n])
            n += 1
        return "".join(out)

    def _subscript(self, name: str, args: str) -> str:
        idx = []
        for (lo, _hi), a in zip(self.bounds[name], _split_top(args)):
            parts = _split_top(a, ":")
            if len(parts) == 1:
                idx.append(f"({self._arrays(a)})-({lo})")
            else:
                b = f"({self._arrays(parts[0])})-({lo})" if parts[0] else ""
                e = f"({self._arrays(parts[1])})-({lo})+1" if parts[1] else ""
                idx.append(f"{b}:{e}")
        return f"{name}[{', '.join(idx)}]"

    # ---- output
    def emit(self, line: str) -> None:
        self.code.append("    " * self.depth + line)

    # ---- declarations
    _DECL = re.compile(r"^(integer|double precision|logical|real|character)\b")

    def declaration(self, s: str) -> bool:
        m = self._DECL.match(s)
        if not m:
            return False
        kind = m.group(1)
        rest = s[m.end():].strip()
        if kind == "character":
            return True
        if rest.startswith("("):                                   # kind selector
            rest = rest[_match_paren(rest, 0) + 1:].strip()
        attrs, ents = "", rest
        if "::" in rest:
            attrs, ents = rest.split("::", 1)
        is_param = "parameter" in attrs
        dims = None
        dm = re.search(r"dimension\s*\(", attrs)
        if dm:
            o = attrs.index("(", dm.start())
            dims = attrs[o + 1:_match_paren(attrs, o)]
        real = kind in ("double precision", "real")
        zero = "0" if kind == "integer" else ("False" if kind == "logical" else ("np.float32(0.0)" if self.single else "0.0"))
        dtype = "np.int64" if kind == "integer" else ("np.float32" if self.single else "np.float64")
        for ent in _split_top(_rename_keywords(ents)):
            if not ent:
                continue
            if "=" in ent and is_param:
                name, val = ent.split("=", 1)
                name = name.strip()
                val = self.overrides.get(name, val)
                if kind == "integer":
                    self.emit(f"{name} = _toint({self.expr(val)})")
                elif real and self.single:
                    self.emit(f"{name} = np.float32({self.expr(val)})")
                else:
                    self.emit(f"{name} = {self.expr(val)}")
                continue
            edims = dims
            name = ent.strip()
            if "(" in name:
                o = name.index("(")
                edims = name[o + 1:_match_paren(name, o)]
                name = name[:o].strip()
            if edims is None:
                if kind == "integer":
                    self.int_scalars.add(name)
                elif real:
                    self.real_scalars.add(name)
                self.emit(f"{name} = {zero}")
            else:
                b = []
                for d in _split_top(edims):
                    p = _split_top(d, ":")
                    b.append(("1", p[0]) if len(p) == 1 else (p[0], p[1]))
                self.bounds[name] = [(self.expr(lo), self.expr(hi)) for lo, hi in b]
                shape = ", ".join(f"({hi})-({lo})+1" for lo, hi in self.bounds[name])
                self.emit(f"{name} = np.zeros(({shape},), dtype={dtype})")
        return True

    # ---- MPI
    def mpi_call(self, name: str, args: list[str]) -> None:
        if name in ("mpi_init", "mpi_finalize"):
            return
        if name == "mpi_barrier":
            self.emit("_mpi.barrier()")
        elif name == "mpi_comm_size":
            self.emit(f"{args[1]} = _mpi.size")
        elif name == "mpi_comm_rank":
            self.emit(f"{args[1]} = _mpi.rank")
        elif name == "mpi_sendrecv":
            send, dest, recv, src = args[0], args[3], args[5], args[8]
            self.emit(f"_t = _mpi.sendrecv({self.expr(send)}, {self.expr(dest)}, {self.expr(src)})")
            self.emit("if _t is not None:")
            self.emit(f"    {self.expr(recv)} = _t")
        elif name == "mpi_reduce":
            val, target, op, root = args[0], args[1], args[4], args[5]
            self.emit(f"_t = _mpi.reduce({self.expr(val)}, '{op.strip()}', {self.expr(root)})")
            self.emit("if _t is not None:")
            self.emit(f"    {self.expr(target)} = _t")
        else:
            raise NotImplementedError(name)

    # ---- statements
    SKIP_CALLS = ("date_and_time", "write_seismograms", "create_color_image", "create_2d_image", "system_clock", "cpu_time",
                  "system")

    def statement(self, s: str) -> None:
        if self.declaration(s):
            return
        if re.match(r"^(implicit|include|use|print|write|open|close|format|\d+\s+format)\b", s):
            return
        m = re.match(r"^do\s+(\w+)\s*=\s*(.*)$", s)
        if m:
            r = _split_top(m.group(2))
            step = f", {self.expr(r[2])}" if len(r) > 2 else ""
            self.emit(f"for {m.group(1)} in _range({self.expr(r[0])}, {self.expr(r[1])}{step}):")
            self.depth += 1
            return
        if re.match(r"^end\s*do$", s):
            self.depth -= 1
            return
        if re.match(r"^end\s*if$", s):
            self.depth -= 1
            return
        if s == "else":
            self.depth -= 1
            self.emit("else:")
            self.depth += 1
            return
        m = re.match(r"^(else\s*)?if\s*\(", s)
        if m:
            o = s.index("(", m.start())
            c = _match_paren(s, o)
            cond, tail = s[o + 1:c], s[c + 1:].strip()
            if tail == "then":
                if m.group(1):
                    self.depth -= 1
                    self.emit(f"elif {self.expr(cond)}:")
                else:
                    self.emit(f"if {self.expr(cond)}:")
                self.depth += 1
            else:
                self.emit(f"if {self.expr(cond)}:")
                self.depth += 1
                n0 = len(self.code)
                self.statement(tail)
                if len(self.code) == n0:
                    self.emit("pass")
                self.depth -= 1
            return
        m = re.match(r"^stop\b(.*)$", s)
        if m:
            self.emit(f"raise FortranStop({m.group(1).strip() or repr('stop')})")
            return
        m = re.match(r"^call\s+(\w+)\s*(\(.*\))?$", s)
        if m:
            name = m.group(1)
            if name in self.SKIP_CALLS:
                return
            args = _split_top(m.group(2)[1:-1]) if m.group(2) else []
            if name.startswith("mpi_"):
                self.mpi_call(name, args)
                return
            if name in self.externals:        # arrays go by reference (numpy), scalars by value: inputs only
                self.emit(f"_ext_{name}({', '.join(self.expr(a) for a in args)})")
                return
            raise NotImplementedError("call " + name)
        # assignment: the first '=' at depth 0 that is not part of ==, /=, <=, >=
        depth = 0
        for n, ch in enumerate(s):
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "=" and depth == 0 and s[n + 1:n + 2] != "=" and s[n - 1] not in "=/<>":
                lhs, rhs = s[:n].strip(), s[n + 1:].strip()
                if lhs in self.int_scalars:
                    self.emit(f"{lhs} = _toint({self.expr(rhs)})")
                elif self.single and _rename_keywords(lhs) in self.real_scalars:
                    self.emit(f"{_rename_keywords(lhs)} = np.float32({self.expr(rhs)})")
                else:
                    self.emit(f"{self._arrays(_rename_keywords(lhs))} = {self.expr(rhs)}")
                return
        raise NotImplementedError("statement: " + s)

    _ASSIGN = re.compile(r"^([a-z_]\w*)\s*(\(.*?\))?\s*=(?!=)")

    def _vector_loop(self, lines: list[str], n: int):
        """If lines[n] opens a loop -- or a perfect nest of two loops -- whose body can run as ONE array operation, emit it
        and return the index of its last `enddo`; else None.  Conditions: the body is plain assignments; an array
        written in the body is only ever referenced at the loop indices themselves (no i+1 of something written here);
        a scalar that is read before the iteration assigns it must be an accumulation `s = s + expr`, which keeps its
        order (inner loop fastest) through a cumulative sum.  The loop variables become integer index arrays -- the
        inner one a column, the outer one a row -- so every element sees the same IEEE operations as in the loop."""
        DO = re.compile(r"^do\s+(\w+)\s*=\s*(.*)$")
        loops = []
        pos = n
        while len(loops) < 2:
            m = DO.match(lines[pos]) if pos < len(lines) else None
            if not m:
                break
            rng = _split_top(m.group(2))
            if len(rng) != 2:
                return None
            loops.append((m.group(1), rng))
            pos += 1
        if not loops:
            return None
        # the body: assignments, and loops over OTHER variables whose bodies are assignments (do i_sls = 1,N_SLS): those
        # stay loops, with vector statements inside
        end, items, body = pos, [], []
        while end < len(lines) and not re.match(r"^end\s*do$", lines[end]):
            m2 = DO.match(lines[end])
            if m2:
                r2 = _split_top(m2.group(2))
                k = end + 1
                while k < len(lines) and not re.match(r"^end\s*do$", lines[k]):
                    if re.match(r"^(do|if|else|end\s*if|call|stop|print|write)\b", lines[k]):
                        return None
                    k += 1
                if len(r2) != 2 or k >= len(lines) or k == end + 1:
                    return None
                items.append((m2.group(1), r2, lines[end + 1:k]))
                body += lines[end + 1:k]
                end = k + 1
                continue
            if re.match(r"^(if|else|end\s*if|call|stop|print|write)\b", lines[end]):
                return None
            items.append((None, None, [lines[end]]))
            body.append(lines[end])
            end += 1
        # a nest of two must be perfect: the outer loop closes right after the inner one
        last = end
        if len(loops) == 2:
            if end + 1 >= len(lines) or not re.match(r"^end\s*do$", lines[end + 1]):
                return None
            last = end + 1
        if not body:
            return None
        names = [v for v, _ in loops]
        uses = lambda text, name: re.search(r"(?<![\w.])" + re.escape(name) + r"(?!\w)", text) is not None
        for v2, r2, _ in items:              # inner loops: another variable, bounds that do not depend on the vector ones
            if v2 is not None and (v2 in names or any(uses(b, v) for b in r2 for v in names)):
                return None
        inner_statements = {st for v2, _, sts in items if v2 is not None for st in sts}
        parsed = []
        for st in body:
            a = self._ASSIGN.match(st)
            if not a:
                return None
            parsed.append((st, a.group(1), st[a.end():], (a.group(2) or "")))
        assigned_scalars = {lhs for _, lhs, _, _ in parsed if lhs not in self.bounds}
        written_arrays, accum, defined = set(), {}, set()
        for st, lhs, rhs, lhs_args in parsed:
            carried = [x for x in assigned_scalars - defined if uses(rhs, x)]
            if lhs in self.bounds:
                if carried:
                    return None
                if not any(uses(lhs_args, v) for v in names):
                    # the same element in every iteration: only the accumulation A(k) = A(k) + expr is handled
                    ref = lhs + lhs_args
                    if self.single or not rhs.strip().startswith(ref) or st in inner_statements:
                        return None
                    rest = rhs.strip()[len(ref):].lstrip()
                    if not rest.startswith("+") or uses(rest, lhs):
                        return None
                    accum[st] = (ref, rest[1:])
                    continue
                written_arrays.add(lhs)
                continue
            if lhs in self.int_scalars or lhs in names:
                return None
            if carried:
                mm = re.match(r"^\s*" + re.escape(lhs) + r"\s*\+(.*)$", rhs)
                if carried != [lhs] or not mm or self.single or uses(mm.group(1), lhs) or st in inner_statements:
                    return None
                accum[st] = (lhs, mm.group(1))
            else:
                defined.add(lhs)
        for st in body:
            for w in written_arrays:
                for ref in re.finditer(r"(?<![\w.])" + re.escape(w) + r"\s*\(", st):
                    o = ref.end() - 1
                    for arg in _split_top(st[o + 1:_match_paren(st, o)]):
                        if any(uses(arg, v) for v in names) and arg.strip() not in names:
                            return None
        bounds = [(self.expr(r[0]), self.expr(r[1])) for _, r in loops]
        inner, outer = names[-1], (names[0] if len(names) == 2 else None)
        # numpy arrays are indexed [i, j, ...] like the Fortran ones: whichever variable is used in an earlier dimension
        # does not matter for broadcasting as long as the two index arrays have different axes
        self.emit(f"{inner} = np.arange({bounds[-1][0]}, ({bounds[-1][1]}) + 1)" + (".reshape(-1, 1)" if outer else ""))
        if outer:
            self.emit(f"{outer} = np.arange({bounds[0][0]}, ({bounds[0][1]}) + 1).reshape(1, -1)")
        size = f"{inner}.size * {outer}.size" if outer else f"{inner}.size"
        self.emit(f"if {size}:")
        self.depth += 1
        for v2, r2, sts in items:
            if v2 is not None:
                self.emit(f"for {v2} in _range({self.expr(r2[0])}, {self.expr(r2[1])}):")
                self.depth += 1
            for st in sts:
                if st in accum:
                    lhs, term = accum[st]
                    shape = f"({inner}.size, {outer}.size)" if outer else f"({inner}.size,)"
                    target = self._arrays(_rename_keywords(lhs))          # a scalar or one array element
                    self.emit(f"{target} = _accumulate({target}, {self.expr(term)}, {shape})")
                else:
                    self.statement(st)
            if v2 is not None:
                self.depth -= 1
        self.depth -= 1
        self.emit(f"{inner} = ({bounds[-1][1]}) + 1")
        if outer:
            self.emit(f"{outer} = ({bounds[0][1]}) + 1")
        return last

    def translate(self, lines: list[str]) -> str:
        skip_to = -1
        for n, s in enumerate(lines):
            if n <= skip_to:
                continue
            if self.vectorize and re.match(r"^do\s+\w+\s*=", s):
                end = self._vector_loop(lines, n)
                if end is not None:
                    skip_to = end
                    continue
            n0 = len(self.code)
            self.statement(s)
            # a block opener whose body turned out empty (only skipped statements) needs a `pass`: add one after every
            # opener and let it be harmless
            if len(self.code) > n0 and self.code[-1].rstrip().endswith(":"):
                self.emit("pass")
        assert self.depth == 0, "unbalanced blocks"
        return "\n".join(self.code) + "\n"


# ------------------------------------------------------------------------------------------------ run time

def _div(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)) and not isinstance(a, bool):
        q = abs(int(a)) // abs(int(b))
        return q if (a >= 0) == (b >= 0) else -q          # Fortran integer division truncates towards zero
    if isinstance(a, (np.ndarray, np.floating)) or isinstance(b, (np.ndarray, np.floating)):
        return a / b                                      # numpy kinds: real / real stays real, real / double is double
    return np.float64(a) / np.float64(b)


def _toint(x):
    """Assignment of a real value to an integer truncates towards zero."""
    return x if isinstance(x, (int, np.integer)) else int(x)


def _pow(a, b):
    if isinstance(b, (int, np.integer)) or float(b) == int(b):
        n = int(b)
        if isinstance(a, (int, np.integer)) and n >= 0:
            return int(a) ** n
        if n == 2:
            return a * a
        if n == 1:
            return a
        if n == 3:
            return a * a * a
    return type(a)(math.pow(float(a), float(b))) if isinstance(a, np.floating) else np.float64(math.pow(float(a), float(b)))


def _range(a, b, step=1):
    return range(int(a), int(b) + (1 if step > 0 else -1), int(step))


def _mod(a, b):
    if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)):
        return int(math.fmod(int(a), int(b)))
    return np.float64(math.fmod(a, b))


def _elementwise(f_scalar, f_array):
    def f(x):
        if isinstance(x, np.ndarray) or isinstance(x, np.float32):
            return f_array(x)                             # the 4-byte routine for a 4-byte argument
        return np.float64(f_scalar(x))
    return f


def _accumulate(s, terms, shape):
    """s = s + term for every iteration of the loop (nest), in loop order: the inner loop variable -- axis 0 of the
    terms -- runs fastest; a cumulative sum adds left to right."""
    t = np.broadcast_to(np.asarray(terms, dtype=np.float64), shape)
    return np.cumsum(np.concatenate(([np.float64(s)], t.T.ravel())))[-1]


def _sum(a):
    """SUM of an array expression the way a compiler without -ffast-math does it: one accumulator, array element order
    (first index fastest)."""
    a = np.asarray(a)
    return np.cumsum(a.T.ravel(), dtype=a.dtype)[-1] if a.size else a.dtype.type(0.0)


RUNTIME = {
    "np": np, "FortranStop": FortranStop, "_div": _div, "_pow": _pow, "_range": _range, "_toint": _toint,
    "_accumulate": _accumulate,
    "dble": lambda x: np.float64(x), "real": lambda x: np.float64(x), "sngl": lambda x: np.float32(x), "int": lambda x: int(x),
    "_dbl": np.float64,
    "exp": _elementwise(math.exp, np.exp), "log": _elementwise(math.log, np.log), "sqrt": _elementwise(math.sqrt, np.sqrt),
    "sin": _elementwise(math.sin, np.sin), "cos": _elementwise(math.cos, np.cos), "abs": abs, "max": max, "min": min,
    "dsqrt": _elementwise(math.sqrt, np.sqrt), "dexp": _elementwise(math.exp, np.exp), "dlog": _elementwise(math.log, np.log),
    "dsin": _elementwise(math.sin, np.sin), "dcos": _elementwise(math.cos, np.cos), "dabs": abs, "atan": _elementwise(math.atan, np.arctan),
    "dmax1": max, "dmin1": min, "nint": lambda x: int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5)), "float": lambda x: np.float64(x),
    "mod": _mod, "sum": _sum, "maxval": lambda a: np.max(a), "minval": lambda a: np.min(a),
    "mpi_proc_null": -1, "mpi_comm_world": 0, "mpi_double_precision": 0, "mpi_sum": "mpi_sum", "mpi_max": "mpi_max",
    "mpi_status_size": 1,
}


class _Comm:
    """MPI_COMM_WORLD of NPROC threads."""

    def __init__(self, size: int):
        self.size = size
        self.p2p = {(s, d): queue.Queue() for s in range(size) for d in range(size)}
        self.red: dict[int, dict[int, object]] = {}       # reduction number -> {rank: value}
        self.red_cv = threading.Condition()
        self.bar = threading.Barrier(size)


class _Rank:
    def __init__(self, comm: _Comm, rank: int):
        self.comm, self.rank, self.size = comm, rank, comm.size
        self.nred = 0                                      # reductions this rank has entered (collectives are ordered)

    def sendrecv(self, sendbuf, dest, src):
        if dest >= 0:
            self.comm.p2p[(self.rank, dest)].put(np.array(sendbuf, copy=True))
        return self.comm.p2p[(src, self.rank)].get(timeout=600) if src >= 0 else None

    def reduce(self, value, op, root):
        seq, self.nred = self.nred, self.nred + 1
        with self.comm.red_cv:
            self.comm.red.setdefault(seq, {})[self.rank] = value
            self.comm.red_cv.notify_all()
            if self.rank != root:
                return None
            if not self.comm.red_cv.wait_for(lambda: len(self.comm.red[seq]) == self.size, timeout=600):
                raise RuntimeError("MPI_REDUCE: a rank never arrived")
            vals = self.comm.red.pop(seq)
        acc = vals[0]
        for r in range(1, self.size):                     # rank order (two ranks: the only order there is)
            acc = acc + vals[r] if op == "mpi_sum" else max(acc, vals[r])
        return acc

    def barrier(self):
        self.comm.bar.wait()


def run_program(path: str, overrides: dict[str, str] | None = None, nproc: int = 1, externals: dict | None = None,
                edits: list[tuple[str, str]] | None = None, single: bool = False, vectorize: bool = False) -> list[dict]:
    """Executes the main program of `path`; returns the variables of every rank after `end program`.
    overrides: {parameter name: Fortran expression} replacing the value of a `parameter` declaration;
    edits: [(regex, replacement)] applied to the (lower-cased) statements first -- for values the reference sets by
    assignment rather than by parameter (the viscoelastic program's receiver positions);
    vectorize: run dependence-free innermost loops as one numpy operation (same bits, ~100x faster);
    single: the single-precision build the reference endorses (every `double precision` entity a 4-byte real);
    externals: {subroutine name: Python callable} for subroutines that live in another file of the reference (the
    SolvOpt attenuation fit, which is pinned separately); array arguments are passed by reference."""
    externals = externals or {}
    tr = Translator(overrides, externals, single, vectorize)
    lines = logical_lines(path)
    for pat, repl in edits or []:
        lines = [re.sub(pat, repl, l) for l in lines]
    src = tr.translate(lines)
    code = compile(src, "<transliterated " + path.rsplit("/", 1)[-1] + ">", "exec")
    comm = _Comm(nproc)
    spaces, errors = [dict(RUNTIME) for _ in range(nproc)], []
    for sp in spaces:
        for name, fn in externals.items():
            sp["_ext_" + name.lower()] = fn

    def body(r):
        try:
            spaces[r]["_mpi"] = _Rank(comm, r)
            with np.errstate(all="ignore"):
                exec(code, spaces[r])
        except BaseException as e:                       # noqa: BLE001 -- reported to the caller below
            errors.append((r, e))

    threads = [threading.Thread(target=body, args=(r,)) for r in range(nproc)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise RuntimeError(f"rank {errors[0][0]} failed: {errors[0][1]!r}") from errors[0][1]
    for sp in spaces:
        sp["_bounds"] = tr.bounds
    return spaces
