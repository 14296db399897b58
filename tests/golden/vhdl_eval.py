"""vhdl_eval - a small evaluator for the synthesizable VHDL subset the reference's receive chain
is written in.  TEST INFRASTRUCTURE ONLY (used by tests/golden/make_golden.py in the build
container, where /root/reference is mounted; never imported by the product).

Purpose ("pin the filter arithmetic to the reference's text"): the reference
holds no vectors for the window / biquad / cascade and no VHDL simulator exists in the image, so
the fixtures in tests/golden/rtl_vectors.npz are produced by PARSING the reference's own source
files - NEW/filter_iir_cust.vhd, NEW/filter_iir12_cust.vhd, IMP/filter_iir.vhd,
IMP/filter_iir12.vhd, IMP/filter_pkg.vhd, NEW/hann8192.vhd, NEW/hann.vhd - and executing their
concurrent assignments, clocked processes, generate statements and component instantiations
with ieee.numeric_std semantics:

  signed * signed        -> width L + R, exact
  signed +/- signed      -> width max(L, R), two's-complement wrap
  signed + std_logic     -> the bit added as 0 / 1, width of the vector operand
  x(hi downto lo)        -> that bit field, same base type (a slice of signed is signed)
  resize(signed, n)      -> n < width: sign bit kept + low n-1 bits;  n > width: sign extension
  signed(), unsigned(), std_logic_vector()  -> the same bits, reinterpreted
  to_signed(i, n), to_integer(), unsigned + integer (wrap), integer + - mod, = /= < > <= >=,
  not / and / or / xor on std_logic, aggregates (others => ...), bit strings x"00"

Simulation: one clock domain.  tick() = evaluate every clocked process on the pre-edge values,
apply all updates at once, then re-evaluate the concurrent assignments (including port
associations, which are elaborated as concurrent assignments) to a fixed point.  Registers
without an initial value start at 0 where a simulator would show 'U'.
"""
from __future__ import annotations

import re


# ------------------------------------------------------------------------------------ lexer
_TOKEN = re.compile(r"""
    (?P<ws>\s+|--[^\n]*)
  | (?P<bitstr>[xXbBoO]"[0-9a-fA-F_]*")
  | (?P<str>"[^"]*")
  | (?P<char>'[01UXZ-]')
  | (?P<num>\d[\d_]*)
  | (?P<id>[A-Za-z][A-Za-z0-9_]*)
  | (?P<sym><=|=>|:=|/=|>=|\*\*|[()\[\];:,+\-*/&=<>'.|])
""", re.X)


def lex(text):
    out, pos = [], 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise SyntaxError(f"cannot tokenise at {text[pos:pos + 30]!r}")
        pos = m.end()
        k = m.lastgroup
        if k == "ws":
            continue
        v = m.group(k)
        if k == "id":
            v = v.lower()
        out.append((k, v))
    out.append(("eof", ""))
    return out


# ----------------------------------------------------------------------------------- values
class Vec:
    """signed / unsigned / std_logic_vector: `w` bits, `v` = the bit pattern as a non-negative int."""
    __slots__ = ("kind", "w", "v")

    def __init__(self, kind, w, v):
        self.kind, self.w, self.v = kind, w, v & ((1 << w) - 1)

    def sint(self):
        return self.v - (1 << self.w) if self.v >> (self.w - 1) else self.v

    def num(self):
        return self.sint() if self.kind == "signed" else self.v

    def __eq__(self, o):
        return isinstance(o, Vec) and (self.kind, self.w, self.v) == (o.kind, o.w, o.v)

    def __repr__(self):
        return f"{self.kind}{self.w}({self.num()})"


class Bit:
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = int(v) & 1

    def __eq__(self, o):
        return isinstance(o, Bit) and self.v == o.v

    def __repr__(self):
        return f"'{self.v}'"


class Enum:
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __eq__(self, o):
        return isinstance(o, Enum) and self.name == o.name

    def __repr__(self):
        return self.name


class Arr:
    """array(lo..hi) of element; index by integer."""
    __slots__ = ("lo", "hi", "el")

    def __init__(self, lo, hi, el):
        self.lo, self.hi, self.el = lo, hi, el

    def __eq__(self, o):
        return isinstance(o, Arr) and self.el == o.el

    def copy(self):
        return Arr(self.lo, self.hi, list(self.el))


class Agg:
    """aggregate awaiting its target type: positional / named choices + others"""

    def __init__(self, named, others):
        self.named, self.others = named, others


def default_value(t):
    k = t[0]
    if k == "bit":
        return Bit(0)
    if k == "vec":
        return Vec(t[1], t[2] - t[3] + 1, 0)
    if k == "int":
        return 0
    if k == "enum":
        return Enum(t[1][0])
    if k == "array":
        return Arr(t[1], t[2], [default_value(t[3]) for _ in range(t[2] - t[1] + 1)])
    raise ValueError(t)


def coerce(val, t):
    """value -> type t (resolves aggregates, checks widths)."""
    k = t[0]
    if isinstance(val, Agg):
        if k == "vec":
            w = t[2] - t[3] + 1
            bits = 0
            for i in range(w):
                b = val.named.get(i + t[3], val.others)
                bits |= coerce(b, ("bit",)).v << i
            return Vec(t[1], w, bits)
        if k == "array":
            return Arr(t[1], t[2], [coerce(val.named.get(i, val.others), t[3]) for i in range(t[1], t[2] + 1)])
        raise TypeError("aggregate for scalar")
    if k == "vec":
        if isinstance(val, Vec):
            if val.w != t[2] - t[3] + 1:
                raise TypeError(f"width mismatch: {val} into {t}")
            return Vec(t[1], val.w, val.v)
        raise TypeError(f"{val!r} into {t}")
    if k == "bit":
        if isinstance(val, Bit):
            return val
        raise TypeError(f"{val!r} into std_logic")
    if k == "int":
        if isinstance(val, int) and not isinstance(val, bool):
            return val
        raise TypeError(f"{val!r} into integer")
    if k == "array":
        if isinstance(val, Arr):
            return Arr(t[1], t[2], [coerce(e, t[3]) for e in val.el])
        raise TypeError(f"{val!r} into array")
    return val


# ----------------------------------------------------------------------------------- parser
class Parser:
    def __init__(self, text):
        self.t = lex(text)
        self.i = 0

    def peek(self, k=0):
        return self.t[self.i + k]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def accept(self, v):
        if self.t[self.i][1] == v and self.t[self.i][0] in ("id", "sym"):
            self.i += 1
            return True
        return False

    def expect(self, v):
        if not self.accept(v):
            raise SyntaxError(f"expected {v!r}, got {self.t[self.i]} near token {self.i}")

    def ident(self):
        k, v = self.next()
        if k != "id":
            raise SyntaxError(f"identifier expected, got {v!r}")
        return v

    # ---- design units
    def design(self):
        units = {"entities": {}, "archs": {}, "packages": {}}
        while self.peek()[0] != "eof":
            if self.accept("library") or self.accept("use"):
                while not self.accept(";"):
                    self.next()
            elif self.accept("entity"):
                e = self.entity()
                units["entities"][e["name"]] = e
            elif self.accept("architecture"):
                a = self.architecture()
                units["archs"][a["of"]] = a
            elif self.accept("package"):
                p = self.package()
                units["packages"][p["name"]] = p
            else:
                raise SyntaxError(f"unexpected {self.peek()}")
        return units

    def entity(self):
        name = self.ident()
        self.expect("is")
        generics, ports = [], []
        if self.accept("generic"):
            self.expect("(")
            while True:
                names = [self.ident()]
                while self.accept(","):
                    names.append(self.ident())
                self.expect(":")
                t = self.type_mark()
                d = self.expr() if self.accept(":=") else None
                generics += [(n, t, d) for n in names]
                if not self.accept(";"):
                    break
            self.expect(")")
            self.expect(";")
        if self.accept("port"):
            self.expect("(")
            while True:
                names = [self.ident()]
                while self.accept(","):
                    names.append(self.ident())
                self.expect(":")
                mode = self.ident()
                t = self.type_mark()
                if self.accept(":="):
                    self.expr()
                ports += [(n, mode, t) for n in names]
                if not self.accept(";"):
                    break
            self.expect(")")
            self.expect(";")
        self.expect("end")
        self.accept("entity")
        if self.peek()[0] == "id":
            self.next()
        self.expect(";")
        return {"name": name, "generics": generics, "ports": ports}

    def type_mark(self):
        """returns an unresolved type: ('name', id, constraint-or-None)"""
        name = self.ident()
        if name == "integer" or name == "natural" or name == "positive":
            if self.accept("range"):
                self.expr()
                if not self.accept("to"):
                    self.expect("downto")
                self.expr()
            return ("name", "integer", None)
        if self.peek()[1] == "(":
            self.next()
            hi = self.expr()
            down = self.accept("downto")
            if not down:
                self.expect("to")
            lo = self.expr()
            self.expect(")")
            return ("name", name, (hi, lo) if down else (lo, hi))
        return ("name", name, None)

    def decls(self, end_words):
        types, consts, signals = {}, [], []
        while self.peek()[1] not in end_words:
            if self.accept("type"):
                name = self.ident()
                self.expect("is")
                if self.accept("array"):
                    self.expect("(")
                    a = self.expr()
                    down = self.accept("downto")
                    if not down:
                        self.expect("to")
                    b = self.expr()
                    self.expect(")")
                    self.expect("of")
                    el = self.type_mark()
                    types[name] = ("arraydef", (b, a) if down else (a, b), el)
                else:
                    self.expect("(")
                    lits = [self.ident()]
                    while self.accept(","):
                        lits.append(self.ident())
                    self.expect(")")
                    types[name] = ("enumdef", lits)
                self.expect(";")
            elif self.accept("constant"):
                name = self.ident()
                self.expect(":")
                t = self.type_mark()
                self.expect(":=")
                consts.append((name, t, self.expr()))
                self.expect(";")
            elif self.accept("signal"):
                names = [self.ident()]
                while self.accept(","):
                    names.append(self.ident())
                self.expect(":")
                t = self.type_mark()
                init = self.expr() if self.accept(":=") else None
                self.expect(";")
                signals += [(n, t, init) for n in names]
            elif self.accept("component"):
                while not (self.peek()[1] == "end" and self.peek(1)[1] == "component"):
                    self.next()
                self.next(); self.next()
                if self.peek()[0] == "id":
                    self.next()
                self.expect(";")
            elif self.accept("attribute"):
                while not self.accept(";"):
                    self.next()
            else:
                raise SyntaxError(f"declaration: unexpected {self.peek()}")
        return types, consts, signals

    def package(self):
        name = self.ident()
        self.expect("is")
        types, consts, _ = self.decls(("end",))
        self.expect("end")
        self.accept("package")
        if self.peek()[0] == "id":
            self.next()
        self.expect(";")
        return {"name": name, "types": types, "consts": consts}

    def architecture(self):
        name = self.ident()
        self.expect("of")
        of = self.ident()
        self.expect("is")
        types, consts, signals = self.decls(("begin",))
        self.expect("begin")
        stmts = self.conc_stmts(("end",))
        self.expect("end")
        self.accept("architecture")
        if self.peek()[0] == "id":
            self.next()
        self.expect(";")
        return {"name": name, "of": of, "types": types, "consts": consts, "signals": signals, "stmts": stmts}

    # ---- concurrent statements
    def conc_stmts(self, end_words):
        out = []
        while self.peek()[1] not in end_words:
            label = None
            if self.peek()[0] == "id" and self.peek(1)[1] == ":" and self.peek(2)[1] != "=":
                label = self.ident()
                self.expect(":")
            if self.accept("process"):
                if self.accept("("):
                    while not self.accept(")"):
                        self.next()
                self.accept("is")
                self.expect("begin")
                body = self.seq_stmts(("end",))
                self.expect("end")
                self.expect("process")
                if self.peek()[0] == "id":
                    self.next()
                self.expect(";")
                out.append(("process", label, body))
            elif self.accept("if"):
                cond = self.expr()
                self.expect("generate")
                body = self.conc_stmts(("end",))
                self.expect("end"); self.expect("generate")
                if self.peek()[0] == "id":
                    self.next()
                self.expect(";")
                out.append(("ifgen", cond, body))
            elif self.accept("for"):
                var = self.ident()
                self.expect("in")
                a = self.expr()
                self.expect("to")
                b = self.expr()
                self.expect("generate")
                body = self.conc_stmts(("end",))
                self.expect("end"); self.expect("generate")
                if self.peek()[0] == "id":
                    self.next()
                self.expect(";")
                out.append(("forgen", var, a, b, body))
            elif label is not None and self.peek()[0] == "id" and self.peek(1)[1] in ("generic", "port"):
                ent = self.ident()
                gmap, pmap = [], []
                if self.accept("generic"):
                    self.expect("map")
                    gmap = self.assoc_list()
                self.expect("port")
                self.expect("map")
                pmap = self.assoc_list()
                self.expect(";")
                out.append(("inst", label, ent, gmap, pmap))
            else:
                target = self.target()
                self.expect("<=")
                e = self.expr()
                self.expect(";")
                out.append(("assign", target, e))
        return out

    def assoc_list(self):
        self.expect("(")
        out = []
        while True:
            formal = self.ident()
            self.expect("=>")
            if self.accept("open"):
                out.append((formal, None))
            else:
                out.append((formal, self.expr()))
            if not self.accept(","):
                break
        self.expect(")")
        return out

    def target(self):
        name = self.ident()
        idx = None
        if self.accept("("):
            idx = self.expr()
            self.expect(")")
        return (name, idx)

    # ---- sequential statements
    def seq_stmts(self, end_words):
        out = []
        while self.peek()[1] not in end_words:
            if self.accept("if"):
                branches = []
                cond = self.expr()
                self.expect("then")
                body = self.seq_stmts(("elsif", "else", "end"))
                branches.append((cond, body))
                els = []
                while True:
                    if self.accept("elsif"):
                        c = self.expr()
                        self.expect("then")
                        branches.append((c, self.seq_stmts(("elsif", "else", "end"))))
                    elif self.accept("else"):
                        els = self.seq_stmts(("end",))
                    else:
                        break
                self.expect("end"); self.expect("if"); self.expect(";")
                out.append(("if", branches, els))
            elif self.accept("null"):
                self.expect(";")
            else:
                target = self.target()
                self.expect("<=")
                e = self.expr()
                self.expect(";")
                out.append(("assign", target, e))
        return out

    # ---- expressions (VHDL precedence: logical < relational < adding < sign < multiplying < not)
    def expr(self):
        left = self.relation()
        while self.peek()[0] == "id" and self.peek()[1] in ("and", "or", "xor", "nand", "nor"):
            op = self.next()[1]
            left = ("bin", op, left, self.relation())
        return left

    def relation(self):
        left = self.simple()
        if self.peek()[0] == "sym" and self.peek()[1] in ("=", "/=", "<", ">", ">=", "<="):
            op = self.next()[1]
            left = ("bin", op, left, self.simple())
        return left

    def simple(self):
        if self.peek()[1] == "-" and self.peek()[0] == "sym":
            self.next()
            left = ("neg", self.term())
        elif self.peek()[1] == "+" and self.peek()[0] == "sym":
            self.next()
            left = self.term()
        else:
            left = self.term()
        while self.peek()[0] == "sym" and self.peek()[1] in ("+", "-", "&"):
            op = self.next()[1]
            left = ("bin", op, left, self.term())
        return left

    def term(self):
        left = self.factor()
        while (self.peek()[0] == "sym" and self.peek()[1] in ("*", "/")) or \
              (self.peek()[0] == "id" and self.peek()[1] in ("mod", "rem")):
            op = self.next()[1]
            left = ("bin", op, left, self.factor())
        return left

    def factor(self):
        if self.peek() == ("id", "not"):
            self.next()
            return ("not", self.primary())
        return self.primary()

    def primary(self):
        k, v = self.peek()
        if k == "num":
            self.next()
            return ("int", int(v.replace("_", "")))
        if k == "char":
            self.next()
            return ("bit", 1 if v[1] == "1" else 0)
        if k == "bitstr":
            self.next()
            base = {"x": 4, "b": 1, "o": 3}[v[0].lower()]
            digits = v[2:-1].replace("_", "")
            return ("vec", "slv", base * len(digits), int(digits, 1 << base) if digits else 0)
        if k == "str":
            self.next()
            return ("vec", "slv", len(v) - 2, int(v[1:-1], 2) if len(v) > 2 else 0)
        if v == "(" and k == "sym":
            self.next()
            # aggregate or parenthesised expression
            if self.peek() == ("id", "others"):
                self.next()
                self.expect("=>")
                e = self.expr()
                self.expect(")")
                return ("agg", [], e)
            first = self.expr()
            if self.accept("=>"):
                named = [(first, self.expr())]
                others = None
                while self.accept(","):
                    if self.accept("others"):
                        self.expect("=>")
                        others = self.expr()
                    else:
                        c = self.expr()
                        self.expect("=>")
                        named.append((c, self.expr()))
                self.expect(")")
                return ("agg", named, others)
            self.expect(")")
            return first
        if k == "id":
            self.next()
            node = ("name", v)
            while self.peek() == ("sym", "("):
                self.next()
                args = []
                while True:
                    a = self.expr()
                    if self.accept("downto"):
                        a = ("range", a, self.expr())
                    elif self.accept("to"):
                        b = self.expr()
                        a = ("range", b, a)
                    args.append(a)
                    if not self.accept(","):
                        break
                self.expect(")")
                node = ("call", node, args)
            return node
        raise SyntaxError(f"primary: unexpected {self.peek()}")


# -------------------------------------------------------------------------------- evaluation
class Scope:
    """one elaborated instance: its signal name prefix, constants (generics, generate variables,
    package / architecture constants) and resolved types."""

    def __init__(self, design, prefix, consts, types):
        self.d, self.prefix, self.consts, self.types = design, prefix, consts, types

    def child(self, extra):
        c = dict(self.consts)
        c.update(extra)
        return Scope(self.d, self.prefix, c, self.types)


class Design:
    def __init__(self, sources):
        """sources: list of VHDL texts (packages, entities, architectures)."""
        self.entities, self.archs, self.packages = {}, {}, {}
        for text in sources:
            u = Parser(text).design()
            self.entities.update(u["entities"])
            self.archs.update(u["archs"])
            self.packages.update(u["packages"])
        self.signals = {}          # full name -> value
        self.sigtypes = {}
        self.assigns = []          # (scope, (name, idx_expr), expr)
        self.procs = []            # (scope, body)
        self.pkg_consts, self.pkg_types = {}, {}
        for p in self.packages.values():
            for name, tdef in p["types"].items():
                self.pkg_types[name] = tdef
            sc = Scope(self, "", self.pkg_consts, self.pkg_types)
            for name, t, e in p["consts"]:
                self.pkg_consts[name] = coerce(self.eval(e, sc), self.resolve_type(t, sc))

    # ---- types
    def resolve_type(self, t, sc):
        _, name, con = t
        if name == "std_logic" or name == "std_ulogic":
            return ("bit",)
        if name in ("std_logic_vector", "signed", "unsigned"):
            kind = {"std_logic_vector": "slv"}.get(name, name)
            hi, lo = self.eval(con[0], sc), self.eval(con[1], sc)
            return ("vec", kind, hi, lo)
        if name == "integer":
            return ("int",)
        tdef = sc.types.get(name)
        if tdef is None:
            raise NameError(f"type {name}")
        if tdef[0] == "enumdef":
            return ("enum", tdef[1])
        lo, hi = self.eval(tdef[1][0], sc), self.eval(tdef[1][1], sc)
        return ("array", lo, hi, self.resolve_type(tdef[2], sc))

    # ---- elaboration
    def elaborate(self, entity, prefix="", generics=None):
        ent, arch = self.entities[entity], self.archs[entity]
        types = dict(self.pkg_types)
        types.update(arch["types"])
        consts = dict(self.pkg_consts)
        for lits in (t[1] for t in types.values() if t[0] == "enumdef"):
            for lit in lits:
                consts[lit] = Enum(lit)
        sc = Scope(self, prefix, consts, types)
        for name, t, d in ent["generics"]:
            if generics and name in generics:
                consts[name] = generics[name]
            elif d is not None:
                consts[name] = self.eval(d, sc)
        for name, t, e in arch["consts"]:
            consts[name] = coerce(self.eval(e, sc), self.resolve_type(t, sc))
        for name, mode, t in ent["ports"]:
            rt = self.resolve_type(t, sc)
            self.sigtypes[prefix + name] = rt
            self.signals[prefix + name] = default_value(rt)
        for name, t, init in arch["signals"]:
            rt = self.resolve_type(t, sc)
            self.sigtypes[prefix + name] = rt
            self.signals[prefix + name] = coerce(self.eval(init, sc), rt) if init is not None else default_value(rt)
        self._elab_stmts(arch["stmts"], sc)
        return sc

    def _elab_stmts(self, stmts, sc):
        for st in stmts:
            k = st[0]
            if k == "assign":
                self.assigns.append((sc, st[1], st[2]))
            elif k == "process":
                self.procs.append((sc, st[2]))
            elif k == "ifgen":
                if self.eval(st[1], sc) is True:
                    self._elab_stmts(st[2], sc)
            elif k == "forgen":
                for i in range(self.eval(st[2], sc), self.eval(st[3], sc) + 1):
                    self._elab_stmts(st[4], sc.child({st[1]: i}))
            elif k == "inst":
                _, label, ent, gmap, pmap = st
                child_prefix = sc.prefix + label + "."
                generics = {f: self.eval(e, sc) for f, e in gmap}
                child = self.elaborate(ent, child_prefix, generics)
                modes = {n: m for n, m, _ in self.entities[ent]["ports"]}
                for formal, actual in pmap:
                    if actual is None:
                        continue
                    if modes[formal] == "in":
                        # child's port <= actual, evaluated in the parent's scope
                        self.assigns.append((("port_in", child, sc), (formal, None), actual))
                    else:
                        self.assigns.append((("port_out", sc, child), self._as_target(actual), ("name", formal)))

    @staticmethod
    def _as_target(e):
        if e[0] == "name":
            return (e[1], None)
        if e[0] == "call" and e[1][0] == "name" and len(e[2]) == 1:
            return (e[1][1], e[2][0])
        raise SyntaxError(f"unsupported port actual {e}")

    # ---- expression evaluation
    def lookup(self, name, sc):
        if name in sc.consts:
            return sc.consts[name]
        full = sc.prefix + name
        if full in self.signals:
            return self.signals[full]
        raise NameError(f"{name} (scope {sc.prefix!r})")

    def eval(self, e, sc):
        k = e[0]
        if k == "int":
            return e[1]
        if k == "bit":
            return Bit(e[1])
        if k == "vec":
            return Vec(e[1], e[2], e[3])
        if k == "name":
            if e[1] == "true":
                return True
            if e[1] == "false":
                return False
            return self.lookup(e[1], sc)
        if k == "agg":
            named = {self.eval(c, sc): self.eval(v, sc) for c, v in e[1]}
            return Agg(named, self.eval(e[2], sc) if e[2] is not None else None)
        if k == "neg":
            v = self.eval(e[1], sc)
            return -v if isinstance(v, int) else Vec(v.kind, v.w, -v.v)
        if k == "not":
            v = self.eval(e[1], sc)
            if isinstance(v, bool):
                return not v
            if isinstance(v, Bit):
                return Bit(1 - v.v)
            return Vec(v.kind, v.w, ~v.v)
        if k == "bin":
            return self.binop(e[1], self.eval(e[2], sc), self.eval(e[3], sc))
        if k == "call":
            return self.call(e, sc)
        raise ValueError(e)

    @staticmethod
    def binop(op, a, b):
        if op in ("and", "or", "xor", "nand", "nor"):
            if isinstance(a, bool):
                r = {"and": a and b, "or": a or b, "xor": a != b, "nand": not (a and b), "nor": not (a or b)}[op]
                return r
            f = {"and": lambda x, y: x & y, "or": lambda x, y: x | y, "xor": lambda x, y: x ^ y,
                 "nand": lambda x, y: ~(x & y), "nor": lambda x, y: ~(x | y)}[op]
            if isinstance(a, Bit):
                return Bit(f(a.v, b.v))
            return Vec(a.kind, a.w, f(a.v, b.v))
        if op in ("=", "/=", "<", ">", "<=", ">="):
            def key(x):
                if isinstance(x, Vec):
                    return x.num()
                if isinstance(x, Bit):
                    return x.v
                if isinstance(x, Enum):
                    return x.name
                return x
            x, y = key(a), key(b)
            return {"=": x == y, "/=": x != y, "<": x < y, ">": x > y, "<=": x <= y, ">=": x >= y}[op]
        if op in ("+", "-"):
            sgn = 1 if op == "+" else -1
            if isinstance(a, int) and isinstance(b, int):
                return a + sgn * b
            if isinstance(a, Vec) and isinstance(b, Vec):
                if a.kind != b.kind:
                    raise TypeError("mixed vector kinds in +/-")
                w = max(a.w, b.w)                       # numeric_std: result length = max of the operands
                return Vec(a.kind, w, a.num() + sgn * b.num())
            if isinstance(a, Vec) and isinstance(b, Bit):   # numeric_std (2008): the bit counts as 0 / 1
                return Vec(a.kind, a.w, a.num() + sgn * b.v)
            if isinstance(a, Vec) and isinstance(b, int):
                return Vec(a.kind, a.w, a.num() + sgn * b)
            if isinstance(a, int) and isinstance(b, Vec):
                return Vec(b.kind, b.w, a + sgn * b.num())
            raise TypeError(f"{op} on {a!r}, {b!r}")
        if op == "*":
            if isinstance(a, int) and isinstance(b, int):
                return a * b
            if isinstance(a, Vec) and isinstance(b, Vec) and a.kind == b.kind:
                return Vec(a.kind, a.w + b.w, a.num() * b.num())      # numeric_std: length = L + R
            raise TypeError(f"* on {a!r}, {b!r}")
        if op == "mod":
            return a % b
        if op == "rem":
            return int(abs(a) % abs(b)) * (1 if a >= 0 else -1)
        if op == "/":
            return int(a / b)
        if op == "&":
            def bits(x):
                return (x.v, x.w) if isinstance(x, Vec) else (x.v, 1)
            (av, aw), (bv, bw) = bits(a), bits(b)
            kind = a.kind if isinstance(a, Vec) else b.kind
            return Vec(kind, aw + bw, (av << bw) | bv)
        raise ValueError(op)

    def call(self, e, sc):
        _, head, args = e
        if head[0] == "name":
            name = head[1]
            # conversions and numeric_std functions
            if name in ("signed", "unsigned", "std_logic_vector"):
                v = self.eval(args[0], sc)
                return Vec({"std_logic_vector": "slv"}.get(name, name), v.w, v.v)
            if name == "to_signed":
                return Vec("signed", self.eval(args[1], sc), self.eval(args[0], sc))
            if name == "to_unsigned":
                return Vec("unsigned", self.eval(args[1], sc), self.eval(args[0], sc))
            if name == "to_integer":
                return self.eval(args[0], sc).num()
            if name == "resize":
                v, n = self.eval(args[0], sc), self.eval(args[1], sc)
                if v.kind == "signed":
                    if n >= v.w:
                        return Vec("signed", n, v.sint())
                    sign = v.v >> (v.w - 1)              # numeric_std RESIZE: sign bit + (n - 1) rightmost bits
                    return Vec("signed", n, (sign << (n - 1)) | (v.v & ((1 << (n - 1)) - 1)))
                return Vec(v.kind, n, v.v)
            if name == "rising_edge":
                return True
            base = None
            try:
                base = self.lookup(name, sc)
            except NameError:
                raise NameError(f"unknown function or object {name}")
        else:
            base = self.eval(head, sc)
        a0 = args[0]
        if a0[0] == "range":
            hi, lo = self.eval(a0[1], sc), self.eval(a0[2], sc)
            if not isinstance(base, Vec):
                raise TypeError("slice of non-vector")
            return Vec(base.kind, hi - lo + 1, base.v >> lo)
        idx = self.eval(a0, sc)
        if isinstance(base, Vec):
            return Bit((base.v >> idx) & 1)
        if isinstance(base, Arr):
            return base.el[idx - base.lo]
        raise TypeError(f"cannot index {base!r}")

    # ---- simulation
    def _store(self, full, idx, val, pending=None):
        """immediate store (concurrent assignments; returns whether the signal changed) or, with
        `pending`, a scheduled one: element-granular, because the elements of one array signal may
        be driven by different processes (ve(0), ve(1), ve(2) of the biquad)."""
        t = self.sigtypes[full]
        if idx is None:
            new = coerce(val, t)
        elif t[0] == "array":
            new = coerce(val, t[3])
        elif t[0] == "vec":
            new = coerce(val, ("bit",))
        else:
            raise TypeError("indexed store into scalar")
        if pending is not None:
            pending[(full, idx)] = new
            return True
        return self._commit(full, idx, new)

    def _commit(self, full, idx, new):
        t = self.sigtypes[full]
        cur = self.signals[full]
        if idx is None:
            pass
        elif t[0] == "array":
            arr = cur.copy()
            arr.el[idx - t[1]] = new
            new = arr
        else:
            new = Vec(cur.kind, cur.w, (cur.v & ~(1 << (idx - t[3]))) | (new.v << (idx - t[3])))
        changed = not (cur == new)
        self.signals[full] = new
        return changed

    def settle(self):
        for _ in range(64):
            changed = False
            for sc, (name, idx_e), e in self.assigns:
                if isinstance(sc, tuple):
                    kind, a, b = sc
                    if kind == "port_in":            # a = child scope (target), b = parent scope (expression)
                        val = self.eval(e, b)
                        changed |= self._store(a.prefix + name, None, val)
                    else:                             # port_out: a = parent scope (target), b = child scope
                        val = self.eval(e, b)
                        idx = self.eval(idx_e, a) if idx_e is not None else None
                        changed |= self._store(a.prefix + name, idx, val)
                else:
                    val = self.eval(e, sc)
                    idx = self.eval(idx_e, sc) if idx_e is not None else None
                    changed |= self._store(sc.prefix + name, idx, val)
            if not changed:
                return
        raise RuntimeError("combinational loop did not settle")

    def _run(self, body, sc, pending):
        for st in body:
            if st[0] == "assign":
                (name, idx_e), e = st[1], st[2]
                idx = self.eval(idx_e, sc) if idx_e is not None else None
                self._store(sc.prefix + name, idx, self.eval(e, sc), pending)
            else:
                _, branches, els = st
                for cond, b in branches:
                    c = self.eval(cond, sc)
                    if c is True:
                        self._run(b, sc, pending)
                        break
                else:
                    self._run(els, sc, pending)

    def set(self, name, value):
        self.signals[name] = coerce(value, self.sigtypes[name])

    def get(self, name):
        return self.signals[name]

    def tick(self):
        """one rising clock edge: every process sees the pre-edge values"""
        updates = {}
        for sc, body in self.procs:
            self._run(body, sc, updates)
        for (full, idx), v in updates.items():
            self._commit(full, idx, v)
        self.settle()
