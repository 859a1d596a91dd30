#!/usr/bin/env python
"""f90toc.py -- mechanical translator from the free-form Fortran 90 subset that
ericmjalbert/PDE-ODEsolver is written in to C (TEST INFRASTRUCTURE).

Why: this image has no Fortran compiler (no gfortran / f951 / flang), so the
reference cannot be built the way runAll.sh:6 builds it.  This tool reads the
reference's .f90 files WHERE THEY LIE (/root/reference, never copied into the
repo), translates them statement by statement following the language rules --
it knows nothing about the algorithm -- and writes one C file into
oracle/_ref/ (git-ignored), which oracle/Makefile compiles with gcc -O3 into

    oracle/_ref/ref_serial, ref_omp      the reference program (runAll.sh:6 / runAll_paral.sh:6)
    oracle/_ref/libref.so, libref_omp.so every subroutine, Fortran calling convention
                                         (by reference, lower-case name + '_')

so that the hand-written oracle (oracle/pdeode_oracle.c) and the CUDA path can
be checked against the reference's own source text instead of against a
reading of it.

Language rules implemented (gfortran -O3 -fdefault-real-8 semantics):
  * `real` and real literals are 8 bytes (-fdefault-real-8); `integer` is int32.
  * real ** integer  ->  __builtin_powi, the very builtin gfortran lowers it to
    (libgcc __powidf2 for a variable exponent, multiplication chain for a constant).
  * integer / integer truncates; INT() truncates; MOD is C's %.
  * sum(), maxval() and whole-array assignments are scalarised into one serial
    loop in element order, as gfortran's scalariser does (no reassociation
    without -ffast-math).
  * arguments are passed by reference; scalar dummies are copied in and out.
  * automatic and allocatable arrays come from calloc: what the reference reads
    before setting (Mnew, maxIters, maxNit, total) is zero, as fresh
    pages are in practice.
  * !$omp directives become #pragma omp (only compiled in with -fopenmp).
  * I/O, system_clock, random_seed / random_number go to a small runtime (f90rt.c).
  * ISO_C_BINDING (for a host that keeps its Fortran and calls a C library): a module made
    of `parameter` constants, bind(C) derived types and an interface block of bind(C)
    functions is read, not translated -- its functions become C prototypes by the
    interoperability rules (a VALUE dummy by value, every other dummy a pointer to the
    interoperable type; type(c_ptr) = void *; a bind(C) type = the struct with the same
    components in order); `use` of it makes the names visible; kinds c_int, c_long_long,
    c_double; structure constructors become compound literals; iand() is `&`.

Usage: f90toc.py -o out.c file1.f90 file2.f90 ...
"""
import re
import sys

# ----------------------------------------------------------------- source -> statements


def strip_comment(line):
    """Cut a trailing ! comment (not inside a string)."""
    q = None
    for k, ch in enumerate(line):
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == '!':
            return line[:k]
    return line


def split_semicolons(text):
    out, cur, q = [], [], None
    for ch in text:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            cur.append(ch)
        elif ch == ';':
            out.append(''.join(cur))
            cur = []
        else:
            cur.append(ch)
    out.append(''.join(cur))
    return [s.strip() for s in out if s.strip()]


def read_statements(path):
    """-> list of (kind, text, lineno); kind 'stmt' or 'omp'."""
    stmts = []
    pend, pend_line = None, 0
    with open(path) as f:
        for ln, raw in enumerate(f, 1):
            s = raw.rstrip('\n')
            st = s.strip()
            low = st.lower()
            if low.startswith('!$omp'):
                body = st[5:].strip()
                if pend is not None and pend_kind == 'omp':
                    pend += ' ' + body
                else:
                    pend, pend_line, pend_kind = body, ln, 'omp'
                if pend.endswith('&'):
                    pend = pend[:-1].rstrip()
                    continue
                stmts.append(('omp', pend, pend_line))
                pend = None
                continue
            s = strip_comment(s).strip()
            if not s:
                continue
            if s.startswith('&'):
                s = s[1:].lstrip()
            if pend is not None:
                pend += ' ' + s
            else:
                pend, pend_line, pend_kind = s, ln, 'stmt'
            if pend.endswith('&'):
                pend = pend[:-1].rstrip()
                continue
            for part in split_semicolons(pend):
                stmts.append(('stmt', part, pend_line))
            pend = None
    return stmts


# ----------------------------------------------------------------- lexer

TOK = re.compile(r"""
    (?P<ws>\s+)
  | (?P<str>'(?:[^']|'')*'|"(?:[^"]|"")*")
  | (?P<dotop>\.(?:le|lt|ge|gt|eq|ne|and|or|not|true|false)\.)
  | (?P<num>(?:\d+\.\d*|\.\d+|\d+)(?:[edED][+-]?\d+)?)
  | (?P<name>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<op>\*\*|==|/=|<=|>=|::|[-+*/(),=<>:%])
""", re.X | re.I)


def lex(text):
    toks, pos = [], 0
    while pos < len(text):
        # a number directly followed by a dotted operator: "1.le.x" -- keep "1" and ".le."
        m = TOK.match(text, pos)
        if not m:
            raise SyntaxError(f"cannot lex at {text[pos:pos+20]!r} in {text!r}")
        kind = m.lastgroup
        val = m.group(kind)
        if kind == 'num' and '.' in val and not re.search(r'[edED]', val):
            # "0 .LE." is fine; "x.le." cannot happen; but "1.and." -> num grabbed "1." wrongly
            rest = text[m.start() + val.index('.'):]
            if re.match(r'\.(?:le|lt|ge|gt|eq|ne|and|or|not)\.', rest, re.I):
                val = val[:val.index('.')]
                toks.append(('num', val))
                pos = m.start() + len(val)
                continue
        pos = m.end()
        if kind == 'ws':
            continue
        if kind == 'name':
            toks.append(('name', val.lower()))
        elif kind == 'dotop':
            toks.append(('op', val.lower()))
        else:
            toks.append((kind, val))
    return toks


# ----------------------------------------------------------------- expressions

class Node:
    def __init__(self, k, **kw):
        self.k = k
        self.__dict__.update(kw)


class Parser:
    def __init__(self, toks):
        self.t = toks
        self.i = 0

    def peek(self, off=0):
        j = self.i + off
        return self.t[j] if j < len(self.t) else ('eof', '')

    def next(self):
        tk = self.peek()
        self.i += 1
        return tk

    def accept(self, val):
        if self.peek()[1] == val and self.peek()[0] in ('op', 'name'):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise SyntaxError(f"expected {val!r}, got {self.peek()} in {self.t}")

    def at_end(self):
        return self.i >= len(self.t)

    # precedence climbing: .or. < .and. < .not. < relational < +- < */ < unary < **
    def expr(self):
        return self.p_or()

    def p_or(self):
        a = self.p_and()
        while self.peek() == ('op', '.or.'):
            self.next()
            a = Node('bin', op='||', a=a, b=self.p_and())
        return a

    def p_and(self):
        a = self.p_not()
        while self.peek() == ('op', '.and.'):
            self.next()
            a = Node('bin', op='&&', a=a, b=self.p_not())
        return a

    def p_not(self):
        if self.peek() == ('op', '.not.'):
            self.next()
            return Node('un', op='!', a=self.p_not())
        return self.p_rel()

    REL = {'.le.': '<=', '.lt.': '<', '.ge.': '>=', '.gt.': '>', '.eq.': '==', '.ne.': '!=',
           '<=': '<=', '<': '<', '>=': '>=', '>': '>', '==': '==', '/=': '!='}

    def p_rel(self):
        a = self.p_add()
        tk = self.peek()
        if tk[0] == 'op' and tk[1] in self.REL:
            self.next()
            return Node('bin', op=self.REL[tk[1]], a=a, b=self.p_add(), rel=True)
        return a

    def p_add(self):
        tk = self.peek()
        if tk == ('op', '-'):
            self.next()
            a = Node('un', op='-', a=self.p_mul())
        elif tk == ('op', '+'):
            self.next()
            a = self.p_mul()
        else:
            a = self.p_mul()
        while self.peek() in (('op', '+'), ('op', '-')):
            op = self.next()[1]
            a = Node('bin', op=op, a=a, b=self.p_mul())
        return a

    def p_mul(self):
        a = self.p_pow()
        while True:
            tk = self.peek()
            if tk == ('op', '*'):
                self.next()
                a = Node('bin', op='*', a=a, b=self.p_pow())
            elif tk == ('op', '/') and self.peek(1) != ('op', ')'):     # "/)" closes a constructor
                self.next()
                a = Node('bin', op='/', a=a, b=self.p_pow())
            else:
                return a

    def p_pow(self):
        a = self.p_primary()
        if self.peek() == ('op', '**'):
            self.next()
            # right associative; a unary minus may follow **
            if self.peek() == ('op', '-'):
                self.next()
                b = Node('un', op='-', a=self.p_pow())
            else:
                b = self.p_pow()
            return Node('pow', a=a, b=b)
        return a

    def p_primary(self):
        kind, val = self.next()
        if kind == 'num':
            isreal = bool(re.search(r'[.edED]', val))
            return Node('num', v=val, real=isreal)
        if kind == 'str':
            q = val[0]
            return Node('str', v=val[1:-1].replace(q + q, q))
        if kind == 'op' and val in ('.true.', '.false.'):
            return Node('num', v='1' if val == '.true.' else '0', real=False)
        if kind == 'op' and val == '(':
            if self.peek() == ('op', '/'):          # array constructor (/ ... /)
                self.next()
                items = []
                while True:
                    items.append(self.cons_item())
                    if self.accept(','):
                        continue
                    break
                self.expect('/')
                self.expect(')')
                return Node('cons', items=items)
            e = self.expr()
            self.expect(')')
            return Node('paren', a=e)
        if kind == 'name':
            if self.peek() == ('op', '('):
                self.next()
                args = []
                if not self.accept(')'):
                    while True:
                        args.append(self.arg())
                        if self.accept(','):
                            continue
                        self.expect(')')
                        break
                return Node('ref', name=val, args=args)
            return Node('ref', name=val, args=None)
        raise SyntaxError(f"unexpected token {kind} {val!r} in {self.t}")

    def arg(self):
        # section ':' or keyword argument name=expr, or expression
        if self.peek() == ('op', ':'):
            self.next()
            return Node('colon')
        if self.peek()[0] == 'name' and self.peek(1) == ('op', '='):
            kw = self.next()[1]
            self.next()
            return Node('kw', name=kw, a=self.expr())
        return self.expr()

    def cons_item(self):
        # implied do: ( expr, var = lo, hi )
        if self.peek() == ('op', '('):
            save = self.i
            try:
                self.next()
                e = self.expr()
                self.expect(',')
                kind, var = self.next()
                if kind != 'name':
                    raise SyntaxError('not an implied do')
                self.expect('=')
                lo = self.expr()
                self.expect(',')
                hi = self.expr()
                self.expect(')')
                return Node('ido', e=e, var=var, lo=lo, hi=hi)
            except SyntaxError:
                self.i = save
        return self.expr()


def parse_expr(text):
    p = Parser(lex(text))
    e = p.expr()
    if not p.at_end():
        raise SyntaxError(f"trailing tokens in expression {text!r}")
    return e


def split_top(text, sep=','):
    """Split on top-level separators (outside parentheses and strings)."""
    out, cur, depth, q = [], [], 0, None
    for ch in text:
        if q:
            cur.append(ch)
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == '(':
            depth += 1
        elif ch == ')':
            depth -= 1
        if ch == sep and depth == 0:
            out.append(''.join(cur).strip())
            cur = []
        else:
            cur.append(ch)
    last = ''.join(cur).strip()
    if last or out:
        out.append(last)
    return out


def match_paren(text, start):
    """text[start] == '(' -> index of the matching ')'."""
    depth, q = 0, None
    for k in range(start, len(text)):
        ch = text[k]
        if q:
            if ch == q:
                q = None
            continue
        if ch in "'\"":
            q = ch
        elif ch == '(':
            depth += 1
        elif ch == ')':
            depth -= 1
            if depth == 0:
                return k
    raise SyntaxError(f"unbalanced parentheses in {text!r}")


# ----------------------------------------------------------------- program units

class Sym:
    def __init__(self, name):
        self.name = name
        self.type = None          # 'int' | 'real' | 'char'
        self.dims = None          # list of extent expression texts, or [':'] for allocatable
        self.dummy = False
        self.intent = None
        self.allocatable = False
        self.charlen = None       # expression text


class Unit:
    def __init__(self, kind, name, args, line, path):
        self.kind, self.name, self.args, self.line, self.path = kind, name, args, line, path
        self.syms = {}
        self.body = []            # (kind, text, lineno)


class _CType(dict):
    def __missing__(self, key):              # 'struct:name' -> the bind(C) type
        if key.startswith('struct:'):
            return 'struct ' + key[7:]
        raise KeyError(key)


CTYPE = _CType({'int': 'int', 'real': 'double', 'char': 'char', 'i64': 'long long', 'cptr': 'void *'})
KINDS = {('integer', 'c_int'): 'int', ('integer', 'c_long_long'): 'i64', ('real', 'c_double'): 'real',
         ('type', 'c_ptr'): 'cptr'}


def kind_type(base, sel):
    """integer(c_int) / real(c_double) / type(c_ptr) / type(name) -> type key."""
    sel = sel.strip().lower()
    sel = re.sub(r'^kind\s*=\s*', '', sel)
    if (base, sel) in KINDS:
        return KINDS[(base, sel)]
    if base == 'type':
        return 'struct:' + sel
    raise SyntaxError(f"unsupported kind {base}({sel})")


class Module:
    """A module of constants, bind(C) types and bind(C) function interfaces (see the header)."""
    def __init__(self, name, path):
        self.name, self.path = name, path
        self.consts = {}          # name -> (type, value text)
        self.types = {}           # name -> [(type, component)]
        self.funcs = {}           # fortran name -> (C name, result type, [(dummy, type, by_value, is_array, intent)])


def parse_module(name, body, path):
    m = Module(name, path)
    k = 0
    while k < len(body):
        kind, text, ln = body[k]
        low = text.lower()
        k += 1
        if re.match(r'(use\s|implicit\s+none\b|interface\b|end\s*interface\b|contains\b)', low):
            continue
        mt = re.match(r'type\s*,\s*bind\s*\(\s*c\s*\)\s*::\s*(\w+)$', low)
        if mt:
            comps = []
            while not re.match(r'end\s*type\b', body[k][1].lower()):
                d = DECL.match(body[k][1])
                base, rest = d.group(1).lower(), d.group(2).strip()
                e = match_paren(rest, 0)
                t = kind_type(base, rest[1:e])
                ents = rest[e + 1:].split('::', 1)[1]
                for ent in split_top(ents, ','):
                    comps.append((t, ent.split('=')[0].strip()))
                k += 1
            k += 1
            m.types[mt.group(1)] = comps
            continue
        mp = re.match(r'(integer|real)\s*\(\s*(\w+)\s*\)\s*,\s*parameter\s*::\s*(\w+)\s*=\s*(.+)$', text, re.I)
        if mp:
            m.consts[mp.group(3).lower()] = (kind_type(mp.group(1).lower(), mp.group(2)), mp.group(4).strip())
            continue
        mf = re.match(r'(integer|real)\s*\(\s*(\w+)\s*\)\s*function\s+(\w+)\s*\((.*?)\)\s*'
                      r'bind\s*\(\s*c\s*,\s*name\s*=\s*["\'](\w+)["\']\s*\)$', text, re.I)
        if mf:
            ret = kind_type(mf.group(1).lower(), mf.group(2))
            dummies = [a.strip().lower() for a in mf.group(4).split(',') if a.strip()]
            decl = {}
            while not re.match(r'end\s*function\b', body[k][1].lower()):
                t2 = body[k][1]
                k += 1
                if re.match(r'import\b', t2, re.I):
                    continue
                d = DECL.match(t2)
                base, rest = d.group(1).lower(), d.group(2).strip()
                e = match_paren(rest, 0)
                t = kind_type(base, rest[1:e])
                attrs, ents = rest[e + 1:].split('::', 1)
                attrs = [a.lower().replace(' ', '') for a in split_top(attrs.strip().lstrip(','), ',') if a]
                intent = next((a[7:-1] for a in attrs if a.startswith('intent(')), None)
                for ent in split_top(ents, ','):
                    decl[ent.strip().lower()] = (t, 'value' in attrs, any(a.startswith('dimension(') for a in attrs), intent)
            k += 1
            m.funcs[mf.group(3).lower()] = (mf.group(5), ret, [(a,) + decl[a] for a in dummies])
            continue
        raise SyntaxError(f"{path}:{ln}: unsupported statement in module {name}: {text}")
    return m


INTRINSICS = {'abs', 'sqrt', 'real', 'int', 'mod', 'max0', 'min0', 'max', 'min', 'sum', 'maxval',
              'len'}


def parse_units(stmts, path):
    units, cur, in_iface = [], None, False
    mod_name, mod_body = None, []
    for kind, text, ln in stmts:
        low = text.lower()
        if kind == 'stmt' and cur is None:
            mm = re.match(r'module\s+(\w+)\s*$', low)
            if mm and mod_name is None:
                mod_name, mod_body = mm.group(1), []
                continue
            if mod_name is not None:
                if re.match(r'end\s*module\b', low):
                    units.append(parse_module(mod_name, mod_body, path))
                    mod_name = None
                else:
                    mod_body.append((kind, text, ln))
                continue
        if kind == 'stmt':
            if re.match(r'interface\b', low):
                in_iface = True
                continue
            if re.match(r'end\s*interface\b', low):
                in_iface = False
                continue
            if in_iface:
                continue
            m = re.match(r'program\s+(\w+)', low)
            if m and cur is None:
                cur = Unit('program', m.group(1), [], ln, path)
                continue
            m = re.match(r'subroutine\s+(\w+)\s*(?:\((.*)\))?\s*$', low)
            if m and cur is None:
                args = [a.strip() for a in (m.group(2) or '').split(',') if a.strip()]
                cur = Unit('subroutine', m.group(1), args, ln, path)
                continue
            if re.match(r'end\s*(program|subroutine)?\b(\s+\w+)?\s*$', low) and \
                    not re.match(r'end\s*(if|do)\b', low) and cur is not None:
                units.append(cur)
                cur = None
                continue
        if cur is None:
            raise SyntaxError(f"{path}:{ln}: statement outside a program unit: {text}")
        cur.body.append((kind, text, ln))
    if cur is not None:
        raise SyntaxError(f"{path}: unterminated program unit {cur.name}")
    return units


DECL = re.compile(r'^(integer|real|character|type)\b(.*)$', re.I)


def parse_decl(u, text):
    """A type declaration statement -> symbols.  Returns False if not a declaration."""
    m = DECL.match(text)
    if not m:
        return False
    base = m.group(1).lower()
    rest = m.group(2).strip()
    charlen = None
    ktype = None
    if base == 'type' and not rest.startswith('('):
        return False                      # `type, bind(C) :: name` is a module matter
    if base in ('integer', 'real', 'type') and rest.startswith('('):
        e = match_paren(rest, 0)
        ktype = kind_type(base, rest[1:e])
        rest = rest[e + 1:].strip()
    if base == 'character':
        charlen = '1'
        if rest.startswith('('):
            e = match_paren(rest, 0)
            inner = rest[1:e].strip()
            inner = re.sub(r'^len\s*=\s*', '', inner, flags=re.I)
            charlen = inner
            rest = rest[e + 1:].strip()
    if '::' in rest:
        attrs, ents = rest.split('::', 1)
    else:
        attrs, ents = '', rest
    attrs = [a for a in split_top(attrs.strip().lstrip(','), ',') if a]
    dims, intent, alloc = None, None, False
    for a in attrs:
        al = a.lower().replace(' ', '')
        if al.startswith('dimension('):
            dims = split_top(a[a.index('(') + 1:a.rindex(')')], ',')
        elif al.startswith('intent('):
            intent = al[7:-1]
        elif al == 'allocatable':
            alloc = True
        else:
            raise SyntaxError(f"unsupported attribute {a!r} in {text!r}")
    for ent in split_top(ents, ','):
        if not ent:
            continue
        m2 = re.match(r'(\w+)\s*(?:\((.*)\))?\s*$', ent)
        if not m2:
            raise SyntaxError(f"cannot parse entity {ent!r} in {text!r}")
        name = m2.group(1).lower()
        s = u.syms.setdefault(name, Sym(name))
        s.type = ktype or {'integer': 'int', 'real': 'real', 'character': 'char'}[base]
        s.dims = split_top(m2.group(2), ',') if m2.group(2) else dims
        s.intent = intent
        s.allocatable = alloc
        s.charlen = charlen
    return True


# ----------------------------------------------------------------- code generation

class Gen:
    def __init__(self, units):
        self.modules = {u.name: u for u in units if isinstance(u, Module)}
        self.units = {u.name: u for u in units if not isinstance(u, Module)}
        self.out = []
        self.u = None
        self.ind = 1
        self.tmp = 0

    def emit(self, s):
        self.out.append('    ' * self.ind + s)

    # ---- names
    def v(self, name):
        return name + '_'

    def used(self, what, name):
        """(module entity) the current unit sees through `use`: what = consts | types | funcs."""
        for mn in getattr(self.u, 'uses', ()):
            d = getattr(self.modules[mn], what)
            if name in d:
                return d[name]
        return None

    def sym(self, name):
        s = self.u.syms.get(name)
        if s is None:
            raise SyntaxError(f"{self.u.name}: undeclared name {name!r}")
        return s

    def is_array(self, name):
        s = self.u.syms.get(name)
        return s is not None and s.dims is not None

    def size_of(self, s):
        """Total element count of array symbol s, as C text."""
        if s.dims == [':']:
            return f"{self.v(s.name)}n"
        return ' * '.join(f"(long)({self.cx(parse_expr(d))})" for d in s.dims)

    # ---- types
    def typeof(self, e):
        k = e.k
        if k == 'num':
            return 'real' if e.real else 'int'
        if k == 'str':
            return 'char'
        if k == 'paren':
            return self.typeof(e.a)
        if k == 'un':
            return 'int' if e.op == '!' else self.typeof(e.a)
        if k == 'bin':
            if e.op in ('&&', '||') or getattr(e, 'rel', False):
                return 'int'
            ta, tb = self.typeof(e.a), self.typeof(e.b)
            return 'real' if 'real' in (ta, tb) else ('i64' if 'i64' in (ta, tb) else 'int')
        if k == 'pow':
            return self.typeof(e.a)
        if k == 'cons':
            it = e.items[0]
            return self.typeof(it.e if it.k == 'ido' else it)
        if k == 'ref':
            if e.name in self.u.syms:
                return self.sym(e.name).type
            if self.used('consts', e.name):
                return self.used('consts', e.name)[0]
            if self.used('funcs', e.name):
                return self.used('funcs', e.name)[1]
            if self.used('types', e.name):
                return 'struct:' + e.name
            if e.name == 'iand':
                return 'int'
            if e.name in ('real', 'sqrt'):
                return 'real'
            if e.name in ('int', 'mod', 'max0', 'min0', 'len'):
                return 'int' if e.name != 'mod' else self.typeof(e.args[0])
            if e.name in ('abs', 'sum', 'maxval', 'max', 'min'):
                return self.typeof(e.args[0])
            raise SyntaxError(f"{self.u.name}: unknown function {e.name!r}")
        raise SyntaxError(f"typeof: {k}")

    # ---- array-ness of an expression: returns the size text of the first whole array in it
    def extent(self, e):
        k = e.k
        if k in ('num', 'str', 'colon'):
            return None
        if k in ('paren', 'un'):
            return self.extent(e.a)
        if k in ('bin', 'pow'):
            return self.extent(e.a) or self.extent(e.b)
        if k == 'cons':
            if len(e.items) == 1 and e.items[0].k == 'ido':
                it = e.items[0]
                return f"((long)({self.cx(it.hi)}) - (long)({self.cx(it.lo)}) + 1)"
            return str(len(e.items))
        if k == 'ref':
            if e.name in self.u.syms and self.is_array(e.name):
                if e.args is None or all(a.k == 'colon' for a in e.args):
                    return self.size_of(self.sym(e.name))
                return None
            if e.name in ('sum', 'maxval'):
                return None                       # reductions are scalars
            if e.args:
                for a in e.args:
                    x = self.extent(a)
                    if x:
                        return x
            return None
        return None

    # ---- expression -> C.  idx: text of the 0-based flat element index when scalarising.
    def cx(self, e, idx=None):
        k = e.k
        if k == 'num':
            v = re.sub(r'[dD]', 'e', e.v)
            return v
        if k == 'str':
            return '"' + e.v.replace('\\', '\\\\').replace('"', '\\"') + '"'
        if k == 'paren':
            return '(' + self.cx(e.a, idx) + ')'
        if k == 'un':
            return f"({e.op}{self.cx(e.a, idx)})"
        if k == 'bin':
            return f"({self.cx(e.a, idx)} {e.op} {self.cx(e.b, idx)})"
        if k == 'pow':
            ta, tb = self.typeof(e.a), self.typeof(e.b)
            if tb != 'int':
                return f"pow({self.cx(e.a, idx)}, {self.cx(e.b, idx)})"
            if ta == 'real':
                return f"__builtin_powi({self.cx(e.a, idx)}, {self.cx(e.b, idx)})"
            return f"f90_ipow({self.cx(e.a, idx)}, {self.cx(e.b, idx)})"
        if k == 'cons':
            if idx is None:
                raise SyntaxError("array constructor in scalar context")
            if len(e.items) == 1 and e.items[0].k == 'ido':
                it = e.items[0]
                return (f"({{ int {self.v(it.var)} = ({self.cx(it.lo)}) + (int)({idx}); "
                        f"({self.cx(it.e)}); }})")
            ct = CTYPE[self.typeof(e)]
            return f"(({ct}[]){{{', '.join(self.cx(i) for i in e.items)}}})[{idx}]"
        if k == 'ref':
            return self.cref(e, idx)
        raise SyntaxError(f"cx: {k}")

    def elem_index(self, s, args):
        """0-based flat (column-major) index text for s(args)."""
        if len(args) != len(s.dims):
            raise SyntaxError(f"{self.u.name}: rank mismatch for {s.name}")
        terms, stride = [], None
        for d, a in zip(s.dims, args):
            t = f"((long)({self.cx(a)}) - 1)"
            terms.append(t if stride is None else f"{t} * ({stride})")
            ext = f"(long)({self.cx(parse_expr(d))})" if d != ':' else f"{self.v(s.name)}n"
            stride = ext if stride is None else f"{stride} * {ext}"
        return ' + '.join(terms)

    def cref(self, e, idx):
        name = e.name
        if name in self.u.syms:
            s = self.sym(name)
            if s.dims is not None:
                if e.args is None or all(a.k == 'colon' for a in e.args):
                    if idx is None:
                        raise SyntaxError(f"{self.u.name}: whole array {name} in scalar context")
                    return f"{self.v(name)}[{idx}]"
                return f"{self.v(name)}[{self.elem_index(s, e.args)}]"
            if e.args is not None:
                raise SyntaxError(f"{self.u.name}: {name} is not an array")
            return self.v(name)
        a = e.args or []
        if self.used('consts', name):
            return '(' + self.used('consts', name)[1] + ')'
        if self.used('funcs', name):
            cname, _, dummies = self.used('funcs', name)
            if len(a) != len(dummies):
                raise SyntaxError(f"{self.u.name}: {name} takes {len(dummies)} arguments")
            return f"{cname}({', '.join(self.cx(x) if d[2] else self.actual(x) for x, d in zip(a, dummies))})"
        if self.used('types', name):                  # structure constructor, components in order
            comps = self.used('types', name)
            if len(a) != len(comps):
                raise SyntaxError(f"{self.u.name}: {name}(...) needs {len(comps)} components")
            return f"((struct {name}){{{', '.join(self.cx(x) for x in a)}}})"
        if name == 'iand':
            return f"(({self.cx(a[0], idx)}) & ({self.cx(a[1], idx)}))"
        if name == 'abs':
            f = 'fabs' if self.typeof(a[0]) == 'real' else 'abs'
            return f"{f}({self.cx(a[0], idx)})"
        if name == 'sqrt':
            return f"sqrt({self.cx(a[0], idx)})"
        if name == 'real':
            return f"((double)({self.cx(a[0], idx)}))"
        if name == 'int':
            return f"((int)({self.cx(a[0], idx)}))"
        if name == 'mod':
            if self.typeof(a[0]) == 'real' or self.typeof(a[1]) == 'real':
                return f"fmod({self.cx(a[0], idx)}, {self.cx(a[1], idx)})"
            return f"(({self.cx(a[0], idx)}) % ({self.cx(a[1], idx)}))"
        if name in ('max0', 'max'):
            return f"F90_MAX({self.cx(a[0], idx)}, {self.cx(a[1], idx)})"
        if name in ('min0', 'min'):
            return f"F90_MIN({self.cx(a[0], idx)}, {self.cx(a[1], idx)})"
        if name == 'len':
            return f"({self.charlen(a[0])})"
        if name in ('sum', 'maxval'):
            ext = self.extent(a[0])
            if ext is None:
                raise SyntaxError(f"{name}() of a scalar")
            ct = CTYPE[self.typeof(a[0])]
            self.tmp += 1
            k = f"k{self.tmp}_"
            body = self.cx(a[0], k)
            if name == 'sum':          # serial, element order, starting from 0
                return (f"({{ {ct} acc_ = 0; for (long {k} = 0; {k} < {ext}; {k}++) "
                        f"acc_ = acc_ + {body}; acc_; }})")
            return (f"({{ {ct} acc_ = 0; int first_ = 1; for (long {k} = 0; {k} < {ext}; {k}++) "
                    f"{{ {ct} t_ = {body}; if (first_ || t_ > acc_) {{ acc_ = t_; first_ = 0; }} }} acc_; }})")
        raise SyntaxError(f"{self.u.name}: unknown function {name!r}")

    def charlen(self, e):
        if e.k == 'str':
            return str(len(e.v))
        if e.k == 'ref' and e.name in self.u.syms and self.sym(e.name).type == 'char':
            return self.cx(parse_expr(self.sym(e.name).charlen))
        raise SyntaxError("len() of a non-character")

    # ---- actual argument -> pointer text
    def actual(self, e):
        if e.k == 'ref' and e.name in self.u.syms:
            s = self.sym(e.name)
            if s.dims is not None:
                if e.args is None:
                    return self.v(e.name)
                return f"&{self.v(e.name)}[{self.elem_index(s, e.args)}]"
            if s.type == 'char':
                return self.v(e.name)
            return f"&{self.v(e.name)}"
        if e.k == 'str':
            return self.cx(e)
        ct = CTYPE[self.typeof(e)]
        return f"&({ct}){{{self.cx(e)}}}"

    # ---- statements
    def assign(self, lhs_text, rhs_text):
        lhs = parse_expr(lhs_text)
        rhs = parse_expr(rhs_text)
        if lhs.k != 'ref' or lhs.name not in self.u.syms:
            raise SyntaxError(f"bad assignment target {lhs_text!r}")
        s = self.sym(lhs.name)
        whole = s.dims is not None and (lhs.args is None or all(a.k == 'colon' for a in lhs.args))
        if whole:
            n = self.size_of(s)
            self.emit(f"for (long ia_ = 0; ia_ < {n}; ia_++) {self.v(s.name)}[ia_] = {self.cx(rhs, 'ia_')};")
        elif s.type == 'char':
            raise SyntaxError("character assignment not supported")
        else:
            self.emit(f"{self.cx(lhs)} = {self.cx(rhs)};")

    def call(self, text):
        m = re.match(r'call\s+(\w+)\s*(?:\((.*)\))?\s*$', text, re.I | re.S)
        name = m.group(1).lower()
        args = [Parser(lex(a)).arg() for a in split_top(m.group(2) or '', ',') if a]
        if name == 'system_clock':
            kw = {'count': 'NULL', 'count_rate': 'NULL', 'count_max': 'NULL'}
            order = ['count', 'count_rate', 'count_max']
            for i, a in enumerate(args):
                if a.k == 'kw':
                    kw[a.name] = self.actual(a.a)
                else:
                    kw[order[i]] = self.actual(a)
            self.emit(f"f90_system_clock({kw['count']}, {kw['count_rate']}, {kw['count_max']});")
            return
        if name == 'random_seed':
            for a in args:
                if a.k != 'kw':
                    raise SyntaxError("random_seed needs keywords")
                if a.name == 'size':
                    self.emit(f"f90_random_seed_size({self.actual(a.a)});")
                elif a.name == 'put':
                    s = self.sym(a.a.name)
                    self.emit(f"f90_random_seed_put({self.v(s.name)}, {self.size_of(s)});")
                else:
                    raise SyntaxError(f"random_seed({a.name}=)")
            return
        if name == 'random_number':
            self.emit(f"f90_random_number({self.actual(args[0])});")
            return
        if name not in self.units:
            raise SyntaxError(f"{self.u.name}: call of unknown subroutine {name}")
        self.emit(f"{name}_({', '.join(self.actual(a) for a in args)});")

    def io_control(self, inner):
        """Control list of read/write/open/close -> (positional list, keyword dict) of texts."""
        pos, kw = [], {}
        for item in split_top(inner, ','):
            m = re.match(r'(\w+)\s*=\s*(.*)$', item, re.S)
            if m and not re.match(r'^\s*[\'"]', item):
                kw[m.group(1).lower()] = m.group(2).strip()
            else:
                pos.append(item.strip())
        return pos, kw

    def str_arg(self, text):
        """A character expression -> (pointer, length) C texts; None -> (NULL, 0)."""
        if text is None:
            return 'NULL', '0'
        e = parse_expr(text)
        if e.k == 'str':
            return self.cx(e), str(len(e.v))
        return self.actual(e), self.charlen(e)

    def write_items(self, items):
        for it in items:
            e = parse_expr(it)
            t = self.typeof(e)
            if t == 'char':
                p, n = self.str_arg(it)
                self.emit(f"f90_w_str({p}, {n});")
            elif t == 'real':
                self.emit(f"f90_w_real({self.cx(e)});")
            else:
                self.emit(f"f90_w_int({self.cx(e)});")

    def io_stmt(self, text):
        low = text.lower()
        m = re.match(r'(write|read|open|close)\s*\(', low)
        op = m.group(1)
        p0 = text.index('(')
        p1 = match_paren(text, p0)
        pos, kw = self.io_control(text[p0 + 1:p1])
        rest = text[p1 + 1:].strip()
        items = [x for x in split_top(rest, ',') if x] if rest else []
        if op in ('write', 'read'):
            unit = kw.get('unit', pos[0] if pos else '*')
            fmt = kw.get('fmt', pos[1] if len(pos) > 1 else '*')
            unit_c = '-1' if unit.strip() == '*' else self.cx(parse_expr(unit))
            if fmt.strip() == '*':
                fmt_c = 'NULL'
            else:
                fe = parse_expr(fmt)
                if fe.k != 'str':
                    raise SyntaxError("format must be a literal")
                fmt_c = self.cx(fe)
            adv = kw.get('advance')
            noadv = 1 if adv and parse_expr(adv).v.lower() == 'no' else 0
            if op == 'write':
                self.emit(f"f90_w_begin({unit_c}, {fmt_c}, {noadv});")
                self.write_items(items)
                self.emit("f90_w_end();")
            else:
                self.emit(f"f90_r_begin({unit_c});")
                for it in items:
                    e = parse_expr(it)
                    t = self.typeof(e)
                    if t == 'char':
                        p, n = self.str_arg(it)
                        self.emit(f"f90_r_str({p}, {n});")
                    elif t == 'real':
                        self.emit(f"f90_r_real({self.actual(e)});")
                    else:
                        self.emit(f"f90_r_int({self.actual(e)});")
                self.emit("f90_r_end();")
            return
        if op == 'open':
            unit = kw.get('unit', pos[0] if pos else None)
            fp, fn = self.str_arg(kw.get('file'))
            sp, sn = self.str_arg(kw.get('status'))
            ap, an = self.str_arg(kw.get('action'))
            pp, pn = self.str_arg(kw.get('position'))
            ios = self.actual(parse_expr(kw['iostat'])) if 'iostat' in kw else 'NULL'
            self.emit(f"f90_open({self.cx(parse_expr(unit))}, {fp}, {fn}, {sp}, {sn}, {ap}, {an}, "
                      f"{pp}, {pn}, {ios});")
            return
        if op == 'close':
            unit = kw.get('unit', pos[0] if pos else None)
            sp, sn = self.str_arg(kw.get('status'))
            self.emit(f"f90_close({self.cx(parse_expr(unit))}, {sp}, {sn});")
            return

    def omp(self, text):
        low = ' '.join(text.lower().split())
        def clauses(s):
            out = []
            for m in re.finditer(r'(private|shared|reduction)\s*\(([^)]*)\)', s):
                kind, lst = m.group(1), m.group(2)
                if kind == 'reduction':
                    op, names = lst.split(':')
                    out.append(f"reduction({op.strip()}:{', '.join(self.v(n.strip()) for n in names.split(','))})")
                else:
                    out.append(f"{kind}({', '.join(self.v(n.strip()) for n in lst.split(','))})")
            return ' '.join(out)
        if low.startswith('end parallel do') or low.startswith('enddo') or low.startswith('end do'):
            return
        if low.startswith('end parallel'):
            self.ind -= 1
            self.emit('}')
            return
        if low.startswith('parallel do'):
            self.pending_pragma = f"#pragma omp parallel for {clauses(low)}".rstrip()
            return
        if low.startswith('parallel'):
            self.out.append(f"#pragma omp parallel {clauses(low)}".rstrip())
            self.emit('{')
            self.ind += 1
            return
        if low.startswith('do'):
            self.pending_pragma = "#pragma omp for"
            return
        raise SyntaxError(f"unsupported OpenMP directive: {text}")

    def statement(self, text, ln):
        low = text.lower()
        self.emit(f"/* {self.u.path.split('/')[-1]}:{ln} */")
        # block ends
        if re.match(r'end\s*if\b', low):
            self.ind -= 1
            self.emit('}')
            return
        if re.match(r'end\s*do\b', low):
            self.ind -= 1
            self.emit('}')
            if self.do_stack.pop():
                self.ind -= 1
                self.emit('}')
            return
        if re.match(r'else\s*if\s*\(', low):
            p0 = text.index('(')
            p1 = match_paren(text, p0)
            self.ind -= 1
            self.emit(f"}} else if ({self.cx(parse_expr(text[p0 + 1:p1]))}) {{")
            self.ind += 1
            return
        if low == 'else':
            self.ind -= 1
            self.emit('} else {')
            self.ind += 1
            return
        if re.match(r'if\s*\(', low):
            p0 = text.index('(')
            p1 = match_paren(text, p0)
            cond = self.cx(parse_expr(text[p0 + 1:p1]))
            rest = text[p1 + 1:].strip()
            if rest.lower() == 'then':
                self.emit(f"if ({cond}) {{")
                self.ind += 1
            else:
                self.emit(f"if ({cond}) {{")
                self.ind += 1
                self.statement(rest, ln)
                self.ind -= 1
                self.emit('}')
            return
        m = re.match(r'do\s*,?\s*while\s*\(', low)
        if m:
            p0 = text.index('(')
            p1 = match_paren(text, p0)
            self.emit(f"while ({self.cx(parse_expr(text[p0 + 1:p1]))}) {{")
            self.ind += 1
            self.do_stack.append(False)
            return
        m = re.match(r'do(?:\s+|\s*,\s*)(\w+)\s*=\s*(.*)$', text, re.I | re.S)
        if m:
            var = m.group(1).lower()
            parts = split_top(m.group(2), ',')
            lo, hi = parts[0], parts[1]
            step = parts[2] if len(parts) > 2 else None
            if step is not None:
                raise SyntaxError("do with a stride is not supported")
            self.emit('{')
            self.ind += 1
            self.emit(f"const int hi_ = {self.cx(parse_expr(hi))};")
            if self.pending_pragma:
                self.out.append(self.pending_pragma)
                self.pending_pragma = None
            v = self.v(var)
            self.emit(f"for ({v} = {self.cx(parse_expr(lo))}; {v} <= hi_; {v}++) {{")
            self.ind += 1
            self.do_stack.append(True)
            return
        if low == 'stop':
            self.emit("f90_stop();")
            return
        if low.startswith('call '):
            self.call(text)
            return
        if re.match(r'(write|read|open|close)\s*\(', low):
            self.io_stmt(text)
            return
        m = re.match(r'allocate\s*\((.*)\)\s*$', text, re.I | re.S)
        if m and low.startswith('allocate'):
            for item in split_top(m.group(1), ','):
                m2 = re.match(r'(\w+)\s*\((.*)\)\s*$', item)
                s = self.sym(m2.group(1).lower())
                n = self.cx(parse_expr(m2.group(2)))
                self.emit(f"{self.v(s.name)}n = {n};")
                self.emit(f"{self.v(s.name)} = ({CTYPE[s.type]} *)f90_calloc({self.v(s.name)}n, sizeof({CTYPE[s.type]}));")
            return
        m = re.match(r'deallocate\s*\((.*)\)\s*$', text, re.I | re.S)
        if m and low.startswith('deallocate'):
            for item in split_top(m.group(1), ','):
                s = self.sym(item.strip().lower())
                self.emit(f"free({self.v(s.name)}); {self.v(s.name)} = NULL;")
            return
        # assignment: first top-level '=' that is not part of ==, <=, >=, /=
        depth, q = 0, None
        for k, ch in enumerate(text):
            if q:
                if ch == q:
                    q = None
                continue
            if ch in "'\"":
                q = ch
            elif ch == '(':
                depth += 1
            elif ch == ')':
                depth -= 1
            elif ch == '=' and depth == 0 and text[k + 1:k + 2] != '=' and text[k - 1] not in '=<>/':
                self.assign(text[:k].strip(), text[k + 1:].strip())
                return
        raise SyntaxError(f"{self.u.path}:{ln}: cannot translate statement: {text}")

    # ---- units
    def signature(self, u):
        if u.kind == 'program':
            return "int f90_program(void)"
        ps = []
        for a in u.args:
            s = u.syms.get(a)
            if s is None or s.type is None:
                raise SyntaxError(f"{u.name}: dummy {a} has no declared type")
            if s.dims is not None or s.type == 'char':
                ps.append(f"{CTYPE[s.type]} *{self.v(a)}")
            else:
                ps.append(f"{CTYPE[s.type]} *{a}_p")
        return f"void {u.name}_({', '.join(ps) if ps else 'void'})"

    def unit(self, u):
        self.u = u
        self.do_stack = []
        self.pending_pragma = None
        # declarations first
        body = []
        u.uses = []
        for kind, text, ln in u.body:
            if kind == 'stmt':
                mu = re.match(r'use\s+(\w+)', text, re.I)
                if mu and mu.group(1).lower() in self.modules:
                    u.uses.append(mu.group(1).lower())
                if re.match(r'implicit\s+none\b', text, re.I) or re.match(r'use\s', text, re.I):
                    continue
                if parse_decl(u, text):
                    continue
            body.append((kind, text, ln))
        for a in u.args:
            u.syms[a].dummy = True
        self.out.append('')
        self.out.append(f"/* {u.kind} {u.name}: {u.path}:{u.line} */")
        self.out.append(self.signature(u))
        self.out.append('{')
        self.ind = 1
        # scalar dummies: copy in
        scal_dummies = [u.syms[a] for a in u.args if u.syms[a].dims is None and u.syms[a].type != 'char']
        for s in scal_dummies:
            self.emit(f"{CTYPE[s.type]} {self.v(s.name)} = *{s.name}_p;")
        # locals: scalars first (array extents may use them ... only dummies in practice)
        locs = [s for s in u.syms.values() if not s.dummy]
        for s in locs:
            if s.dims is None and s.type.startswith('struct:'):
                self.emit(f"{CTYPE[s.type]} {self.v(s.name)} = {{0}};")
            elif s.dims is None and s.type != 'char':
                self.emit(f"{CTYPE[s.type]} {self.v(s.name)} = 0;")
            elif s.dims is None:
                n = self.cx(parse_expr(s.charlen))
                self.emit(f"char {self.v(s.name)}[{n}]; memset({self.v(s.name)}, ' ', {n});")
        frees = []
        for s in locs:
            if s.dims is None:
                continue
            ct = CTYPE[s.type]
            if s.dims == [':']:
                self.emit(f"{ct} *{self.v(s.name)} = NULL; long {self.v(s.name)}n = 0;")
            else:
                self.emit(f"{ct} *{self.v(s.name)} = ({ct} *)f90_calloc({self.size_of(s)}, sizeof({ct}));")
            frees.append(s)
        for s in u.syms.values():
            self.emit(f"(void){self.v(s.name)};")
        for kind, text, ln in body:
            if kind == 'omp':
                self.omp(text)
            else:
                self.statement(text, ln)
        if self.do_stack or self.ind != 1:
            raise SyntaxError(f"{u.name}: unbalanced blocks")
        for s in frees:
            self.emit(f"free({self.v(s.name)});")
        for s in scal_dummies:
            if s.intent != 'in':
                self.emit(f"*{s.name}_p = {self.v(s.name)};")
        if u.kind == 'program':
            self.emit("f90_flush_all();")
            self.emit("return 0;")
        self.out.append('}')

    def run(self):
        head = ['/* GENERATED by oracle/f90c/f90toc.py from the reference sources; do not commit. */',
                '#include "f90rt.h"', '']
        # pre-declare: need declarations parsed -> do a declaration pass now
        for u in self.units.values():
            for kind, text, ln in u.body:
                if kind == 'stmt':
                    parse_decl(u, text)
        protos = []
        for mod in self.modules.values():
            protos.append(f"/* module {mod.name}: {mod.path} (interoperability rules) */")
            for tn, comps in mod.types.items():
                protos.append(f"struct {tn} {{ " + ' '.join(f"{CTYPE[t]} {c};" for t, c in comps) + " };")
            for cname, ret, dummies in mod.funcs.values():
                ps = []
                for _, t, by_value, _, intent in dummies:
                    ct = CTYPE[t]
                    ps.append(ct if by_value else (("const " if intent == 'in' else "") + ct + " *"))
                protos.append(f"extern {CTYPE[ret]} {cname}({', '.join(ps) if ps else 'void'});")
        for u in self.units.values():
            self.u = u
            protos.append(self.signature(u) + ';')
        for u in self.units.values():
            self.unit(u)
        return '\n'.join(head + protos + self.out) + '\n'


def main(argv):
    if len(argv) < 4 or argv[1] != '-o':
        sys.exit(__doc__)
    out = argv[2]
    units = []
    for path in argv[3:]:
        units += parse_units(read_statements(path), path)
    code = Gen(units).run()
    with open(out, 'w') as f:
        f.write(code)


if __name__ == '__main__':
    main(sys.argv)
