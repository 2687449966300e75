"""Minimal SIAL front-end for the block-contraction hot path.

The reference compiles SIAL to `.siox` bytecode with a Java compiler (not available here) and interprets it one
opcode at a time (src/sip/worker/interpreter.cpp:98-910).  This module is NOT that interpreter: it understands just the
statement subset that the amplitude equations of the coupled-cluster programs use (src/sialx/qm/cc/rlccd_rhf.sialx,
rlccsd_rhf.sialx, rccsd_rhf.sialx: every procedure of an iteration except DIIS; tests/golden/*_program.sialx) and turns them into the per-block calls the interpreter would issue -- in the same order, with the same temp-block
lifetimes (temps die at the end of the loop iteration that created them, BlockManager::leave_scope) and the same
pardo work distribution (iteration k of the where-true iterations of a barrier section runs on worker k mod nworkers,
BalancedTaskAllocParallelPardoLoop::do_update, loop_manager.cpp:468-499; first index fastest, :434-452).

Those calls go to a *backend*:
  * DeviceBackend  -- libsipgpu through the C ABI; with `record=True` every pardo is wrapped in
                      sipgpu_wl_begin()/sipgpu_wl_end(), so the op-at-a-time stream is scheduled into a few batched
                      launches by aces4_b200/csrc/worklist.cu; with `record=False` every op is its own launch (the
                      naive port of the interpreter onto a device ABI) -- the comparison the work-list exists for;
  * the tests' OracleBackend (tests/sial_oracle_backend.py) -- the CPU oracle, for parity.

Supported statements (case-insensitive keywords, `#` comments):
    pardo i, j, ...  /  endpardo ...        do i / enddo i        where i < j   (<, <=, >, >=, ==, != on indices/ints)
    request|get A[...]       put|prepare A[...] = T[...]      put|prepare A[...] += T[...]     put|prepare A[...] = number
    T[...] = number      T[...] = X[...]    T[...] = X[...] * Y[...]        T[...] += X[...]     T[...] -= X[...]
    T[...] = X[...] ^ Y[...] (outer product: the same opcode)               T[...] *= number     T[...] *= s
    s = X[...] * Y[...]   s = number    s += t    s -= t    s *= number
    execute energy_denominator_rhf T[...] fock      sip_barrier | server_barrier      collective s += t
    proc NAME ... endproc NAME        call NAME        allocate L[a,*,b,j] / deallocate L[a,*,b,j]
    set_persistent A "label"          restore_persistent A "label"       (arrays and scalars)
    if X <op> Y ... endif (X, Y: ints, scalars, indices, numbers, `(int)` / `(scalar)` casts)      exit (leaves the innermost do)
    int n      n = n + 1      x = (int)Xijk[NT,i7]      t = a/(scalar)n      (host arithmetic on ints / scalars)
    index ii = 1: naocc (simple index: integer loop variable; a block dimension it labels has extent 1)
    execute stripi X[..] Y[..]      execute set_ijk_aab Xijk      execute return_sval Xijk[Nt,i7] s      print / create: ignored
Indices are declared `moaindex i = baocc: eaocc` / `moaindex a = bavirt: eavirt` / `moaindex p = baocc: eavirt` /
`aoindex mu = 1: norb`; arrays `served|distributed|temp|local|static NAME[i,j,..]`, scalars `scalar s`.  Anything else
raises SialSyntaxError.

An array declared over a `p` index (occupied followed by virtual, e.g. `served Vpiqj[p,i,q,j]`, `static ca[mu,p]`) is
addressed with occupied or virtual labels as in the reference (`Vpiqj[a,i,b,j]`, `Vpiqj[i,i1,j,j1]`, `ca[mu,b]`): the
segment number of a virtual label is shifted by the number of occupied segments.  `static` arrays are read block-wise
through the backend (`array_block`), i.e. as slices of a resident array (contiguous_array_manager.cpp:162-230).

`allocate` creates the blocks of a `local` array for every segment of the `*` dimensions, zero-filled, and they live
until `deallocate` (not until the end of the loop iteration like temps; block_manager.cpp allocate_local /
deallocate_local) -- here a block is created, zero-filled, when it is first touched.

`set_persistent` hands an array (or scalar) over to the backend's label registry and `restore_persistent` adopts it into
the array of that name declared by the program that follows -- how the reference chains its programs (scf -> tran -> cc,
worker_persistent_array_manager.cpp:34-155).  As in the reference the hand-over happens when the program ENDS (Walker.run /
run_proc return), so an array stays usable after its `set_persistent` statement.

Procedures: statements between `proc NAME` and `endproc` form a procedure; `call NAME` runs it; `Walker.run()` runs
the main program (the statements outside procedures) and `Walker.run_proc(NAME)` one procedure (the test drivers use
it for the `do kiter` loop of a CC program, whose `if`/`exit` are not in the subset).  A file that consists of
procedures only (a fragment) is run in textual order.
"""
import itertools
import math
import re


class SialSyntaxError(ValueError):
    pass


_KIND_BY_RANGE = {("baocc", "eaocc"): "o", ("bavirt", "eavirt"): "v", ("baocc", "eavirt"): "p", ("1", "norb"): "ao",
                  ("bocc", "eocc"): "o", ("bvirt", "evirt"): "v",
                  # the defs files' own index of the static arrays ca / fock_a: ALL molecular orbital segments
                  ("1", "eavirt"): "pa", ("1", "ebvirt"): "pa"}
_REF = r"([A-Za-z_]\w*)\s*\[([^\]]*)\]"
_NUM = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eEdD][-+]?\d+)?"


def _labels(s):
    """subscripts of an array reference, lower-cased (SIAL names are case-insensitive); the one-value range `x:x` that addresses
    a contiguous array (`CLRB1_a[kstart:kstart,a:a,i:i]`) is the block `x`"""
    out = []
    for x in s.split(","):
        x = x.strip().lower()
        if not x:
            continue
        m = re.match(r"\(\s*index\s*\)\s*(\w+)$", x)
        if m:      # `a[(index)j]`: the VALUE of j addresses a simple-index dimension (cast_indices_to_simple.sialx)
            x = "@" + m.group(1)
        if ":" in x:
            lo, hi = (y.strip() for y in x.split(":", 1))
            if lo == hi:
                x = lo
        out.append(x)
    return tuple(out)


_TOK = re.compile(r"\s*(?:(\d+\.?\d*(?:[ed][-+]?\d+)?|\.\d+)|([a-z_]\w*)|(\[[^\]]*\])|(.))", re.I)


def parse_expr(text):
    """int / scalar expression -> AST: ('num', v) | ('var', name) | ('elem', array, labels) | ('cast', 'int'|'scalar', e) |
    ('neg', e) | (op, l, r) with op in + - * / **   (SIAL casts are prefix: `(int)Xijk[NT,i7]`, `tcomb/(scalar)nsects`)"""
    toks = []
    for num, name, br, other in _TOK.findall(text.strip().lower()):
        if num:      # `42` is an int, `42.0` / `1.d-3` a scalar (int / int truncates as in the SIP: sial_math / interpreter.cpp)
            toks.append(("num", int(num) if num.isdigit() else float(num.lower().replace("d", "e"))))
        elif name:
            toks.append(("name", name))
        elif br:
            toks.append(("br", _labels(br[1:-1])))
        elif other.strip():
            toks.append(("op", other))
    pos = [0]

    def peek():
        return toks[pos[0]] if pos[0] < len(toks) else (None, None)

    def take():
        pos[0] += 1
        return toks[pos[0] - 1]

    def primary():
        k, v = take() if pos[0] < len(toks) else (None, None)
        if k == "num":
            return ("num", v)
        if k == "name":
            if v in ("sqrt", "abs") and peek()[0] in ("num", "name") or v in ("sqrt", "abs") and peek() == ("op", "("):
                return ("fn", v, primary())      # `sr0 = sqrt 256.0`
            if peek()[0] == "br":
                return ("elem", v, take()[1])
            return ("var", v)
        if (k, v) == ("op", "-"):
            return ("neg", primary())
        if (k, v) == ("op", "("):
            if peek() in (("name", "int"), ("name", "scalar")) and pos[0] + 1 < len(toks) and toks[pos[0] + 1] == ("op", ")"):
                kind = take()[1]
                take()
                return ("cast", kind, primary())
            e = expr()
            if take() != ("op", ")"):
                raise SialSyntaxError("unbalanced parenthesis in expression")
            return e
        raise SialSyntaxError(f"bad expression {text!r}")

    def power():
        e = primary()
        while peek() == ("op", "*") and pos[0] + 1 < len(toks) and toks[pos[0] + 1] == ("op", "*"):
            take()
            take()
            e = ("**", e, primary())
        return e

    def term():
        e = power()
        while peek() in (("op", "*"), ("op", "/")) and not (peek() == ("op", "*") and pos[0] + 1 < len(toks) and toks[pos[0] + 1] == ("op", "*")):
            op = take()[1]
            e = (op, e, power())
        return e

    def expr():
        e = term()
        while peek() in (("op", "+"), ("op", "-")):
            op = take()[1]
            e = (op, e, term())
        return e

    out = expr()
    if pos[0] != len(toks):
        raise SialSyntaxError(f"trailing tokens in expression {text!r}")
    return out


class _Exit(Exception):
    """`exit`: leaves the innermost do loop (interpreter.cpp handles it as a jump to the loop's end)"""


class _Ones:
    """segment extents of a simple index: every value labels a block dimension of extent 1"""

    def __getitem__(self, i):
        return 1


class Program:
    """Parsed SIAL fragment: declarations + a statement tree."""

    def __init__(self, text):
        self.index_kind, self.arrays, self.scalars = {}, {}, set()
        self.ints = set()         # names declared `int` (integer arithmetic: `/` truncates, `(int)` rounds like lrint)
        self.simple_range = {}    # simple index name -> (lo, hi) as written (numbers or predefined constants such as naocc)
        self.procs = {}           # name -> statement list (textual order kept: dicts are ordered)
        self.predefined = set()   # `predefined int|scalar X`: the value comes with the job (the .dat file), Walker(constants=...)
        self.contiguous = set()   # `contiguous local X[...]` arrays (held block-wise like `local` ones)
        self.var_labels = set()   # int variables that occur as subscripts
        main = []
        stack = [main]
        opens = []
        in_proc = None
        for ln, raw in enumerate(text.splitlines(), 1):
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            low = line.lower()
            tok = low.split()
            try:
                st = self._parse(line, low, tok)
            except SialSyntaxError as e:
                raise SialSyntaxError(f"line {ln}: {e}: {raw.strip()!r}") from None
            if st is None:
                continue
            label_slots = {"assign": (2, 4), "add": (2, 4), "fill": (2,), "fill_expr": (2,), "scale": (2,), "scale_expr": (2,),
                           "contract": (2, 4, 6), "put": (2, 5), "request": (2,)}.get(st[0])
            if label_slots:
                # an int VARIABLE used as a subscript (`CLRB1_a[kstart:kstart,a:a,i:i]`, kstart = kstate + roots): it labels a
                # one-element dimension like a simple index, its value is read when the statement executes
                var_labels = tuple(sorted({x for k in label_slots for x in st[k]
                                           if x in self.scalars and (x not in self.index_kind or x in self.var_labels)}))
                if var_labels:
                    for x in var_labels:
                        self.var_labels.add(x)
                        self.index_kind[x] = "s"
                        self.simple_range[x] = ("1", "1")
                    st = ("with_vars", var_labels, st)
            if st[0] == "proc":
                if in_proc is not None or opens or st[1] in self.procs:
                    raise SialSyntaxError(f"line {ln}: misplaced or duplicate proc {st[1]}")
                in_proc = st[1]
                self.procs[in_proc] = []
                stack = [self.procs[in_proc]]
            elif st[0] == "endproc":
                if in_proc is None or opens:
                    raise SialSyntaxError(f"line {ln}: misplaced endproc")
                in_proc = None
                stack = [main]
            elif st[0] in ("pardo", "do", "if"):
                st = st + ([],)
                stack[-1].append(st)
                stack.append(st[-1])
                opens.append(st[0])
            elif st[0] == "else":
                if not opens or opens[-1] != "if":
                    raise SialSyntaxError(f"line {ln}: else outside an if")
                stack.pop()
                st = ("else", [])
                stack[-1].append(st)
                stack.append(st[-1])
            elif st[0] in ("endpardo", "enddo", "endif"):
                if not opens or opens.pop() != st[0][3:]:
                    raise SialSyntaxError(f"line {ln}: unbalanced {st[0]}")
                stack.pop()
            else:
                stack[-1].append(st)
        if opens or in_proc is not None:
            raise SialSyntaxError("unterminated " + (opens[-1] if opens else "proc " + in_proc))
        # a fragment (procedures only) runs in textual order; otherwise the body is the main program
        self.body = main if main else [st for body in self.procs.values() for st in body]

    def _parse(self, line, low, tok):
        kw = tok[0]
        if kw == "print":      # `print a[i,j]`: handed to Walker.print_hook (the reference's tests compare printed blocks with fixtures)
            m = re.match(r"print\s+" + _REF + r"\s*$", line, re.I)
            if m:
                return ("print_block", m.group(1).lower(), _labels(m.group(2)))
            return None
        if kw in ("sial", "endsial", "import", "print", "println", "create", "delete", "destroy", "special", "broadcast_from", "assert_same",
                  "gpu_on", "gpu_off", "gpu_put", "gpu_get", "gpu_allocate", "gpu_free"):
            return None           # arrays exist (zero) from the start; nothing is printed; super-instruction signatures are not
                                  # needed; the gpu_* statements of the reference's dormant CUDA path (host <-> device moves of one
                                  # block, interpreter.cpp:636-651) have nothing to do: every block is resident;
                                  # needed; broadcast_from / assert_same: every worker computes the replicated statics itself
        if kw == "predefined":
            if len(tok) < 3 or tok[1] not in ("int", "scalar"):
                raise SialSyntaxError("bad predefined declaration")
            self.predefined.add(tok[2])
            self.scalars.add(tok[2])
            if tok[1] == "int":
                self.ints.add(tok[2])
            return None
        if kw == "contiguous":     # a dense local array addressed by ranges; the programs here address it block by block (`x:x`)
            m = re.match(r"contiguous\s+local\s+" + _REF, line, re.I)
            if not m:
                raise SialSyntaxError("bad contiguous declaration")
            self.contiguous.add(m.group(1).lower())
            self.arrays[m.group(1).lower()] = ("local", _labels(m.group(2)))
            return None
        if kw in ("allocate", "deallocate") and len(tok) > 1 and tok[1] == "contiguous":
            m = re.match(r"\w+\s+contiguous\s+([A-Za-z_]\w*)\s*\[", line, re.I)
            if not m:
                raise SialSyntaxError("bad " + kw + " contiguous")
            return (kw, m.group(1).lower(), ("*",))
        if kw == "index":
            m = re.match(r"index\s+(\w+)\s*=\s*(-?\s*\w+)\s*:\s*(-?\s*\w+)", line, re.I)
            if not m:
                raise SialSyntaxError("bad simple index declaration")
            self.index_kind[m.group(1).lower()] = "s"
            self.simple_range[m.group(1).lower()] = (m.group(2).lower().replace(" ", ""), m.group(3).lower().replace(" ", ""))
            return None
        if kw == "int":
            self.scalars.add(tok[1])
            self.ints.add(tok[1])
            if "=" in low:      # `int e0 = 77`
                return ("sexpr", tok[1], "=", parse_expr(low.split("=", 1)[1]))
            return None
        if kw == "if":
            m = re.match(r"if\s+(.+?)\s*(<=|>=|==|!=|<|>)\s*(.+)$", low)
            if not m:
                raise SialSyntaxError("unsupported if condition")
            return ("if", parse_expr(m.group(1)), m.group(2), parse_expr(m.group(3)))
        if kw == "endif":
            return ("endif",)
        if kw == "else" and len(tok) == 1:
            return ("else",)
        if kw == "exit":
            return ("exit",)
        if kw in ("proc", "call"):
            if len(tok) != 2:
                raise SialSyntaxError("bad " + kw)
            return (kw, tok[1])
        if kw == "endproc":
            return ("endproc",)
        if kw in ("moaindex", "aoindex", "moindex", "mobindex"):
            m = re.match(r"\w+\s+(\w+)\s*=\s*(\w+)\s*:\s*(\w+)", line)
            if not m:
                raise SialSyntaxError("bad index declaration")
            lo, hi = m.group(2).lower(), m.group(3).lower()
            kind = _KIND_BY_RANGE.get((lo, hi))
            if kind is None:     # any other segment range of the index type (`aoindex i = 1:1`, `moaindex j = 2: 3`): a kind of its own
                kind = f"{'ao' if kw == 'aoindex' else 'mo'}:{lo}:{hi}"
            self.index_kind[m.group(1).lower()] = kind
            return None
        if kw in ("served", "distributed", "temp", "local", "static"):
            m = re.match(r"\w+\s+" + _REF, line)
            if not m:
                raise SialSyntaxError("bad array declaration")
            self.arrays[m.group(1).lower()] = (kw, _labels(m.group(2)))
            return None
        if kw == "scalar":
            self.scalars.add(tok[1])
            if "=" in low:      # `scalar total = 0.0`
                return ("sexpr", tok[1], "=", parse_expr(low.split("=", 1)[1]))
            return None
        if kw == "pardo" or kw == "do":
            return (kw, _labels(re.sub(r'"[^"]*"', "", line.split(None, 1)[1])))     # `pardo i, j "pragma text"`
        if kw in ("endpardo", "enddo"):
            return (kw,)
        if kw == "where":
            m = re.match(r"where\s+(\w+)\s*(<=|>=|==|!=|<|>)\s*(\w+)\s*$", low)
            if m:
                return ("where", m.group(1), m.group(2), m.group(3))
            m = re.match(r"where\s+(.+?)\s*(<=|>=|==|!=|<|>)\s*(.+)$", low)     # `where k == ((i-1)*norb + (j-1)) + 1`
            if not m:
                raise SialSyntaxError("unsupported where clause")
            return ("where", parse_expr(m.group(1)), m.group(2), parse_expr(m.group(3)))
        if kw in ("request", "get"):
            m = re.match(r"\w+\s+" + _REF, line)
            if not m:
                raise SialSyntaxError("bad request")
            return ("request", m.group(1).lower(), _labels(m.group(2)))
        if kw in ("put", "prepare"):
            m = re.match(r"\w+\s+" + _REF + r"\s*=\s*(" + _NUM + r")\s*$", line)
            if m:
                return ("put_init", m.group(1).lower(), _labels(m.group(2)), float(m.group(3).lower().replace("d", "e")))
            m = re.match(r"\w+\s+" + _REF + r"\s*\*=\s*(.+)$", line)
            if m:     # `prepare RB1_a[davidson,a,i] *= etemp`: the owner scales its block
                return ("put_scale", m.group(1).lower(), _labels(m.group(2)), parse_expr(m.group(3)))
            m = re.match(r"\w+\s+" + _REF + r"\s*(\+?=)\s*([^\[\]]+)$", line)
            if m:     # `put a[i,j] = x` / `put a[i,j] += x` with a scalar: put_initialize / put_increment (sial_ops_parallel.cpp:412-528)
                return ("put_value", m.group(1).lower(), _labels(m.group(2)), m.group(3), parse_expr(m.group(4)))
            m = re.match(r"\w+\s+" + _REF + r"\s*(\+?=)\s*" + _REF + r"\s*$", line)
            if not m:
                raise SialSyntaxError("bad put/prepare")
            return ("put", m.group(1).lower(), _labels(m.group(2)), m.group(3), m.group(4).lower(), _labels(m.group(5)))
        if kw == "execute":
            if len(tok) < 2:
                raise SialSyntaxError("bad execute")
            args = [(m.group(1).lower(), _labels(m.group(2))) for m in re.finditer(_REF, line)]
            bare = [t for t in re.sub(_REF, " ", line).split()[2:]]
            return ("execute", tok[1], args, [b.lower() for b in bare])
        if kw in ("set_persistent", "restore_persistent"):
            m = re.match(r'\w+\s+([A-Za-z_]\w*)\s+"([^"]+)"\s*$', line)
            if not m:
                raise SialSyntaxError("bad " + kw)
            return (kw, m.group(1).lower(), m.group(2))
        if kw in ("allocate", "deallocate"):
            m = re.match(r"\w+\s+([A-Za-z_]\w*)\s*\[([^\]]*)\]\s*$", line)
            if not m:
                m = re.match(r"deallocate\s+([A-Za-z_]\w*)\s*$", line, re.I)     # `deallocate LGmat`: the whole array
                if m:
                    return ("deallocate", m.group(1).lower(), ("*",))
                raise SialSyntaxError("bad " + kw)
            return (kw, m.group(1).lower(), _labels(m.group(2)))
        if kw in ("sip_barrier", "server_barrier"):
            return ("barrier",)
        if kw == "collective":
            m = re.match(r"collective\s+(\w+)\s*\+=\s*(?:\(\s*(?:scalar|int)\s*\)\s*)?(\w+)\s*$", low)
            if not m:
                raise SialSyntaxError("bad collective")
            return ("collective", m.group(1), m.group(2))
        # block statements
        m = re.match(_REF + r"\s*(\+=|-=|\*=|=)\s*(.+)$", line)
        if m:
            name, labs, op, rhs = m.group(1).lower(), _labels(m.group(2)), m.group(3), m.group(4).strip()
            mm = re.match(_REF + r"\s*[\*\^]\s*" + _REF + r"\s*$", rhs)   # `^` (outer product) is the same opcode
            if mm and op == "=":
                return ("contract", name, labs, mm.group(1).lower(), _labels(mm.group(2)), mm.group(3).lower(),
                        _labels(mm.group(4)))
            mm = re.match(_REF + r"\s*([+-])\s*" + _REF + r"\s*$", rhs)
            if mm and op == "=":     # `a[i,j] = b[i,j] + c[i,j]` (handle_block_add / subtract, interpreter.cpp:1874-1997)
                return ("addsub", name, labs, mm.group(1).lower(), _labels(mm.group(2)), mm.group(4).lower(), _labels(mm.group(5)),
                        1.0 if mm.group(3) == "+" else -1.0)
            mm = re.match(_REF + r"\s*$", rhs)
            if mm and op in ("=", "+=", "-="):
                return ("assign" if op == "=" else "add", name, labs, mm.group(1).lower(), _labels(mm.group(2)),
                        -1.0 if op == "-=" else 1.0)
            if re.match(_NUM + r"$", rhs) and op in ("=", "*="):
                return ("fill" if op == "=" else "scale", name, labs, float(rhs.lower().replace("d", "e")))
            if re.match(r"[A-Za-z_]\w*$", rhs) and op == "*=":
                return ("scale_by", name, labs, rhs.lower())
            if op in ("=", "*=") and "[" not in rhs:      # `T4kai[davidson,a,i] = omega`, `Tkai[kstate,a,i] *= (0.5)**(0.5)`
                return ("fill_expr" if op == "=" else "scale_expr", name, labs, parse_expr(rhs))
            if op in ("+=", "-=") and "[" not in rhs:     # `dipole[ixyz] -= dsum`: every element incremented by a number
                return ("incr_expr", name, labs, -1.0 if op == "-=" else 1.0, parse_expr(rhs))
            raise SialSyntaxError("unsupported block statement")
        # scalar statements
        m = re.match(r"(\w+)\s*(\+=|-=|\*=|=)\s*(.+)$", line)
        if m:
            name, op, rhs = m.group(1).lower(), m.group(2), m.group(3).strip()
            mm = re.match(_REF + r"\s*\*\s*" + _REF + r"\s*$", rhs)
            if mm and op == "=":
                return ("sdot", name, mm.group(1).lower(), _labels(mm.group(2)), mm.group(3).lower(), _labels(mm.group(4)))
            if re.match(_NUM + r"$", rhs):
                return ("sset" if op == "=" else {"+=": "sinc", "-=": "sinc", "*=": "smul"}[op], name,
                        float(rhs) * (-1.0 if op == "-=" else 1.0))
            if re.match(r"\w+$", rhs) and op in ("+=", "-=", "="):
                return ("sadd" if op != "=" else "scopy", name, rhs.lower(), -1.0 if op == "-=" else 1.0)
        m = re.match(r"(\w+)\s*(\+=|-=|\*=|/=|=)\s*(.+)$", line)
        if m:   # host arithmetic on ints / scalars: counters, casts, elements of index tables
            return ("sexpr", m.group(1).lower(), m.group(2), parse_expr(m.group(3).lower()))
        raise SialSyntaxError("unsupported statement")


def set_ijk_aab(moa_seg_ranges, baocc, eaocc, maxi=5, ordered_k=False):
    """The batch table of the reference's (T) programs (super_instructions/qm/qm-generic/set_ijk_aab.F:60-160): every
    occupied segment is cut into pieces of at most `maxi` orbitals (maxi shrinks by one for every occupied segment that is
    not longer than it, as in the Fortran), and row Nt = (segment i, first orbital, last orbital, segment j >= i, first,
    last, segment k) for every pair of pieces and every k; the row after the last one is all -1.  Host-only control logic
    (which loop iterations exist), not arithmetic.  ordered_k: set_ijk_aaa.F, the same table restricted to j <= k.
    Returns {(Nt, column): value}, 1-based."""
    nseg = eaocc - baocc + 1
    for i in range(1, nseg + 1):
        if maxi >= moa_seg_ranges[baocc + i - 2]:
            maxi -= 1
    start, end = [0] * (nseg + 1), [0] * (nseg + 1)
    start[1], end[1] = 1, moa_seg_ranges[0]
    for i in range(2, baocc + 1):
        start[1] = end[1] + 1
        end[1] = start[1] + moa_seg_ranges[i - 1] - 1
    for i in range(2, nseg + 1):
        start[i] = end[i - 1] + 1
        end[i] = start[i] + moa_seg_ranges[baocc + i - 2] - 1
    pieces = {}
    for i in range(1, nseg + 1):
        n = end[i] - start[i] + 1
        np_ = n // maxi + (1 if maxi * (n // maxi) < n else 0)
        sizes = [n // np_] * (np_ - 1)
        sizes.append(n - sum(sizes))
        pieces[i] = sizes
    table, nt = {}, 0
    ie = start[1] - 1
    for i in range(1, nseg + 1):
        for ni in pieces[i]:
            is_, ie = ie + 1, ie + ni
            je = start[1] - 1
            for j in range(1, nseg + 1):
                for nj in pieces[j]:
                    js, je = je + 1, je + nj
                    if i <= j:
                        for k in range(1, nseg + 1):
                            if ordered_k and j > k:     # set_ijk_aaa.F:99: the AAA combination takes i <= j <= k only
                                continue
                            nt += 1
                            for col, v in enumerate((i, is_, ie, j, js, je, k), 1):
                                table[(nt, col)] = float(v)
    for col in range(1, 8):
        table[(nt + 1, col)] = -1.0
    return table


def compute_diis(B):
    """The reference's `execute compute_diis BB` (super_instructions/qm/utility/compute_diis.F -> form_R.F:8-110): the upper
    triangle of the error-overlap matrix BB is symmetrised, its all-zero rows (history slots not filled yet) are dropped from
    the end, the matrix is bordered with -1 / 0 and R c = (0, ..., 0, -1) is solved with LAPACK dgesv; the coefficients come back
    on the DIAGONAL of BB (zero for the unused slots).  Host control logic of the CC iterations (order <= 30).  -> [c_1 .. c_n]"""
    import numpy as np

    B = np.array(B, dtype=float)
    n = B.shape[0]
    R = np.triu(B) + np.triu(B, 1).T
    ndim = n - sum(1 for i in range(n) if not np.any(np.abs(R[i, :]) > 0.0))
    M = np.zeros((ndim + 1, ndim + 1))
    M[:ndim, :ndim] = R[:ndim, :ndim]
    M[:ndim, ndim] = M[ndim, :ndim] = -1.0
    v = np.zeros(ndim + 1)
    v[ndim] = -1.0
    c = np.linalg.solve(M, v)
    out = np.zeros(n)
    out[:ndim] = c[:ndim]
    return out


def eigen_calc(A):
    """The reference's `execute eigen_calc A V` (super_instructions/qm/utility/eigen_calc.F dsyev_wrapper): LAPACK DSYEV('V','L'),
    each eigenvector's sign fixed so that its first component above 1e-5 in magnitude is positive, ascending eigenvalues; A comes
    back as diag(eigenvalues), V holds the eigenvectors in its columns.  -> (A_out, V)"""
    import numpy as np

    A = np.array(A, dtype=float)
    w, v = np.linalg.eigh(np.tril(A) + np.tril(A, -1).T, UPLO="L")
    for j in range(v.shape[1]):
        big = np.nonzero(np.abs(v[:, j]) > 1.0e-5)[0]
        if big.size and v[big[0], j] < 0.0:
            v[:, j] = -v[:, j]
    order = np.argsort(w, kind="stable")
    return np.diag(w[order]), v[:, order]


def cis_unit_guess(hdiag, nsub):
    """The reference's `execute cis_unit_guess C1_a SHDiag` (super_instructions/qm/excited/cis_unit_guess.F unitvecguess): unit
    vectors on the `nsub` smallest elements of the Hamiltonian diagonal hdiag[a,i], smallest first; on ties the LAST element
    in (i outer, a inner) order wins (`.le.`).  -> B[nsub, a, i]"""
    import numpy as np

    work = np.array(hdiag, dtype=float)
    nv, no = work.shape
    B = np.zeros((nsub, nv, no))
    for k in range(nsub):
        ee, sav = 1.0e8, None
        for i in range(no):
            for a in range(nv):
                if work[a, i] <= ee:
                    ee, sav = work[a, i], (a, i)
        B[k, sav[0], sav[1]] = 1.0
        work[sav] = 1.0e9
    return B


def gen_eigen_calc(A):
    """The reference's `execute gen_eigen_calc G L R E` (super_instructions/qm/utility/gen_eigen_calc.F:60-78 and its
    dgeev_wrapper :80-230): LAPACK DGEEV('V','V') of the (zero-padded) subspace matrix, real parts of the eigenvalues,
    eigenvalues below 1e-12 in magnitude (the padding) sorted to the END and reported as 0, left / right eigenvector columns
    reordered with the ascending eigenvalues.  Host control logic of the Davidson solver (the matrix is at most 60 x 60), not
    block arithmetic.  -> (VL, VR, eigenvalues)"""
    import numpy as np
    from scipy.linalg import eig

    A = np.array(A, dtype=float, order="F")
    n = A.shape[0]
    w, vl, vr = eig(A, left=True, right=True)

    def lapack_columns(v):       # DGEEV stores a complex pair as (real part, imaginary part) in consecutive columns
        out = np.zeros((n, n))
        j = 0
        while j < n:
            if abs(w[j].imag) > 0.0 and j + 1 < n:
                out[:, j], out[:, j + 1] = v[:, j].real, v[:, j].imag
                j += 2
            else:
                out[:, j] = v[:, j].real
                j += 1
        return out

    L, R = lapack_columns(vl), lapack_columns(vr)
    ev = w.real.copy()
    ev[np.abs(ev) < 1e-12] = 999.0
    order = np.argsort(ev, kind="stable")
    ev = ev[order]
    ev[np.abs(ev) == 999.0] = 0.0
    return L[:, order], R[:, order], ev


class Walker:
    """Executes a Program against a backend: the per-block call stream of one worker."""
    host_registry = {}       # persistent host tables (label -> {index values: number}), shared by consecutive programs

    def __init__(self, program, backend, segs, rank=0, world=1, index_base=None, constants=None, host_data=None, seg_tables=None,
                 extra_si=None):
        """segs: {'o': [extents of the occupied segments], 'v': [...], 'ao': [...]}; 'p' = o followed by v.
        index_base: {'o': baocc - 1, 'v': bavirt - 1, ...}: what to add to a loop's segment number to get the absolute
        segment number of its index type -- the index values a super-instruction receives (the reference's loops run
        over baocc..eaocc / bavirt..eavirt directly, interpreter.cpp:1011-1208)."""
        self.p, self.be, self.rank, self.world = program, backend, rank, world
        self.host_data = dict(host_data or {})     # tables an out-of-scope engine would compute (dipole integrals, ...)
        self.extra_si = dict(extra_si or {})       # further super-instructions by name: f(walker, args, bare) (the reference's test helpers)
        self.index_base = dict(index_base or {})
        self.segs = dict(segs)
        if "p" not in self.segs and "o" in self.segs and "v" in self.segs:
            self.segs["p"] = list(self.segs["o"]) + list(self.segs["v"])
        if "pa" not in self.segs and "p" in self.segs and not self.index_base.get("o", 0):
            self.segs["pa"] = list(self.segs["p"])       # all-electron job: every MO segment is active
        self.segs["s"] = _Ones()
        # predefined ints (naocc, eom_roots, ...) and scalars (eom_tol, ...)
        self.constants = {k.lower(): (int(v) if float(v).is_integer() else float(v)) for k, v in (constants or {}).items()}
        # index ranges outside the named ones: segments lo..hi of the index type's table (`seg_tables` = {'ao': [...], 'mo': [...]},
        # default: segs['ao'] / segs['pa'])
        tables = dict(seg_tables or {})
        self._range_lo = {}
        for kind in set(program.index_kind.values()):
            if ":" in kind:
                typ, lo, hi = kind.split(":")
                lo, hi = (int(x) if x.isdigit() else int(self.constants[x]) for x in (lo, hi))
                full = tables.get(typ, self.segs.get("ao" if typ == "ao" else "pa"))
                if full is None or hi > len(full):
                    raise SialSyntaxError(f"index range {kind}: no segment table of that length for index type {typ}")
                self.segs[kind] = list(full[lo - 1: hi])
                self._range_lo[kind] = lo
        self.tables = {}         # static arrays over simple indices only (index tables such as Xijk): {(i, j): value}
        self.idx = _Idx()        # index name -> current segment number (1-based)
        self.scopes = [dict()]   # temp blocks per open loop iteration: (name, segs) -> handle
        self.locals = {}         # allocated local arrays: name -> {segs: handle}
        self._aseg_plan, self._ext_of, self._segkey_plan = {}, {}, {}   # memoised label resolution (per distinct reference in the text)
        self.iteration = 0       # pardo iteration counter of the current barrier section
        self.scalars = {s: 0.0 for s in program.scalars}
        for name in program.predefined:
            if name in self.constants:
                self.scalars[name] = float(self.constants[name])
        self._own_memo = {}
        self._pending_persist = []   # (name, label) of the set_persistent statements executed so far: handed over at the end
        self.own_static = {}     # static arrays the program itself fills (St1a[a,i], SHDiag[a,i], ...): name -> {segs: handle}

    # ---- helpers -------------------------------------------------------------------------------------
    def _kind(self, lab):
        if lab[0] == "@":        # `(index)j`
            return "s"
        try:
            return self.p.index_kind[lab]
        except KeyError:
            raise SialSyntaxError(f"undeclared index {lab}") from None

    def _nseg(self, lab):
        return len(self.segs[self._kind(lab)])

    def _range(self, lab):
        """the values a loop over `lab` takes: segment numbers 1..nseg, or lo..hi of a simple index"""
        if self._kind(lab) != "s":
            return range(1, self._nseg(lab) + 1)
        lo, hi = (int(x) if x.lstrip('-').isdigit() else self.constants[x] for x in self.p.simple_range[lab])
        return range(lo, hi + 1)

    def _core(self, labs):
        """labels without the simple indices (their dimensions have extent 1 and do not affect the layout)"""
        kinds = self.p.index_kind
        return tuple(lab for lab in labs if kinds.get(lab) != "s")

    def _is_table(self, name):
        a = self.p.arrays.get(name)
        return a is not None and a[0] == "static" and all(self.p.index_kind.get(d) == "s" for d in a[1])

    def _eval(self, e):
        """host value of an int / scalar expression (parse_expr AST)"""
        k = e[0]
        if k == "num":
            return e[1]
        if k == "var":
            n = e[1]
            if n in self.idx:
                return self.idx[n]
            if n in self.scalars:
                v = self.be.value(self.scalars[n])
                return int(round(v)) if n in self.p.ints and float(v).is_integer() else v
            if n in self.constants:
                return self.constants[n]
            raise SialSyntaxError(f"undefined name {n} in expression")
        if k == "elem":
            if not self._is_table(e[1]):
                if all(self._kind(x) == "s" for x in e[2]):      # a one-element block of a local array over simple indices
                    return self.be.block_value(self._read(e[1], tuple(e[2]))[0])
                raise SialSyntaxError(f"{e[1]}: only arrays over simple indices can be read as numbers")
            return self.tables.get(e[1], {}).get(tuple(self.idx[x] for x in e[2]), 0.0)
        if k == "cast":
            v = self._eval(e[2])
            return int(round(v)) if e[1] == "int" else float(v)      # cast_double_to_int = lrint (sial_math.cpp:39)
        if k == "neg":
            return -self._eval(e[1])
        if k == "fn":
            v = self._eval(e[2])
            return math.sqrt(v) if e[1] == "sqrt" else abs(v)
        a, b = self._eval(e[1]), self._eval(e[2])
        if k == "/" and type(a) is int and type(b) is int:
            return int(math.copysign(abs(a) // abs(b), a * b)) if b else a / b      # C integer division
        return a + b if k == "+" else a - b if k == "-" else a * b if k == "*" else a ** b if k == "**" else a / b

    def _segs_of(self, labs):
        """identity of a temp / local block: ABSOLUTE segment numbers, as in the SIP (a block of `tpp[p,p2]` addressed as
        tpp[a,a2] and as tpp[b,k1] must not be confused when the loop values coincide: a virtual segment number counts
        on from the occupied ones; AO and simple indices live in ranges of their own)"""
        plan = self._segkey_plan.get(labs)
        if plan is None:
            base = {"o": 0, "p": 0, "v": len(self.segs.get("o", ())), "ao": 1 << 20, "s": 1 << 21}

            def key_base(kind):
                if ":" in kind:      # absolute segment number inside the index type's own range of keys
                    return (1 << 20 if kind.startswith("ao:") else 0) + self._range_lo[kind] - 1
                return base.get(kind, 0)

            plan = self._segkey_plan[labs] = [(lab, key_base(self._kind(lab))) for lab in labs]
        idx = self.idx
        return tuple(idx[lab] + b for lab, b in plan)

    def _shape(self, labs):
        ext = self._ext_of.get(labs)
        if ext is None:         # the extent table of each label's index type, resolved once per label tuple
            ext = self._ext_of[labs] = [(lab, self.segs[self._kind(lab)]) for lab in labs]
        idx = self.idx
        return tuple(e[idx[lab] - 1] for lab, e in ext)

    def _is_remote(self, name):
        return self.p.arrays.get(name, ("temp",))[0] in ("served", "distributed", "static")

    def _array_segs(self, name, labs):
        """block coordinates of `name[labs]` in the segment numbering of the array's DECLARED indices"""
        plan = self._aseg_plan.get((name, labs))
        if plan is None:        # (label, shift) per dimension, resolved once per distinct reference in the program text
            decl = self.p.arrays[name][1]
            if len(decl) != len(labs):
                raise SialSyntaxError(f"{name} has rank {len(decl)}")
            plan = []
            for d, lab in zip(decl, labs):
                dk, k = self._kind(d), self._kind(lab)
                if dk == "p" and k == "v":
                    plan.append((lab, len(self.segs["o"])))
                elif dk == "pa" and k in ("o", "v", "p", "pa"):    # all MO segments: absolute segment numbers
                    base = self.index_base.get("o", 0) if k in ("o", "p") else 0
                    plan.append((lab, self.index_base.get("v", len(self.segs["o"])) if k == "v" else base))
                elif dk == "s" and k == "s":       # block coordinate = value - first value + 1
                    lo = self.p.simple_range[d][0]
                    plan.append((lab, 1 - (int(lo) if lo.lstrip('-').isdigit() else self.constants[lo])))
                elif dk == k or (dk == "p" and k == "o"):
                    plan.append((lab, 0))
                elif (":" in dk or dk == "ao") and (":" in k or k == "ao") and dk.split(":")[0] == k.split(":")[0]:
                    plan.append((lab, self._range_lo.get(k, 1) - self._range_lo.get(dk, 1)))    # same index type, other range
                else:
                    raise SialSyntaxError(f"index {lab} ({k}) cannot address dimension {d} ({dk}) of {name}")
            self._aseg_plan[name, labs] = plan
        idx = self.idx
        return tuple(idx[lab] + shift for lab, shift in plan)

    def block_of(self, name, values):
        """handle of the block of a local / static array at the given values of its DECLARED indices (what the reference's test
        controllers read with `local_block(name, indices)`), or None"""
        decl = self.p.arrays[name][1]
        saved = self.idx.copy()
        self.idx.update(dict(zip(decl, values)))
        try:
            key = self._segs_of(decl)
            return self.locals.get(name, {}).get(key) or self.own_static.get(name, {}).get(key)
        finally:
            self.idx = saved

    def _own_static(self, name):
        """a `static` array that is not handed in by the harness: the program fills it itself (zero until then)"""
        own = self._own_memo.get(name)
        if own is None:
            a = self.p.arrays.get(name)
            own = self._own_memo[name] = (a is not None and a[0] == "static" and not self._is_table(name)
                                          and not self.be.has_array(name))
        return own

    def _find(self, name, labs):
        key = (name, self._segs_of(labs))
        if name in self.own_static or self._own_static(name):
            blocks = self.own_static.setdefault(name, {})
            if key[1] not in blocks:
                blocks[key[1]] = self.be.new_block(self._shape(labs))
                self.be.fill(blocks[key[1]], 0.0)
            return blocks[key[1]]
        if name not in self.locals and self.p.arrays.get(name, ("",))[0] == "local" and name not in self.p.contiguous:
            self.locals[name] = {}         # a local array written without `allocate` (put_test.sialx: `result[k] = x`): created on first touch
        if name in self.locals:            # allocated local array: zero-filled blocks that outlive the loop scopes
            blocks = self.locals[name]
            if key[1] not in blocks:
                blocks[key[1]] = self.be.new_block(self._shape(labs))
                self.be.fill(blocks[key[1]], 0.0)
            return blocks[key[1]]
        for sc in reversed(self.scopes):
            if key in sc:
                return sc[key]
        return None

    def _read(self, name, labs):
        """(handle, labels it is stored with) of an operand block"""
        if self._is_table(name):       # an element of a host table as a one-element block
            h = self.be.new_block(self._shape(labs))
            self.be.fill(h, self.tables.get(name, {}).get(tuple(self.idx[x] for x in labs), 0.0))
            self.scopes[-1][("%table", name, len(self.scopes[-1]))] = h      # dies with the loop iteration
            return h, labs
        if self._is_remote(name) and not self._own_static(name):
            fetch = self.be.static_block if self.p.arrays[name][0] == "static" else self.be.array_block
            return fetch(name, self._array_segs(name, labs), self._shape(labs)), labs
        h = self._find(name, labs)
        if h is None:
            raise SialSyntaxError(f"block {name}{list(labs)} read before it was written")
        return h, labs

    def _write(self, name, labs):
        if self._is_remote(name) and not self._own_static(name):
            # a block statement ON a distributed array (`Gmi_a[i1,i] = 0.0`, rlambda_rhf.sialx:1118): the interpreter writes the
            # worker's cached copy of the block, which never reaches the owner and is dropped at the next barrier
            # (sial_ops_parallel.cpp:41-47) -- here: a scratch block that dies with the loop iteration
            key = ("%cache", name, self._array_segs(name, labs))
            if key not in self.scopes[-1]:
                self.scopes[-1][key] = self.be.new_block(self._shape(labs))
            return self.scopes[-1][key]
        h = self._find(name, labs)
        if h is None:
            h = self.be.new_block(self._shape(labs))
            self.scopes[-1][(name, self._segs_of(labs))] = h
        return h

    def _leave_scope(self):
        for h in self.scopes.pop().values():
            self.be.free(h)

    # ---- execution -----------------------------------------------------------------------------------
    def run(self):
        self._block(self.p.body)
        self._hand_over()
        return self.scalars

    def run_proc(self, name):
        self._x_call(name)
        self._hand_over()
        return self.scalars

    def _hand_over(self):
        """`set_persistent` takes effect when the program ends (worker_persistent_array_manager.cpp:34-88: the arrays are
        moved to the persistent manager in `save_marked_arrays` after the last instruction), so `restore_persistent ca "ca"`
        directly followed by `set_persistent ca "ca"` (rccsd_rhf.sialx:229-232) keeps the array usable in between"""
        pending, self._pending_persist = self._pending_persist, []
        for name, label in pending:
            if self._is_table(name):           # a host table (e.g. the converged roots SEk0): handed over by value
                self.host_registry[label] = dict(self.tables.get(name, {}))
            elif name in self.own_static:
                self.host_registry[label] = self.own_static.pop(name)
            elif name in self.p.scalars:
                self.be.persist_scalar(label, self.be.value(self.scalars[name]))
            else:
                self.be.set_persistent(name, label)

    def _x_call(self, name):
        if name not in self.p.procs:
            raise SialSyntaxError(f"call of undefined proc {name}")
        self._block(self.p.procs[name])

    def _block(self, stmts):
        for st in stmts:
            if getattr(self, "_x_" + st[0])(*st[1:]) is False:
                return False
        return True

    def _x_with_vars(self, names, st):
        for n in names:
            self.idx[n] = int(self.be.value(self.scalars[n]))
        try:
            return getattr(self, "_x_" + st[0])(*st[1:])
        finally:
            for n in names:
                self.idx.pop(n, None)

    def _x_where(self, a, op, b):
        if isinstance(a, tuple):
            va, vb = self._eval(a), self._eval(b)
        else:
            va = self.idx[a] if a in self.idx else int(a) if a.isdigit() else self._eval(("var", a))
            vb = self.idx[b] if b in self.idx else int(b) if b.isdigit() else self._eval(("var", b))
        return {"<": va < vb, "<=": va <= vb, ">": va > vb, ">=": va >= vb, "==": va == vb, "!=": va != vb}[op]

    def _x_pardo(self, labs, body):
        wheres = [s for s in body if s[0] == "where"]
        rest = [s for s in body if s[0] != "where"]
        self.be.begin_pardo()
        try:
            ranges = [self._range(lab) for lab in reversed(labs)]   # first index fastest
            for combo in itertools.product(*ranges):
                for lab, v in zip(reversed(labs), combo):
                    self.idx[lab] = v
                if not all(self._x_where(*w[1:]) for w in wheres):
                    continue
                self.iteration += 1
                if (self.iteration - 1) % self.world != self.rank:
                    continue
                self.scopes.append({})
                self._block(rest)
                self._leave_scope()
            for lab in labs:
                del self.idx[lab]
        except BaseException:
            try:
                self.be.end_pardo()      # do not leave a recording open behind a failing program
            except Exception:
                pass
            raise
        self.be.end_pardo()

    def _x_do(self, labs, body):
        lab = labs[0]
        for v in self._range(lab):
            self.idx[lab] = v
            self.scopes.append({})
            try:
                self._block(body)      # a false `where` skips the rest of this iteration
            except _Exit:
                self._leave_scope()
                break
            self._leave_scope()
        del self.idx[lab]

    print_hook = None      # callable(name, index values, host array) for every `print <block>` statement

    def _x_print_block(self, name, labs):
        if self.print_hook is None:
            return
        h = self._read(name, labs)[0]
        self.print_hook(name, tuple(self.idx[x] for x in labs), self.be.host_array(h))

    def _x_if(self, lhs, op, rhs, body):
        va, vb = self._eval(lhs), self._eval(rhs)
        self._if_taken = taken = {"<": va < vb, "<=": va <= vb, ">": va > vb, ">=": va >= vb, "==": va == vb, "!=": va != vb}[op]
        if taken:
            res = self._block(body)   # a false `where` inside propagates
            self._if_taken = True     # (an if / else inside the body has used the flag)
            return res

    def _x_else(self, body):          # directly follows its `if` in the statement list
        if not self._if_taken:
            return self._block(body)

    def _x_exit(self):
        raise _Exit()

    def _x_sexpr(self, name, op, e):
        v = self._eval(e)
        if op != "=":
            cur = self.be.value(self.scalars[name])
            v = cur + v if op == "+=" else cur - v if op == "-=" else cur * v if op == "*=" else cur / v
        self.scalars[name] = self.be.scalar_set(self.scalars.get(name), v)

    def _x_allocate(self, name, labs):
        if self.p.arrays.get(name, ("",))[0] != "local":
            raise SialSyntaxError(f"allocate of {name}: not a local array")
        if "*" not in labs or any(x != "*" for x in labs):
            # one block, named by the current index values (rccsdpt_aab.sialx:2447), or one row of blocks (`allocate a[i,*]` inside
            # `do i`, local_arrays_wild.sialx): created, zero-filled, on first touch
            self.locals.setdefault(name, {})
            return
        if name in self.locals:
            if self.locals[name] and name not in self.p.contiguous:     # (empty: only single blocks were allocated, all freed)
                raise SialSyntaxError(f"allocate of {name}: already allocated")
            self._x_deallocate(name, labs)
        self.locals[name] = {}

    def _x_deallocate(self, name, labs):
        if name not in self.locals:
            if name in self.p.contiguous:
                return
            raise SialSyntaxError(f"deallocate of {name}: not allocated")
        if "*" not in labs:
            h = self.locals[name].pop(self._segs_of(labs), None)
            if h is not None:
                self.be.free(h)
            return
        for h in self.locals.pop(name).values():
            self.be.free(h)

    def _x_set_persistent(self, name, label):
        if not (self._is_table(name) or name in self.p.scalars or self._is_remote(name)):
            raise SialSyntaxError(f"set_persistent of {name}: not a scalar, served, distributed or static array")
        self._pending_persist = [(n, lab) for n, lab in self._pending_persist if lab != label] + [(name, label)]

    def _x_restore_persistent(self, name, label):
        if self._is_table(name):
            self.tables[name] = dict(self.host_registry.pop(label))
        elif self.p.arrays.get(name, ("",))[0] == "static" and label in self.host_registry:      # a static array a program filled itself
            self.own_static[name] = self.host_registry.pop(label)
            self._own_memo[name] = True
        elif name in self.p.scalars:
            self.scalars[name] = self.be.restore_scalar(label)
        elif self._is_remote(name):
            self.be.restore_persistent(name, label)
        else:
            raise SialSyntaxError(f"restore_persistent of {name}: not a scalar, served, distributed or static array")

    def _x_request(self, name, labs):
        self.be.request(name, self._array_segs(name, labs), self._shape(labs))

    def _x_fill(self, name, labs, v):
        if self._is_table(name):
            return self._table_set(name, labs, v)
        self.be.fill(self._write(name, labs), v)

    def _x_scale(self, name, labs, f):
        self.be.scale(self._write(name, labs), f)

    def _x_scale_by(self, name, labs, scalar):
        if scalar not in self.scalars:
            raise SialSyntaxError(f"undeclared scalar {scalar}")
        self.be.scale(self._write(name, labs), self.be.value(self.scalars[scalar]))

    def _static_blocks(self, name):
        """every block of a program-owned static array: [(per-dimension 0-based start, handle)], plus the dense extents"""
        decl = self.p.arrays[name][1]
        ranges, starts, dims = [], [], []
        for d in decl:
            vals = list(self._range(d))
            ext = [self.segs[self._kind(d)][v - 1] for v in vals] if self._kind(d) != "s" else [1] * len(vals)
            ranges.append(vals)
            starts.append([sum(ext[:k]) for k in range(len(vals))])
            dims.append(sum(ext))
        saved = self.idx.copy()
        out = []
        for combo in itertools.product(*[range(len(r)) for r in ranges]):
            for d, r, k in zip(decl, ranges, combo):
                self.idx[d] = r[k]
            out.append((tuple(st[k] for st, k in zip(starts, combo)), self._find(name, decl)))
        self.idx = saved
        return out, tuple(dims)

    def _static_dense(self, name):
        import numpy as np
        blocks, dims = self._static_blocks(name)
        full = np.zeros(dims, order="F")
        for start, h in blocks:
            a = self.be.host_array(h)
            full[tuple(slice(s, s + e) for s, e in zip(start, a.shape))] = a
        return full

    def _static_scatter(self, name, full):
        import numpy as np
        blocks, _ = self._static_blocks(name)
        for start, h in blocks:
            shape = self.be.host_array(h).shape
            self.be.set_from_host(h, np.asfortranarray(full[tuple(slice(s, s + e) for s, e in zip(start, shape))]))

    def _table_set(self, name, labs, v):
        self.tables.setdefault(name, {})[tuple(self.idx[x] for x in labs)] = float(v)

    def _x_fill_expr(self, name, labs, e):
        v = self._eval(e)
        if self._is_table(name):
            self._table_set(name, labs, v)
        else:
            self.be.fill(self._write(name, labs), v)

    def _x_incr_expr(self, name, labs, sign, e):
        v = sign * self._eval(e)
        if self._is_table(name):
            key = tuple(self.idx[x] for x in labs)
            t = self.tables.setdefault(name, {})
            t[key] = t.get(key, 0.0) + v
        else:
            self.be.increment(self._write(name, labs), v)    # Block::increment_elements, block.cpp:258-268

    def _x_scale_expr(self, name, labs, e):
        self.be.scale(self._write(name, labs), self._eval(e))

    def _x_put_scale(self, arr, alabs, e):
        self.be.put_scale(arr, self._array_segs(arr, alabs), self._shape(alabs), self._eval(e))

    def _x_put_value(self, arr, alabs, op, e):
        v = self._eval(e)
        if op == "=":
            self.be.put_initialize(arr, self._array_segs(arr, alabs), self._shape(alabs), v)
        else:
            self.be.put_increment(arr, self._array_segs(arr, alabs), self._shape(alabs), v)

    def _x_put_init(self, arr, alabs, v):
        self.be.put_initialize(arr, self._array_segs(arr, alabs), self._shape(alabs), v)

    def _views(self, d, labs, s, slabs):
        """Blocks whose labels differ only by simple indices (extent-1 dimensions, e.g. t1ppp[a2,i1,a] = tppps[a2,i1,a,jj])
        are the same run of elements: compare / permute on the labels that carry data."""
        cd, cs = self._core(labs), self._core(slabs)
        if cd != tuple(labs) or cs != tuple(slabs):
            d = self.be.reshaped(d, self._shape(cd))
            s = self.be.reshaped(s, self._shape(cs))
        return d, cd, s, cs

    def _x_assign(self, name, labs, src, slabs, _sign):
        if self._is_table(name):       # `GSmat[ksub1,ksub] = Tkk[ksub1,ksub]`, `SEkold[kstate] = SEk0[kstate]`: one number
            if self._is_table(src):
                v = self.tables.get(src, {}).get(tuple(self.idx[x] for x in slabs), 0.0)
            else:
                v = self.be.block_value(self._read(src, slabs)[0])
            return self._table_set(name, labs, v)
        s, sl = self._read(src, slabs)
        d, dl, s, sl = self._views(self._write(name, labs), labs, s, sl)
        self.be.copy(d, dl, s, sl)

    def _x_add(self, name, labs, src, slabs, sign):
        s, sl = self._read(src, slabs)
        d, dl, s, sl = self._views(self._write(name, labs), labs, s, sl)
        if tuple(sl) == tuple(dl):
            self.be.axpy(d, s, sign)
        else:   # handle_block_add: permute the rhs into a temp first (interpreter.cpp:1874-1997)
            t = self.be.new_block(self._shape(dl))
            self.be.copy(t, dl, s, sl)
            self.be.axpy(d, t, sign)
            self.be.free(t)

    def _x_addsub(self, name, labs, lname, llabs, rname, rlabs, sign):
        L, ll = self._read(lname, llabs)
        R, rl = self._read(rname, rlabs)
        if tuple(ll) != tuple(labs):
            raise SialSyntaxError("block add: the first operand must carry the destination's labels")
        d = self._write(name, labs)
        if tuple(rl) != tuple(labs):     # the interpreter permutes the second operand into a temp first
            t = self.be.new_block(self._shape(labs))
            self.be.copy(t, labs, R, rl)
            self.be.add_sub(d, L, t, sign)
            self.be.free(t)
        else:
            self.be.add_sub(d, L, R, sign)

    def _x_contract(self, name, labs, lname, llabs, rname, rlabs):
        L, ll = self._read(lname, llabs)
        R, rl = self._read(rname, rlabs)
        self.be.contract(self._write(name, labs), labs, L, ll, R, rl)

    def _x_put(self, arr, alabs, op, src, slabs):
        s, sl = self._read(src, slabs)
        if tuple(sl) != tuple(alabs):
            if self._core(sl) != self._core(alabs):
                raise SialSyntaxError("put/prepare needs matching labels on both sides")
            s = self.be.reshaped(s, self._shape(alabs))     # `PREPARE Daibj[a,i,b,j,kiter] = Taibj[a,i,b,j]`: extent-1 dimensions
        (self.be.put_accumulate if op == "+=" else self.be.put)(arr, self._array_segs(arr, alabs), s)

    def _x_execute(self, fname, args, bare):
        if fname in self.extra_si:
            return self.extra_si[fname](self, args, bare)
        if fname in ("compute_int_scratchmem", "print_block", "print_scalar", "test_print_block", "print_static_array"):
            return                                   # integral-engine scratch sizing / output: nothing on this path
        if fname in ("set_ijk_aab", "set_ijk_aaa"):  # occupied-triplet batches of the (T) programs: host logic
            self.tables[bare[0]] = set_ijk_aab(self.be.moa_seg_ranges(), self.constants["baocc"], self.constants["eaocc"],
                                               ordered_k=fname == "set_ijk_aaa")
            return
        if fname == "get_my_rank":
            self.scalars[bare[0]] = self.be.scalar_set(self.scalars.get(bare[0]), float(self.rank))
            return
        if fname == "compute_dipole_integrals":      # the dipole integral engine (OED package, out of scope): a resident table here
            # `execute compute_dipole_integrals DAOINT[mu,nu] ncount1 ncount2`: block = <mu| r_c |nu>, c = (int)ncount2; ncount1 = the
            # nuclear dipole component (compute_dipole_integrals.F: data_1(1) = dnuc).  Walker.host_data["dipole_integrals"] [3,nao,nao],
            # ["nuclear_dipole"] [3] come with the job like the AO repulsion integrals do.
            import numpy as np
            (name, labs), (s1, s2) = args[0], bare
            c = int(round(self.be.value(self.scalars[s2]))) - 1
            D = self.host_data["dipole_integrals"][c]
            sl = []
            for lab in labs:
                ext = self.segs[self._kind(lab)]
                lo = sum(ext[: self.idx[lab] - 1])
                sl.append(slice(lo, lo + ext[self.idx[lab] - 1]))
            self.be.set_from_host(self._write(name, labs), np.asfortranarray(D[tuple(sl)]))
            self.scalars[s1] = self.be.scalar_set(self.scalars.get(s1), float(self.host_data["nuclear_dipole"][c]))
            return
        if fname == "compute_diis":                  # the DIIS equations of the CC iterations: host LAPACK (dgesv), as in the reference
            B = bare[0]
            lo, hi = (int(x) if x.lstrip('-').isdigit() else self.constants[x] for x in self.p.simple_range[self.p.arrays[B][1][1]])
            n = hi - lo + 1
            t = self.tables.setdefault(B, {})
            c = compute_diis([[t.get((i + lo, j + lo), 0.0) for j in range(n)] for i in range(n)])
            for i in range(n):
                t[(i + lo, i + lo)] = float(c[i])
            return
        if fname == "eigen_calc":                    # dsyev of the (symmetric) CIS subspace matrix: host LAPACK, as in the reference
            G, V = bare
            lo, hi = (int(x) if x.lstrip('-').isdigit() else self.constants[x] for x in self.p.simple_range[self.p.arrays[G][1][0]])
            n = hi - lo + 1
            t = self.tables.get(G, {})
            a_out, vec = eigen_calc([[t.get((i + lo, j + lo), 0.0) for j in range(n)] for i in range(n)])
            self.tables[G] = {(i + lo, j + lo): float(a_out[i][j]) for i in range(n) for j in range(n)}
            self.tables[V] = {(i + lo, j + lo): float(vec[i][j]) for i in range(n) for j in range(n)}
            return
        if fname == "cis_unit_guess":                # unit starting vectors on the smallest Hamiltonian diagonal elements
            B, H = bare
            hd = self._static_dense(H)
            nsub = len(list(self._range(self.p.arrays[B][1][0])))
            self._static_scatter(B, cis_unit_guess(hd, nsub))
            return
        if fname == "gen_eigen_calc":                # dgeev of the Davidson subspace matrix: host LAPACK, as in the reference
            G, L, R, E = bare
            lo, hi = (int(x) if x.lstrip('-').isdigit() else self.constants[x] for x in self.p.simple_range[self.p.arrays[G][1][0]])
            n = hi - lo + 1
            t = self.tables.get(G, {})
            A = [[t.get((i + lo, j + lo), 0.0) for j in range(n)] for i in range(n)]
            vl, vr, ev = gen_eigen_calc(A)
            self.tables[L] = {(i + lo, j + lo): float(vl[i][j]) for i in range(n) for j in range(n)}
            self.tables[R] = {(i + lo, j + lo): float(vr[i][j]) for i in range(n) for j in range(n)}
            self.tables[E] = {(i + lo,): float(ev[i]) for i in range(n)}
            self.tables[G] = {}                      # dgeev overwrites its input
            return
        if fname == "return_sval" and args and self._is_table(args[0][0]):
            name, labs = args[0]
            v = self.tables.get(name, {}).get(tuple(self.idx[x] for x in labs), 0.0)
            self.scalars[bare[0]] = self.be.scalar_set(self.scalars.get(bare[0]), v)
            return
        if fname == "return_sval" and args and all(self._kind(x) == "s" for x in args[0][1]):     # a one-element block of a local array
            v = self.be.block_value(self._read(args[0][0], tuple(args[0][1]))[0])
            self.scalars[bare[0]] = self.be.scalar_set(self.scalars.get(bare[0]), v)
            return
        if fname == "energy_ty_denominator_rhf":     # `execute energy_ty_denominator_rhf T[a,i,b,j] fock_a shift`: shift by value
            bare = [bare[0], self.be.value(self.scalars[bare[1]])]
        blocks = [self._read(n, labs)[0] if self._is_remote(n) and not self._own_static(n) else self._write(n, labs) for n, labs in args]
        kinds = [[self._kind(x) for x in labs] for _, labs in args]
        segs = [tuple(self.idx[x] + (self._range_lo[k] - 1 if k in self._range_lo else self.index_base.get(k, 0)) for x, k in zip(labs, ks))
                for (_, labs), ks in zip(args, kinds)]
        self.be.execute(fname, blocks, segs, kinds, bare)

    def _x_sdot(self, name, lname, llabs, rname, rlabs):
        L, ll = self._read(lname, llabs)
        R, rl = self._read(rname, rlabs)
        self.scalars[name] = self.be.dot(L, ll, R, rl, self.scalars.get(name))

    def _x_sset(self, name, v):
        self.scalars[name] = self.be.scalar_set(self.scalars.get(name), v)

    def _x_sinc(self, name, v):
        self.scalars[name] = self.be.scalar_axpy(self.scalars[name], v, None)

    def _x_smul(self, name, v):
        self.scalars[name] = self.be.scalar_scale(self.scalars[name], v)

    def _x_sadd(self, name, other, sign):
        if other not in self.scalars:      # an index value or a predefined constant
            return self._x_sinc(name, sign * self._eval(("var", other)))
        self.scalars[name] = self.be.scalar_axpy(self.scalars[name], sign, self.scalars[other])

    def _x_scopy(self, name, other, _sign):
        if other not in self.scalars:      # `maxdav = davidson`: an index value or a predefined constant
            return self._x_sset(name, self._eval(("var", other)))
        self.scalars[name] = self.be.scalar_set(self.scalars.get(name), self.scalars[other])

    def _x_barrier(self):
        self.iteration = 0           # interpreter.cpp:205-209
        self.be.barrier()

    def _x_collective(self, a, b):
        self.scalars[a] = self.be.collective_sum(self.scalars[a], self.scalars[b])


class _Idx(dict):
    """index name -> current value; `@j` (an `(index)j` cast) reads j"""

    def __missing__(self, key):
        if key[:1] == "@":
            return self[key[1:]]
        raise KeyError(key)

    def copy(self):
        return _Idx(self)


_LABEL_NUMBERS = {}


def label_numbers(*label_lists):
    """labels -> small ints shared across the operands (the index-table slots the interpreter passes); memoised per
    distinct label pattern (a program has a few hundred at most), callers must not modify the result"""
    key = tuple(tuple(labs) for labs in label_lists)
    out = _LABEL_NUMBERS.get(key)
    if out is None:
        num = {}
        for labs in key:
            for x in labs:
                num.setdefault(x, len(num) + 1)
        out = _LABEL_NUMBERS[key] = [[num[x] for x in labs] for labs in key]
    return out


class StaticSlice:
    """A block of a static (contiguous) array that stays where it is: parent device block + 0-based first element +
    extents.  Only contractions take it (sipgpu_block_contract_sliced reads the parent through its strides), which is
    all the CC programs do with `ca[mu,p]`; the reference extracts a copy per use (contiguous_array_manager.cpp:202-230)."""

    def __init__(self, parent, beg, shape):
        self.parent, self.beg, self.shape = parent, tuple(beg), tuple(shape)


class DeviceBackend:
    """The per-block calls of the walker on libsipgpu (C ABI).  Holds no arithmetic."""

    def __init__(self, api, arrays, record=True, rank=0, world=1, barrier=None, allreduce=None, static_in_place=None):
        """arrays: name -> api.DistArray (served/distributed arrays, created by the caller).
        static_in_place: name -> (resident api.DeviceBlock holding the WHOLE static array, [segment extents per
        dimension]): blocks of these arrays are read in place by the contraction kernel instead of through `arrays`."""
        self.api, self.arrays, self.record = api, arrays, record
        self.static = dict(static_in_place or {})
        self.rank, self.world = rank, world
        self._barrier, self._allreduce = barrier, allreduce
        self.fock = None    # resident Fock diagonal block for `execute energy_denominator_rhf` (set by the caller)
        self.seg_ranges = []  # predefined int array moa_seg_ranges (set by the caller next to sipgpu_set_predefined_int_array)
        self.cache = {}     # remote blocks fetched since the last barrier (sial_ops_parallel.cpp:41-47)
        self.stats = []     # work-list statistics per pardo

    # scalars are host floats except while a pardo accumulates into them on the device
    def new_block(self, shape):
        return self.api.DeviceBlock(shape)

    def free(self, b):
        b.free()

    def begin_pardo(self):
        if self.record:
            self.api.wl_begin()

    def end_pardo(self):
        if self.record:
            self.stats.append(self.api.wl_end())

    def request(self, name, segs, shape):
        self.array_block(name, segs, shape)

    def array_block(self, name, segs, shape):
        A = self.arrays[name]
        if A.owner(segs) == self.rank:
            return A.block_view(segs)          # resident: no copy at all
        key = (name, segs)
        if key not in self.cache:
            self.cache[key] = A.get(segs)      # peer read over NVLink into the worker-side cache
        return self.cache[key]

    def static_block(self, name, segs, shape):
        if name not in self.static:
            return self.array_block(name, segs, shape)
        parent, seg_ext = self.static[name]
        beg = [sum(ext[: s - 1]) for ext, s in zip(seg_ext, segs)]
        return StaticSlice(parent, beg, shape)

    def fill(self, b, v):
        b.fill(v)

    def scale(self, b, f):
        b.scale(f)

    def axpy(self, d, s, f):
        d.axpy(s, f)

    def increment(self, b, v):
        b.increment(v)

    def add_sub(self, d, l, r, sign):
        d.set_add_sub(l, r, sign)

    def copy(self, d, dlabs, s, slabs):
        if tuple(dlabs) == tuple(slabs):
            d.scale_and_copy(s, 1.0)
        else:
            dn, sn = label_numbers(dlabs, slabs)
            self.api.permute_labels(dn, sn, s, out=d)

    def contract(self, d, dlabs, L, llabs, R, rlabs):
        dn, ln, rn = label_numbers(dlabs, llabs, rlabs)
        if isinstance(L, StaticSlice) or isinstance(R, StaticSlice):
            ptrn, ierr = self.api.get_contraction_ptrn(dn, ln, rn)
            if ierr:
                raise SialSyntaxError(f"illegal contraction pattern {dlabs} = {llabs} * {rlabs} (ierr {ierr})")
            lb, Lb = (L.beg, L.parent) if isinstance(L, StaticSlice) else (None, L)
            rb, Rb = (R.beg, R.parent) if isinstance(R, StaticSlice) else (None, R)
            self.api.contract_sliced(ptrn, Lb, L.shape, lb, Rb, R.shape, rb, d.shape, out=d)
        else:
            self.api.contract_labels(dn, d.shape, ln, L, rn, R, out=d)

    def put(self, arr, segs, b):
        self.arrays[arr].put(segs, b)

    def put_accumulate(self, arr, segs, b):
        self.arrays[arr].put_accumulate(segs, b)

    def put_initialize(self, arr, segs, shape, v):
        self.arrays[arr].put_initialize(segs, v)

    def put_increment(self, arr, segs, shape, v):
        self.arrays[arr].put_increment(segs, v)

    def put_scale(self, arr, segs, shape, f):
        """`prepare A[...] *= s`: the block is scaled where it lives (one writer per block and section, like `put`)"""
        self.arrays[arr].put_scale(segs, f)      # SialOpsParallel::put_scale, sial_ops_parallel.cpp:412-528

    def has_array(self, name):
        return name in self.arrays or name in self.static

    def block_value(self, b):
        """the single element of a one-element block, on the host (Davidson control logic: subspace matrix elements)"""
        return float(b.to_numpy().reshape(-1)[0])

    def host_array(self, b):
        return b.to_numpy()

    def set_from_host(self, b, a):
        import numpy as np
        a = np.asfortranarray(a, dtype=np.float64)
        self.api._check(self.api.lib().sipgpu_h2d(b.ptr, self.api._hp(a), b.size), "sipgpu_h2d")
        self.api.sync()

    # persistence: the array object (with its HBM slab) moves to the library's label registry and back -- no copy
    def set_persistent(self, name, label):
        self.cache.clear()
        self.arrays[name].persist(label)

    def restore_persistent(self, name, label):
        self.arrays[name].restore(label)

    def persist_scalar(self, label, value):
        self.api.persist_scalar(label, value)

    def restore_scalar(self, label):
        return self.api.restore_scalar(label)

    def execute(self, fname, blocks, segs, kinds, bare):
        if fname == "energy_denominator_rhf":
            self.api.si_energy_denominator_rhf(blocks[0], segs[0], self.fock)
        elif fname == "energy_ty_denominator_rhf":      # bare = (fock_a, shift): the scalar travels as a one-element device block
            shift = self.api.DeviceBlock((1,)).fill(float(bare[1]))
            if self.api.si_energy_ty_denominator_rhf(blocks[0], segs[0], self.fock, shift) != 0:
                raise SialSyntaxError("energy_ty_denominator_rhf failed")
            shift.free()
        elif fname == "stripi":
            if self.api.si_stripi(blocks[0], segs[0], blocks[1], segs[1]) != 0:
                raise SialSyntaxError("stripi: " + self.api.lib().sipgpu_last_error().decode(errors="replace"))
        elif fname in ("anti_symm_o", "anti_symm_v", "return_diagonal_elements"):
            if getattr(self.api, "si_" + fname)(blocks[0], segs[0]) != 0:
                raise SialSyntaxError(fname + " failed")
        elif fname == "invert_diagonal":
            if self.api.si_invert_diagonal(blocks[0], blocks[1]) != 0:
                raise SialSyntaxError("invert_diagonal failed")
        elif fname == "invert_diagonal_asym":
            if self.api.si_invert_diagonal_asym(blocks[0], segs[0], blocks[1]) != 0:
                raise SialSyntaxError("invert_diagonal_asym failed")
        else:
            raise SialSyntaxError(f"super-instruction {fname} is not on the device path")

    def reshaped(self, b, shape):
        """the same elements under other extents (extent-1 dimensions of simple indices dropped or added)"""
        return b if tuple(b.shape) == tuple(shape) else self.api.DeviceBlock(shape, ptr=b.ptr, owned=False)

    def moa_seg_ranges(self):
        return list(self.seg_ranges)

    def dot(self, L, llabs, R, rlabs, prev):
        acc = prev if isinstance(prev, self.api.DeviceBlock) else self.api.DeviceBlock((1,), zero=True)
        if tuple(llabs) != tuple(rlabs):
            t = self.api.DeviceBlock(R.shape)
            ln, rn = label_numbers(llabs, rlabs)
            self.api.permute_labels(rn, ln, L, out=t)
            L = t
        acc.fill(0.0)
        self.api._check(self.api.lib().sipgpu_block_dot_accumulate(L.ptr, R.ptr, L.size, acc.ptr))
        return acc

    def _val(self, s):
        return float(s.to_numpy()[0]) if isinstance(s, self.api.DeviceBlock) else float(s)

    def scalar_set(self, prev, v):
        if isinstance(v, self.api.DeviceBlock):
            out = prev if isinstance(prev, self.api.DeviceBlock) and prev is not v else self.api.DeviceBlock((1,))
            out.scale_and_copy(v, 1.0)
            return out
        if isinstance(prev, self.api.DeviceBlock):
            return prev.fill(v)
        return float(v)

    def scalar_axpy(self, s, f, other):
        if other is None:
            return s.increment(f) if isinstance(s, self.api.DeviceBlock) else s + f
        if isinstance(other, self.api.DeviceBlock):
            if not isinstance(s, self.api.DeviceBlock):
                s = self.api.DeviceBlock((1,)).fill(float(s))
            return s.axpy(other, f)
        return s.increment(f * other) if isinstance(s, self.api.DeviceBlock) else s + f * other

    def scalar_scale(self, s, f):
        return s.scale(f) if isinstance(s, self.api.DeviceBlock) else s * f

    def barrier(self):
        self.api.sync()
        if self._barrier:
            self._barrier()
        self.cache.clear()

    def collective_sum(self, a, b):
        v = self._val(b)
        if self._allreduce:
            v = self._allreduce(v)
        return self._val(a) + v

    def value(self, s):
        return self._val(s)
