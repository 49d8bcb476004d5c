"""Recognise the combiner a user passed to reduceByKey / combineByKey.

The reference calls an opaque Python function per row
(dpark/task.py:222-226, dpark/shuffle.py:604-605).  The CUDA path needs a
closed op set, so the function is traced ONCE with symbolic operands and the
resulting expression is matched against {sum, prod, min, max, and, or, xor};
the match is then cross-checked on concrete samples.  Anything else raises
NotImplementedError -- there is no CPU fallback for the shuffle
(BASELINE.json north_star).
"""
import operator

OPS = ("sum", "min", "max", "prod", "and", "or", "xor")


class _Unsupported(Exception):
    pass


class _Sym(object):
    """Symbolic operand.  Arithmetic builds an expression tuple; a comparison
    is answered from `script` (one bool per comparison, in call order) and
    recorded, so `min`, `max` and `a if a < b else b` can be traced."""
    __slots__ = ("expr", "ctx")

    def __init__(self, expr, ctx):
        self.expr, self.ctx = expr, ctx

    def _bin(self, op, other, swap=False):
        o = other.expr if isinstance(other, _Sym) else ("const", other)
        a, b = (o, self.expr) if swap else (self.expr, o)
        return _Sym((op, a, b), self.ctx)

    def __add__(self, o): return self._bin("sum", o)
    def __radd__(self, o): return self._bin("sum", o, True)
    def __mul__(self, o): return self._bin("prod", o)
    def __rmul__(self, o): return self._bin("prod", o, True)
    def __and__(self, o): return self._bin("and", o)
    def __rand__(self, o): return self._bin("and", o, True)
    def __or__(self, o): return self._bin("or", o)
    def __ror__(self, o): return self._bin("or", o, True)
    def __xor__(self, o): return self._bin("xor", o)
    def __rxor__(self, o): return self._bin("xor", o, True)

    def _cmp(self, op, other):
        if not isinstance(other, _Sym):
            raise _Unsupported("comparison with a constant")
        ctx = self.ctx
        i = len(ctx["cmps"])
        ans = ctx["script"][i] if i < len(ctx["script"]) else False
        ctx["cmps"].append((op, self.expr, other.expr, ans))
        return ans

    def __lt__(self, o): return self._cmp("lt", o)
    def __le__(self, o): return self._cmp("le", o)
    def __gt__(self, o): return self._cmp("gt", o)
    def __ge__(self, o): return self._cmp("ge", o)

    def __bool__(self):
        raise _Unsupported("truth value of a symbolic operand")

    def __getattr__(self, name):
        raise _Unsupported("attribute %r of a combiner operand" % name)

    def __getitem__(self, i):
        raise _Unsupported("indexing a combiner operand (tuple-valued combiners are not supported yet)")


def _trace(func, script):
    ctx = {"cmps": [], "script": script}
    x, y = _Sym("x", ctx), _Sym("y", ctx)
    r = func(x, y)
    return (r.expr if isinstance(r, _Sym) else ("const", r)), ctx["cmps"]


_SAMPLES_INT = [(3, 5), (5, 3), (4, 4), (-7, 2), (0, -1), (6, 10), (12, 7)]
_SAMPLES_FLT = [(0.5, 2.25), (2.25, 0.5), (-1.5, -1.25), (3.0, 3.0)]
_CONCRETE = {
    "sum": operator.add, "prod": operator.mul, "min": min, "max": max,
    "and": operator.and_, "or": operator.or_, "xor": operator.xor,
}


def _verify(func, op):
    for a, b in _SAMPLES_INT:
        if func(a, b) != _CONCRETE[op](a, b):
            return False
    if op in ("sum", "prod", "min", "max"):
        for a, b in _SAMPLES_FLT:
            if func(a, b) != _CONCRETE[op](a, b):
                return False
    return True


def recognize_binary(func):
    """Name of the op `func(x, y)` computes, or raise NotImplementedError."""
    if func in (operator.add, operator.iadd):
        return "sum"
    if func in (operator.mul, operator.imul):
        return "prod"
    if func is min:
        return "min"
    if func is max:
        return "max"
    if func in (operator.and_, operator.iand):
        return "and"
    if func in (operator.or_, operator.ior):
        return "or"
    if func in (operator.xor, operator.ixor):
        return "xor"
    why = "could not trace"
    try:
        expr, cmps = _trace(func, [False])
        op = None
        if not cmps:
            if isinstance(expr, tuple) and len(expr) == 3 and expr[0] in OPS and \
                    {expr[1], expr[2]} == {"x", "y"}:
                op = expr[0]
            else:
                why = "expression %r is not a single commutative op of both operands" % (expr,)
        elif len(cmps) == 1:
            expr_t, _ = _trace(func, [True])
            # which operand does it return when the comparison is false / true?
            c_op, lhs, rhs, _ = cmps[0]
            if {expr, expr_t} == {"x", "y"} and {lhs, rhs} == {"x", "y"}:
                # comparison `lhs <c_op> rhs` true -> returns expr_t
                smaller_if_true = lhs if c_op in ("lt", "le") else rhs
                op = "min" if expr_t == smaller_if_true else "max"
            else:
                why = "comparison-based combiner does not return one of its operands"
        else:
            why = "more than one comparison"
        if op is not None and _verify(func, op):
            return op
        if op is not None:
            why = "traced as %s but disagrees on concrete samples" % op
    except _Unsupported as e:
        why = str(e)
    except NotImplementedError:
        raise
    except Exception as e:  # the function did something the tracer cannot follow
        why = "%s: %s" % (type(e).__name__, e)
    raise NotImplementedError(
        "reduceByKey/combineByKey function %r is not recognised as one of %s (%s); the B200 shuffle has no "
        "CPU fallback for opaque combiners" % (getattr(func, "__name__", func), ", ".join(OPS), why))


def _is_identity(f):
    try:
        token = object()
        return f(token) is token
    except Exception:
        return False


_LIST_OPS_FORBIDDEN = ("POP_JUMP", "JUMP", "COMPARE_OP", "CONTAINS_OP", "IS_OP", "BINARY_SLICE", "BUILD_SLICE",
                       "FOR_ITER", "GET_ITER", "BINARY_SUBSCR", "STORE_SUBSCR", "DELETE_SUBSCR", "MAKE_FUNCTION",
                       "LOAD_GLOBAL", "LOAD_NAME", "LOAD_DEREF", "LOAD_CLOSURE", "IMPORT", "YIELD", "RAISE")


def _straight_line_list_code(fn):
    """The function's bytecode has no branch, comparison, slice, subscript, loop, nested function or reference to
    anything outside its arguments, and touches no attribute but append/extend: whatever it does, it does the same
    for lists of every length and content.  (A combiner like `c + [v] if len(c) < 5 else c` or a de-duplicating one
    behaves like append on small probes; its bytecode gives it away.)"""
    import dis
    code = getattr(fn, "__code__", None) or getattr(getattr(fn, "__func__", None), "__code__", None)
    if code is None:
        return False
    for ins in dis.get_instructions(code):
        if any(ins.opname.startswith(bad) for bad in _LIST_OPS_FORBIDDEN):
            return False
        if ins.opname in ("LOAD_ATTR", "LOAD_METHOD") and ins.argval not in ("append", "extend"):
            return False
    return True


def _builds_lists(create, merge_value, merge_combiners):
    """True for the list-collecting aggregator written by hand -- createCombiner(v) == [v], mergeValue appends,
    mergeCombiners concatenates (e.g. the usual `combineByKey(lambda v: [v], lambda c, v: c + [v], lambda a, b: a + b)`):
    that is a groupByKey, whose GPU path yields every key's values as a list in (map split, position) order -- the list
    these functions build when they meet the rows in that order.

    Sound by construction, not by sampling alone: all three functions must be straight-line code over their
    arguments (_straight_line_list_code) AND behave as append/concatenate on probes of several sizes with repeated
    elements.  Anything else (bounded lists, de-duplication, sorting ...) is NOT a group-by and falls through to the
    NotImplementedError of the reduce-style recogniser."""
    if not all(_straight_line_list_code(f) for f in (create, merge_value, merge_combiners)):
        return False
    a, b = object(), object()
    try:
        one = create(a)
        if type(one) is not list or len(one) != 1 or one[0] is not a:
            return False
        for n in (1, 2, 7, 64, 200):
            left = [a, b, a][:min(n, 3)] + [object() for _ in range(max(0, n - 3))]
            v = left[0]                                   # a value already in the list (de-duplication would drop it)
            got = merge_value(list(left), v)
            if type(got) is not list or len(got) != n + 1 or any(x is not y for x, y in zip(got, left + [v])):
                return False
            right = list(reversed(left)) + [a]
            both = merge_combiners(list(left), list(right))
            if type(both) is not list or len(both) != 2 * n + 1 or any(x is not y for x, y in zip(both, left + right)):
                return False
        return True
    except Exception:
        return False


def recognize_aggregator(agg):
    """-> ("reduce", op) | ("group", None) for an Aggregator-like object
    (dpark/dependency.py:107-161), else NotImplementedError."""
    from .dependency import AddAggregator, GroupByAggregator, MergeAggregator
    if isinstance(agg, (GroupByAggregator, MergeAggregator)):
        return "group", None
    from . import bagel
    if isinstance(agg, bagel.DefaultListCombiner):       # known list collector (dpark/bagel.py:40-49)
        return "group", None
    if isinstance(agg, AddAggregator):
        return "reduce", "sum"
    create = getattr(agg, "createCombiner", None)
    mv = getattr(agg, "mergeValue", None)
    mc = getattr(agg, "mergeCombiners", None)
    if create is None or mv is None or mc is None:
        raise NotImplementedError("aggregator %r lacks createCombiner/mergeValue/mergeCombiners" % (agg,))
    if _builds_lists(create, mv, mc):
        return "group", None
    if not _is_identity(create):
        raise NotImplementedError(
            "createCombiner of %r is not the identity; only reduce-style aggregators (identity, f, f) and the "
            "group-by aggregators are supported on the B200 path" % (agg,))
    op1 = recognize_binary(mv)
    op2 = recognize_binary(mc)
    if op1 != op2:
        raise NotImplementedError("mergeValue (%s) and mergeCombiners (%s) differ" % (op1, op2))
    return "reduce", op1
