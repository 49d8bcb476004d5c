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


def _builds_lists(create, merge_value, merge_combiners):
    """True for the list-collecting aggregator written by hand -- createCombiner(v) == [v], mergeValue appends,
    mergeCombiners concatenates (e.g. dpark.bagel.DefaultListCombiner, or the usual
    `combineByKey(lambda v: [v], lambda c, v: c + [v], lambda a, b: a + b)`): that is a groupByKey, whose GPU
    path yields every key's values as a list in (map split, position) order -- the list these functions build
    when they meet the rows in that order."""
    a, b, c, d = object(), object(), object(), object()
    try:
        one = create(a)
        if type(one) is not list or len(one) != 1 or one[0] is not a:
            return False
        two = merge_value([a], b)
        if type(two) is not list or len(two) != 2 or two[0] is not a or two[1] is not b:
            return False
        four = merge_combiners([a, b], [c, d])
        return type(four) is list and len(four) == 4 and all(x is y for x, y in zip(four, (a, b, c, d)))
    except Exception:
        return False


def recognize_aggregator(agg):
    """-> ("reduce", op) | ("group", None) for an Aggregator-like object
    (dpark/dependency.py:107-161), else NotImplementedError."""
    from .dependency import AddAggregator, GroupByAggregator, MergeAggregator
    if isinstance(agg, (GroupByAggregator, MergeAggregator)):
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
