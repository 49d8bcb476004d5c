"""Per-operation shuffle flags, with the reference's names (dpark/conf.py:59-129).

Only `ordered_group` has an effect and the GPU path always behaves as if it were
True: groupByKey values come out ordered by (map split, position), which is the
deterministic member of the outcomes the reference's unordered mode allows
(dpark/shuffle.py:626-646).  The CPU-memory knobs (disk_merge, sort_merge,
iter_group, dump_mem_ratio) select out-of-core CPU mergers in the reference;
HBM-resident buffers are sized exactly from the histogram pass, so they are
accepted and ignored.
"""
OP_UDF, OP_GROUPBY, OP_COGROUP = "udf", "groupby", "cogroup"


class RDDConf(object):
    ATTRS = dict(disk_merge=False, sort_merge=False, iter_group=False, ordered_group=False,
                 dump_mem_ratio=0.9, op=OP_UDF)

    def __init__(self, **kw):
        for k, v in self.ATTRS.items():
            setattr(self, k, kw.get(k, v))
        unknown = set(kw) - set(self.ATTRS)
        if unknown:
            raise TypeError("unknown rddconf fields: %s" % sorted(unknown))

    def dup(self, **kw):
        d = dict((k, getattr(self, k)) for k in self.ATTRS)
        d.update(kw)
        return RDDConf(**d)

    @property
    def is_groupby(self):
        return self.op == OP_GROUPBY

    @property
    def is_cogroup(self):
        return self.op == OP_COGROUP

    def __repr__(self):
        return "RDDConf(%s)" % ", ".join("%s=%r" % (k, getattr(self, k)) for k in sorted(self.ATTRS))


default_rddconf = RDDConf()


def rddconf(disk_merge=None, sort_merge=None, iter_group=False, ordered_group=None, dump_mem_ratio=None,
            op=OP_UDF):
    """New RDDConf based on the defaults; keyword arguments only in spirit
    (dpark/conf.py:114-129)."""
    kw = dict(disk_merge=disk_merge, sort_merge=sort_merge, iter_group=iter_group,
              ordered_group=ordered_group, dump_mem_ratio=dump_mem_ratio, op=op)
    return default_rddconf.dup(**dict((k, v) for k, v in kw.items() if v is not None))
