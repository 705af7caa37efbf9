"""Sub-circuit base class, mirroring the reference's ``module/subckt.py``.

A Subckt is a flat list of ``(name, device)`` pairs; local nodes / devices are
made unique with a process-global instance counter (``<id>@node``,
``<id>#device``; reference module/subckt.py:29,54-66), which fixes the row
order -- and therefore the sparsity pattern -- of circuits built from IBIS
pins.
"""

_subckt_count = 0


def reset_counter(value=0):
    """Restart the global instance counter (tests want reproducible names)."""
    global _subckt_count
    _subckt_count = value


class Subckt(list):
    def __new__(cls, *args, **kwargs):
        global _subckt_count
        obj = list.__new__(cls)
        obj.__dict__["subcktCnt"] = _subckt_count
        _subckt_count += 1
        return obj

    def node(self, name):
        """Name of a node local to this sub-circuit instance."""
        return "%s@%s" % (self.subcktCnt, name)

    def device(self, name):
        """Name of a device local to this sub-circuit instance."""
        return "%s#%s" % (self.subcktCnt, name)

    def __setattr__(self, name, value):
        self.__dict__[name] = value
        if isinstance(value, Subckt):
            for pair in value:
                self.append(pair)
        else:
            self.append((self.device(name), value))
