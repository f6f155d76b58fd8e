"""Host values of device scalars that are fetched when they are USED, not when they are produced.

The reference's ``elbo`` returns its reconstruction terms as Python floats (``.item()`` / ``.tolist()``,
src/prob_unet.py:273-317, :325-381): a host synchronisation between the forward and the backward pass.  Its training
loop only appends them to a list and averages the list at the end of the epoch (src/train_prob_unet_model.py:144, :204).
``LazyScalar`` keeps that calling code working unchanged -- it converts, adds, compares and formats like a float and
numpy accepts it as a scalar -- while the device-to-host copy is enqueued asynchronously into pinned memory right where
the reference would have blocked; the host only waits (on that copy's event, not on the whole stream) when the value
is first looked at.  ``ProbabilisticUNet.sync_scalars = "lazy"`` selects it; ``True`` keeps real floats."""
import operator

import numpy as np
import torch


class HostFetch:
    """One asynchronous device -> pinned-host copy of a small tensor, shared by the LazyScalars that index into it."""
    __slots__ = ("_host", "_event", "_vals")

    def __init__(self, t):
        t = t.detach().reshape(-1)
        self._vals = None
        if t.is_cuda:
            self._host = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
            self._host.copy_(t, non_blocking=True)
            self._event = torch.cuda.Event()
            self._event.record()
        else:
            self._host, self._event = t.clone(), None

    def get(self, i):
        if self._vals is None:
            if self._event is not None:
                self._event.synchronize()
            self._vals = self._host.tolist()
            self._host = self._event = None
        return self._vals[i]


def _binary(op, reflected=False):
    if reflected:
        return lambda self, other: op(other, float(self))
    return lambda self, other: op(float(self), float(other) if isinstance(other, LazyScalar) else other)


class LazyScalar:
    """float-like view of element ``index`` of a HostFetch (or of a tensor)."""
    __slots__ = ("_fetch", "_index")
    __array_priority__ = 1000          # numpy scalars / arrays defer to the reflected operators below

    def __init__(self, source, index=0):
        self._fetch = source if isinstance(source, HostFetch) else HostFetch(source)
        self._index = index

    # ---- conversions
    def __float__(self):
        return float(self._fetch.get(self._index))

    def item(self):
        return float(self)

    def __int__(self):
        return int(float(self))

    def __bool__(self):
        return bool(float(self))

    def __round__(self, n=None):
        return round(float(self), n)

    def __array__(self, dtype=None, copy=None):
        return np.asarray(float(self), dtype=dtype)

    def __repr__(self):
        return repr(float(self))

    __str__ = __repr__

    def __format__(self, spec):
        return format(float(self), spec)

    def __hash__(self):
        return hash(float(self))

    # ---- arithmetic and comparisons: results are plain Python floats / bools
    __add__, __radd__ = _binary(operator.add), _binary(operator.add, True)
    __sub__, __rsub__ = _binary(operator.sub), _binary(operator.sub, True)
    __mul__, __rmul__ = _binary(operator.mul), _binary(operator.mul, True)
    __truediv__, __rtruediv__ = _binary(operator.truediv), _binary(operator.truediv, True)
    __floordiv__, __rfloordiv__ = _binary(operator.floordiv), _binary(operator.floordiv, True)
    __mod__, __rmod__ = _binary(operator.mod), _binary(operator.mod, True)
    __pow__, __rpow__ = _binary(operator.pow), _binary(operator.pow, True)
    __lt__, __le__ = _binary(operator.lt), _binary(operator.le)
    __gt__, __ge__ = _binary(operator.gt), _binary(operator.ge)
    __eq__, __ne__ = _binary(operator.eq), _binary(operator.ne)

    def __neg__(self):
        return -float(self)

    def __pos__(self):
        return float(self)

    def __abs__(self):
        return abs(float(self))


def lazy_list(t):
    """[LazyScalar, ...] over the elements of a small tensor, one shared copy (the lazy form of ``t.tolist()``)."""
    fetch = HostFetch(t)
    return [LazyScalar(fetch, i) for i in range(t.numel())]
