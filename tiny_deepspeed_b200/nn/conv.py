"""Convolution layers are NOT part of the engine — exactly like the reference, whose `core/module/conv.py` and
`core/module/ops/conv{1,2,3}d.py` are empty placeholders.  The names exist so that code probing for
them gets a clear error instead of an AttributeError."""


class _Unsupported:
    def __init__(self, *a, **k):
        raise NotImplementedError(f"{type(self).__name__}: convolution layers are placeholders in Tiny-DeepSpeed and here; "
                                  "supported parameterised layers are Linear, LayerNorm and Embedding")


class Conv1d(_Unsupported):
    pass


class Conv2d(_Unsupported):
    pass


class Conv3d(_Unsupported):
    pass
