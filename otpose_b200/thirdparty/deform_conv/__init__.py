"""Drop-in for the reference package ``thirdparty.deform_conv`` (its ``__all__``,
thirdparty/deform_conv/__init__.py:8-13).  Only the modulated deformable
convolution is on the OTPose hot path; the other names exist so that
``from thirdparty.deform_conv import DeformConv, ModulatedDeformConv`` (reference
model/OTPose.py:16) keeps working, and raise ``NotImplementedError`` when used.
"""
from .deform_conv import (DeformConv, DeformConvPack, ModulatedDeformConv, ModulatedDeformConvPack,
                          deform_conv, modulated_deform_conv)


def _not_built(name):
    class _Stub:
        def __init__(self, *a, **k):
            raise NotImplementedError(f"{name} is never constructed by OTPose and is not built (SURVEY.md 2)")
    _Stub.__name__ = name
    return _Stub


DeformRoIPooling = _not_built("DeformRoIPooling")
DeformRoIPoolingPack = _not_built("DeformRoIPoolingPack")
ModulatedDeformRoIPoolingPack = _not_built("ModulatedDeformRoIPoolingPack")


def deform_roi_pooling(*a, **k):
    raise NotImplementedError("deform_roi_pooling is never called by OTPose and is not built")


__all__ = [
    'DeformConv', 'DeformConvPack', 'ModulatedDeformConv',
    'ModulatedDeformConvPack', 'DeformRoIPooling', 'DeformRoIPoolingPack',
    'ModulatedDeformRoIPoolingPack', 'deform_conv', 'modulated_deform_conv',
    'deform_roi_pooling'
]
