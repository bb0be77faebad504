"""Drop-in ``DeformableCONV`` shim (reference model/layers.py:9-29)."""
from torch import nn

from ..thirdparty.deform_conv import ModulatedDeformConv


def modulated_deform_conv(n_channels, kernel_height, kernel_width, deformable_dilation, deformable_groups):
    return ModulatedDeformConv(n_channels, n_channels, (kernel_height, kernel_width), stride=1,
                               padding=int(kernel_height / 2) * deformable_dilation,
                               dilation=deformable_dilation, deformable_groups=deformable_groups)


class DeformableCONV(nn.Module):
    def __init__(self, num_joints, k, dilation):
        super().__init__()
        self.deform_conv = modulated_deform_conv(num_joints, k, k, dilation, num_joints)

    def forward(self, x, offsets, mask):
        return self.deform_conv(x, offsets, mask)
