"""otpose_b200 -- B200-native (sm_100a) temporal fusion head of OTPose.

Drop-in module interfaces of the reference's hot path, computing through a
C-ABI CUDA library (include/otpose_b200.h):

    otpose_b200.model.OTPose.OTPose                        model/OTPose.py
    otpose_b200.model.ConvVideoTransformer.ConvTransformer model/ConvVideoTransformer.py
    otpose_b200.model.blocks.TransformerBlock ...          model/blocks.py
    otpose_b200.model.RSB.CHAIN_RSB_BLOCKS                 model/RSB.py
    otpose_b200.model.layers.DeformableCONV                model/layers.py
    otpose_b200.thirdparty.deform_conv.ModulatedDeformConv thirdparty/deform_conv
    otpose_b200.utils.heatmap.get_final_preds              utils/heatmap.py
"""
__version__ = "0.1.0"
