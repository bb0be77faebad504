"""Time the fused HRNet.final_layer + frame-sum kernel (a0 + a1) at bench size against its algorithmic bytes:
read frames*B*Cin*T*e (e = element size) + write (frames + 1)*B*17*T*4 + B*T*4.

    python scripts/time_final_layer.py [clips]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from otpose_b200 import _lib  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
frames, cin, j, h, w = 5, 48, 17, 96, 72
t = h * w
lib = _lib.load()
g = torch.Generator().manual_seed(0)
wt = (torch.randn(j, cin, generator=g) / cin ** 0.5).cuda()
bias = torch.randn(j, generator=g).cuda()
rough = torch.empty((frames * b, j, h, w), device="cuda")
total_b = torch.empty((b, j, h, w), device="cuda")
squeezed = torch.empty((b, 1, h, w), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, td, nhwc in (("fp32 NCHW", torch.float32, 0), ("bf16 NHWC", torch.bfloat16, 1), ("fp16 NHWC", torch.float16, 1),
                       ("fp32 NHWC", torch.float32, 1)):
    feats = torch.randn(frames * b, cin, h, w, generator=g).to(td).cuda()
    if nhwc:
        feats = feats.contiguous(memory_format=torch.channels_last)
    code = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[td]

    def run():
        _lib.check(lib.otp_final_layer_fusion_sum(feats.data_ptr(), code, nhwc, wt.data_ptr(), bias.data_ptr(), frames, b,
                                                  cin, j, t, rough.data_ptr(), total_b.data_ptr(), squeezed.data_ptr(),
                                                  None), "otp_final_layer_fusion_sum")
    for _ in range(3):
        run()
    ms = []
    for _ in range(10):
        flush.zero_()                       # 256 MB > L2: the next run reads from HBM
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms = sorted(ms)[len(ms) // 2]
    nbytes = frames * b * cin * t * feats.element_size() + (frames + 1) * b * j * t * 4 + b * t * 4
    print(f"{name}: {ms * 1e3:7.1f} us  {nbytes / ms / 1e6:7.0f} GB/s algorithmic ({nbytes / 1e6:.0f} MB)")
