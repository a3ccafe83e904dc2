"""Probe: the torch.nn.LSTM drop-in (forward-only, 2 x 512, the MD_VAE recipes' main RNN) vs cuDNN bf16 at 64 x 500 frames, fwd + bwd."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ml_vae_b200.modules import LSTM

dev = torch.device("cuda:0")
torch.manual_seed(0)
for B, T, In, H, layers, bidir in [(64, 500, 128, 512, 2, False), (128, 500, 128, 512, 2, False), (64, 500, 64, 512, 2, True)]:
    ref = torch.nn.LSTM(In, H, layers, batch_first=True, bidirectional=bidir).to(dev)
    m = LSTM(In, H, layers, batch_first=True, bidirectional=bidir).to(dev)
    m.load_state_dict(ref.state_dict())
    refb = torch.nn.LSTM(In, H, layers, batch_first=True, bidirectional=bidir).to(dev).bfloat16()
    refb.load_state_dict({k: v.bfloat16() for k, v in ref.state_dict().items()})
    x = torch.randn(B, T, In, device=dev).bfloat16()
    gy = torch.randn(B, T, H * (2 if bidir else 1), device=dev).bfloat16()

    def run(mod):
        xx = x.clone().requires_grad_(True)
        mod(xx)[0].backward(gy)

    def timed(mod, n=5):
        for _ in range(2):
            run(mod)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            run(mod)
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    t_ours, t_cudnn = timed(m), timed(refb)
    print(f"B={B} T={T} {In}->{H} x{layers} {'bi' if bidir else 'uni'}directional, fwd+bwd incl. all weight gradients: "
          f"drop-in {t_ours:.2f} ms, cuDNN bf16 {t_cudnn:.2f} ms (x{t_cudnn / t_ours:.1f})", flush=True)
