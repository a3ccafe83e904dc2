"""One persistent-LSTM layer fwd+bwd at benchmark size for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ml_vae_b200.lstm import bilstm_layer
dev = torch.device("cuda:0")
B, T, In, H = 64, 500, int(os.environ.get('NCU_IN', 64)), 512          # the benched shape (BASELINE configs[1]), layer 0 or 1 (NCU_IN=1024)
torch.manual_seed(0)
ref = torch.nn.LSTM(In, H, 1, bidirectional=True, batch_first=True).to(dev)
names = [f"{k}_l0{s}" for s in ("", "_reverse") for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
ps = [getattr(ref, n).detach().clone().requires_grad_(True) for n in names]
x = torch.randn(B, T, In, device=dev).bfloat16().requires_grad_(True)
gy = torch.randn(B, T, 2 * H, device=dev).bfloat16()
for _ in range(int(os.environ.get('NCU_PASSES', 2))):
    y = bilstm_layer(x, *ps, training=True)
    y.backward(gy)
torch.cuda.synchronize()
print("done")
