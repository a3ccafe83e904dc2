"""Probe: persistent LSTM fwd / bwd kernel times alone (B=64, T=500, H=512).  Run under `timeout`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from ml_vae_b200 import _lib as L
L.LIB_PATH = os.environ.get('AB_LIB', L.LIB_PATH)      # A/B against another build of the library
from ml_vae_b200.lstm import _gate_perm

dev = torch.device("cuda:0")
B, T, H, In = int(os.environ.get("LSTM_B", 64)), 500, 512, 64
torch.manual_seed(1)
lstm = torch.nn.LSTM(In, H, 1, bidirectional=True, batch_first=True).to(dev)
x = torch.randn(B, T, In, device=dev)
wih = torch.cat([lstm.weight_ih_l0, lstm.weight_ih_l0_reverse], 0)
bias = torch.cat([lstm.bias_ih_l0 + lstm.bias_hh_l0, lstm.bias_ih_l0_reverse + lstm.bias_hh_l0_reverse], 0)
perm, _ = _gate_perm(H, dev)
with torch.no_grad():
    P0 = (x @ wih[perm].t() + bias[perm]).bfloat16().contiguous()
    ref, _ = lstm.bfloat16().float()(x)
whh = torch.stack([lstm.weight_hh_l0, lstm.weight_hh_l0_reverse], 0).bfloat16().contiguous()
Y = torch.empty(B, T, 2 * H, device=dev, dtype=torch.bfloat16)
C = torch.empty(B, T, 2 * H, device=dev)
dY = torch.randn(B, T, 2 * H, device=dev).bfloat16()
db = torch.empty((B + 15) // 16, 2, 4 * H, device=dev)
scratch = torch.empty(L.lib().mlvae_lstm_scratch_bytes(B, H), dtype=torch.uint8, device=dev)
lib = L.lib()


def fwd(P):
    L.check(lib.mlvae_lstm_fwd(L.ptr(P), L.ptr(whh), L.ptr(Y), L.ptr(C), B, T, H, 1, L.ptr(scratch), L.stream_ptr()), "f")


def bwd(G):
    L.check(lib.mlvae_lstm_bwd(L.ptr(G), L.ptr(C), L.ptr(dY), L.ptr(whh), L.ptr(db), B, T, H, L.ptr(scratch), L.stream_ptr()), "b")


def timed(fn, make, n=5):
    bufs = [make() for _ in range(n + 2)]
    for i in range(2):
        fn(bufs[i])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(bufs[2 + i])
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


G0 = None
base = None
for delay in (sys.argv[1:] or ["600:0", "600:0"]):          # "fwd:bwd" poll delays in cycles (or one number for both)
    df, db_ = (delay.split(":") + [delay])[:2]
    if lib.mlvae_debug_set_option(3, int(df)) != 0 or lib.mlvae_debug_set_option(4, int(db_)) != 0:
        print("(library has no poll-delay knob)")
    P = P0.clone(); fwd(P); torch.cuda.synchronize()
    err = float((Y.float() - ref).abs().max() / ref.abs().max())
    if G0 is None:
        G0 = P.clone()
    G = G0.clone(); bwd(G); torch.cuda.synchronize()
    sig = float(G.float().abs().sum()) + float(db.abs().sum())
    base = base or sig
    tf = timed(fwd, lambda: P0.clone())
    tb = timed(bwd, lambda: G0.clone())
    print(f"run {delay}: fwd {tf:.3f} ms  bwd {tb:.3f} ms  (y err {err:.2e}, bwd checksum drift {abs(sig - base) / base:.1e})", flush=True)
