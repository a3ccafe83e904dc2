"""Stand-alone probe: persistent LSTM layer fwd+bwd vs torch (run under `timeout`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_vae_b200 import _lib as L
from ml_vae_b200.lstm import bilstm_layer

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30))


def run(B, T, In, H, time_it=False):
    torch.manual_seed(B + T + H)
    ref = torch.nn.LSTM(In, H, 1, bidirectional=True, batch_first=True).to(dev)
    with torch.no_grad():
        for p_ in ref.parameters():
            p_.copy_(p_.bfloat16().float())
    x = torch.randn(B, T, In, device=dev).bfloat16()
    gy = torch.randn(B, T, 2 * H, device=dev).bfloat16()
    xr = x.float().requires_grad_(True)
    yr, _ = ref(xr)
    (yr * gy.float()).sum().backward()
    names = ["weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0", "weight_ih_l0_reverse", "weight_hh_l0_reverse",
             "bias_ih_l0_reverse", "bias_hh_l0_reverse"]
    ps = [getattr(ref, n).detach().clone().requires_grad_(True) for n in names]
    xm = x.clone().requires_grad_(True)
    y = bilstm_layer(xm, *ps, training=True)
    (y.float() * gy.float()).sum().backward()
    torch.cuda.synchronize()
    errs = {"y": rel(y, yr), "dx": rel(xm.grad, xr.grad)}
    for n, p_ in zip(names, ps):
        errs["d" + n.replace("weight_", "w").replace("bias_", "b").replace("_l0", "").replace("_reverse", "R")] = rel(p_.grad, getattr(ref, n).grad)
    worst = max(errs.values())
    print(f"B={B} T={T} In={In} H={H}: " + " ".join(f"{k}={v:.1e}" for k, v in errs.items()), "OK" if worst < 3e-2 else "MISMATCH", flush=True)
    if time_it:
        def step():
            xm.grad = None
            yy = bilstm_layer(xm, *ps, training=True)
            yy.backward(gy)
        for _ in range(2): step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): step()
        b.record(); torch.cuda.synchronize()
        ours = a.elapsed_time(b) / 3
        prof = torch.zeros(128, dtype=torch.int64, device=dev)
        yy = bilstm_layer(xm, *ps, training=True)
        L.check(L.lib().mlvae_debug_set_profile_buffer(L.ptr(prof)), "prof")
        yy.backward(gy); torch.cuda.synchronize()
        L.check(L.lib().mlvae_debug_set_profile_buffer(None), "prof")
        print("   bwd cycles/step:", dict(zip(["exchange wait", "reduce", "gate grads + hand-off", "mma completion + scatter", "issuer 0: wait dA", "issuer 0: issue + commit", "issuer 1: wait dA", "issuer 1: issue + commit"], [round(v / T) for v in prof.cpu().tolist()[:8]])))
        pc = prof.cpu().tolist()
        print("   per CTA of group 0 (gate warp 0, chain 0) [wait, reduce, gates, mma+scatter]:")
        for cta in range(H // 32):
            print("     cta", cta, [round(v / T) for v in pc[12 + 4 * cta: 16 + 4 * cta]])
        print("   per gate warp of CTA 0 chain 0:")
        for w in range(1, 8):
            print("     warp", w, [round(v / T) for v in pc[76 + 4 * w: 80 + 4 * w]])
        lb = ref.bfloat16()
        def rstep():
            xb = x.clone().requires_grad_(True)
            yy, _ = lb(xb)
            yy.backward(gy)
        for _ in range(2): rstep()
        torch.cuda.synchronize(); a.record()
        for _ in range(3): rstep()
        b.record(); torch.cuda.synchronize()
        print(f"   layer fwd+bwd: ours {ours:.2f} ms vs cuDNN bf16 {a.elapsed_time(b) / 3:.2f} ms", flush=True)
    return worst < 3e-2


import sys
if len(sys.argv) > 1:                      # python tests/lstm_bwd_probe.py B T  -> instrumented timing of that shape only
    run(int(sys.argv[1]), int(sys.argv[2]), 64, 512, time_it=True)
    sys.exit(0)
ok = True
ok &= run(4, 6, 16, 32)
ok &= run(16, 20, 24, 64)
ok &= run(20, 33, 64, 128)
ok &= run(7, 40, 48, 256)
ok &= run(64, 50, 64, 512)
ok &= run(64, 500, 64, 512, time_it=True)
ok &= run(64, 500, 1024, 512, time_it=True)
print("ALL OK" if ok else "FAILED")
