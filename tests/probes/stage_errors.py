"""Probe: where does the bf16 step deviate from the float32 oracle?  Stage-by-stage errors at configs[1] shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import fbank_ref, vae_ref
from ml_vae_b200 import ops
from ml_vae_b200.features import Fbank
from ml_vae_b200.modules import Decoder, VanillaVAE
from ml_vae_b200.normalizer import InputNormalization
from ml_vae_b200.dense import linear, linear_chain

cuda = torch.device("cuda:0")
torch.manual_seed(123456)
B, seconds, latent, H, D = int(os.environ.get("PB", 64)), 5.0, 64, 512, 80
n = int(seconds * 16000)
g = torch.Generator().manual_seed(B + n)
lens_abs = (n * (0.5 + 0.5 * torch.rand(B, generator=g))).long() // 160 * 160
lens_abs[0] = n
lens_abs, _ = torch.sort(lens_abs, descending=True)
wav = torch.zeros(B, n)
for b in range(B):
    wav[b, : lens_abs[b]] = 0.1 * torch.randn(int(lens_abs[b]), generator=g)
enc = VanillaVAE([D, 64, 64], latent, materialize_loss=True).to(cuda)
dec = Decoder(latent, H, 2, 0.0, [2 * H, 64, 64, D], materialize_loss=True).to(cuda)
fb = Fbank(deltas=False, hop_length=10, n_mels=80)
norm = InputNormalization().to(cuda)
feats32, rel = fb(wav.to(cuda), lens_abs.to(cuda), truncate=True)
xb = norm(feats32, rel, epoch=0, out_dtype=torch.bfloat16)
T = xb.shape[1]
f_ref, frames = fbank_ref.batched_features(wav, lens_abs, deltas_=False, hop_length=10, n_mels=80)
rel_ref = frames.float() / T
x = vae_ref.GlobalNormRef()(f_ref, rel_ref)
mask = vae_ref.length_mask(rel_ref, T)                       # (B, T) valid frames


def report(name, a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    d = (a - b)
    dv = d[mask]
    bv = b[mask]
    print(f"{name:28s} all: max|d|/max|b| {float(d.abs().max() / b.abs().max()):.3e}   valid frames: max {float(dv.abs().max() / bv.abs().max()):.3e} "
          f"rms(d)/rms(b) {float(dv.pow(2).mean().sqrt() / bv.pow(2).mean().sqrt()):.3e}  mean(d)/rms(b) {float(dv.mean() / bv.pow(2).mean().sqrt()):.3e}", flush=True)


report("fbank", feats32, f_ref)
report("normalised feats (bf16)", xb, x)
eps = ops.philox_normal((B, T, latent), 77, 0)
ep = {k: v.detach().cpu() for k, v in enc.state_dict().items()}
dp = {k: v.detach().cpu() for k, v in dec.state_dict().items()}
with torch.no_grad():
    eo = enc(xb, eps=eps.bfloat16())
    er = vae_ref.encoder_forward(ep, x, eps.cpu())
    ws, bs = enc._trunk()
    h_ours = linear_chain(xb, ws, bs, end_activation=True)
    h_ref = torch.nn.functional.leaky_relu(vae_ref.fc_stack(ep, "fc.0.blocks", x), 0.01)
    report("enc trunk h", h_ours, h_ref)
    report("enc mean", eo["mean"], er["mean"])
    report("enc log_var", eo["log_var"], er["log_var"])
    report("sampled_h", eo["sampled_h"], er["sampled_h"])
    report("kl elem", eo["loss"], er["loss"])
    # same encoder math in float32 on the GPU modules (library path) as a cross-check of the oracle plumbing
    e32 = enc(x.to(cuda), eps=eps)
    report("enc mean (fp32 path)", e32["mean"], er["mean"])
    # decoder on the ORACLE's sampled_h (isolates the decoder) and on ours
    z_ref = er["sampled_h"]
    r1 = vae_ref.bilstm(dp, z_ref, H, 1)
    from ml_vae_b200 import lstm as lstm_mod
    ps = [getattr(dec.rnn, f"{kind}_l0{sfx}") for sfx in ("", "_reverse") for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    y1 = lstm_mod.bilstm_layer(z_ref.to(cuda).bfloat16(), *ps, training=False)
    report("lstm layer 1 (oracle z)", y1, r1)
    dref = vae_ref.decoder_forward(dp, z_ref, x, H, 2)
    do = dec(z_ref.to(cuda).bfloat16(), xb)
    report("decoder rnn_out->mean", do["mean"], dref["mean"])
    report("decoder log_var", do["log_var"], dref["log_var"])
    report("recon elem", do["losses"]["recon_loss"], dref["losses"]["recon_loss"])
    d32 = dec(z_ref.to(cuda), x.to(cuda))
    report("decoder mean (fp32 path)", d32["mean"], dref["mean"])
    kl_o = vae_ref.masked_reduce(er["loss"], rel_ref); kl = ops.masked_reduce(eo["loss"], rel) if hasattr(ops, "masked_reduce") else None
    print("kl mean ref", float(kl_o), "ours(elem,masked)", None if kl is None else float(kl))
    rc_o = vae_ref.masked_reduce(dref["losses"]["recon_loss"], rel_ref)
    rc = ops.masked_reduce(do["losses"]["recon_loss"], rel) if hasattr(ops, "masked_reduce") else None
    print("recon mean ref", float(rc_o), "ours", None if rc is None else float(rc))
