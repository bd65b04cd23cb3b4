"""Generate the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference classes.

Runs only where /root/reference exists (the build container); the fixtures it writes are committed and are what the
`-m "not gpu"` (oracle) and `-m gpu` (CUDA) parity tests read. It also pins the oracle:
  * oracle vs the reference classes (fp64) on every fixture;
  * the oracle's timm-block restatement vs nn.TransformerEncoderLayer(norm_first=True, gelu, eps 1e-6) and vs
    HF transformers' ViTLayer (fp64) — the two independent pre-norm ViT blocks available offline.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import reference_math as R  # noqa: E402
from oracle import timm_shim  # noqa: E402


def import_reference():
    timm_shim.install()
    sys.path.insert(0, REF)
    # the reference packages are top-level `models_fer_vit` and `modules`
    import importlib
    mods = {}
    for name in ("models_fer_vit.latent_vit", "models_fer_vit.latent_vit_v2", "models_fer_vit.hybrid_latent_vit",
                 "models_fer_vit.image_vit", "modules"):
        mods[name] = importlib.import_module(name)
        assert mods[name].__file__.startswith(REF), mods[name].__file__
    return mods


def to64(sd):
    return {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}


def run_reference(model, x, y, weight, smoothing, dtype):
    model = model.to(dtype)
    model.zero_grad()
    logits = model(x.to(dtype))
    crit = nn.CrossEntropyLoss(weight=weight.to(dtype) if weight is not None else None, label_smoothing=smoothing)
    loss = crit(logits, y)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    return logits.detach(), loss.detach(), grads


def save(name, sd, x, y, weight, smoothing, logits, loss, grads, meta):
    out = {"x": x.numpy(), "y": y.numpy(), "logits": logits.float().numpy(), "loss": np.float32(loss.item()),
           "label_smoothing": np.float32(smoothing)}
    if weight is not None:
        out["class_weight"] = weight.numpy()
    for k, v in sd.items():
        out["sd/" + k] = v.detach().numpy()
    for k, v in grads.items():
        out["grad/" + k] = v.float().numpy()
    for k, v in meta.items():
        out["meta/" + k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"wrote {name}.npz: {len(sd)} tensors, {len(grads)} grads, loss {loss.item():.6f}")


def relerr(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def check_oracle(name, oracle_fn, sd, x, y, weight, smoothing, ref_logits, ref_loss, ref_grads):
    sd64 = {k: (v.detach().double().requires_grad_(k in ref_grads) if v.is_floating_point() else v) for k, v in sd.items()}
    logits = oracle_fn(sd64, x.double())
    loss = R.cross_entropy(logits, y, weight.double() if weight is not None else None, smoothing)
    grads = R.grads_of(loss, sd64)
    e_l = relerr(logits, ref_logits)
    e_g = max(relerr(grads[k], ref_grads[k]) for k in ref_grads)
    print(f"  oracle vs reference [{name}] fp64: logits {e_l:.2e}, loss {abs(loss.item() - ref_loss.item()):.2e}, "
          f"worst grad {e_g:.2e}")
    assert e_l < 1e-10 and e_g < 1e-8, "oracle disagrees with the reference"


def pin_timm_block():
    """oracle timm_block vs two independent pre-norm implementations, fp64."""
    torch.manual_seed(7)
    E, H = 64, 2
    sd = {}
    g = torch.Generator().manual_seed(3)
    for n, shp in [("norm1.weight", (E,)), ("norm1.bias", (E,)), ("attn.qkv.weight", (3 * E, E)), ("attn.qkv.bias", (3 * E,)),
                   ("attn.proj.weight", (E, E)), ("attn.proj.bias", (E,)), ("norm2.weight", (E,)), ("norm2.bias", (E,)),
                   ("mlp.fc1.weight", (4 * E, E)), ("mlp.fc1.bias", (4 * E,)), ("mlp.fc2.weight", (E, 4 * E)),
                   ("mlp.fc2.bias", (E,))]:
        sd["b." + n] = torch.randn(*shp, generator=g, dtype=torch.float64) * (0.2 if "weight" in n else 0.1)
    sd["b.norm1.weight"] += 1.0
    sd["b.norm2.weight"] += 1.0
    x = torch.randn(3, 19, E, generator=g, dtype=torch.float64)
    ours = R.timm_block(x, sd, "b.", H)
    lay = nn.TransformerEncoderLayer(E, H, 4 * E, dropout=0.0, activation="gelu", layer_norm_eps=1e-6, batch_first=True,
                                     norm_first=True).double()
    with torch.no_grad():
        lay.self_attn.in_proj_weight.copy_(sd["b.attn.qkv.weight"]); lay.self_attn.in_proj_bias.copy_(sd["b.attn.qkv.bias"])
        lay.self_attn.out_proj.weight.copy_(sd["b.attn.proj.weight"]); lay.self_attn.out_proj.bias.copy_(sd["b.attn.proj.bias"])
        lay.linear1.weight.copy_(sd["b.mlp.fc1.weight"]); lay.linear1.bias.copy_(sd["b.mlp.fc1.bias"])
        lay.linear2.weight.copy_(sd["b.mlp.fc2.weight"]); lay.linear2.bias.copy_(sd["b.mlp.fc2.bias"])
        lay.norm1.weight.copy_(sd["b.norm1.weight"]); lay.norm1.bias.copy_(sd["b.norm1.bias"])
        lay.norm2.weight.copy_(sd["b.norm2.weight"]); lay.norm2.bias.copy_(sd["b.norm2.bias"])
    lay.train()
    e1 = relerr(ours, lay(x))
    print(f"  timm-block restatement vs nn.TransformerEncoderLayer(norm_first): {e1:.2e}")
    assert e1 < 1e-12
    try:
        from transformers import ViTConfig
        from transformers.models.vit.modeling_vit import ViTLayer
        cfg = ViTConfig(hidden_size=E, num_attention_heads=H, intermediate_size=4 * E, hidden_act="gelu",
                        layer_norm_eps=1e-6, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        cfg._attn_implementation = "eager"
        hf = ViTLayer(cfg).double().eval()
        with torch.no_grad():
            wq, wk, wv = sd["b.attn.qkv.weight"].chunk(3, 0)
            bq, bk, bv = sd["b.attn.qkv.bias"].chunk(3, 0)
            att = hf.attention.attention
            att.query.weight.copy_(wq); att.key.weight.copy_(wk); att.value.weight.copy_(wv)
            att.query.bias.copy_(bq); att.key.bias.copy_(bk); att.value.bias.copy_(bv)
            hf.attention.output.dense.weight.copy_(sd["b.attn.proj.weight"]); hf.attention.output.dense.bias.copy_(sd["b.attn.proj.bias"])
            hf.intermediate.dense.weight.copy_(sd["b.mlp.fc1.weight"]); hf.intermediate.dense.bias.copy_(sd["b.mlp.fc1.bias"])
            hf.output.dense.weight.copy_(sd["b.mlp.fc2.weight"]); hf.output.dense.bias.copy_(sd["b.mlp.fc2.bias"])
            hf.layernorm_before.weight.copy_(sd["b.norm1.weight"]); hf.layernorm_before.bias.copy_(sd["b.norm1.bias"])
            hf.layernorm_after.weight.copy_(sd["b.norm2.weight"]); hf.layernorm_after.bias.copy_(sd["b.norm2.bias"])
        out = hf(x)
        out = out[0] if isinstance(out, (tuple, list)) else out
        e2 = relerr(ours, out)
        print(f"  timm-block restatement vs HF transformers ViTLayer: {e2:.2e}")
        assert e2 < 1e-12
    except ImportError as exc:  # pragma: no cover
        print("  HF transformers not importable, skipped:", exc)


def main():
    print("pinning the timm-block restatement")
    pin_timm_block()   # before the timm shim is installed (transformers probes for timm on import)
    mods = import_reference()
    LatentViT = mods["models_fer_vit.latent_vit"].LatentViT
    LatentViTv2 = mods["models_fer_vit.latent_vit_v2"].LatentViTv2
    hyb = mods["models_fer_vit.hybrid_latent_vit"]
    ImageViT = mods["models_fer_vit.image_vit"].ImageViT
    def perturb(model, seed):
        # move LayerNorm / bias parameters off their trivial init so every gradient path is exercised
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if p.dim() == 1 and "alpha" not in n and "gate" not in n and "layer_weights" not in n:
                    p.add_(0.1 * torch.randn(p.shape, generator=g))

    # ---------------- LatentViT ----------------
    torch.manual_seed(42)
    m = LatentViT(latent_dim=64, seq_len=18, embed_dim=64, depth=2, heads=2, mlp_dim=128, num_classes=7, dropout=0.0)
    perturb(m, 1)
    m.train()
    g = torch.Generator().manual_seed(42)
    x = torch.randn(6, 18, 64, generator=g); y = torch.randint(0, 7, (6,), generator=g)
    w = torch.rand(7, generator=g) + 0.5
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    l32 = run_reference(m, x, y, w, 0.1, torch.float32)
    l64 = run_reference(m, x, y, w, 0.1, torch.float64)
    print(f"  reference fp32 vs fp64 logits: {relerr(l32[0], l64[0]):.2e}")
    check_oracle("latent_vit", lambda s, xx: R.latent_vit_forward(s, xx, 2, 2), sd, x, y, w, 0.1, *l64)
    save("latent_vit", sd, x, y, w, 0.1, l64[0], l64[1], l64[2], {"depth": 2, "heads": 2})

    # ---------------- LatentViTv2 ----------------
    torch.manual_seed(43)
    m = LatentViTv2(latent_dim=64, seq_len=18, embed_dim=64, depth=2, heads=2, mlp_dim=128, num_classes=7, dropout=0.0,
                    use_lwn=True, use_lwn_residual=True, use_spe=True, use_leam=True)
    perturb(m, 2)
    with torch.no_grad():
        m.lwn.gate.add_(4.0 + torch.randn(18))   # move the gate away from sigmoid(-5) ~ 0 so LWN matters
        m.leam.layer_weights.add_(0.3 * torch.randn(18))
    m.train()
    x = torch.randn(6, 18, 64, generator=g) * 0.7 + 0.3; y = torch.randint(0, 7, (6,), generator=g)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    l64 = run_reference(m, x, y, None, 0.0, torch.float64)
    check_oracle("latent_vit_v2", lambda s, xx: R.latent_vit_v2_forward(s, xx, 2, 2, True, True, True, True), sd, x, y,
                 None, 0.0, *l64)
    save("latent_vit_v2", sd, x, y, None, 0.0, l64[0], l64[1], l64[2], {"depth": 2, "heads": 2})

    # ---------------- HybridLatentViT (frozen blocks + adapters), eval-mode head (Dropout(0.1) is hard-coded) ----
    torch.manual_seed(44)
    m = hyb.HybridLatentViT(latent_dim=64, seq_len=18, pretrained_model_name="vit_test_patch16_224", num_classes=7,
                            use_pretrained=False, freeze_transformer=True, adapter_dim=16)
    perturb(m, 3)
    with torch.no_grad():
        for i, a in enumerate(m.adapters):
            a.alpha.fill_(0.1 + 0.05 * i)
    m.eval()
    x = torch.randn(6, 18, 64, generator=g); y = torch.randint(0, 7, (6,), generator=g)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    l64 = run_reference(m, x, y, None, 0.0, torch.float64)
    check_oracle("hybrid", lambda s, xx: R.hybrid_forward(s, xx, 2, 2, True), sd, x, y, None, 0.0, *l64)
    save("hybrid_adapter", sd, x, y, None, 0.0, l64[0], l64[1], l64[2], {"depth": 2, "heads": 2, "adapter_dim": 16})

    # ---------------- HybridLatentViT full fine-tune, no adapter ----------------
    torch.manual_seed(45)
    m = hyb.HybridLatentViT(latent_dim=64, seq_len=18, pretrained_model_name="vit_test_patch16_224", num_classes=7,
                            use_pretrained=False, freeze_transformer=False, adapter_dim=None)
    perturb(m, 4)
    m.eval()
    x = torch.randn(5, 18, 64, generator=g); y = torch.randint(0, 7, (5,), generator=g)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    l64 = run_reference(m, x, y, w, 0.0, torch.float64)
    check_oracle("hybrid_full", lambda s, xx: R.hybrid_forward(s, xx, 2, 2, False), sd, x, y, w, 0.0, *l64)
    save("hybrid_full", sd, x, y, w, 0.0, l64[0], l64[1], l64[2], {"depth": 2, "heads": 2, "adapter_dim": 0})

    # ---------------- ImageViT ----------------
    torch.manual_seed(46)
    m = ImageViT(img_size=32, patch_size=16, in_channels=3, embed_dim=64, depth=2, heads=2, mlp_dim=128, num_classes=7,
                 dropout=0.0)
    perturb(m, 5)
    m.train()
    x = torch.randn(4, 3, 32, 32, generator=g); y = torch.randint(0, 7, (4,), generator=g)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    l64 = run_reference(m, x, y, None, 0.1, torch.float64)
    check_oracle("image_vit", lambda s, xx: R.image_vit_forward(s, xx, 2, 2, 16), sd, x, y, None, 0.1, *l64)
    save("image_vit", sd, x, y, None, 0.1, l64[0], l64[1], l64[2], {"depth": 2, "heads": 2, "patch": 16})


if __name__ == "__main__":
    main()
