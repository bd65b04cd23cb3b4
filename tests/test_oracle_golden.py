"""CPU: the oracle reproduces the golden vectors the UNMODIFIED reference classes produced
(tests/golden/make_golden.py), in fp64 to round-off and in fp32 to fp32 round-off."""
import pytest
import torch

from oracle import reference_math as R
from tests.util import FIXTURES, load_golden, oracle_forward, relerr


@pytest.mark.parametrize("name", FIXTURES)
@pytest.mark.parametrize("dtype,tol_l,tol_g", [(torch.float64, 1e-6, 1e-5), (torch.float32, 2e-5, 2e-4)])
def test_oracle_matches_reference_golden(name, dtype, tol_l, tol_g):
    g = load_golden(name)
    sd = {k: (v.to(dtype).requires_grad_(k in g["grad"]) if v.is_floating_point() else v) for k, v in g["sd"].items()}
    logits = oracle_forward(name, sd, g["x"].to(dtype))
    w = g["class_weight"].to(dtype) if g["class_weight"] is not None else None
    loss = R.cross_entropy(logits, g["y"], w, g["label_smoothing"])
    grads = R.grads_of(loss, sd)
    # golden logits are stored in fp32, so 1e-6 is the floor even for the fp64 run
    assert relerr(logits, g["logits"]) < tol_l
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    assert set(grads) == set(g["grad"])
    for k in g["grad"]:
        assert relerr(grads[k], g["grad"][k]) < tol_g, k
    assert torch.equal(logits.argmax(-1), g["logits"].argmax(-1))


def test_cross_entropy_matches_torch():
    torch.manual_seed(0)
    z = torch.randn(33, 7, dtype=torch.float64, requires_grad=True)
    y = torch.randint(0, 7, (33,))
    w = torch.rand(7, dtype=torch.float64) + 0.5
    for weight in (None, w):
        for eps in (0.0, 0.1):
            ours = R.cross_entropy(z, y, weight, eps)
            ref = torch.nn.functional.cross_entropy(z, y, weight=weight, label_smoothing=eps)
            assert abs(ours.item() - ref.item()) < 1e-12
            g1, = torch.autograd.grad(ours, z)
            g2, = torch.autograd.grad(ref, z)
            assert (g1 - g2).abs().max() < 1e-12


def test_timm_block_restatement_matches_torch_prenorm_layer():
    """The timm Block restatement vs nn.TransformerEncoderLayer(norm_first=True, gelu, eps 1e-6): same block."""
    torch.manual_seed(1)
    E, H = 64, 2
    lay = torch.nn.TransformerEncoderLayer(E, H, 4 * E, dropout=0.0, activation="gelu", layer_norm_eps=1e-6,
                                           batch_first=True, norm_first=True).double().train()
    with torch.no_grad():
        for p in lay.parameters():
            p.add_(0.1 * torch.randn_like(p))
    sd = {"b.norm1.weight": lay.norm1.weight, "b.norm1.bias": lay.norm1.bias,
          "b.attn.qkv.weight": lay.self_attn.in_proj_weight, "b.attn.qkv.bias": lay.self_attn.in_proj_bias,
          "b.attn.proj.weight": lay.self_attn.out_proj.weight, "b.attn.proj.bias": lay.self_attn.out_proj.bias,
          "b.norm2.weight": lay.norm2.weight, "b.norm2.bias": lay.norm2.bias,
          "b.mlp.fc1.weight": lay.linear1.weight, "b.mlp.fc1.bias": lay.linear1.bias,
          "b.mlp.fc2.weight": lay.linear2.weight, "b.mlp.fc2.bias": lay.linear2.bias}
    x = torch.randn(3, 19, E, dtype=torch.float64)
    assert relerr(R.timm_block(x, sd, "b.", H), lay(x)) < 1e-12


def test_post_norm_layer_restatement_matches_torch_layer():
    torch.manual_seed(2)
    E, H = 64, 2
    for act_name, act in (("relu", R.relu), ("gelu", R.gelu)):
        lay = torch.nn.TransformerEncoderLayer(E, H, 128, dropout=0.0, activation=act_name, batch_first=True).double().train()
        sd = {"l." + k: v for k, v in lay.state_dict().items()}
        x = torch.randn(3, 19, E, dtype=torch.float64)
        assert relerr(R.torch_encoder_layer(x, sd, "l.", H, act), lay(x)) < 1e-12


def test_relu_selector_injection_is_the_reference_relu():
    """The oracle's injected ReLU active set (parity tests hand it the selector the kernels used): with the oracle's OWN
    selector it is the plain ReLU path bit for bit, value and gradients; with one unit flipped it differs."""
    from oracle import reference_math as R
    g = load_golden("latent_vit")
    sd = {k: v.double().requires_grad_(True) for k, v in g["sd"].items()}
    x, y = g["x"].double(), g["y"]
    masks = {"trace": {}}
    ref = R.latent_vit_forward(sd, x, 2, 2, masks)
    gref = R.grads_of(R.cross_entropy(ref, y, g["class_weight"], g["label_smoothing"]), sd)
    sel = {("relu", i): (masks["trace"][("ffn_pre", i)] > 0).double() for i in range(2)}
    out = R.latent_vit_forward(sd, x, 2, 2, sel)
    gout = R.grads_of(R.cross_entropy(out, y, g["class_weight"], g["label_smoothing"]), sd)
    assert torch.equal(out, ref)
    assert all(torch.equal(gout[k], gref[k]) for k in gref)
    flipped = {k: v.clone() for k, v in sel.items()}
    flipped[("relu", 0)][0, 0, 0] = 1.0 - flipped[("relu", 0)][0, 0, 0]
    assert not torch.equal(R.latent_vit_forward(sd, x, 2, 2, flipped), ref)
