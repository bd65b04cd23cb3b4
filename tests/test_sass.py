"""CPU: the shipped library's hot kernels really are tcgen05 / TMA / TMEM code (SASS mnemonics of
/opt/skills/guides/B200_PROFILING.md), not a recompiled legacy path. Needs cuobjdump, no GPU."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fer_vit_b200", "libfervit_b200.so")


@pytest.fixture(scope="module")
def sass_by_kernel():
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    text = subprocess.run([cuobjdump, "-sass", LIB], capture_output=True, text=True, check=True, timeout=900).stdout
    out, cur = {}, None
    for line in text.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur is not None:
            out[cur].append(line)
    return {k: "\n".join(v) for k, v in out.items()}


def _kernels(sass, needle):
    ks = {k: v for k, v in sass.items() if needle in k}
    assert ks, needle
    return ks


def test_cta_pair_gemms_are_tcgen05_with_tma_and_tmem(sass_by_kernel):
    for needle in ("gemm_tc2_kernel", "wgrad2_kernel"):
        for name, body in _kernels(sass_by_kernel, needle).items():
            assert "UTCHMMA.2CTA" in body, name             # tcgen05.mma.cta_group::2
            assert "UTMALDG.2D.2CTA" in body, name          # TMA loads counted on the leader's mbarrier
            assert "LDTM" in body, name                      # tcgen05.ld: accumulators come from TMEM
            assert "UTCBAR.2CTA.MULTICAST" in body, name    # tcgen05.commit to both CTAs
            assert not re.search(r"[^C]HMMA", body), name    # no legacy mma.sync in the GEMMs
    # the forward / dgrad kernel stores through TMA as well
    assert any("UTMASTG" in b for b in _kernels(sass_by_kernel, "gemm_tc2_kernel").values())


def test_single_cta_gemm_and_adapter_kernels_are_tcgen05(sass_by_kernel):
    for needle in ("gemm_tc_kernel", "adapter_kernel"):
        for name, body in _kernels(sass_by_kernel, needle).items():
            assert "UTCHMMA" in body and "UTMALDG" in body and "LDTM" in body, name


def test_gelu_epilogues_use_packed_fp32(sass_by_kernel):
    # KIND 1 = GELU forward (epilogue.cuh: EPK_GELU): the polynomial runs on FFMA2 (fma.rn.f32x2)
    gelu = {k: v for k, v in _kernels(sass_by_kernel, "gemm_tc2_kernel").items() if re.search(r"ILi256ELi1E", k)}
    assert gelu
    for name, body in gelu.items():
        assert "FFMA2" in body, name


def test_attention_kernels_are_warp_level_tensor_core_code(sass_by_kernel):
    for needle in ("attn_tc_fwd2_kernel", "attn_tc_bwd2_kernel", "attn_long_fwd_kernel", "attn_long_bwd_kernel"):
        for name, body in _kernels(sass_by_kernel, needle).items():
            assert "HMMA.16816.F32.BF16" in body and "LDSM" in body and "LDGSTS" in body, name
