"""Static checks on the compiled sm_100a code (no GPU needed): the hot kernels really are tcgen05 / TMA / TMEM code, and
the role loops of the tensor-core conv contain no function call (a `printf` in the bounded mbarrier wait once forced all
MMA-issue state into vector registers: 14 R2UR per 4 MMAs — DESIGN.md §5.1)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def _sass(obj):
    from maua_stylegan2_b200 import build

    build.build()
    path = os.path.join(ROOT, "maua_stylegan2_b200", "build", obj)
    out = subprocess.run([CUOBJDUMP, "-sass", path], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out or "SM100" in out.upper() or "EF_CUDA_SM100" in out, "not an sm_100a cubin"
    return out


def _functions(sass):
    """{mangled name: body text}"""
    parts = re.split(r"\n\s*Function : ", sass)
    return {p.split("\n", 1)[0].strip(): p for p in parts[1:]}


@pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not installed")
def test_conv_kernels_are_tcgen05_tma_tmem_code():
    for obj, kernel in (("modconv_tc2.o", "modconv_tc2_kernel"), ("modconv_tc.o", "modconv_tc_kernel")):
        fns = {k: v for k, v in _functions(_sass(obj)).items() if kernel in k}
        assert fns, obj
        for name, body in fns.items():
            for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS"):
                assert mnemonic in body, (name, mnemonic)
            # CALL.REL = compiler-local subroutines (integer division) are fine; CALL.ABS = an external call (vprintf)
            assert "CALL.ABS" not in body, f"{name}: an external call (printf?) breaks uniform-register MMA issue"
    v2 = {k: v for k, v in _functions(_sass("modconv_tc2.o")).items() if "modconv_tc2_kernel" in k}
    # {same-res, transposed} x {bf16: 1 product, 3 products, concat; fp16: two N=BN MMAs, one N=2*BN concat MMA}
    # + the CTA-pair (cta_group::2) form of the 3-product kernel
    assert len(v2) == 12
    pairs = {k: v for k, v in v2.items() if "UTCHMMA.2CTA" in v}
    assert len(pairs) == 2, sorted(pairs)
    for name, body in pairs.items():
        for mnemonic in ("UTMALDG.4D.2CTA", "UTMALDG.3D.2CTA", "UTCBAR.2CTA.MULTICAST", "UTCATOMSWS.2CTA", "UCGABAR_ARV"):
            assert mnemonic in body, (name, mnemonic)
    for name, body in v2.items():
        assert "FFMA2" in body or "FMUL2" in body, name   # packed fp32 epilogue
        # the unrolled R = 4 tap issues its MMAs back to back: at least one run of >= 8 UTCHMMA within 24 instructions
        lines = [l for l in body.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        idx = [i for i, l in enumerate(lines) if "UTCHMMA" in l]
        assert any(idx[i + 7] - idx[i] <= 24 for i in range(len(idx) - 7)), name


@pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not installed")
def test_memory_bound_kernels_use_tma_and_vector_accesses():
    blur = _functions(_sass("blur_act_nhwc.o"))
    tma = [v for k, v in blur.items() if "blur_act_nhwc_tma_kernel" in k]
    assert tma
    for body in tma:
        assert "UTMALDG" in body and "LDS.128" in body and "FFMA2" in body
        assert "LDL" not in body and "LD.E.128" not in body     # tile reads are shared-memory loads, nothing spilled
    ufd = _sass("upfirdn2d.o")
    assert "STG.E.128" in ufd and "LDS.128" in ufd               # float4 HBM writes (north_star), taps/tiles staged in smem
