"""What the shipped binary is made of, checked without a GPU: `cuobjdump -sass` of the cubins inside libnoiseflow_b200.so.

The mnemonics are the ones /opt/skills/guides/B200_PROFILING.md names as proof of the Blackwell paths: UTCHMMA = tcgen05.mma
(fp16 / bf16 kind), UTMALDG = TMA tensor load (cp.async.bulk.tensor), LDTM / STTM = tcgen05.ld / st (tensor memory), FFMA2 =
packed fma.rn.f32x2; LDL / STL = local-memory (spill) traffic, which the dominant kernels must not have."""
import os
import re
import subprocess
import tempfile

import pytest

from noise_flow_b200 import _lib


def _functions(cubin_name):
    """{demangled-ish function name: [opcodes]} of one cubin embedded in the library."""
    with tempfile.TemporaryDirectory() as td:
        r = subprocess.run(["cuobjdump", "-xelf", cubin_name, os.path.abspath(_lib.LIB_PATH)], cwd=td, capture_output=True, text=True)
        files = [f for f in os.listdir(td) if f.endswith(".cubin")]
        assert files, "cubin %s not found in the library: %s" % (cubin_name, r.stderr)
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(td, files[0])], capture_output=True, text=True).stdout
    out = {}
    for chunk in sass.split("Function : ")[1:]:
        name = chunk.split("\n", 1)[0].strip()
        ops = [re.sub(r"^@!?U?P\d+\s+", "", m.group(1)).split()[0].rstrip(";")
               for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", chunk)]
        out[name] = ops
    return out


def _count(ops, prefix):
    return sum(1 for o in ops if o.startswith(prefix))


@pytest.fixture(scope="module")
def lib_built():
    _lib.load()
    return True


def test_default_chain_kernel_is_tmem_resident_packed_fp32_and_spill_free(lib_built):
    fns = _functions("nf_wino.sm_100a.cubin")
    kernels = {n: o for n, o in fns.items() if "nf_chain_wino_kernel" in n}
    assert len(kernels) == 4                                   # <INV> x <TMEM_COLS 512 / 256>
    for name, ops in kernels.items():
        assert _count(ops, "LDTM") > 0 and _count(ops, "STTM") > 0, name      # patches live in tensor memory
        assert _count(ops, "FFMA2") > 1000, name                              # packed fp32 FMAs carry the convolutions
        assert _count(ops, "LDL") == 0 and _count(ops, "STL") == 0, name      # no local-memory spills
        assert _count(ops, "UTCHMMA") == 0, name                              # all-fp32: no tensor-core MMA in this kernel


def test_hybrid_and_wide_kernels_use_tcgen05_mma_and_tma(lib_built):
    hyb = _functions("nf_hybrid.sm_100a.cubin")
    for name, ops in hyb.items():
        if "nf_chain_hyb_kernel" in name:
            assert _count(ops, "UTCHMMA") > 0 and _count(ops, "LDTM") > 0, name
    assert any("nf_chain_hyb_kernel" in n for n in hyb)
    for cubin, kern in (("nf_wide_tc.sm_100a.cubin", "nf_wide_tc_kernel"), ("nf_wide_tcs.sm_100a.cubin", "nf_wide_tcs_kernel")):
        fns = {n: o for n, o in _functions(cubin).items() if kern in n}
        assert fns, cubin
        for name, ops in fns.items():
            assert _count(ops, "UTCHMMA") > 0, name            # tcgen05.mma
            assert _count(ops, "UTMALDG") > 0, name            # weights arrive by TMA tensor copies
            assert _count(ops, "LDTM") > 0, name               # accumulators read back from tensor memory
