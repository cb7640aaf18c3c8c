"""Static check of the built library's SASS (cuobjdump, no GPU): the contractions of the hot path are tcgen05 / TMEM / TMA code
(north_star: "dense contractions run on tcgen05/TMEM tiles fed by TMA"), not mma.sync, and the HBM-bound prep kernel of the loss
carries no per-element fp32 divisions (they made it instruction-bound: DESIGN 4.3)."""
import collections
import re
import shutil
import subprocess

import pytest

import dclip_b200

TENSOR_KERNELS = ["k_stream_tc", "k_gemm_pair", "k_gemm_tc", "k_infonce_bwd_pair", "k_infonce_bwd_tc", "k_attn_fwd_tc", "k_attn_bwd_tc"]


@pytest.fixture(scope="module")
def sass_counts(built_library):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "-sass", dclip_b200.lib_path()], capture_output=True, text=True, timeout=600, check=True).stdout
    except (OSError, subprocess.SubprocessError) as e:
        pytest.skip(f"cuobjdump not usable here: {e}")
    counts = collections.defaultdict(collections.Counter)
    fn = None
    op = re.compile(r"^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)")
    for line in out.splitlines():
        if "Function :" in line:
            fn = line.split("Function :", 1)[1].strip()
            continue
        m = op.match(line)
        if m and fn:
            counts[fn][m.group(1)] += 1
    assert counts, "no SASS found in the library (built for another architecture?)"
    return counts


def _kernels(counts, name):
    return {fn: c for fn, c in counts.items() if name in fn}


@pytest.mark.parametrize("name", TENSOR_KERNELS)
def test_contractions_are_tcgen05_tmem_tma(sass_counts, name):
    ks = _kernels(sass_counts, name)
    assert ks, f"{name} is not in the library"
    for fn, c in ks.items():
        assert c["UTCHMMA"] > 0, f"{fn}: no tcgen05.mma (UTCHMMA)"
        assert c["UTMALDG"] > 0, f"{fn}: operands are not loaded by TMA (UTMALDG)"
        assert c["LDTM"] > 0, f"{fn}: accumulators are not read back from TMEM (LDTM)"
        assert c["HMMA"] == 0, f"{fn}: mma.sync (HMMA) in a tcgen05 kernel"


def test_no_mma_sync_anywhere(sass_counts):
    offenders = [fn for fn, c in sass_counts.items() if c["HMMA"] or c["IMMA"] or c["DMMA"]]
    assert not offenders, offenders


def test_16bit_operand_prep_has_no_fp32_division(sass_counts):
    ks = {fn: c for fn, c in _kernels(sass_counts, "k_loss_prep_pair_nc").items() if "6__half" in fn.split("k_loss_prep_pair_nc", 1)[1][:40]}
    assert ks, "k_loss_prep_pair_nc<*, __half, *> is not in the library"
    for fn, c in ks.items():
        assert c["FCHK"] <= 6, f"{fn}: {c['FCHK']} fp32 divisions (per-row scalars only are expected)"
    # the FP32 exact tier keeps F.normalize's division
    exact = {fn: c for fn, c in _kernels(sass_counts, "k_loss_prep").items() if fn.rstrip("E").endswith("fEEvNS_8PrepArgs")}
    assert exact and all(c["FCHK"] > 0 for c in exact.values())
