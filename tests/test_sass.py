"""SASS of the built library (CPU, needs cuobjdump from the CUDA toolkit): the tensor-core kernels are Blackwell-native
(tcgen05 / TMA mnemonics, no legacy mma.sync) and their TMA / MMA issue loops are warp-uniform — no R2UR.BROADCAST
waterfall loop around UTCHMMA / UTMALDG ("second half": ~90 cycles per MMA, ~170 per TMA operation)."""
import collections
import re
import shutil
import subprocess

import pytest

FAMILIES = ("tc_gemm_kernel", "edge_pair_kernel", "node_chain_t_kernel", "node_chain_kernel")


def _sass_counts(lib_path):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "-sass", lib_path], capture_output=True, text=True, timeout=600)
    except (OSError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump not available")
    if out.returncode != 0:
        pytest.skip("cuobjdump failed: %s" % out.stderr[:200])
    per = {f: collections.Counter() for f in FAMILIES}
    cur = None
    for line in out.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next((f for f in FAMILIES if f in m.group(1)), None)
            continue
        if cur:
            mm = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if mm:
                per[cur][mm.group(1).split(".")[0] if not mm.group(1).startswith("R2UR") else mm.group(1)] += 1
    return per


def test_tensor_core_kernels_are_native_and_issue_uniformly():
    from matinvent_b200.csrc.build import build
    per = _sass_counts(build())
    for fam, c in per.items():
        assert c["UTCHMMA"] > 0, (fam, "no tcgen05.mma")
        assert c["UTMALDG"] > 0, (fam, "no TMA loads")
        assert c["LDTM"] > 0, (fam, "no tcgen05.ld")
        assert c["HMMA"] == 0, (fam, "legacy mma.sync")
        assert c["R2UR.BROADCAST"] == 0, (fam, c["R2UR.BROADCAST"], "waterfall loop around a uniform-operand instruction")
    assert per["edge_pair_kernel"]["UTMASTG"] > 0          # the first per-edge block leaves through TMA stores
    assert per["node_chain_t_kernel"]["STAS"] > 0 and per["node_chain_kernel"]["STAS"] > 0     # st.async statistics exchange
