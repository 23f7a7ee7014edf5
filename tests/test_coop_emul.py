"""The cooperative kernels' arithmetic without a GPU: tests/cpp/coop_emul.cpp runs the SAME source the kernels compile
(cuzk_b200/csrc/coop16.cuh and both layouts of coop.cuh) on the host, one thread per lane in lockstep, and compares every
multiply / add / MDS layer / permutation / sponge with the oracle.  A unit whose group raised no flag must be bit-exact; the
inputs include values crafted around multiples of p, all-ones words and 64-bit leaves, so a good share of the units is flagged
(those would take the exact one-thread path on the GPU) -- and the plain inputs must raise no flag at all."""
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(binary, units, seed, plain):
    path = os.path.join(HERE, "cpp", binary)
    if not os.path.exists(path):
        subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "cpp"), binary])
    res = subprocess.run([path, str(units), str(seed), str(plain)], capture_output=True, text=True, timeout=600)
    rows = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and len(rows) == 6, res.stdout[-2000:] + res.stderr[-2000:]
    return rows


@pytest.mark.parametrize("binary", ["coop_emul16k", "coop_emul16", "coop_emul8"])
def test_cooperative_source_matches_oracle_on_crafted_inputs(binary):
    rows = _run(binary, 12, 901, 0)
    for r in rows:
        assert r["unflagged_mismatches"] == 0, r
        assert r["checked"] > 0
    assert sum(r["flagged"] for r in rows) > 0   # the crafted inputs do reach the undecidable cases


@pytest.mark.parametrize("binary", ["coop_emul16k", "coop_emul8"])
def test_cooperative_source_raises_no_flag_on_plain_inputs(binary):
    rows = _run(binary, 12, 902, 1)
    for r in rows:
        assert r["unflagged_mismatches"] == 0 and r["flagged"] == 0, r
