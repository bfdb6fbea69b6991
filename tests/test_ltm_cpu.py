"""CPU: the long-term memory's bookkeeping (diversity filter, burden / diversity metrics, baseline) against the CPU
restatement of memory/ltm.py on the same key streams (oracle/ltm_oracle.py).  The keys themselves come from the device
kernel in the GPU test; here integer keys stand in for the formulas."""
import numpy as np
import pytest
import torch

from matinvent_b200.memory.ltm import LongTimeMem
from oracle.ltm_oracle import LongTimeMemOracle


def _stream(rng, n, n_comp, n_ele):
    comp = rng.integers(0, n_comp, n)
    ele = comp % n_ele
    return comp, ele, rng.random(n)


@pytest.mark.parametrize("method", ["composition", "element_comb"])
def test_div_filter_and_metrics_match_reference_bookkeeping(method):
    rng = np.random.default_rng(5)
    mine, ref = LongTimeMem(device="cpu"), LongTimeMemOracle()
    for step in range(12):
        comp, ele, rew = _stream(rng, 160, 40, 9)
        mine.extend_keys(torch.as_tensor(comp), torch.as_tensor(ele), rew, step)
        ref.extend(["c%d" % c for c in comp], [("e", int(e)) for e in ele], rew, step)
        qc, qe, qr = _stream(rng, 64, 60, 12)                   # some keys never seen (occurrence 0)
        got = mine.div_filter_keys(torch.as_tensor(qc if method == "composition" else qe), qr, tol=10, buff=20, method=method)
        want = ref.div_filter(["c%d" % c for c in qc] if method == "composition" else [("e", int(e)) for e in qe], qr,
                              tol=10, buff=20, method=method)
        assert np.array_equal(got[0], want[0]) and got[1] == want[1] and got[2:] == want[2:]
        assert len(mine) == len(ref.memory) and mine.unique_comps.numel() == len(ref.unique_comps)
        for thred, cand in ((0.5, 10), (0.9, 100), (0.99, 30)):
            b1, d1 = mine.calc_metrics(thred, budget=1500, num_candidate=cand)
            b2, d2 = ref.calc_metrics(thred, budget=1500, num_candidate=cand)
            assert (b1 is None) == (b2 is None) and (b1 is None or abs(b1 - b2) < 1e-12)
            assert (d1 is None) == (d2 is None) and (d1 is None or abs(d1 - d2) < 1e-12)
        assert abs(mine.get_baseline(step) - ref.get_baseline(step)) < 1e-12
    # best crystal of every composition (ltm.py:139-149)
    assert sorted(mine.deduplicate_indices().tolist()) == ref.best_row_per_comp()


def test_empty_memory_and_empty_query():
    m = LongTimeMem(device="cpu")
    assert len(m) == 0 and m.calc_metrics(0.5) == (None, None)
    r, pen, a, b = m.div_filter_keys(torch.as_tensor([3, 3, 4]), [0.1, 0.2, 0.3])
    assert np.allclose(r, [0.1, 0.2, 0.3]) and pen == [] and (a, b) == (0, 0)
    m.extend_keys(torch.as_tensor([1]), torch.as_tensor([1]), [0.5], 0)
    r, pen, a, b = m.div_filter_keys(torch.zeros(0, dtype=torch.int64), [])
    assert r.shape == (0,) and pen == []
