"""Pin the oracle (oracle/vsr_oracle.py) to the golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import vsr_oracle as O
from common import load_golden, checksum


def _same(ref_list, got_list, exact_float=False):
    for r, g in zip(ref_list, got_list):
        assert r.shape == g.shape
        if r.dtype == torch.long:
            assert torch.equal(r, g)
        else:
            assert torch.allclose(r, g, rtol=0, atol=0 if exact_float else 1e-6)


@pytest.mark.parametrize("name", ["small_a.pt", "small_b.pt"])
def test_oracle_matches_reference_small(name):
    fx = load_golden(name)
    d, W = fx["dims_obj"], fx["weights"]
    det, ds = fx["det"], fx["det_seqs"]
    with torch.no_grad():
        c = fx["cases"]["bsv_gt_k3"]
        o, lp = O.beam_search(W, d, (det, ds, fx["verbs_gt"]), c["eos"], c["beam"], c["out_size"],
                              use_verbs=True, gt=True)
        _same(c["out"] + c["lp"], o + lp)
        c = fx["cases"]["bsv_tab_k5"]
        o, lp = O.beam_search(W, d, (det, ds, fx["verbs_tab"]), c["eos"], c["beam"], c["out_size"],
                              use_verbs=True, gt=False, verb_table=fx["verb_table"])
        _same(c["out"] + c["lp"], o + lp)
        c = fx["cases"]["bs_freeze_k4"]
        o, lp = O.beam_search(W, d, (det, ds), c["eos"], c["beam"], c["out_size"])
        _same(c["out"] + c["lp"], o + lp)
        c = fx["cases"]["bsv_gt_b1"]
        o, lp = O.beam_search(W, d, (det[:1], ds[:1], fx["verbs_gt"][:1]), c["eos"], c["beam"], 1,
                              use_verbs=True, gt=True)
        _same(c["out"] + c["lp"], o + lp)
        out, gate = O.forward_teacher(W, d, (det,), (fx["captions"], fx["ctrl"]))
        _same([fx["cases"]["forward"]["out"], fx["cases"]["forward"]["gate"]], [out, gate])
        w, g = O.greedy_test(W, d, (det, ds))
        _same([fx["cases"]["greedy"]["words"], fx["cases"]["greedy"]["gates"]], [w, g])


def test_oracle_step_v_small():
    fx = load_golden("small_a.pt")
    d, W, c = fx["dims_obj"], fx["weights"], fx["cases"]["step_v"]
    st = O.init_state(d, fx["det"].size(0))
    statics = (fx["det"], fx["det_seqs"], fx["verbs_gt"])
    with torch.no_grad():
        (o0, g0), st = O.decoder_step(W, d, 0, st, None, statics, None, "feedback", use_verbs=True, gt=True)
        (o1, g1), st = O.decoder_step(W, d, 1, st, [c["prev_word"], c["prev_gate"]], statics, None,
                                      "feedback", use_verbs=True, gt=True)
    _same([c["out0"], c["gate0"], c["out1"], c["gate1"], c["h1"], c["c1"], c["h2"], c["c2"], c["ptr"]],
          [o0, g0, o1, g1, st[0][0], st[0][1], st[1][0], st[1][1], st[2]])


def test_oracle_step_v_teacher_forcing_raises():
    fx = load_golden("small_a.pt")
    d, W = fx["dims_obj"], fx["weights"]
    st = O.init_state(d, fx["det"].size(0))
    with pytest.raises(NameError):
        O.decoder_step(W, d, 0, st, None, (fx["det"], fx["det_seqs"], fx["verbs_gt"]),
                       (fx["captions"], fx["ctrl"]), "teacher_forcing", use_verbs=True)
    with pytest.raises(AssertionError):
        O.decoder_step(W, d, 0, st, None, (fx["det"],), None, "sampling")


def test_oracle_full_config1():
    """BASELINE config 1 at full model size.  Weights/inputs are re-drawn from the seeds; if this
    host's RNG stream differs from the build container's the checksums catch it and we skip."""
    fx = load_golden("full_cfg1.pt")
    d = fx["dims_obj"]
    W = O.init_weights(d, seed=fx["seed_w"])
    for k, v in fx["weight_checksums"].items():
        if checksum(W[k]) != v:
            pytest.skip(f"seeded weight stream differs on this host ({k})")
    det, ds, verbs = O.synth_inputs(**fx["synth"])
    for k, t in (("det", det), ("det_seqs", ds), ("verbs", verbs)):
        if checksum(t) != fx["input_checksums"][k]:
            pytest.skip("seeded input stream differs on this host")
    with torch.no_grad():
        c = fx["cases"]["bsv_gt_k3"]
        o, lp = O.beam_search(W, d, (det, ds, verbs), c["eos"], c["beam"], 1, use_verbs=True, gt=True)
        _same(c["out"] + c["lp"], o + lp)
        W["out_fc.weight"] = W["out_fc.weight"] * 100.0
        c = fx["cases"]["bsv_gt_k3_sharp100"]
        o, lp = O.beam_search(W, d, (det, ds, verbs), c["eos"], c["beam"], 1, use_verbs=True, gt=True)
        _same(c["out"] + c["lp"], o + lp)
