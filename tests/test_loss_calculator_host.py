"""Host logic of the LossCalculator mirror against facts recorded from the reference (tests/golden/host_facts.npz):
constructor rules, errors, get_control_output (including the F8 typo), dict/attribute surface.  CPU only."""
import contextlib
import io

import numpy as np
import pytest

from conftest import golden
from distillclip_b200.model import LOSSNAME, IMAGE_TEXT_LOSS, LossCalculator, ControlOutput


def make(*a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return LossCalculator(*a, **k)


def flags(co):
    return [co.need_emb, co.need_attn_score, co.need_value_map, co.need_attn_prob, co.need_rep]


def test_control_output_matches_reference():
    f = golden("host_facts")
    c = make(["embedding_mse", "hidden_rep_mse", "attention_probs_kl", "attention_probs_mse"])
    assert flags(c.get_control_output()) == [bool(x) for x in f["control_all4"]]
    co = make(["attention_probs_kl"]).get_control_output()
    assert isinstance(co, ControlOutput)
    assert flags(co) == [bool(x) for x in f["control_kl_only"]]           # F8: need_attn_prob stays False
    assert hasattr(co, "attention_probs_mse") == bool(f["control_kl_only_has_typo_attr"])


def test_percent_rules_and_errors_match_reference():
    f = golden("host_facts")
    c = make(["hard_label", "soft_label"], temperature=1.0, percent={"hard_label": 0.7})
    assert list(c.percent.keys()) == [str(k) for k in f["percent_fill_keys"]]
    assert np.allclose(list(c.percent.values()), f["percent_fill_vals"], rtol=0, atol=1e-15)
    with pytest.raises(ValueError, match=str(f["invalid_name_error"])):
        make(["nope"])
    assert str(f["neg_percent_error"]) == "ValueError"
    with pytest.raises(ValueError):
        make(["hard_label", "hidden_rep_mse"], percent={"hard_label": 1.0})
    assert str(f["fill_3names_1given"]) == "AssertionError"
    with pytest.raises(AssertionError):
        make(["hard_label", "soft_label", "hidden_rep_mse"], temperature=1.0, percent={"hard_label": 0.5})


def test_surface():
    c = make(["hard_label", "soft_label", "hidden_rep_mse", "out_l1", "out_cos", "cos_diff"], temperature=2.0,
             loss_scale={"soft_label": 0.5})
    assert list(c.loss.keys()) == ["hard_label", "soft_label", "hidden_rep_mse", "out_l1", "out_cos", "cos_diff"]
    assert c.loss_scale["soft_label"] == 0.5 and c.loss_scale["hard_label"] == 1
    assert abs(sum(c.percent.values()) - 1) < 1e-9
    assert len(list(c.parameters())) == 0 and len(list(c.buffers())) == 0          # checkpoints keep round-tripping
    c.set_percent({"hard_label": 1.0})
    c.set_scale({"hard_label": 2.0})
    assert c.percent == {"hard_label": 1.0} and c.loss_scale == {"hard_label": 2.0}
    assert len(LOSSNAME) == 15 and "smdhard_label" in LOSSNAME                       # F9: the reference's missing comma
    assert IMAGE_TEXT_LOSS == ['hard_label', 'soft_label', 'logits_mse', 'fine_grain', 'cos_diff']


def test_out_of_scope_names_are_recognised():
    for n in ("vit_kd", "fine_grain", "smd"):
        with pytest.raises(NotImplementedError):
            make([n], temperature=1.0)
