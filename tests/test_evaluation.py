"""Row f4 of SURVEY.md §8f: evaluate_for_arti_axis / evaluate_for_recognition / compute_ap against outputs
of the reference itself (tests/golden/eval/*.json, oracle/gen_golden.py --eval)."""
import json
import os

import pytest
import torch

from articulation3d_b200 import evaluation as ev
from tests import golden_util as gu

EVAL = os.path.join(gu.ROOT, "tests", "golden", "eval")


def _load(name):
    with open(os.path.join(EVAL, f"{name}.json")) as f:
        d = json.load(f)
    records = [{"image_id": r["image_id"], "instances": r["instances"],
                "pred_plane": torch.tensor(r["pred_plane"], dtype=torch.float32).reshape(-1, 3),
                "pred_rot_axis": torch.tensor(r["pred_rot_axis"], dtype=torch.float32).reshape(-1, 3),
                "pred_tran_axis": torch.tensor(r["pred_tran_axis"], dtype=torch.float32).reshape(-1, 2)}
               for r in d["records"]]
    return d, records


@pytest.mark.parametrize("name", ["case_a", "case_b", "case_c"])
def test_arti_axis_ap_matches_reference(name):
    d, records = _load(name)
    got = ev.evaluate_for_arti_axis(records, ev.CocoGT(d["gt"]), ev.Metadata(), d["filter_iou"])
    assert set(got) == set(d["arti_axis"])
    for k, want in d["arti_axis"].items():
        assert float(got[k]) == pytest.approx(want, rel=1e-5, abs=1e-7), k
    # every criterion is a restriction of 'bbox'
    for cat in ("arti_rot", "arti_tran"):
        if f"bbox - {cat}" in got:
            assert float(got[f"bbox+normal+axis - {cat}"]) <= float(got[f"bbox+axis - {cat}"]) + 1e-7
            assert float(got[f"bbox+axis - {cat}"]) <= float(got[f"bbox - {cat}"]) + 1e-7


@pytest.mark.parametrize("name", ["case_a", "case_b", "case_c"])
def test_recognition_matches_reference(name):
    d, records = _load(name)
    got = ev.evaluate_for_recognition(records, ev.CocoGT(d["gt"]))
    for k, want in d["recognition"].items():
        assert float(got[k]) == pytest.approx(want, rel=1e-9), k


def test_compute_ap_known_answers():
    s = torch.tensor([0.9, 0.8, 0.7, 0.6])
    assert ev.compute_ap(torch.zeros(0), torch.zeros(0, dtype=torch.uint8), 3.0) == 0.0
    assert float(ev.compute_ap(s, torch.tensor([1, 1, 1, 1], dtype=torch.uint8), 4.0)) == pytest.approx(1.0)
    assert float(ev.compute_ap(s, torch.tensor([0, 0, 0, 0], dtype=torch.uint8), 4.0)) == 0.0
    # ranks: tp fp tp fp, npos 2 -> recall .5 .5 1 1, precision 1 .5 .667 .5 -> AP = .5*1 + .5*.6667
    assert float(ev.compute_ap(s, torch.tensor([1, 0, 1, 0], dtype=torch.uint8), 2.0)) == pytest.approx(0.5 + 0.5 * 2 / 3)
    # order of the input does not matter, only the scores
    p = torch.tensor([2, 0, 3, 1])
    assert float(ev.compute_ap(s[p], torch.tensor([1, 0, 1, 0], dtype=torch.uint8)[p], 2.0)) == pytest.approx(0.5 + 1 / 3)


def test_records_without_ground_truth_or_instances_are_skipped():
    d, records = _load("case_a")
    gt = ev.CocoGT(d["gt"])
    base = ev.evaluate_for_arti_axis(records, gt, ev.Metadata(), 0.0)
    extra = records + [{"image_id": 10_000, "instances": [], "pred_plane": torch.zeros(0, 3),
                        "pred_rot_axis": torch.zeros(0, 3), "pred_tran_axis": torch.zeros(0, 2)}]
    again = ev.evaluate_for_arti_axis(extra, gt, ev.Metadata(), 0.0)
    assert {k: float(v) for k, v in base.items()} == {k: float(v) for k, v in again.items()}
