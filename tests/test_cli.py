"""CLI / text-report compatibility: CSV loading incl. CR-only line endings, the
reference's report files, pickling of each layer.  The report writers are exercised on CPU with a stand-in model built
from a golden fixture; the end-to-end command runs on the GPU."""
import os
import pickle
import types

import numpy as np
import pytest

from conftest import load_golden, golden_moments


def _write_csv(path, x, cr_only=True, row_names=False):
    nl = "\r" if cr_only else "\n"
    cols = ["q%d" % i for i in range(x.shape[1])]
    lines = [",".join((["name"] if row_names else []) + cols)]
    for r, row in enumerate(x):
        lines.append(",".join((["s%d" % r] if row_names else []) + [repr(float(v)) for v in row]))
    with open(path, "w", newline="") as fh:
        fh.write(nl.join(lines) + nl)


def test_load_csv_cr_only_and_names(tmp_path):
    from linearcorex_b200 import cli
    z, _, x = load_golden("big5_l0_f64")
    p = str(tmp_path / "big5.csv")
    _write_csv(p, x[:50], cr_only=True)
    got, samples, variables = cli.load_csv(p, no_row_names=True)
    np.testing.assert_array_equal(got, x[:50])
    assert samples is None and variables[:2] == ["q0", "q1"] and len(variables) == 50
    _write_csv(p, x[:7], cr_only=False, row_names=True)
    got, samples, variables = cli.load_csv(p)
    np.testing.assert_array_equal(got, x[:7])
    assert samples == ["s%d" % i for i in range(7)] and len(variables) == 50


def test_report_writers_on_golden_model(tmp_path):
    from linearcorex_b200 import cli
    z, _, x = load_golden("big5_l0_f64")
    mdl = types.SimpleNamespace(ws=z["ws"], moments=golden_moments(z), mis=z["mis"], tc=float(z["m_TC"]), tcs=z["m_TCs"])
    labels = ["v%d" % i for i in range(50)]
    out = str(tmp_path / "out")
    cli.write_groups(mdl, labels, out)
    cli.write_labels(z["transform"][:5], ["a", "b", "c", "d", "e"], out)
    cli.write_hierarchy_summary([mdl], out)
    groups = open(out + "/summary/groups_no_overlaps.txt").read().splitlines()
    assert groups[0] == "variable, weight, MI" and groups[1].startswith("Group num: 0, TC(X;Y_j): %0.6f" % z["m_TCs"][0])
    # every variable appears exactly once in the no-overlap file, under the factor with the largest |W|
    members = [l.split(",")[0] for l in groups[1:] if l.startswith("v")]
    assert sorted(members) == sorted(labels)
    summary = open(out + "/summary/summary.txt").read()
    assert "Total: %f" % z["m_TCs"].sum() in summary
    assert len(open(out + "/summary/labels.txt").read().splitlines()) == 5
    assert open(out + "/summary/higher_layer_group_tcs.txt").read().startswith("At layer: 0, Total TC: %0.3f" % float(z["m_TC"]))


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path):
    from linearcorex_b200 import cli
    z0, _, x = load_golden("big5_l0_f64")
    z1, _, _ = load_golden("big5_l1_f64")
    p = str(tmp_path / "big5.csv")
    _write_csv(p, x, cr_only=True)
    out = str(tmp_path / "run")
    np.random.seed(0)  # the CLI never seeds (vis_corex.py:535-545): the global RNG state decides W0
    assert cli.main([p, "--layers=5,1", "--no_row_names", "-o", out]) == 0
    layers = [pickle.load(open(out + "/layer_%d.dat" % l, "rb")) for l in range(2)]
    assert layers[0].ws.shape == (5, 50) and layers[1].ws.shape == (1, 5)
    c = layers[0].clusters()
    assert len(set(c)) == 5 and all(c[i] == c[i % 5] for i in range(50))  # known answer: trait = i mod 5
    assert abs(layers[0].tc - float(z0["m_TC"])) < 1e-3 * float(z0["m_TC"])  # unseeded start, same optimum
    for f in ("groups.txt", "groups_no_overlaps.txt", "summary.txt", "labels.txt", "higher_layer_group_tcs.txt"):
        assert os.path.getsize(out + "/summary/" + f) > 0
    assert len(open(out + "/summary/labels.txt").read().splitlines()) == 2000
    # --regraph re-reads the pickles (binary mode; the reference's text-mode open fails on python 3, vis_corex.py:551)
    assert cli.main([p, "--layers=5,1", "--no_row_names", "-o", out, "-q"]) == 0
