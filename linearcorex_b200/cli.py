"""Command line front end: the non-graphical part of the reference CLI (vis_corex.py:414-551, :55-93, :215-225).

    python -m linearcorex_b200.cli data.csv --layers=5,1 --no_row_names -o big5
    python -m linearcorex_b200.cli adni_blood.csv --layers=30,5,1 --missing=-1e6 -o adni

Same flags as the reference (`-t -f -m -d -g -l -w -a -o -v -e -q`; `-n/--gpu` is accepted and ignored -- the GPU is
the only path).  It loads the CSV (first row = variable names, first column = sample names unless told otherwise; CR-only
line endings like tests/data/test_big5.csv are handled), fits the layer stack on the device (hierarchy.fit_layers),
pickles each layer to `<out>/layer_<l>.dat` and writes the reference's text reports:
`summary/groups.txt`, `summary/groups_no_overlaps.txt`, `summary/summary.txt`, `summary/labels.txt`,
`summary/higher_layer_group_tcs.txt`.  Plots and graphviz export (vis_corex.py:96-210, :252-411) are out of scope.
"""
import csv
import io
import os
import pickle
import sys
from optparse import OptionGroup, OptionParser
from time import time

import numpy as np


def load_csv(filename, delimiter=",", no_column_names=False, no_row_names=False):
    """vis_corex.py:494-512 (universal newlines so CR-only files parse)."""
    with open(filename, "r", newline=None) as fh:
        text = fh.read()
    reader = csv.reader(io.StringIO(text), delimiter=delimiter)
    skip = 0 if no_row_names else 1
    variable_names = None if no_column_names else next(reader)[skip:]
    sample_names, data = ([] if not no_row_names else None), []
    for row in reader:
        if not row:
            continue
        if sample_names is not None:
            sample_names.append(row[0])
        data.append(row[skip:])
    return np.array(data, dtype=float), sample_names, variable_names


def _open(path, mode):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    return open(path, mode)


def write_groups(model, column_label, prefix):
    """output_groups (vis_corex.py:55-84): membership by explained variance (alpha) and by argmax |W|."""
    ws, moments, mis = model.ws, model.moments, model.mis
    tcs = moments["TCs"]
    alpha = (moments["rho"] * moments["X_i Z_j"].T) > 0.05  # vis_rep, vis_corex.py:36-37
    m = ws.shape[0]
    owner = np.argmax(np.abs(ws), axis=0)
    with _open(prefix + "/summary/groups.txt", "w") as f, _open(prefix + "/summary/groups_no_overlaps.txt", "w") as g, \
            _open(prefix + "/summary/summary.txt", "w") as h:
        h.write("Group, TC\n")
        f.write("variable, weight, MI\n")
        g.write("variable, weight, MI\n")
        for j in range(m):
            f.write("Group num: %d, TC(X;Y_j): %0.6f\n" % (j, tcs[j]))
            g.write("Group num: %d, TC(X;Y_j): %0.6f\n" % (j, tcs[j]))
            h.write("%d, %0.6f\n" % (j, tcs[j]))
            for out, members in ((f, np.where(alpha[j] > 0)[0]), (g, np.where(owner == j)[0])):
                for ind in members[np.argsort(-np.abs(ws)[j][members])]:
                    out.write(column_label[ind] + ", {:.3f}, {:.3f}\n".format(ws[j][ind], mis[j][ind]))
        h.write("Total: {:f}\n".format(np.sum(tcs)))
        h.write("The total of individual TCs should approximately equal the objective: {:f}\n".format(moments["TC"]))
        h.write("If not, this signals redundancy/synergy in the final solution (measured by additivity: {:f}".format(
            moments["additivity"]))


def write_labels(labels, row_label, prefix):
    """output_labels (vis_corex.py:87-92)."""
    with _open(prefix + "/summary/labels.txt", "w") as f:
        for name, row in zip(row_label, labels):
            f.write(name + "," + ",".join(map(str, row)) + "\n")


def write_hierarchy_summary(models, prefix):
    """The text part of vis_hierarchy (vis_corex.py:219-225)."""
    with _open(prefix + "/summary/higher_layer_group_tcs.txt", "w") as f:
        for j, mdl in enumerate(models):
            f.write("At layer: %d, Total TC: %0.3f\n" % (j, mdl.tc))
            f.write("Individual TCS:" + str(mdl.tcs) + "\n")


def build_parser():
    parser = OptionParser(usage="usage: %prog [options] data_file.csv \n"
                                "It is assumed that the first row and first column of the data CSV file are labels.\n"
                                "Use options to indicate otherwise.")
    group = OptionGroup(parser, "Input Data Format Options")
    group.add_option("-t", "--no_column_names", action="store_true", dest="nc", default=False)
    group.add_option("-f", "--no_row_names", action="store_true", dest="nr", default=False)
    group.add_option("-m", "--missing", action="store", dest="missing", type="float", default=-1e6)
    group.add_option("-d", "--delimiter", action="store", dest="delimiter", type="string", default=",")
    group.add_option("-g", "--gaussianize", action="store", dest="gaussianize", type="string", default="standard")
    parser.add_option_group(group)
    group = OptionGroup(parser, "CorEx Options")
    group.add_option("-l", "--layers", dest="layers", type="string", default="2,1")
    group.add_option("-w", "--max_iter", action="store", dest="max_iter", type="int", default=10000)
    group.add_option("-a", "--additive", action="store_false", dest="additive", default=True)
    group.add_option("-p", "--precision", action="store", dest="precision", type="string", default="auto",
                     help="auto (fp64_split, or fp64 for tiny problems), fp64 (DMMA), fp64_split (int8 digit products on tcgen05, same 1e-9 parity; fp64_split5 / fp64_split7 = "
                          "5 / 7 digits) or fast")
    parser.add_option_group(group)
    group = OptionGroup(parser, "Computational Options")
    group.add_option("-n", "--gpu", action="store_true", dest="gpu", default=False, help="accepted for compatibility")
    parser.add_option_group(group)
    group = OptionGroup(parser, "Output Options")
    group.add_option("-o", "--output", action="store", dest="output", type="string", default="corex_output")
    group.add_option("-v", "--verbose", action="store", dest="verbose", type="int", default=0)
    group.add_option("-e", "--edges", action="store", dest="max_edges", type="int", default=200)
    group.add_option("-q", "--regraph", action="store_true", dest="regraph", default=False)
    parser.add_option_group(group)
    return parser


def main(argv=None):
    from .hierarchy import fit_layers
    options, args = build_parser().parse_args(argv)
    if len(args) != 1:
        print("Run with '-h' option for usage help.")
        return 1
    np.set_printoptions(precision=3, suppress=True)
    layers = list(map(int, options.layers.split(",")))
    X, sample_names, variable_names = load_csv(args[0], options.delimiter, options.nc, options.nr)
    if options.verbose:
        print("\nData summary: X has %d rows and %d columns" % X.shape)
    n_layers = len(layers) + (0 if layers[-1] == 1 else 1)
    if not options.regraph:
        t0 = time()
        models = fit_layers(X, layers, missing_values=options.missing, gaussianize=options.gaussianize,
                            discourage_overlap=options.additive, max_iter=options.max_iter, verbose=options.verbose,
                            precision=options.precision)
        print("Time for all layers: %0.2f" % (time() - t0))
        for l, mdl in enumerate(models):
            print("TC at layer %d is: %0.3f" % (l, mdl.tc))
            with _open(options.output + "/layer_" + str(l) + ".dat", "wb") as fh:
                pickle.dump(mdl, fh)
    else:  # the reference opens these files in text mode and fails on python 3 (vis_corex.py:551); binary mode here
        models = [pickle.load(open(options.output + "/layer_" + str(l) + ".dat", "rb")) for l in range(n_layers)]
    column_label = variable_names if variable_names is not None else list(map(str, range(X.shape[1])))
    row_label = sample_names if sample_names is not None else list(map(str, range(len(X))))
    print("Variable groups in summary/groups.txt")
    write_groups(models[0], column_label, options.output)
    print("Latent factors for each sample in summary/labels.txt")
    write_labels(models[0].transform(X), row_label, options.output)
    write_hierarchy_summary(models, options.output)
    return 0


if __name__ == "__main__":
    sys.exit(main())
