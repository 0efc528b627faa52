"""Re-derives the expected values in reference_vectors.json that are computable without the
reference (scipy / closed form) and fails if the committed literals disagree.  The literals
themselves are transcribed from the reference's tests (each entry cites file:line); the reference
cannot be imported or run in this image (Julia + TensorFlow 1.15), so there is nothing to execute.

    python tests/golden/make_golden.py
"""
import json
import os

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    g = json.load(open(os.path.join(HERE, "reference_vectors.json")))
    c = g["compress"]
    idx = np.array(c["indices_1based"]) - 1
    A = sp.coo_matrix((c["v"], (idx[:, 0], idx[:, 1])), shape=(3, 3)).tocsr()
    A.sum_duplicates()
    coo = A.tocoo()
    assert np.array_equal(np.stack([coo.row, coo.col], 1), np.array(c["expect_indices_0based"]))
    assert np.allclose(coo.data, c["expect_values"])
    for key in ("assembler_tol_filter", "assembler_duplicates", "assembler_docstring_example2"):
        e = g[key]
        D = np.zeros((e["m"], e["ncol"]))
        for a in e["adds"]:
            for col, v in zip(a["cols"], a["vals"]):
                if abs(v) > e["tol"]:
                    D[a["row"] - 1, col - 1] += v
        assert np.allclose(D, np.array(e["expect_dense"], dtype=float)), key
    n = g["newton_cuberoot_adjoint"]
    th = np.array(n["theta"])
    assert np.allclose(np.cbrt(th), n["expect_x"]) and np.allclose(th ** (-2 / 3) / 3, n["expect_grad"])
    ls = g["least_squares_adjacent"]                               # test/sparse.jl:131-139
    A = sp.coo_matrix((ls["vv"], (np.array(ls["ii"]) - 1, np.array(ls["jj"]) - 1)), shape=(ls["m"], ls["n"])).toarray()
    assert np.allclose(np.linalg.lstsq(A, np.array(ls["ff"], float), rcond=None)[0], ls["expect"], atol=ls["tol"])
    gi = g["get_index_bool_mask"]                                  # test/sparse.jl:316-321
    M = np.diag(gi["diag"])
    sel = np.flatnonzero(gi["idof"])
    assert M[np.ix_(sel, sel)].tolist() == gi["expect_dense"]
    print("golden vectors consistent")


if __name__ == "__main__":
    main()
