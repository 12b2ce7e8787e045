"""CPU checks of table arithmetic the EAM kernels rely on (no GPU): the generation-3 force pass (eam_variant = 2,
md-bench_b200/csrc/eam_kernels.cuh) keeps only (value, slope) per knot and derives the cubic's coefficients in registers;
that derivation must reproduce the reference's 7-coefficient spline rows (common/eam_utils.c:253-284, here the tables the
reference itself produced for data/Cu_u3.eam, fixture tests/golden/eam_cu_nx5.npz) to rounding."""
import os

import numpy as np
import pytest

from conftest import ROOT


@pytest.mark.parametrize("name", ["eam_rhor_spline", "eam_z2r_spline"])
def test_value_slope_rows_reproduce_the_reference_spline_rows(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "eam_cu_nx5.npz"))
    nr, rdr = int(g["eam_nr"]), float(g["eam_rdr"])
    a = g[name]
    rows = len(a) // 7
    S = a[:rows * 7].reshape(rows, 7)
    assert rows >= nr + 1                     # the force pass reads the knots m and m + 1 with m <= nr - 1
    m = np.arange(1, nr)
    f0, s0, f1, s1 = S[m, 6], S[m, 5], S[m + 1, 6], S[m + 1, 5]
    d = f1 - f0
    c4 = 3.0 * d - 2.0 * s0 - s1              # eam_utils.c:269-271
    c3 = s0 + s1 - 2.0 * d                    # eam_utils.c:272-273
    scale = np.abs(S[m, 3:7]).max()
    assert np.abs(c3 - S[m, 3]).max() <= 4e-16 * scale and np.abs(c4 - S[m, 4]).max() <= 4e-16 * scale
    # value and derivative at random points of every interval: table form (force_eam.c:170-180) vs the derived form
    p = np.random.default_rng(7).uniform(0.0, 1.0, len(m))
    val_ref = ((S[m, 3] * p + S[m, 4]) * p + S[m, 5]) * p + S[m, 6]
    val_new = ((c3 * p + c4) * p + s0) * p + f0
    der_ref = (S[m, 0] * p + S[m, 1]) * p + S[m, 2]
    der_new = ((3.0 * c3 * p + 2.0 * c4) * p + s0) * rdr
    assert np.abs(val_new - val_ref).max() <= 1e-15 * np.abs(val_ref).max()
    assert np.abs(der_new - der_ref).max() <= 2e-15 * np.abs(der_ref).max()
