"""Host-side mirror of the reference interface (acquisition constructors and `update`, ensemble weights, domains, surrogate
accessors) with the reference's own test vectors
(test/test_acquisition.jl:10-18, 45-64, 67-72, 97-114, 117-124, 152-157, 183-201, 204-221, 255-277).  No device needed:
`update(acq, ys, surrogate)` only looks at the surrogate's TYPE (StandardGP.jl:418, GradientGP.jl:1044)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def abo():
    import __graft_entry__ as g
    g.build()
    import abo_b200
    return abo_b200


def test_constructors_keep_their_fields(abo):
    ei = abo.ExpectedImprovement(0.01, 1.0)
    assert ei.xi == 0.01 and ei.best_y == 1.0
    pi = abo.ProbabilityImprovement(0.01, 1.0)
    assert pi.xi == 0.01 and pi.best_y == 1.0
    assert abo.UpperConfidenceBound(2.0).beta == 2.0
    assert abo.GradientNormUCB(2.0).beta == 2.0


def test_update_takes_the_minimum_of_the_new_data(abo):
    gp = abo.StandardGP(abo.SqExponentialKernel(), 0.1)
    ei = abo.ExpectedImprovement(0.01, 1.0)
    up = ei.update([2.0, 1.5, 0.8], gp)
    assert up.xi == 0.01 and up.best_y == 0.8 and ei.best_y == 1.0          # functional: the old object is untouched
    up = abo.ProbabilityImprovement(0.01, 1.0).update([2.0, 1.5, 0.8], gp)
    assert up.best_y == 0.8
    ucb = abo.UpperConfidenceBound(2.0)
    assert ucb.update([1.0, 2.0], gp) is ucb                                  # "updated_ucb === ucb"
    ggp = abo.GradientGP(abo.SqExponentialKernel(), 3, 0.1)
    gucb = abo.GradientNormUCB(2.0)
    assert gucb.update([[1.0, 0.1, 0.1], [2.0, 0.2, 0.2]], ggp) is gucb
    # a GradientGP's incumbent is the minimum of the VALUE output only (GradientGP.jl:1044)
    assert abo.ExpectedImprovement(0.0, 9.0).update([[1.0, -5.0, 0.1], [2.0, 0.2, -7.0]], ggp).best_y == 1.0


def test_ensemble_weights_and_update(abo):
    ei, ucb = abo.ExpectedImprovement(0.01, 1.0), abo.UpperConfidenceBound(2.0)
    ens = abo.EnsembleAcquisition([0.5, 0.5], [ei, ucb])
    assert np.allclose(ens.weights, [0.5, 0.5]) and len(ens.acquisitions) == 2
    assert np.allclose(abo.EnsembleAcquisition([1.0, 3.0], [ei, ucb]).weights, [0.25, 0.75])
    gp = abo.StandardGP(abo.SqExponentialKernel(), 0.1)
    up = ens.update([2.0, 1.5, 0.8], gp)
    assert np.array_equal(up.weights, ens.weights) and len(up.acquisitions) == 2
    assert up.acquisitions[0].best_y == 0.8 and up.acquisitions[1] is ucb
    with pytest.raises(ValueError):
        abo.EnsembleAcquisition([0.5], [ei, ucb])
    with pytest.raises(ValueError):
        abo.EnsembleAcquisition([-1.0, 2.0], [ei, ucb])
    with pytest.raises(ValueError):
        abo.EnsembleAcquisition([0.0, 0.0], [ei, ucb])


def test_continuous_domain(abo):
    """test/test_domains.jl:6-44 (ArgumentError -> ValueError in the Python mirror)."""
    dom = abo.ContinuousDomain([0.0, -1.0], [1.0, 1.0])
    assert list(dom.lower) == [0.0, -1.0] and list(dom.upper) == [1.0, 1.0] and dom.bounds == [(0.0, 1.0), (-1.0, 1.0)]
    assert abo.ContinuousDomain([0.0], [1.0]).bounds == [(0.0, 1.0)]
    for lo, hi in (([0.0, 1.0], [1.0]), ([0.0], [1.0, 2.0]), ([1.0], [0.0]), ([0.0, 2.0], [1.0, 1.0])):
        with pytest.raises(ValueError):
            abo.ContinuousDomain(lo, hi)
    eq = abo.ContinuousDomain([1.0], [1.0])                                 # equal bounds are valid
    assert list(eq.lower) == [1.0] and list(eq.upper) == [1.0]
    assert len(abo.ContinuousDomain([1e6], [1e7]).bounds) == 1


def test_surrogate_construction_vectors(abo):
    """test/test_surrogates.jl:10-57 (StandardGP) and :174-233 (GradientGP): accessors of freshly constructed priors."""
    for make in (lambda k: abo.StandardGP(k, 0.1), lambda k: abo.GradientGP(k, 3, 0.1)):
        base = abo.SqExponentialKernel() if make(abo.SqExponentialKernel()).p == 1 else abo.ApproxMatern52Kernel()
        gp = make(base)
        assert gp.noise_var == 0.1 and gp.gpx is None
        assert abo.get_lengthscale(gp)[0] == 1.0 and abo.get_scale(gp)[0] == 1.0
        gp_custom = make(2.0 * abo.with_lengthscale(base, 0.5))
        assert abo.get_lengthscale(gp_custom) == [0.5] and abo.get_scale(gp_custom) == [2.0]
        assert gp_custom.noise_var == 0.1 and gp_custom.gpx is None
        gp_ls = make(abo.with_lengthscale(base, 0.3))
        assert abo.get_lengthscale(gp_ls) == [0.3] and abo.get_scale(gp_ls) == [1.0] and gp_ls.gpx is None
        gp_sc = make(3.0 * base)
        assert abo.get_lengthscale(gp_sc) == [1.0] and abo.get_scale(gp_sc) == [3.0] and gp_sc.gpx is None
    assert abo.GradientGP(abo.ApproxMatern52Kernel(), 3, 0.1).p == 3


def test_rescale_output_and_print_info(abo, capsys):
    """rescale_output (BO_utils.jl:162-182) and print_info / show (BO_utils.jl:5-24)."""
    assert abo.rescale_output([0.5, -1.0], (3.0, 2.0)) == [4.0, 1.0]
    out = abo.rescale_output([np.array([1.0, 2.0, 3.0])], (np.array([1.0, 0.0, 0.0]), np.array([2.0, 2.0, 2.0])))
    assert np.array_equal(out[0], [3.0, 4.0, 6.0])
    assert abo.rescale_output([1.0, 2.0], (None, None)) == [1.0, 2.0]
    f = lambda x: float(np.sum(np.asarray(x) ** 2))
    bo = abo.BOStruct(f, abo.ExpectedImprovement(0.01, 0.0), abo.StandardGP(abo.SqExponentialKernel(), 0.1),
                      abo.ContinuousDomain([-2.0], [2.0]), [-1.0, 0.0, 1.0], [1.0, 0.0, 1.0], 3, 0.1)
    abo.print_info(bo)
    text = capsys.readouterr().out
    for line in ("== BOStruct Information ==", "Number of data points: 3", "Max iterations: 3", "Noise level: 0.1"):
        assert line in text
    assert repr(bo).startswith("== BOStruct Information ==")
